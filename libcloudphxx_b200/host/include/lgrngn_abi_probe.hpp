// Describes, as text, everything of the lgrngn API that a caller and the library must agree on at the BINARY level:
// sizes and member offsets of opts_init_t / opts_t / arrinfo_t, the values of the enumerations, and the v-table slot of every
// virtual method of particles_proto_t (Itanium C++ ABI: a pointer to a virtual member function stores 1 + its byte offset in
// the v-table).  Compiled against the REFERENCE's headers (tests, in the development container) and against this
// library's own headers (inside liblgrngn_b200.so: lgrngn_b200_abi_layout()); the two texts must be equal, otherwise a model
// built with the reference's headers would call wrong v-table slots / read wrong option fields when linked to this library.
#pragma once
#include <cstddef>
#include <cstring>
#include <sstream>
#include <string>

#include <libcloudph++/lgrngn/factory.hpp>

namespace lgrngn_abi_probe
{
  namespace lg = libcloudphxx::lgrngn;

  template <class M>
  inline long vslot(M pmf)
  {
    long v[2] = {0, 0};
    static_assert(sizeof(pmf) <= sizeof(v), "pointer to member function larger than expected");
    std::memcpy(v, &pmf, sizeof(pmf));
    return (v[0] & 1) ? (v[0] - 1) / long(sizeof(void *)) : -1;     // -1: not virtual
  }

#define LGP_OFF(T, m) os << " " #m "@" << offsetof(T, m)
#define LGP_SLOT0(name) os << " " #name "=" << vslot(static_cast<void (P::*)()>(&P::name))
#define LGP_SLOTI(name) os << " " #name "=" << vslot(static_cast<void (P::*)(const int &)>(&P::name))
#define LGP_SLOTRR(name) os << " " #name "=" << vslot(static_cast<void (P::*)(const real_t &, const real_t &)>(&P::name))

#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Winvalid-offsetof"
  template <class real_t>
  inline void describe(std::ostream &os, const char *tag)
  {
    typedef lg::opts_init_t<real_t> OI;
    typedef lg::opts_t<real_t> O;
    typedef lg::arrinfo_t<real_t> A;
    typedef lg::particles_proto_t<real_t> P;
    os << tag << " opts_init_t size=" << sizeof(OI);
    LGP_OFF(OI, dry_distros); LGP_OFF(OI, dry_sizes); LGP_OFF(OI, nx); LGP_OFF(OI, ny); LGP_OFF(OI, nz); LGP_OFF(OI, dx); LGP_OFF(OI, dy);
    LGP_OFF(OI, dz); LGP_OFF(OI, dt); LGP_OFF(OI, sstp_cond); LGP_OFF(OI, sstp_coal); LGP_OFF(OI, sstp_cond_act); LGP_OFF(OI, sstp_chem);
    LGP_OFF(OI, x0); LGP_OFF(OI, y0); LGP_OFF(OI, z0); LGP_OFF(OI, x1); LGP_OFF(OI, y1); LGP_OFF(OI, z1); LGP_OFF(OI, sd_conc);
    LGP_OFF(OI, sd_conc_large_tail); LGP_OFF(OI, aerosol_independent_of_rhod); LGP_OFF(OI, variable_dt_switch); LGP_OFF(OI, sd_const_multi);
    LGP_OFF(OI, n_sd_max); LGP_OFF(OI, kernel); LGP_OFF(OI, terminal_velocity); LGP_OFF(OI, adve_scheme); LGP_OFF(OI, RH_formula);
    LGP_OFF(OI, kernel_parameters); LGP_OFF(OI, chem_switch); LGP_OFF(OI, coal_switch); LGP_OFF(OI, sedi_switch); LGP_OFF(OI, subs_switch);
    LGP_OFF(OI, rlx_switch); LGP_OFF(OI, turb_adve_switch); LGP_OFF(OI, turb_cond_switch); LGP_OFF(OI, turb_coal_switch); LGP_OFF(OI, ice_switch);
    LGP_OFF(OI, exact_sstp_cond); LGP_OFF(OI, sstp_cond_mix); LGP_OFF(OI, adaptive_sstp_cond); LGP_OFF(OI, time_dep_ice_nucl);
    LGP_OFF(OI, sstp_cond_adapt_drw2_eps); LGP_OFF(OI, sstp_cond_adapt_drw2_max); LGP_OFF(OI, inp_type); LGP_OFF(OI, chem_rho);
    LGP_OFF(OI, diag_incloud_time); LGP_OFF(OI, RH_max); LGP_OFF(OI, rng_seed); LGP_OFF(OI, rng_seed_init); LGP_OFF(OI, rng_seed_init_switch);
    LGP_OFF(OI, dev_count); LGP_OFF(OI, dev_id); LGP_OFF(OI, w_LS); LGP_OFF(OI, SGS_mix_len); LGP_OFF(OI, aerosol_conc_factor);
    LGP_OFF(OI, rd_min); LGP_OFF(OI, rd_max); LGP_OFF(OI, no_ccn_at_init); LGP_OFF(OI, open_side_walls); LGP_OFF(OI, periodic_topbot_walls);
    LGP_OFF(OI, rc2_T); LGP_OFF(OI, src_type); LGP_OFF(OI, src_x0); LGP_OFF(OI, src_y0); LGP_OFF(OI, src_z0); LGP_OFF(OI, src_x1);
    LGP_OFF(OI, src_y1); LGP_OFF(OI, src_z1); LGP_OFF(OI, rlx_dry_distros); LGP_OFF(OI, rlx_bins); LGP_OFF(OI, rlx_sd_per_bin);
    LGP_OFF(OI, supstp_rlx); LGP_OFF(OI, rlx_timescale); LGP_OFF(OI, th_dry); LGP_OFF(OI, const_p);
    os << "\n" << tag << " opts_t size=" << sizeof(O);
    LGP_OFF(O, adve); LGP_OFF(O, sedi); LGP_OFF(O, subs); LGP_OFF(O, cond); LGP_OFF(O, coal); LGP_OFF(O, src); LGP_OFF(O, rlx); LGP_OFF(O, rcyc);
    LGP_OFF(O, turb_adve); LGP_OFF(O, turb_cond); LGP_OFF(O, turb_coal); LGP_OFF(O, ice_nucl); LGP_OFF(O, RH_max); LGP_OFF(O, chem_dsl);
    LGP_OFF(O, chem_dsc); LGP_OFF(O, chem_rct); LGP_OFF(O, dt); LGP_OFF(O, src_dry_distros); LGP_OFF(O, src_dry_sizes);
    os << "\n" << tag << " arrinfo_t size=" << sizeof(A);
    LGP_OFF(A, data); LGP_OFF(A, strides);
    os << "\n" << tag << " particles_proto_t size=" << sizeof(P);
    LGP_OFF(P, opts_init);
    os << " vslots:";
    os << " init=" << vslot(&P::init) << " step_sync=" << vslot(&P::step_sync) << " sync_in=" << vslot(&P::sync_in)
       << " step_cond=" << vslot(&P::step_cond) << " step_async=" << vslot(&P::step_async);
    LGP_SLOT0(diag_sd_conc); LGP_SLOT0(diag_pressure); LGP_SLOT0(diag_temperature); LGP_SLOT0(diag_RH); LGP_SLOT0(diag_all);
    LGP_SLOT0(diag_rw_ge_rc); LGP_SLOT0(diag_RH_ge_Sc); LGP_SLOTRR(diag_dry_rng); LGP_SLOTRR(diag_wet_rng); LGP_SLOTRR(diag_ice_a_rng);
    LGP_SLOTRR(diag_ice_c_rng); LGP_SLOTRR(diag_kappa_rng); LGP_SLOT0(diag_ice); LGP_SLOT0(diag_water); LGP_SLOTRR(diag_dry_rng_cons);
    LGP_SLOTRR(diag_wet_rng_cons); LGP_SLOTRR(diag_ice_a_rng_cons); LGP_SLOTRR(diag_ice_c_rng_cons); LGP_SLOTRR(diag_kappa_rng_cons);
    LGP_SLOT0(diag_ice_cons); LGP_SLOT0(diag_water_cons); LGP_SLOTI(diag_dry_mom); LGP_SLOTI(diag_wet_mom); LGP_SLOTI(diag_ice_a_mom);
    LGP_SLOTI(diag_ice_c_mom); LGP_SLOT0(diag_ice_mix_ratio); LGP_SLOTRR(diag_wet_mass_dens);
    os << " diag_chem=" << vslot(&P::diag_chem);
    LGP_SLOT0(diag_precip_rate); LGP_SLOT0(diag_precip_rate_ice_mass); LGP_SLOTI(diag_kappa_mom); LGP_SLOTI(diag_up_mom); LGP_SLOTI(diag_vp_mom);
    LGP_SLOTI(diag_wp_mom); LGP_SLOTI(diag_incloud_time_mom); LGP_SLOT0(diag_max_rw); LGP_SLOT0(diag_vel_div);
    os << " diag_puddle=" << vslot(&P::diag_puddle) << " get_attr=" << vslot(&P::get_attr) << " outbuf=" << vslot(&P::outbuf);
    os << "\n";
  }
#pragma GCC diagnostic pop

  inline std::string describe_all()
  {
    std::ostringstream os;
    os << "enums backend:" << int(lg::serial) << int(lg::OpenMP) << int(lg::CUDA) << int(lg::multi_CUDA)
       << " kernel:" << int(lg::kernel_t::geometric) << "," << int(lg::kernel_t::golovin) << "," << int(lg::kernel_t::hall) << ","
       << int(lg::kernel_t::hall_davis_no_waals) << "," << int(lg::kernel_t::Long) << "," << int(lg::kernel_t::onishi_hall) << ","
       << int(lg::kernel_t::onishi_hall_davis_no_waals) << "," << int(lg::kernel_t::hall_pinsky_1000mb_grav) << ","
       << int(lg::kernel_t::hall_pinsky_cumulonimbus) << "," << int(lg::kernel_t::hall_pinsky_stratocumulus) << "," << int(lg::kernel_t::vohl_davis_no_waals)
       << " vt:" << int(lg::vt_t::beard76) << int(lg::vt_t::beard77) << int(lg::vt_t::beard77fast) << int(lg::vt_t::khvorostyanov_spherical)
       << int(lg::vt_t::khvorostyanov_nonspherical)
       << " as:" << int(lg::as_t::implicit) << int(lg::as_t::euler) << int(lg::as_t::pred_corr)
       << " RH:" << int(lg::RH_formula_t::pv_cc) << int(lg::RH_formula_t::rv_cc) << int(lg::RH_formula_t::pv_tet) << int(lg::RH_formula_t::rv_tet)
       << " src:" << int(lg::src_t::off) << int(lg::src_t::simple) << int(lg::src_t::matching)
       << " sizes:" << sizeof(lg::kernel_t) << sizeof(lg::vt_t) << sizeof(lg::as_t) << sizeof(lg::RH_formula_t) << sizeof(lg::src_t) << sizeof(lg::backend_t)
       << "\n";
    describe<float>(os, "float");
    describe<double>(os, "double");
    return os.str();
  }
}
