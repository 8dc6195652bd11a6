// Public interface of the B200-native Lagrangian (super-droplet) microphysics back-end.
//
// This header restates - in one place and with this project's own layout - the caller-facing
// types of libcloudph++'s `lgrngn` component so that a host model written against the
// reference compiles unchanged against this library:
//   enums          <-> reference include/libcloudph++/lgrngn/{backend,kernel,terminal_velocity,
//                      advection_scheme,RH_formula,ccn_source}.hpp
//   arrinfo_t      <-> lgrngn/arrinfo.hpp:11-49
//   distro typedefs<-> lgrngn/distro_t.hpp:10-57
//   opts_t         <-> lgrngn/opts.hpp:20-50
//   opts_init_t    <-> lgrngn/opts_init.hpp:29-253   (same member names / types / defaults)
//   particles_proto_t, factory <-> lgrngn/particles.hpp:17-134, lgrngn/factory.hpp:12-15
// Only names, types, defaults and documented behaviour are shared with the reference; the
// implementation behind `factory` is hand-written sm_100a CUDA reached through the C ABI
// declared in include/lcx_b200.h.
#pragma once

#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../common/unary_function.hpp"
#include "../common/chem.hpp"
#include "../common/output.hpp"

namespace libcloudphxx
{
  namespace common { namespace ice_nucleation { enum class INP_t { mineral }; } }

  namespace lgrngn
  {
    using common::unary_function;
    using common::ice_nucleation::INP_t;

    // ---- enumerations (values and order as in the reference) ---------------------------------
    enum backend_t { undefined, serial, OpenMP, CUDA, multi_CUDA };

    enum class kernel_t {
      undefined, geometric, golovin, hall, hall_davis_no_waals, Long, onishi_hall,
      onishi_hall_davis_no_waals, hall_pinsky_1000mb_grav, hall_pinsky_cumulonimbus,
      hall_pinsky_stratocumulus, vohl_davis_no_waals
    };
    enum class vt_t { undefined, beard76, beard77, beard77fast, khvorostyanov_spherical, khvorostyanov_nonspherical };
    enum class as_t { undefined, implicit, euler, pred_corr };
    enum class RH_formula_t { pv_cc, rv_cc, pv_tet, rv_tet };
    enum class src_t { off, simple, matching };

    const std::unordered_map<backend_t, std::string> backend_name = {
      {undefined, "undefined"}, {serial, "serial"}, {OpenMP, "OpenMP"}, {CUDA, "CUDA"}, {multi_CUDA, "multi_CUDA"}};
    const std::unordered_map<kernel_t, std::string> kernel_name = {
      {kernel_t::undefined, "undefined"}, {kernel_t::geometric, "geometric"}, {kernel_t::golovin, "golovin"},
      {kernel_t::hall, "hall"}, {kernel_t::hall_davis_no_waals, "hall_davis_no_waals"}, {kernel_t::Long, "Long"},
      {kernel_t::onishi_hall, "onishi_hall"}, {kernel_t::onishi_hall_davis_no_waals, "onishi_hall_davis_no_waals"},
      {kernel_t::hall_pinsky_1000mb_grav, "hall_pinsky_1000mb_grav"},
      {kernel_t::hall_pinsky_cumulonimbus, "hall_pinsky_cumulonimbus"},
      {kernel_t::hall_pinsky_stratocumulus, "hall_pinsky_stratocumulus"},
      {kernel_t::vohl_davis_no_waals, "vohl_davis_no_waals"}};
    const std::unordered_map<vt_t, std::string> vt_name = {
      {vt_t::undefined, "undefined"}, {vt_t::beard76, "beard76"}, {vt_t::beard77, "beard77"},
      {vt_t::beard77fast, "beard77fast"}, {vt_t::khvorostyanov_spherical, "khvorostyanov_spherical"},
      {vt_t::khvorostyanov_nonspherical, "khvorostyanov_nonspherical"}};
    const std::unordered_map<as_t, std::string> as_name = {
      {as_t::undefined, "undefined"}, {as_t::implicit, "implicit"}, {as_t::euler, "euler"}, {as_t::pred_corr, "pred_corr"}};
    const std::unordered_map<RH_formula_t, std::string> RH_formula_name = {
      {RH_formula_t::pv_cc, "pv_cc"}, {RH_formula_t::rv_cc, "rv_cc"}, {RH_formula_t::pv_tet, "pv_tet"}, {RH_formula_t::rv_tet, "rv_tet"}};
    const std::unordered_map<src_t, std::string> src_name = {
      {src_t::off, "off"}, {src_t::simple, "simple"}, {src_t::matching, "matching"}};

    // ---- n-dimensional array descriptor: raw pointer + element strides (x[,y],z; z contiguous) ----
    template <typename real_t>
    struct arrinfo_t
    {
      const std::vector<ptrdiff_t> strvec;   // optional local copy of the strides
      real_t * const data;
      const ptrdiff_t *strides;

      arrinfo_t() : data(nullptr), strides(nullptr) {}
      arrinfo_t(real_t * const data_, const ptrdiff_t *strides_) : data(data_), strides(strides_) {}
      arrinfo_t(real_t * const data_, const std::vector<ptrdiff_t> &sv) : strvec(sv), data(data_), strides(strvec.data()) {}
      arrinfo_t(const arrinfo_t &o) : strvec(o.strvec), data(o.data), strides(strvec.empty() ? o.strides : strvec.data()) {}
      arrinfo_t(arrinfo_t &&o) : strvec(std::move(o.strvec)), data(o.data), strides(strvec.empty() ? o.strides : strvec.data()) {}

      bool is_null() const { return data == nullptr || strides == nullptr; }
    };

    // ---- dry-aerosol spectra ---------------------------------------------------------------------
    template <typename real_t>
    struct kappa_rd_insol_t
    {
      real_t kappa, rd_insol;
      kappa_rd_insol_t(real_t k, real_t r) : kappa(k), rd_insol(r) {}
      bool operator<(const kappa_rd_insol_t &o) const { return kappa != o.kappa ? kappa < o.kappa : rd_insol < o.rd_insol; }
    };
    // (kappa, rd_insol) -> n(ln rd) at STP
    template <typename real_t>
    using dry_distros_t = std::map<kappa_rd_insol_t<real_t>, std::shared_ptr<unary_function<real_t>>>;
    // (kappa, rd_insol) -> { radius -> (STP concentration, number of SDs) }
    template <typename real_t>
    using dry_sizes_t = std::map<kappa_rd_insol_t<real_t>, std::map<real_t, std::pair<real_t, int>>>;
    template <typename real_t>
    using src_dry_distros_t = std::map<kappa_rd_insol_t<real_t>, std::tuple<std::shared_ptr<unary_function<real_t>>, int, int>>;
    template <typename real_t>
    using src_dry_sizes_t = std::map<kappa_rd_insol_t<real_t>, std::map<real_t, std::tuple<real_t, int, int>>>;

    // ---- per-step options -----------------------------------------------------------------------
    template <typename real_t>
    struct opts_t
    {
      bool adve = true, sedi = true, subs = false, cond = true, coal = true, src = false, rlx = false, rcyc = false,
           turb_adve = false, turb_cond = false, turb_coal = false, ice_nucl = false;
      real_t RH_max = 44;            // cap on RH seen by condensation (anything > 1.1 means "no cap")
      bool chem_dsl = false, chem_dsc = false, chem_rct = false;
      real_t dt = -1;                // > 0 overrides opts_init.dt for this step
      src_dry_distros_t<real_t> src_dry_distros;
      src_dry_sizes_t<real_t> src_dry_sizes;
    };

    // ---- construction-time options ---------------------------------------------------------------
    template <typename real_t>
    struct opts_init_t
    {
      dry_distros_t<real_t> dry_distros;
      dry_sizes_t<real_t> dry_sizes;

      int nx = 0, ny = 0, nz = 0;
      real_t dx = 1, dy = 1, dz = 1, dt = 0;
      int sstp_cond = 1, sstp_coal = 1, sstp_cond_act = 1, sstp_chem = 1;
      real_t x0 = 0, y0 = 0, z0 = 0, x1 = 1, y1 = 1, z1 = 1;

      unsigned long long sd_conc = 0;
      bool sd_conc_large_tail = false, aerosol_independent_of_rhod = false, variable_dt_switch = false;
      unsigned long long sd_const_multi = 0, n_sd_max = 0;

      kernel_t kernel = kernel_t::undefined;
      vt_t terminal_velocity = vt_t::undefined;
      as_t adve_scheme = as_t::implicit;
      RH_formula_t RH_formula = RH_formula_t::pv_cc;
      std::vector<real_t> kernel_parameters;

      bool chem_switch = false, coal_switch = true, sedi_switch = true, subs_switch = false, rlx_switch = false,
           turb_adve_switch = false, turb_cond_switch = false, turb_coal_switch = false, ice_switch = false,
           exact_sstp_cond = false, sstp_cond_mix = true, adaptive_sstp_cond = false, time_dep_ice_nucl = false;
      real_t sstp_cond_adapt_drw2_eps = 1e-4, sstp_cond_adapt_drw2_max = 4;
      INP_t inp_type = INP_t::mineral;
      real_t chem_rho = 0;
      bool diag_incloud_time = false;

      real_t RH_max = real_t(.95);   // RH cap for the equilibrium wet radius at t=0
      int rng_seed = 44, rng_seed_init = 44;
      bool rng_seed_init_switch = false;
      int dev_count = 0;             // multi_CUDA: number of GPUs (0 = all)
      int dev_id = -1;               // CUDA: device ordinal (-1 = current)

      std::vector<real_t> w_LS, SGS_mix_len, aerosol_conc_factor;
      real_t rd_min = -1, rd_max = -1;   // < 0: detect the dry-radius range automatically
      bool no_ccn_at_init = false, open_side_walls = false, periodic_topbot_walls = false;
      real_t rc2_T = 10;

      src_t src_type = src_t::off;
      real_t src_x0 = 0, src_y0 = 0, src_z0 = 0, src_x1 = 0, src_y1 = 0, src_z1 = 0;

      typedef std::unordered_map<real_t, std::tuple<std::shared_ptr<unary_function<real_t>>,
                                                   std::pair<real_t, real_t>, std::pair<real_t, real_t>>> rlx_dry_distros_t;
      rlx_dry_distros_t rlx_dry_distros;
      unsigned long long rlx_bins = 0;
      real_t rlx_sd_per_bin = 0;
      int supstp_rlx = 1;
      real_t rlx_timescale = 1;

      bool th_dry = true, const_p = false;
    };

    // ---- the particle-system interface -----------------------------------------------------------
    template <typename real_t>
    struct particles_proto_t
    {
      typedef std::map<enum common::chem::chem_species_t, arrinfo_t<real_t>> chem_map_t;
      typedef std::map<enum common::chem::chem_species_t, const arrinfo_t<real_t>> chem_cmap_t;

      virtual void init(
        const arrinfo_t<real_t> th, const arrinfo_t<real_t> rv, const arrinfo_t<real_t> rhod,
        const arrinfo_t<real_t> p = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_x = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_y = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_z = arrinfo_t<real_t>(),
        const chem_cmap_t ambient_chem = chem_cmap_t()) { unsupported("init"); }

      virtual void step_sync(
        const opts_t<real_t> &, arrinfo_t<real_t> th, arrinfo_t<real_t> rv,
        const arrinfo_t<real_t> rhod = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_x = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_y = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_z = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> diss_rate = arrinfo_t<real_t>(),
        chem_map_t ambient_chem = chem_map_t()) { unsupported("step_sync"); }

      virtual void sync_in(
        arrinfo_t<real_t> th, arrinfo_t<real_t> rv,
        const arrinfo_t<real_t> rhod = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_x = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_y = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> courant_z = arrinfo_t<real_t>(),
        const arrinfo_t<real_t> diss_rate = arrinfo_t<real_t>(),
        chem_map_t ambient_chem = chem_map_t()) { unsupported("sync_in"); }

      virtual void step_cond(const opts_t<real_t> &, arrinfo_t<real_t> th, arrinfo_t<real_t> rv,
                             chem_map_t ambient_chem = chem_map_t()) { unsupported("step_cond"); }

      virtual void step_async(const opts_t<real_t> &) { unsupported("step_async"); }

      // The ORDER of the virtual methods is the reference's (lgrngn/particles.hpp:20-125) - it fixes the v-table slots, so a host
      // model compiled against the reference's headers calls the right methods when linked to this library; checked by
      // tests/test_cpu_abi.py with lgrngn_abi_probe.hpp (and at run time through lgrngn_b200_abi_layout()).
      // per-cell fields (result is fetched with outbuf())
      virtual void diag_sd_conc()                                    { unsupported("diag_sd_conc"); }
      virtual void diag_pressure()                                   { unsupported("diag_pressure"); }
      virtual void diag_temperature()                                { unsupported("diag_temperature"); }
      virtual void diag_RH()                                         { unsupported("diag_RH"); }
      // selectors (fill the per-SD mask used by the following *_mom / sd_conc call)
      virtual void diag_all()                                        { unsupported("diag_all"); }
      virtual void diag_rw_ge_rc()                                   { unsupported("diag_rw_ge_rc"); }
      virtual void diag_RH_ge_Sc()                                   { unsupported("diag_RH_ge_Sc"); }
      virtual void diag_dry_rng(const real_t&, const real_t&)        { unsupported("diag_dry_rng"); }
      virtual void diag_wet_rng(const real_t&, const real_t&)        { unsupported("diag_wet_rng"); }
      virtual void diag_ice_a_rng(const real_t&, const real_t&)      { unsupported("diag_ice_a_rng"); }
      virtual void diag_ice_c_rng(const real_t&, const real_t&)      { unsupported("diag_ice_c_rng"); }
      virtual void diag_kappa_rng(const real_t&, const real_t&)      { unsupported("diag_kappa_rng"); }
      virtual void diag_ice()                                        { unsupported("diag_ice"); }
      virtual void diag_water()                                      { unsupported("diag_water"); }
      // "consecutive" selectors: refine the mask left by the immediately preceding selector
      virtual void diag_dry_rng_cons(const real_t&, const real_t&)   { unsupported("diag_dry_rng_cons"); }
      virtual void diag_wet_rng_cons(const real_t&, const real_t&)   { unsupported("diag_wet_rng_cons"); }
      virtual void diag_ice_a_rng_cons(const real_t&, const real_t&) { unsupported("diag_ice_a_rng_cons"); }
      virtual void diag_ice_c_rng_cons(const real_t&, const real_t&) { unsupported("diag_ice_c_rng_cons"); }
      virtual void diag_kappa_rng_cons(const real_t&, const real_t&) { unsupported("diag_kappa_rng_cons"); }
      virtual void diag_ice_cons()                                   { unsupported("diag_ice_cons"); }
      virtual void diag_water_cons()                                 { unsupported("diag_water_cons"); }
      // moments (result is fetched with outbuf())
      virtual void diag_dry_mom(const int&)                          { unsupported("diag_dry_mom"); }
      virtual void diag_wet_mom(const int&)                          { unsupported("diag_wet_mom"); }
      virtual void diag_ice_a_mom(const int&)                        { unsupported("diag_ice_a_mom"); }
      virtual void diag_ice_c_mom(const int&)                        { unsupported("diag_ice_c_mom"); }
      virtual void diag_ice_mix_ratio()                              { unsupported("diag_ice_mix_ratio"); }
      virtual void diag_wet_mass_dens(const real_t&, const real_t&)  { unsupported("diag_wet_mass_dens"); }
      virtual void diag_chem(const enum common::chem::chem_species_t&) { unsupported("diag_chem"); }
      virtual void diag_precip_rate()                                { unsupported("diag_precip_rate"); }
      virtual void diag_precip_rate_ice_mass()                       { unsupported("diag_precip_rate_ice_mass"); }
      virtual void diag_kappa_mom(const int&)                        { unsupported("diag_kappa_mom"); }
      virtual void diag_up_mom(const int&)                           { unsupported("diag_up_mom"); }
      virtual void diag_vp_mom(const int&)                           { unsupported("diag_vp_mom"); }
      virtual void diag_wp_mom(const int&)                           { unsupported("diag_wp_mom"); }
      virtual void diag_incloud_time_mom(const int&)                 { unsupported("diag_incloud_time_mom"); }
      virtual void diag_max_rw()                                     { unsupported("diag_max_rw"); }
      virtual void diag_vel_div()                                    { unsupported("diag_vel_div"); }
      virtual std::map<common::output_t, real_t> diag_puddle()       { unsupported("diag_puddle"); return {}; }
      virtual std::vector<real_t> get_attr(const std::string &)      { unsupported("get_attr"); return {}; }
      virtual real_t *outbuf()                                       { unsupported("outbuf"); return nullptr; }

      opts_init_t<real_t> *opts_init = nullptr;   // the instance's own copy (read by bindings for shapes)

      virtual ~particles_proto_t() {}

      protected:
      [[noreturn]] static void unsupported(const char *what)
      {
        throw std::runtime_error(std::string("libcloudph++: ") + what + "() is not provided by this back-end");
      }
    };

    // Creates a particle system on the chosen back-end; the caller owns the returned object.
    // CUDA and multi_CUDA are served by the B200-native implementation; serial and OpenMP are not
    // part of this library (the reference's CPU back-ends exist only as the test oracle).
    template <typename real_t>
    particles_proto_t<real_t> *factory(const backend_t, opts_init_t<real_t>);
  }
}
