// Flat C binding over the public lgrngn API - see lgrngn_capi.h.
// Compiled against either header tree (this library's or the reference's); nothing below may
// depend on anything that is not in reference include/libcloudph++/lgrngn/*.hpp.
#include "lgrngn_capi.h"

#include <libcloudph++/lgrngn/factory.hpp>

#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace lg = libcloudphxx::lgrngn;
namespace lc = libcloudphxx::common;

typedef lgc_real real;

#if defined(LGC_REFERENCE_BUILD)
// implemented in oracle/ref_internals.cpp (reads the reference's private state)
extern "C" long lgc_ref_internal_get_n(void *proto, int backend, unsigned long long *dst, long cap);
extern "C" long lgc_ref_internal_get_n_f32(void *proto, int backend, unsigned long long *dst, long cap);
#else
// extension of the B200 back-end (host/particles_b200.h): multiplicities as 64-bit integers, either precision
extern "C" int lgrngn_b200_get_n(void *proto, int real_bytes, unsigned long long *dst, long long cap, long long *n_out);
#endif

struct lgc_handle
{
  std::unique_ptr<lg::particles_proto_t<real>> p;
  int backend;
  long n_cell;
};

namespace
{
  thread_local std::string last_error;

  // sum of lognormal modes in ln r
  struct lognormal_sum : lc::unary_function<real>
  {
    std::vector<double> mean_r, stdev, n_tot;
    real funval(const real lnr_) const
    {
      const double lnr = lnr_;
      double res = 0;
      for (std::size_t m = 0; m < mean_r.size(); ++m)
      {
        const double lns = std::log(stdev[m]), d = lnr - std::log(mean_r[m]);
        res += n_tot[m] * std::exp(-(d * d) / 2. / (lns * lns)) / lns / std::sqrt(2 * M_PI);
      }
      return real(res);
    }
  };

  // exponential distribution in droplet volume written in ln r (Shima et al. 2009, Golovin test)
  struct expvolume : lc::unary_function<real>
  {
    double r0, n0;
    real funval(const real lnr) const
    {
      const double r = std::exp(double(lnr)), q = r / r0, q3 = q * q * q;
      return real(n0 * 3. * q3 * std::exp(-q3));
    }
  };

  // the caller's own n(ln r)
  struct callback_distro : lc::unary_function<real>
  {
    double (*fn)(double, void *);
    void *ctx;
    real funval(const real lnr) const { return real(fn(double(lnr), ctx)); }
  };

  lg::arrinfo_t<real> mk(const lgc_arr *a)
  {
    if (a == nullptr || a->data == nullptr) return lg::arrinfo_t<real>();
    return lg::arrinfo_t<real>(a->data, std::vector<ptrdiff_t>{a->strides[0], a->strides[1], a->strides[2]});
  }

  lg::opts_t<real> mk(const lgc_opts *o)
  {
    lg::opts_t<real> r;
    r.adve = o->adve; r.sedi = o->sedi; r.subs = o->subs; r.cond = o->cond; r.coal = o->coal; r.rcyc = o->rcyc;
    r.RH_max = o->RH_max; r.dt = o->dt;
    return r;
  }

  template <class F>
  int guarded(F f)
  {
    try { f(); return 0; }
    catch (const std::exception &e) { last_error = e.what(); return 1; }
    catch (...) { last_error = "unknown exception"; return 2; }
  }
}

extern "C" {

const char *lgc_last_error(void) { return last_error.c_str(); }

const char *lgc_impl_name(void)
{
#if defined(LGC_REFERENCE_BUILD)
  return "reference";
#else
  return "b200";
#endif
}

void lgc_opts_init_defaults(lgc_opts_init *o)
{
  std::memset(o, 0, sizeof(*o));
  const lg::opts_init_t<real> d;
  o->backend = lg::CUDA;
  o->nx = d.nx; o->ny = d.ny; o->nz = d.nz;
  o->dx = d.dx; o->dy = d.dy; o->dz = d.dz; o->dt = d.dt;
  o->sstp_cond = d.sstp_cond; o->sstp_coal = d.sstp_coal;
  o->x0 = d.x0; o->y0 = d.y0; o->z0 = d.z0; o->x1 = d.x1; o->y1 = d.y1; o->z1 = d.z1;
  o->sd_conc = d.sd_conc; o->sd_const_multi = d.sd_const_multi; o->n_sd_max = d.n_sd_max;
  o->kernel = int(d.kernel); o->terminal_velocity = int(d.terminal_velocity);
  o->adve_scheme = int(d.adve_scheme); o->RH_formula = int(d.RH_formula);
  o->coal_switch = d.coal_switch; o->sedi_switch = d.sedi_switch; o->subs_switch = d.subs_switch;
  o->exact_sstp_cond = d.exact_sstp_cond;
  o->RH_max = d.RH_max;
  o->rng_seed = d.rng_seed; o->rng_seed_init = d.rng_seed_init; o->rng_seed_init_switch = d.rng_seed_init_switch;
  o->dev_count = d.dev_count; o->dev_id = d.dev_id;
  o->rd_min = d.rd_min; o->rd_max = d.rd_max;
  o->open_side_walls = d.open_side_walls; o->periodic_topbot_walls = d.periodic_topbot_walls;
  o->variable_dt_switch = d.variable_dt_switch; o->th_dry = d.th_dry; o->const_p = d.const_p;
  o->aerosol_independent_of_rhod = d.aerosol_independent_of_rhod;
  o->sd_conc_large_tail = d.sd_conc_large_tail; o->no_ccn_at_init = d.no_ccn_at_init;
  o->sstp_cond_mix = d.sstp_cond_mix;
  o->adaptive_sstp_cond = d.adaptive_sstp_cond; o->sstp_cond_act = d.sstp_cond_act;
  o->sstp_cond_adapt_drw2_eps = d.sstp_cond_adapt_drw2_eps; o->sstp_cond_adapt_drw2_max = d.sstp_cond_adapt_drw2_max; o->rc2_T = d.rc2_T;
}

void lgc_opts_defaults(lgc_opts *o)
{
  const lg::opts_t<real> d;
  o->adve = d.adve; o->sedi = d.sedi; o->subs = d.subs; o->cond = d.cond; o->coal = d.coal; o->rcyc = d.rcyc;
  o->RH_max = d.RH_max; o->dt = d.dt;
}

int lgc_create(const lgc_opts_init *c, lgc_handle **out)
{
  *out = nullptr;
  return guarded([&] {
    lg::opts_init_t<real> o;
    o.nx = c->nx; o.ny = c->ny; o.nz = c->nz;
    o.dx = c->dx; o.dy = c->dy; o.dz = c->dz; o.dt = c->dt;
    o.sstp_cond = c->sstp_cond; o.sstp_coal = c->sstp_coal;
    o.x0 = c->x0; o.y0 = c->y0; o.z0 = c->z0; o.x1 = c->x1; o.y1 = c->y1; o.z1 = c->z1;
    o.sd_conc = c->sd_conc; o.sd_const_multi = c->sd_const_multi; o.n_sd_max = c->n_sd_max;
    o.kernel = lg::kernel_t(c->kernel); o.terminal_velocity = lg::vt_t(c->terminal_velocity);
    o.adve_scheme = lg::as_t(c->adve_scheme); o.RH_formula = lg::RH_formula_t(c->RH_formula);
    o.kernel_parameters.assign(c->kernel_parameters, c->kernel_parameters + c->n_kernel_parameters);
    o.coal_switch = c->coal_switch; o.sedi_switch = c->sedi_switch; o.subs_switch = c->subs_switch;
    o.exact_sstp_cond = c->exact_sstp_cond;
    o.turb_adve_switch = c->turb_adve_switch; o.turb_cond_switch = c->turb_cond_switch;
    o.turb_coal_switch = c->turb_coal_switch; o.ice_switch = c->ice_switch; o.chem_switch = c->chem_switch;
    o.RH_max = c->RH_max;
    o.rng_seed = c->rng_seed; o.rng_seed_init = c->rng_seed_init; o.rng_seed_init_switch = c->rng_seed_init_switch;
    o.dev_count = c->dev_count; o.dev_id = c->dev_id;
    o.rd_min = c->rd_min; o.rd_max = c->rd_max;
    o.open_side_walls = c->open_side_walls; o.periodic_topbot_walls = c->periodic_topbot_walls;
    o.variable_dt_switch = c->variable_dt_switch; o.th_dry = c->th_dry; o.const_p = c->const_p;
    o.aerosol_independent_of_rhod = c->aerosol_independent_of_rhod;
    if (c->n_w_LS > 0) o.w_LS.assign(c->w_LS, c->w_LS + c->n_w_LS);
    o.sd_conc_large_tail = c->sd_conc_large_tail; o.no_ccn_at_init = c->no_ccn_at_init;
    o.sstp_cond_mix = c->sstp_cond_mix;
    o.adaptive_sstp_cond = c->adaptive_sstp_cond; o.sstp_cond_act = c->sstp_cond_act;
    o.sstp_cond_adapt_drw2_eps = c->sstp_cond_adapt_drw2_eps; o.sstp_cond_adapt_drw2_max = c->sstp_cond_adapt_drw2_max; o.rc2_T = c->rc2_T;
    if (c->n_aerosol_conc_factor > 0) o.aerosol_conc_factor.assign(c->aerosol_conc_factor, c->aerosol_conc_factor + c->n_aerosol_conc_factor);
    for (int i = 0; i < c->n_dry_sizes; ++i)
    {
      const lgc_dry_size &d = c->dry_sizes[i];
      o.dry_sizes[lg::kappa_rd_insol_t<real>(real(d.kappa), real(d.rd_insol))][real(d.radius)] = std::make_pair(real(d.conc), d.count);
    }
    for (int i = 0; i < c->n_distros; ++i)
    {
      const lgc_distro &d = c->distros[i];
      std::shared_ptr<lc::unary_function<real>> f;
      if (d.kind == 0)
      {
        auto s = std::make_shared<lognormal_sum>();
        s->mean_r.assign(d.mean_r, d.mean_r + d.n_modes);
        s->stdev.assign(d.stdev, d.stdev + d.n_modes);
        s->n_tot.assign(d.n_tot, d.n_tot + d.n_modes);
        f = s;
      }
      else if (d.kind == 1)
      {
        auto s = std::make_shared<expvolume>();
        s->r0 = d.r0; s->n0 = d.n0;
        f = s;
      }
      else if (d.kind == 2)
      {
        if (!d.fn) throw std::runtime_error("lgc_create: distro kind 2 without a function");
        auto s = std::make_shared<callback_distro>();
        s->fn = d.fn; s->ctx = d.ctx;
        f = s;
      }
      else throw std::runtime_error("lgc_create: unknown distro kind");
      o.dry_distros.emplace(lg::kappa_rd_insol_t<real>(real(d.kappa), real(d.rd_insol)), f);
    }
    std::unique_ptr<lgc_handle> h(new lgc_handle);
    h->backend = c->backend;
    h->p.reset(lg::factory<real>(lg::backend_t(c->backend), o));
    const lg::opts_init_t<real> &oi = *h->p->opts_init;
    h->n_cell = long(oi.nx ? oi.nx : 1) * (oi.ny ? oi.ny : 1) * (oi.nz ? oi.nz : 1);
    if (c->backend == lg::multi_CUDA) h->n_cell = long(c->nx ? c->nx : 1) * (c->ny ? c->ny : 1) * (c->nz ? c->nz : 1);
    *out = h.release();
  });
}

void lgc_destroy(lgc_handle *h) { delete h; }

int lgc_init(lgc_handle *h, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod, const lgc_arr *p,
             const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz)
{
  return guarded([&] { h->p->init(mk(th), mk(rv), mk(rhod), mk(p), mk(cx), mk(cy), mk(cz)); });
}

int lgc_step_sync(lgc_handle *h, const lgc_opts *o, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod,
                  const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz)
{
  return guarded([&] { h->p->step_sync(mk(o), mk(th), mk(rv), mk(rhod), mk(cx), mk(cy), mk(cz)); });
}

int lgc_sync_in(lgc_handle *h, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod,
                const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz)
{
  return guarded([&] { h->p->sync_in(mk(th), mk(rv), mk(rhod), mk(cx), mk(cy), mk(cz)); });
}

int lgc_step_cond(lgc_handle *h, const lgc_opts *o, const lgc_arr *th, const lgc_arr *rv)
{
  return guarded([&] { h->p->step_cond(mk(o), mk(th), mk(rv)); });
}

int lgc_step_async(lgc_handle *h, const lgc_opts *o)
{
  return guarded([&] { h->p->step_async(mk(o)); });
}

int lgc_diag(lgc_handle *h, int what, double a, double b)
{
  return guarded([&] {
    lg::particles_proto_t<real> &p = *h->p;
    switch (what)
    {
      case LGC_DIAG_ALL:            p.diag_all(); break;
      case LGC_DIAG_RW_GE_RC:       p.diag_rw_ge_rc(); break;
      case LGC_DIAG_RH_GE_SC:       p.diag_RH_ge_Sc(); break;
      case LGC_DIAG_DRY_RNG:        p.diag_dry_rng(a, b); break;
      case LGC_DIAG_WET_RNG:        p.diag_wet_rng(a, b); break;
      case LGC_DIAG_KAPPA_RNG:      p.diag_kappa_rng(a, b); break;
      case LGC_DIAG_DRY_RNG_CONS:   p.diag_dry_rng_cons(a, b); break;
      case LGC_DIAG_WET_RNG_CONS:   p.diag_wet_rng_cons(a, b); break;
      case LGC_DIAG_KAPPA_RNG_CONS: p.diag_kappa_rng_cons(a, b); break;
      case LGC_DIAG_WATER:          p.diag_water(); break;
      case LGC_DIAG_WATER_CONS:     p.diag_water_cons(); break;
      case LGC_DIAG_SD_CONC:        p.diag_sd_conc(); break;
      case LGC_DIAG_PRESSURE:       p.diag_pressure(); break;
      case LGC_DIAG_TEMPERATURE:    p.diag_temperature(); break;
      case LGC_DIAG_RH:             p.diag_RH(); break;
      case LGC_DIAG_DRY_MOM:        p.diag_dry_mom(int(a)); break;
      case LGC_DIAG_WET_MOM:        p.diag_wet_mom(int(a)); break;
      case LGC_DIAG_KAPPA_MOM:      p.diag_kappa_mom(int(a)); break;
      case LGC_DIAG_PRECIP_RATE:    p.diag_precip_rate(); break;
      case LGC_DIAG_MAX_RW:         p.diag_max_rw(); break;
      case LGC_DIAG_VEL_DIV:        p.diag_vel_div(); break;
      case LGC_DIAG_WET_MASS_DENS:  p.diag_wet_mass_dens(a, b); break;
      default: throw std::runtime_error("lgc_diag: unknown selector");
    }
  });
}

long lgc_n_cell(lgc_handle *h) { return h->n_cell; }

int lgc_outbuf(lgc_handle *h, real *dst, long n)
{
  return guarded([&] {
    const real *src = h->p->outbuf();
    std::memcpy(dst, src, sizeof(real) * std::size_t(n < h->n_cell ? n : h->n_cell));
  });
}

int lgc_get_attr(lgc_handle *h, const char *name, real *dst, long cap, long *n_out)
{
  return guarded([&] {
    const std::vector<real> v = h->p->get_attr(name);
    *n_out = long(v.size());
    std::memcpy(dst, v.data(), sizeof(real) * std::size_t(*n_out < cap ? *n_out : cap));
  });
}

int lgc_get_n(lgc_handle *h, unsigned long long *dst, long cap, long *n_out)
{
  return guarded([&] {
#if defined(LGC_REFERENCE_BUILD)
    *n_out = sizeof(real) == 4 ? lgc_ref_internal_get_n_f32(h->p.get(), h->backend, dst, cap) : lgc_ref_internal_get_n(h->p.get(), h->backend, dst, cap);
    if (*n_out < 0) throw std::runtime_error("lgc_get_n: only the serial and OpenMP oracle back-ends expose n");
#else
    // extension of this back-end: multiplicities are readable as 64-bit integers (get_attr("n") would round them to real)
    long long got = 0;
    if (lgrngn_b200_get_n(h->p.get(), int(sizeof(real)), dst, cap, &got) != 0) throw std::runtime_error("lgc_get_n failed");
    *n_out = long(got);
#endif
  });
}

void *lgc_proto(lgc_handle *h) { return h->p.get(); }

int lgc_puddle(lgc_handle *h, double *out14)
{
  return guarded([&] {
    for (int i = 0; i < 14; ++i) out14[i] = 0;
    for (const auto &kv : h->p->diag_puddle()) if (int(kv.first) < 14) out14[int(kv.first)] = kv.second;
  });
}

}  // extern "C"
