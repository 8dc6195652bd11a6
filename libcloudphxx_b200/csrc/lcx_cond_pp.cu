// Per-particle condensation sub-stepping (opts_init.exact_sstp_cond with sstp_cond > 1, no adaptation).
// Reference: src/particles_step.ipp:199-236 and src/impl/condensation/perparticle/
//   calculate_noncond_perparticle_sstp_delta.ipp:13-37, apply_noncond_perparticle_sstp_delta.ipp:13-33,
//   set_perparticle_drwX_to_minus_rwX.ipp:13-38, cond_perparticle_advance_rw2.ipp:33-125, perparticle_advance_rw2.ipp:13-45,
//   add_perparticle_rwX_to_drwX.ipp:13-44, apply_perparticle_drw3_to_perparticle_rv_and_th.ipp:13-62,
//   apply_perparticle_cond_change_to_percell_rv_and_th.ipp:13-26; update_pstate / update_state (update_th_rv.ipp:243-300).
//
// Every SD carries the rv, th, rhod (and p with const_p) it saw at the end of the previous condensation (sstp_tmp_*); the
// change since then - advection of the SD and the Eulerian tendencies - is applied in sstp_cond equal parts, each followed
// by one implicit-Euler growth step in the SD's own thermodynamic state.  With mixing the vapour and heat exchanged by all
// SDs of a cell in a sub-step are added to every SD of the cell; without it each SD keeps its own budget and the cell is
// updated from the change of its liquid water at the end.
//
// One thread per SD and sub-step; the alternate SD buffer set (idle between two re-layouts) is the scratch space.
#ifndef LCX_NO_FAST_MATH
#define LCX_FAST_MATH 1
#endif
#include "lcx_engine.cuh"

#include <cstdlib>

namespace lcx
{
  namespace
  {
    constexpr int TPB = 128;

    struct pp_state
    {
      real_t *rv, *th, *rh, *p;          // the SD's record (sstp_tmp_*)
      real_t *d_rv, *d_th, *d_rh, *d_p;  // change to spread over the sub-steps (sstp_dlt_*)
      real_t *rw3, *drv, *Tp;            // rw^3 left by the previous sub-step, vapour change of this one, temperature
    };

    __global__ void __launch_bounds__(256) k_pp_delta(size_t n, const idx_t *__restrict__ ijk, pp_state S, int const_p,
                                                     const real_t *__restrict__ rv, const real_t *__restrict__ th,
                                                     const real_t *__restrict__ rhod, const real_t *__restrict__ p)
    {
      const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      S.d_rv[i] = rv[c] - S.rv[i];
      S.d_th[i] = th[c] - S.th[i];
      S.d_rh[i] = rhod[c] - S.rh[i];
      if (const_p) S.d_p[i] = p[c] - S.p[i];
    }

    __global__ void __launch_bounds__(256) k_pp_save(size_t first, size_t n, const idx_t *__restrict__ ijk, pp_state S, int const_p,
                                                    const real_t *__restrict__ rv, const real_t *__restrict__ th,
                                                    const real_t *__restrict__ rhod, const real_t *__restrict__ p)
    {
      const size_t i = first + size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      S.rv[i] = rv[c]; S.th[i] = th[c]; S.rh[i] = rhod[c];
      if (const_p) S.p[i] = p[c];
    }

    struct pp_opts { int th_dry, const_p, RH_formula, n_dims, mix, step, sstp; real_t dt_sub, RH_max; };

    template <int MODE>
    __global__ void __launch_bounds__(TPB) k_pp_advance(size_t n, pp_opts O, pp_state S, const idx_t *__restrict__ ijk,
                                                       real_t *__restrict__ rw2, const real_t *__restrict__ rd3, const real_t *__restrict__ kpa,
                                                       const real_t *__restrict__ vt, const n_t *__restrict__ ns,
                                                       const real_t *__restrict__ dv, const real_t *__restrict__ lam_D, const real_t *__restrict__ lam_K)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      const real_t sstp = real_t(O.sstp);
      // the sub-step's share of the non-condensational change
      const real_t rv_i = S.rv[i] + S.d_rv[i] / sstp;
      const real_t th_i = S.th[i] + S.d_th[i] / sstp;
      const real_t rh_i = S.rh[i] + S.d_rh[i] / sstp;
      real_t p_i = 0;
      if (O.const_p) { p_i = S.p[i] + S.d_p[i] / sstp; S.p[i] = p_i; }
      S.rh[i] = rh_i;

      const real_t r2 = rw2[i];
      real_t drw3 = -(O.step > 0 ? S.rw3[i] : real_t(pow(r2, real_t(3) / real_t(2))));

      // the SD's own thermodynamic state
      const real_t T_i = O.th_dry ? T_of_th_dry(th_i, rh_i) : th_i * exner(p_i);
      if (!O.const_p) p_i = p_of_rhod_rv_T(rh_i, rv_i, T_i);
      cond_cell<real_t> cl;
      cl.rhod = rh_i; cl.rv = rv_i; cl.T = T_i; cl.p = p_i;
      cl.RH = RH_of(O.RH_formula, p_i, rv_i, T_i);
      cl.eta = visc(T_i);
      cl.lambda_D = lam_D[c]; cl.lambda_K = lam_K[c];
      const real_t r2n = MODE == COND_EXACT ? advance_rw2(r2, rd3[i], kpa[i], vt[i], cl, O.dt_sub, O.RH_max)
                                        : advance_rw2_fast<MODE == COND_TOMS748>(r2, rd3[i], kpa[i], vt[i], make_cond_consts(cl, O.RH_max), O.dt_sub);
      rw2[i] = r2n;

      const real_t rw3n = pow(r2n, real_t(3) / real_t(2));
      if (O.step < O.sstp - 1) S.rw3[i] = rw3n;
      drw3 = rw3n + drw3;
      // vapour taken from the air by this SD, per unit mass of dry air (rw3diff2drv, cond_common.ipp:24-41)
      const real_t mlt = -cst<real_t>::rho_w() * real_t(4. / 3) * cst<real_t>::pi();
      const real_t nn = real_t(ns[i]);
      const real_t drv = O.n_dims > 0 ? mlt * drw3 * nn / rh_i / dv[c] : mlt * drw3 * nn;
      if (O.mix)
      {
        S.rv[i] = rv_i; S.th[i] = th_i;      // the cell-wide sums are added by k_pp_mix
        S.drv[i] = drv; S.Tp[i] = T_i;
      }
      else
      {
        S.rv[i] = drv + rv_i;
        S.th[i] = drv * d_th_d_rv(T_i, th_i) + th_i;
      }
    }

    // mixing, first half: own heat release from own vapour change and the SD's temperature / theta before the update
    __global__ void __launch_bounds__(256) k_pp_dth(size_t n, pp_state S, real_t *__restrict__ dth)
    {
      const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i < n) dth[i] = S.drv[i] * d_th_d_rv(S.Tp[i], S.th[i]);
    }
    // mixing, second half: every SD of a cell receives the cell's totals (update_pstate)
    __global__ void __launch_bounds__(256) k_pp_mix(size_t n, const idx_t *__restrict__ ijk, pp_state S,
                                                   const real_t *__restrict__ sum_drv, const real_t *__restrict__ sum_dth)
    {
      const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      S.rv[i] = S.rv[i] + sum_drv[c];
      S.th[i] = S.th[i] + sum_dth[c];
    }
    // mixing, end of the step: the cell takes the record of its last SD in storage order (update_state copies from all SDs
    // of a cell to the same place in storage order; they agree to rounding)
    __global__ void __launch_bounds__(256) k_pp_to_cells(size_t n, const idx_t *__restrict__ ijk, const idx_t *__restrict__ sid,
                                                        const real_t *__restrict__ max_sid, pp_state S, real_t *__restrict__ rv, real_t *__restrict__ th)
    {
      const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      if (real_t(sid[i]) == max_sid[c]) { rv[c] = S.rv[i]; th[c] = S.th[i]; }
    }

    // ---- adaptive number of sub-steps per SD (perparticle_nomixing_adaptive_sstp_cond.ipp:49-290) ------------------------
    // One thread runs the whole time step of its SD: it tries 1, 2, 4, ... sub-steps until halving the step no longer
    // changes the first growth increment (relative to rw2) by more than drw2_eps, forces sstp_act sub-steps when the SD
    // would cross its critical radius, then integrates.  Control flow and the order of every update follow the reference's
    // functor statement by statement (including its habits: the temperature left by the last trial is the one used for the
    // heat release of a first step that was "already done" during the trials).
    struct pp_adapt { int sstp_max, sstp_act, th_dry, const_p, RH_formula, n_dims; real_t drw2_eps, drw2_max, dt, RH_max; };

    template <int MODE>
    __global__ void __launch_bounds__(TPB) k_pp_adaptive(size_t n, pp_adapt A, pp_state S, const idx_t *__restrict__ ijk,
                                                        real_t *__restrict__ rw2, const real_t *__restrict__ rd3, const real_t *__restrict__ kpa,
                                                        const real_t *__restrict__ vt, const n_t *__restrict__ ns, const real_t *__restrict__ rc2,
                                                        const real_t *__restrict__ dv, const real_t *__restrict__ lam_D, const real_t *__restrict__ lam_K)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n) return;
      const idx_t c = ijk[i];
      const real_t d_rv = S.d_rv[i], d_th = S.d_th[i], d_rh = S.d_rh[i], d_p = A.const_p ? S.d_p[i] : real_t(0);
      const real_t rd3_i = rd3[i], kpa_i = kpa[i], vt_i = vt[i], nn = real_t(ns[i]), dv_c = dv[c], lD = lam_D[c], lK = lam_K[c];
      real_t t_rv = S.rv[i], t_th = S.th[i], t_rh = S.rh[i], t_p = A.const_p ? S.p[i] : real_t(0);
      real_t r2 = rw2[i], drw2 = 0, Tp = 0, RH = 0;

      auto shift = [&](real_t m) { t_rv += d_rv * m; t_th += d_th * m; t_rh += d_rh * m; if (A.const_p) t_p += d_p * m; };
      auto thermo = [&]() {
        Tp = A.th_dry ? T_of_th_dry(t_th, t_rh) : t_th * exner(t_p);
        if (!A.const_p) t_p = p_of_rhod_rv_T(t_rh, t_rv, Tp);
        RH = RH_of(A.RH_formula, t_p, t_rv, Tp);
      };
      auto grow = [&](real_t dt_sub) -> real_t {
        cond_cell<real_t> cl;
        cl.rhod = t_rh; cl.rv = t_rv; cl.T = Tp; cl.p = t_p; cl.RH = RH; cl.eta = visc(Tp); cl.lambda_D = lD; cl.lambda_K = lK;
        return MODE == COND_EXACT ? advance_rw2(r2, rd3_i, kpa_i, vt_i, cl, dt_sub, A.RH_max)
                                        : advance_rw2_fast<MODE == COND_TOMS748>(r2, rd3_i, kpa_i, vt_i, make_cond_consts(cl, A.RH_max), dt_sub);
      };

      unsigned sstp = unsigned(A.sstp_max);
      real_t frac = 0;
      bool first_done = A.sstp_max == 1;
      {
        real_t drw2_new = 0;
        for (int tr = 1; tr <= A.sstp_max; tr *= 2)
        {
          frac = tr == 1 ? real_t(1) : -real_t(1) / tr;
          shift(frac);
          thermo();
          const real_t d = grow(A.dt / tr) - r2;
          if (tr == 1) drw2 = d; else drw2_new = d;
          if (tr > 1)
          {
            if ((fabs(drw2_new * 2 - drw2) <= A.drw2_eps * r2) && (fabs(drw2) < A.drw2_max * r2))
            {
              sstp = unsigned(tr / 2);
              shift(-frac);
              first_done = true;
              break;
            }
            drw2 = drw2_new;
          }
        }
        if (A.sstp_act > 1)
        {
          const real_t rc = rc2[i];
          if ((r2 < rc && (r2 + sstp * drw2) > rc) || (r2 > rc && (r2 + sstp * drw2) < rc))
          {
            sstp = unsigned(A.sstp_act);
            first_done = false;
          }
        }
        if (!first_done) shift(A.sstp_max == 1 ? -frac : frac);
      }

      frac = real_t(1) / sstp;
      const real_t mlt = -cst<real_t>::rho_w() * real_t(4. / 3) * real_t(3.14159265358979323846264338);
      real_t rw3 = 0;
      for (unsigned step = 0; step < sstp; ++step)
      {
        real_t drw3 = step > 0 ? -rw3 : -real_t(pow(r2, real_t(3) / real_t(2)));
        if (first_done && step == 0) r2 += drw2;
        else
        {
          shift(frac);
          thermo();
          r2 = grow(A.dt / sstp);
        }
        if (step < sstp - 1) { rw3 = pow(r2, real_t(3) / real_t(2)); drw3 += rw3; }
        else drw3 += real_t(pow(r2, real_t(3) / real_t(2)));
        drw3 = A.n_dims > 0 ? mlt * drw3 * nn / t_rh / dv_c : mlt * drw3 * nn;
        t_rv += drw3;
        drw3 = drw3 * d_th_d_rv(Tp, t_th);
        t_th += drw3;
      }
      S.rv[i] = t_rv; S.th[i] = t_th; S.rh[i] = t_rh;
      if (A.const_p) S.p[i] = t_p;
      rw2[i] = r2;
    }

    __global__ void __launch_bounds__(256) k_rc2(size_t n, real_t T, const real_t *__restrict__ rd3, const real_t *__restrict__ kpa, real_t *__restrict__ rc2)
    {
      const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
      if (i < n && rc2[i] == real_t(-1)) rc2[i] = pow(rw3_cr(rd3[i], kpa[i], T), real_t(2. / 3));
    }

    __global__ void k_diff(idx_t n_cell, const real_t *__restrict__ after, const real_t *__restrict__ before, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c < n_cell) out[c] = -before[c] + after[c];
    }

    pp_state state_of(lcx_engine *e)
    {
      sd_arrays &s = e->S(), &a = e->A();
      return pp_state{s.pp_rv.p, s.pp_th.p, s.pp_rh.p, s.pp_p.p, a.pp_rv.p, a.pp_th.p, a.pp_rh.p, a.pp_p.p, a.rd3.p, a.rw2.p, a.kpa.p};
    }
  }

  void pp_save(lcx_engine *e, size_t first)
  {
    if (!e->cfg.exact_sstp_cond || !e->cfg.allow_sstp_cond || e->n_part <= first) return;
    LCX_LAUNCH(e, k_pp_save, div_up(e->n_part - first, 256), 256, 0, first, e->n_part, e->S().ijk.p, state_of(e), e->cfg.const_p,
               e->rv.p, e->th.p, e->rhod.p, e->p.p);
  }

  void hskpng_rc2(lcx_engine *e)
  {
    sd_arrays &s = e->S();
    if (!s.rc2.p || e->n_part == 0) return;
    LCX_LAUNCH(e, k_rc2, div_up(e->n_part, 256), 256, 0, e->n_part, real_t(e->cfg.rc2_T + 273.15), s.rd3.p, s.kpa.p, s.rc2.p);
  }

  void cond_perparticle_adaptive(lcx_engine *e, real_t dt, real_t RH_max, int sstp_max, int sstp_act, real_t drw2_eps, real_t drw2_max)
  {
    if (!e->cfg.exact_sstp_cond || !e->cfg.allow_sstp_cond) throw error("per-particle condensation sub-stepping was not enabled in opts_init");
    if (!e->grouped) throw error("condensation requested while super-droplets are not grouped by cell");
    const size_t n = e->n_part;
    const grid_t &g = e->grid;
    sd_arrays &s = e->S();
    const pp_state S = state_of(e);
    if (sstp_act > 1 && !s.rc2.p) throw error("sstp_cond_act > 1 needs the critical radii (opts_init.sstp_cond_act at construction)");
    const int mode = cond_solver();

    cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->rw_mom3.p);                   // save_liq_ice_content_before_change
    if (n)
    {
      LCX_LAUNCH(e, k_pp_delta, div_up(n, 256), 256, 0, n, s.ijk.p, S, e->cfg.const_p, e->rv.p, e->th.p, e->rhod.p, e->p.p);
      pp_adapt A = {sstp_max, sstp_act, e->cfg.th_dry, e->cfg.const_p, e->cfg.RH_formula, g.n_dims, drw2_eps, drw2_max, dt, RH_max};
      LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_pp_adaptive<M>), div_up(n, TPB), TPB, 0, n, A, S, s.ijk.p, s.rw2.p, s.rd3.p, s.kpa.p, s.vt.p, s.n.p, s.rc2.p,
                                        e->dv.p, e->lambda_D.p, e->lambda_K.p));
    }
    cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->count_mom.p);                 // calc_liq_ice_content_change
    LCX_LAUNCH(e, k_diff, div_up(g.n_cell, 256), 256, 0, g.n_cell, e->count_mom.p, e->rw_mom3.p, e->drw_mom3.p);
    update_th_rv(e);
  }

  void cond_perparticle(lcx_engine *e, real_t dt, real_t RH_max, int sstp, bool mix)
  {
    if (!e->cfg.exact_sstp_cond || !e->cfg.allow_sstp_cond) throw error("per-particle condensation sub-stepping was not enabled in opts_init");
    if (!e->grouped) throw error("condensation requested while super-droplets are not grouped by cell");
    const size_t n = e->n_part;
    const grid_t &g = e->grid;
    sd_arrays &s = e->S();
    const pp_state S = state_of(e);
    const int mode = cond_solver();

    if (!mix) cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->rw_mom3.p);       // save_liq_ice_content_before_change
    if (n)
    {
      LCX_LAUNCH(e, k_pp_delta, div_up(n, 256), 256, 0, n, s.ijk.p, S, e->cfg.const_p, e->rv.p, e->th.p, e->rhod.p, e->p.p);
      pp_opts O = {e->cfg.th_dry, e->cfg.const_p, e->cfg.RH_formula, g.n_dims, int(mix), 0, sstp, dt / sstp, RH_max};
      real_t *dth = e->A().vt.p;
      for (int step = 0; step < sstp; ++step)
      {
        O.step = step;
        LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_pp_advance<M>), div_up(n, TPB), TPB, 0, n, O, S, s.ijk.p, s.rw2.p, s.rd3.p, s.kpa.p, s.vt.p, s.n.p,
                                          e->dv.p, e->lambda_D.p, e->lambda_K.p));
        if (mix)
        {
          LCX_LAUNCH(e, k_pp_dth, div_up(n, 256), 256, 0, n, S, dth);
          cell_sum(e, S.drv, e->count_mom.p);
          cell_sum(e, dth, e->drw_mom3.p);
          LCX_LAUNCH(e, k_pp_mix, div_up(n, 256), 256, 0, n, s.ijk.p, S, e->count_mom.p, e->drw_mom3.p);
        }
      }
    }
    if (mix)
    {
      if (n)
      {
        cell_max_sid(e, e->count_mom.p);
        LCX_LAUNCH(e, k_pp_to_cells, div_up(n, 256), 256, 0, n, s.ijk.p, s.sid.p, e->count_mom.p, S, e->rv.p, e->th.p);
      }
    }
    else
    {
      cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->count_mom.p);               // calc_liq_ice_content_change
      LCX_LAUNCH(e, k_diff, div_up(g.n_cell, 256), 256, 0, g.n_cell, e->count_mom.p, e->rw_mom3.p, e->drw_mom3.p);
      update_th_rv(e);
    }
  }
}
