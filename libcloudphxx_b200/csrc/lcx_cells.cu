// Per-cell thermodynamic housekeeping and per-SD terminal velocities.
// One kernel per reference pass, each a single coalesced sweep (the reference issues 4-7 Thrust
// transforms per pass: src/impl/housekeeping/particles_impl_hskpng_Tpr.ipp:224-304).
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;

    // T, p, RH, eta (and dv = 1/rhod for a parcel): hskpng_Tpr.ipp:219-305
    __global__ void __launch_bounds__(TPB) k_cells_Tpr(idx_t c_begin, idx_t n_cell, int n_dims, int th_dry, int const_p, int RH_formula,
                                                      const real_t *__restrict__ th, const real_t *__restrict__ rv,
                                                      const real_t *__restrict__ rhod, real_t *__restrict__ p,
                                                      real_t *__restrict__ T, real_t *__restrict__ RH, real_t *__restrict__ eta,
                                                      real_t *__restrict__ dv)
    {
      const idx_t c = c_begin + blockIdx.x * TPB + threadIdx.x;      // cells [c_begin, n_cell): the whole grid or one chunk of it
      if (c >= n_cell) return;
      const real_t th_c = th[c], rv_c = rv[c], rhod_c = rhod[c];
      real_t T_c, p_c;
      if (th_dry) T_c = T_of_th_dry(th_c, rhod_c);
      else        T_c = th_c * exner(p[c]);
      if (!const_p) { p_c = p_of_rhod_rv_T(rhod_c, rv_c, T_c); p[c] = p_c; }
      else p_c = p[c];
      T[c] = T_c;
      RH[c] = RH_of(RH_formula, p_c, rv_c, T_c);
      eta[c] = visc(T_c);
      if (n_dims == 0) dv[c] = real_t(1) / rhod_c;
    }

    // mean free paths for the molecular correction: hskpng_mfp.ipp:42-52
    __global__ void __launch_bounds__(TPB) k_cells_mfp(idx_t n_cell, const real_t *__restrict__ T, const real_t *__restrict__ p,
                                                      real_t *__restrict__ lam_D, real_t *__restrict__ lam_K)
    {
      const idx_t c = blockIdx.x * TPB + threadIdx.x;
      if (c >= n_cell) return;
      lam_D[c] = lambda_D(T[c]);
      lam_K[c] = lambda_K(T[c], p[c]);
    }

    // terminal velocity of liquid SDs; `only_invalid` refreshes entries flagged -1: hskpng_vterm.ipp:185-342
    // drops with rw > 40 um seen by a full fall-speed pass - in one CTA of 64, a fixed sample: what the next step's condensation
    // picks its walk order by (lcx_cond.cu, k_cond_classed).  n_large == nullptr: not a full pass.  (Every warp adding to the one
    // counter cost 1.4 ms on a rain-laden slab.)
    constexpr unsigned LARGE_SAMPLE = 64;
    __device__ __forceinline__ void count_large(real_t r2, unsigned int *n_large)
    {
      if (!n_large || (blockIdx.x % LARGE_SAMPLE) != 0) return;
      const unsigned active = __activemask();
      const unsigned m = __ballot_sync(active, r2 > real_t(1.6e-9));
      if (m && (threadIdx.x % 32) == unsigned(__ffs(active) - 1)) atomicAdd(n_large, unsigned(__popc(m)));
    }

    __global__ void __launch_bounds__(TPB) k_vterm(size_t n_part, int formula, int only_invalid,
                                                  const real_t *__restrict__ rw2, const idx_t *__restrict__ ijk,
                                                  const real_t *__restrict__ T, const real_t *__restrict__ p,
                                                  const real_t *__restrict__ rhod, const real_t *__restrict__ eta,
                                                  const real_t *__restrict__ vt0, real_t *__restrict__ vt, unsigned int *__restrict__ n_large)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n_part) return;
      const real_t r2 = rw2[i];
      count_large(r2, n_large);
      if (!(r2 > real_t(0))) return;
      if (only_invalid && !(vt[i] == real_t(-1))) return;
      const idx_t c = ijk[i];
      vt[i] = vt_of(formula, r2, T[c], p[c], rhod[c], eta[c], vt0);
    }

    // Beard (1977) variants: the altitude correction needs four per-cell numbers (two divisions, two square roots);
    // they are evaluated once per cell here instead of once per SD, then k_vterm_beard77 does the per-droplet rest
    __global__ void __launch_bounds__(TPB) k_beard77_cells(idx_t n_cell, const real_t *__restrict__ p, const real_t *__restrict__ rhod,
                                                          const real_t *__restrict__ eta, beard77_cell<real_t> *__restrict__ out)
    {
      const idx_t c = blockIdx.x * TPB + threadIdx.x;
      if (c < n_cell) out[c] = vt_beard77_cell_consts(p[c], rhod[c], eta[c]);
    }

    template <bool FAST>
    __global__ void __launch_bounds__(TPB) k_vterm_beard77(size_t n_part, int only_invalid, const real_t *__restrict__ rw2, const idx_t *__restrict__ ijk,
                                                          const beard77_cell<real_t> *__restrict__ cells, const real_t *__restrict__ vt0, real_t *__restrict__ vt,
                                                          const vt0_bins<real_t> bins, unsigned int *__restrict__ n_large)
    {
      // `bins` comes from the host (the same object the host layer built the table with): constructed here it cost two log()
      // per droplet, a fifth of the kernel's instructions
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n_part) return;
      const real_t r2 = rw2[i];
      count_large(r2, n_large);
      if (!(r2 > real_t(0))) return;
      if (only_invalid && !(vt[i] == real_t(-1))) return;
      const beard77_cell<real_t> k = cells[ijk[i]];
      const real_t r = sqrt(r2);
      vt[i] = vt_beard77_fact(r, k) * (FAST ? vt0[bins.bin_of(r2)] : vt_beard77_v0(r));
    }

    // linear interpolation of the Eulerian state over condensation sub-steps: sstp_percell_step.ipp:7-47
    __global__ void __launch_bounds__(TPB) k_sstp_percell(idx_t n_cell, int step, real_t sstp, real_t *__restrict__ scl, real_t *__restrict__ tmp)
    {
      const idx_t c = blockIdx.x * TPB + threadIdx.x;
      if (c >= n_cell) return;
      if (step == 0)
      {
        const real_t d = scl[c] - tmp[c];
        tmp[c] = d;
        scl[c] = scl[c] - (sstp - 1) * d / sstp;
      }
      else scl[c] = scl[c] + tmp[c] / sstp;
    }

    // rv -= drv, th -= drv * dth/drv with drv from the change of the specific 3rd wet moment: update_th_rv.ipp:74-191
    __global__ void __launch_bounds__(TPB) k_update_th_rv(idx_t n_cell, real_t *__restrict__ drw_mom3, const real_t *__restrict__ T,
                                                         real_t *__restrict__ th, real_t *__restrict__ rv)
    {
      const idx_t c = blockIdx.x * TPB + threadIdx.x;
      if (c >= n_cell) return;
      const real_t drv = drw_mom3[c] * (cst<real_t>::rho_w() * real_t(4. / 3) * cst<real_t>::pi());
      drw_mom3[c] = drv;
      rv[c] = rv[c] - drv;
      const real_t th_c = th[c];
      th[c] = th_c - drv * d_th_d_rv(T[c], th_c);
    }
  }

  void hskpng_Tpr(lcx_engine *e)
  {
    const grid_t &g = e->grid;
    const idx_t c_begin = e->win_end ? e->win_begin : 0, c_end = e->win_end ? e->win_end : g.n_cell;
    LCX_LAUNCH(e, k_cells_Tpr, div_up(c_end - c_begin, TPB), TPB, 0, c_begin, c_end, g.n_dims, e->cfg.th_dry, e->cfg.const_p, e->cfg.RH_formula,
               e->th.p, e->rv.p, e->rhod.p, e->p.p, e->T.p, e->RH.p, e->eta.p, e->dv.p);
  }

  void hskpng_mfp(lcx_engine *e)
  {
    const grid_t &g = e->grid;
    LCX_LAUNCH(e, k_cells_mfp, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->T.p, e->p.p, e->lambda_D.p, e->lambda_K.p);
  }

  void hskpng_vterm(lcx_engine *e, bool only_invalid)
  {
    if (e->n_part == 0) return;
    sd_arrays &s = e->S();
    unsigned int *n_large = only_invalid ? nullptr : &e->scalars.p->n_large;       // a full pass also counts the large drops
    const int formula = e->cfg.terminal_velocity;
    if (formula == VT_BEARD77 || formula == VT_BEARD77FAST)
    {
      const grid_t &g = e->grid;
      beard77_cell<real_t> *cells = reinterpret_cast<beard77_cell<real_t> *>(e->cell_tmp4.p);
      LCX_LAUNCH(e, k_beard77_cells, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->p.p, e->rhod.p, e->eta.p, cells);
      const vt0_bins<real_t> bins;
      if (formula == VT_BEARD77FAST)
        LCX_LAUNCH(e, (k_vterm_beard77<true>), div_up(e->n_part, TPB), TPB, 0, e->n_part, int(only_invalid), s.rw2.p, s.ijk.p, cells, e->vt0.p, s.vt.p, bins, n_large);
      else
        LCX_LAUNCH(e, (k_vterm_beard77<false>), div_up(e->n_part, TPB), TPB, 0, e->n_part, int(only_invalid), s.rw2.p, s.ijk.p, cells, e->vt0.p, s.vt.p, bins, n_large);
      return;
    }
    LCX_LAUNCH(e, k_vterm, div_up(e->n_part, TPB), TPB, 0, e->n_part, e->cfg.terminal_velocity, int(only_invalid),
               s.rw2.p, s.ijk.p, e->T.p, e->p.p, e->rhod.p, e->eta.p, e->vt0.p, s.vt.p, n_large);
  }

  void sstp_percell_step(lcx_engine *e, int step, int sstp, bool var_rho)
  {
    if (sstp == 1) return;
    if (!e->cfg.allow_sstp_cond) throw error("condensation sub-stepping requested but opts_init.sstp_cond was 1");
    const grid_t &g = e->grid;
    real_t *scl[3] = {e->rv.p, e->th.p, e->rhod.p};
    real_t *tmp[3] = {e->sstp_tmp_rv.p, e->sstp_tmp_th.p, e->sstp_tmp_rh.p};
    for (int ix = 0; ix < (var_rho ? 3 : 2); ++ix)
      LCX_LAUNCH(e, k_sstp_percell, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, step, real_t(sstp), scl[ix], tmp[ix]);
  }

  void sstp_save(lcx_engine *e)
  {
    if (!e->cfg.allow_sstp_cond) return;
    if (e->cfg.exact_sstp_cond) { pp_save(e, 0); return; }        // per-particle version: sstp_save.ipp:19-24
    const size_t b = size_t(e->grid.n_cell) * sizeof(real_t);
    LCX_CUDA(cudaMemcpyAsync(e->sstp_tmp_rv.p, e->rv.p, b, cudaMemcpyDeviceToDevice, e->stream));
    LCX_CUDA(cudaMemcpyAsync(e->sstp_tmp_th.p, e->th.p, b, cudaMemcpyDeviceToDevice, e->stream));
    LCX_CUDA(cudaMemcpyAsync(e->sstp_tmp_rh.p, e->rhod.p, b, cudaMemcpyDeviceToDevice, e->stream));
  }

  void update_th_rv(lcx_engine *e)
  {
    const grid_t &g = e->grid;
    LCX_LAUNCH(e, k_update_th_rv, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->drw_mom3.p, e->T.p, e->th.p, e->rv.p);
  }
}
