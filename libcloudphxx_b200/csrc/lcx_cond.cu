// Condensation / evaporation, per-cell sub-stepping path.
// Reference: src/impl/condensation/percell/particles_impl_cond.ipp:13-139 (driver),
//            src/impl/condensation/common/particles_impl_cond_common.ipp:79-338 (advance_rw2 + minfun),
//            src/impl/common/save_liq_ice_content_before_change.ipp:13-58 (3rd moment before),
//            src/impl/common/particles_impl_update_th_rv.ipp:74-191 (vapour / heat feedback).
//
// The reference runs per sub-step: 2 reduce_by_key (3rd wet moment before / after), 1 transform with 9 gathered cell
// fields, and 8 small per-cell transforms.  Here, for grids with small cells, ONE kernel does all of it:
//   k_cond_cells  - a group of 8 lanes owns one cell; its SDs are a contiguous segment, walked 8 at a time;
//                   per SD: n r^3 before, implicit-Euler root solve (TOMS 748), n r^3 after;
//                   per cell: deterministic 8-lane shuffle reduction, then rv -= drv, th -= drv dth/drv.
//                   HBM traffic: 40 B read (n, rw2, rd3, kpa, vt) + 8 B written per SD, cell constants once per cell.
// Cells too populous for that (0-D boxes: one cell with 1e5..1e6 SDs) use a thread-per-SD kernel between two
// chunked per-cell moment reductions (lcx_diag.cu).
// The root solve is FP64-compute bound; three variants are compiled (lcx_set_cond_solver / LCX_COND_SOLVER, lcx_physics.h):
//   toms748 (default) - the reference's TOMS 748, trial point by trial point, growth law in the single-quotient form growth_fast;
//   exact             - the same with the reference's formula transcribed operation by operation (growth_fn), for cross-checks;
//   secant            - opt-in: safeguarded secant that stops once the root is known to the reference's tolerance (half the work).
#ifndef LCX_NO_FAST_MATH
#define LCX_FAST_MATH 1      // see lcx_physics.h: quotients inside the root solve need not be correctly rounded
#endif
#include "lcx_engine.cuh"

#include <cstdlib>
#include <string>

namespace lcx
{
  namespace
  {
    constexpr int TPB = 128;
#ifndef LCX_COND_MINB
#define LCX_COND_MINB 8                      // 8 CTAs of 4 warps per SM (64 registers): measured best of {4,5,6,8}
#endif
    constexpr int GROUP = 8;                 // lanes per cell in k_cond_cells (4 and 16 measured slower)
#ifndef LCX_COND_FUSED_MAX
#define LCX_COND_FUSED_MAX 2048
#endif
    constexpr unsigned FUSED_MAX = LCX_COND_FUSED_MAX;     // largest cell population for the fused kernel

    struct cond_args
    {
      real_t *rw2; const real_t *rd3, *kpa, *vt; const n_t *n; const idx_t *ijk;
      const real_t *rhod, *rv_c, *T, *p, *RH, *eta, *lam_D, *lam_K;
    };

    __device__ __forceinline__ cond_cell<real_t> load_cell(const cond_args &a, idx_t c)
    {
      cond_cell<real_t> cl;
      cl.rhod = a.rhod[c]; cl.rv = a.rv_c[c]; cl.T = a.T[c]; cl.p = a.p[c]; cl.RH = a.RH[c]; cl.eta = a.eta[c];
      cl.lambda_D = a.lam_D[c]; cl.lambda_K = a.lam_K[c];
      return cl;
    }

    // thread per SD (big cells)
    template <int MODE>
    __global__ void __launch_bounds__(TPB) k_cond(size_t n_part, real_t dt, real_t RH_max, cond_args a)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n_part) return;
      const real_t r2 = a.rw2[i];
      if (r2 <= 0) return;
      const cond_cell<real_t> cl = load_cell(a, a.ijk[i]);
      if (MODE == COND_EXACT) a.rw2[i] = advance_rw2(r2, a.rd3[i], a.kpa[i], a.vt[i], cl, dt, RH_max);
      else                    a.rw2[i] = advance_rw2_fast<MODE == COND_TOMS748>(r2, a.rd3[i], a.kpa[i], a.vt[i], make_cond_consts(cl, RH_max), dt);
    }

    // group of 8 lanes per cell: growth + 3rd-moment change + th/rv update
    // Not fused here although it looks tempting: the hskpng_Tpr + hskpng_vterm_all that open step_async.  Measured as an
    // epilogue of the last sub-step it cost 0.86 ms against 0.27 ms for the separate full-occupancy k_vterm (16 M SDs).
    // Lanes of a group stay in lock-step droplet by droplet on purpose: all root solves of a warp are then in the same
    // phase of TOMS 748 and share its (division-heavy) interpolation code; letting early finishers start their next
    // droplet at once (persistent-lane variant, measured) desynchronises the phases and is 35 % slower.
    template <int MODE>
    __global__ void __launch_bounds__(TPB, LCX_COND_MINB) k_cond_cells(idx_t n_cell, const uint32_t *__restrict__ off, real_t dt, real_t RH_max, cond_args a,
                                                       int n_dims, const real_t *__restrict__ dv, int first_step, int keep_after,
                                                       real_t *__restrict__ rw_mom3, real_t *__restrict__ drw_mom3,
                                                       real_t *__restrict__ th, real_t *__restrict__ rv)
    {
      const idx_t c = (blockIdx.x * TPB + threadIdx.x) / GROUP;
      const int l = threadIdx.x % GROUP;
      const bool live = c < n_cell;
      real_t m3_before = 0, m3_after = 0;
      cond_cell<real_t> cl;
      if (live)
      {
        cl = load_cell(a, c);
        const cond_cell_consts<real_t> k = make_cond_consts(cl, RH_max);
        const uint32_t b = off[c], en = off[c + 1];
        for (uint32_t i = b + l; i < en; i += GROUP)
        {
          const real_t r2 = a.rw2[i];
          const real_t nn = real_t(a.n[i]);
          m3_before += nn * (r2 * sqrt(r2));
          real_t r2n = r2;
          if (r2 > 0)
          {
            r2n = MODE == COND_EXACT ? advance_rw2(r2, a.rd3[i], a.kpa[i], a.vt[i], cl, dt, RH_max)
                                        : advance_rw2_fast<MODE == COND_TOMS748>(r2, a.rd3[i], a.kpa[i], a.vt[i], k, dt);
            a.rw2[i] = r2n;
          }
          m3_after += nn * (r2n * sqrt(r2n));
        }
      }
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1)
      {
        m3_before += __shfl_xor_sync(0xffffffffu, m3_before, o, GROUP);
        m3_after += __shfl_xor_sync(0xffffffffu, m3_after, o, GROUP);
      }
      if (live && l == 0)
      {
        // specific moments: divided by the cell volume, then by the dry-air density (moms.ipp:322-350)
        if (n_dims > 0) { m3_before = m3_before / dv[c] / cl.rhod; m3_after = m3_after / dv[c] / cl.rhod; }
        if (!first_step) m3_before = rw_mom3[c];           // value left by the previous sub-step (cond.ipp:38-49)
        if (keep_after) rw_mom3[c] = m3_after;
        const real_t drv = (-m3_before + m3_after) * (cst<real_t>::rho_w() * real_t(4. / 3) * cst<real_t>::pi());
        drw_mom3[c] = drv;
        rv[c] = cl.rv - drv;
        const real_t th_c = th[c];
        th[c] = th_c - drv * d_th_d_rv(cl.T, th_c);
      }
    }

    // out[c] = -before[c] + after[c]
    __global__ void k_mom_diff(idx_t n_cell, const real_t *__restrict__ after, const real_t *__restrict__ before, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c < n_cell) out[c] = -before[c] + after[c];
    }

    int g_solver = -1;
  }

  int cond_solver()
  {
    if (g_solver < 0)
    {
      const char *v = std::getenv("LCX_COND_SOLVER"), *x = std::getenv("LCX_COND_EXACT");
      const std::string m = v ? v : "";
      g_solver = (m == "exact" || (x && x[0] == '1')) ? COND_EXACT : m == "secant" ? COND_SECANT : COND_TOMS748;
    }
    return g_solver;
  }
  void set_cond_solver(int mode) { g_solver = (mode == COND_EXACT || mode == COND_SECANT) ? mode : COND_TOMS748; }

  // one condensation sub-step INCLUDING the th/rv feedback (the host layer no longer calls update_th_rv separately)
  void cond(lcx_engine *e, real_t dt_sub, real_t RH_max, int step, int sstp)
  {
    const grid_t &g = e->grid;
    sd_arrays &s = e->S();
    if (!e->grouped) throw error("condensation requested while super-droplets are not grouped by cell");
    cond_args a = {s.rw2.p, s.rd3.p, s.kpa.p, s.vt.p, s.n.p, s.ijk.p, e->rhod.p, e->rv.p, e->T.p, e->p.p, e->RH.p, e->eta.p, e->lambda_D.p, e->lambda_K.p};
    const int mode = cond_solver();
    const int keep_after = step < sstp - 1;

    if (e->max_count <= FUSED_MAX)
    {
      const unsigned blocks = div_up(size_t(g.n_cell) * GROUP, TPB);
      LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond_cells<M>), blocks, TPB, 0, g.n_cell, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                                        int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p));
      return;
    }

    // big cells: moment before, growth, moment after, feedback
    if (step == 0) cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->rw_mom3.p);
    if (e->n_part)
    {
      LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond<M>), div_up(e->n_part, TPB), TPB, 0, e->n_part, dt_sub, RH_max, a));
    }
    cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->count_mom.p);
    LCX_LAUNCH(e, k_mom_diff, div_up(g.n_cell, 256), 256, 0, g.n_cell, e->count_mom.p, e->rw_mom3.p, e->drw_mom3.p);
    if (keep_after)
      LCX_CUDA(cudaMemcpyAsync(e->rw_mom3.p, e->count_mom.p, size_t(g.n_cell) * sizeof(real_t), cudaMemcpyDeviceToDevice, e->stream));
    update_th_rv(e);
  }
}
