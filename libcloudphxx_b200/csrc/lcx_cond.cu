// Condensation / evaporation, per-cell sub-stepping path.
// Reference: src/impl/condensation/percell/particles_impl_cond.ipp:13-139 (driver),
//            src/impl/condensation/common/particles_impl_cond_common.ipp:79-338 (advance_rw2 + minfun),
//            src/impl/common/save_liq_ice_content_before_change.ipp:13-58 (3rd moment before).
//
// Kernel: one thread per super-droplet; the eight cell scalars are fetched through the SD's cell index -
// SDs are grouped by cell, so a warp touches one or two cells and the loads are L1 broadcasts.  The
// implicit-Euler root solve is FP64-compute bound (about 4.6 growth-rate evaluations per SD); algorithmic
// HBM traffic is 44 B read (rw2, rd3, kpa, vt, ijk) + 8 B written per SD.
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 128;

    __global__ void __launch_bounds__(TPB) k_cond(size_t n_part, real_t dt, real_t RH_max,
                                                 real_t *__restrict__ rw2, const real_t *__restrict__ rd3, const real_t *__restrict__ kpa,
                                                 const real_t *__restrict__ vt, const idx_t *__restrict__ ijk,
                                                 const real_t *__restrict__ rhod, const real_t *__restrict__ rv, const real_t *__restrict__ T,
                                                 const real_t *__restrict__ p, const real_t *__restrict__ RH, const real_t *__restrict__ eta,
                                                 const real_t *__restrict__ lam_D, const real_t *__restrict__ lam_K)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n_part) return;
      const real_t r2 = rw2[i];
      if (r2 <= 0) return;
      const idx_t c = ijk[i];
      cond_cell<real_t> cl;
      cl.rhod = rhod[c]; cl.rv = rv[c]; cl.T = T[c]; cl.p = p[c]; cl.RH = RH[c]; cl.eta = eta[c];
      cl.lambda_D = lam_D[c]; cl.lambda_K = lam_K[c];
      rw2[i] = advance_rw2(r2, rd3[i], kpa[i], vt[i], cl, dt, RH_max);
    }

    // out[c] = a[c] - b[c]  (drw_mom3 = -before + after), optionally keeping `a` for the next sub-step
    __global__ void k_mom_diff(idx_t n_cell, const real_t *__restrict__ after, const real_t *__restrict__ before, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c < n_cell) out[c] = -before[c] + after[c];
    }
  }

  void cond(lcx_engine *e, real_t dt_sub, real_t RH_max, int step, int sstp)
  {
    const grid_t &g = e->grid;
    sd_arrays &s = e->S();
    // 3rd specific wet moment before the change: computed at the first sub-step, afterwards the value left
    // by the previous sub-step is reused as is (it was normalised with that sub-step's rhod - cond.ipp:38-49)
    if (step == 0) cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->rw_mom3.p);
    if (e->n_part)
      LCX_LAUNCH(e, k_cond, div_up(e->n_part, TPB), TPB, 0, e->n_part, dt_sub, RH_max, s.rw2.p, s.rd3.p, s.kpa.p, s.vt.p, s.ijk.p,
                 e->rhod.p, e->rv.p, e->T.p, e->p.p, e->RH.p, e->eta.p, e->lambda_D.p, e->lambda_K.p);
    cell_moment(e, nullptr, s.rw2.p, real_t(3. / 2.), true, e->count_mom.p);
    LCX_LAUNCH(e, k_mom_diff, div_up(g.n_cell, 256), 256, 0, g.n_cell, e->count_mom.p, e->rw_mom3.p, e->drw_mom3.p);
    if (step < sstp - 1)
      LCX_CUDA(cudaMemcpyAsync(e->rw_mom3.p, e->count_mom.p, size_t(g.n_cell) * sizeof(real_t), cudaMemcpyDeviceToDevice, e->stream));
  }
}
