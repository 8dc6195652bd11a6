// Device-side creation of super-droplets for the `sd_conc` initialisation (opts_init.sd_conc > 0 with dry_distros):
// what the reference does in init_SD_with_distros_sd_conc.ipp:16-52 -> init_dry_sd_conc.ipp:25-66 (dry radii stratified in
// ln(rd), one random offset per SD), init_wet.ipp:18-74 (equilibrium wet radius at min(RH, RH_max)), init_xyz.ipp:16-73
// (positions uniform within the part of the cell inside the domain), init_ijk.ipp:36-52 (per_cell SDs in every cell).
// The multiplicities need the caller's spectrum - an arbitrary host functor (opts_init.dry_distros) - so the dry radii go
// back to the host, which evaluates it (as the reference does: init_n.ipp:48-137) and sends n with lcx_sd_set_n.
//
// Used by the host layer when the random stream is the counter-based one (Philox): the initial state is then a different
// sample of the same distributions than the reference's mt19937 would give, like every later collision draw.  With the
// replayed mt19937 stream the host layer keeps creating the SDs on the host, bit-identical to the reference.
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;

    // uniform in [0, 1): 53 bits for double; 24 bits for float, so that the value never rounds to 1
    template <class T> __device__ __forceinline__ T u01_of(uint32_t w0, uint32_t w1)
    {
      if constexpr (sizeof(T) == 8) return T(philox_u01(w0, w1));
      else return T(w0 >> 8) * T(1.0f / 16777216.0f);
    }

    struct init_args
    {
      size_t first, count;
      unsigned per_cell;
      real_t log_rd_min, log_rd_max, kappa, RH_max;
      uint32_t seed, stream, call_lo, call_hi;
    };

    __global__ void __launch_bounds__(TPB) k_init_sd_conc(init_args A, grid_t g, const real_t *__restrict__ RH, const real_t *__restrict__ T,
                                                         n_t *__restrict__ n, real_t *__restrict__ rd3, real_t *__restrict__ rw2, real_t *__restrict__ kpa,
                                                         real_t *__restrict__ vt, real_t *__restrict__ rc2, real_t *__restrict__ x, real_t *__restrict__ y,
                                                         real_t *__restrict__ z, idx_t *__restrict__ sid, idx_t *__restrict__ ijk)
    {
      const size_t s = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (s >= A.count) return;
      const idx_t c = idx_t(s / A.per_cell);
      const unsigned q = unsigned(s - size_t(c) * A.per_cell);
      // two Philox blocks per SD: (dry radius, x) and (y, z); counter = (SD index, block tag, call)
      uint32_t b0[4] = {uint32_t(s), 0x1000u + uint32_t(s >> 32) * 2u, A.call_lo, A.call_hi};
      uint32_t b1[4] = {uint32_t(s), 0x1001u + uint32_t(s >> 32) * 2u, A.call_lo, A.call_hi};
      philox4x32_10(b0, philox_key{A.seed, A.stream});
      philox4x32_10(b1, philox_key{A.seed, A.stream});
      const real_t u_r = u01_of<real_t>(b0[0], b0[1]);
      const real_t u_p[3] = {u01_of<real_t>(b0[2], b0[3]), u01_of<real_t>(b1[0], b1[1]), u01_of<real_t>(b1[2], b1[3])};

      const real_t lnrd = A.log_rd_min + (real_t(q) + u_r) * (A.log_rd_max - A.log_rd_min) / real_t(A.per_cell);
      const real_t rd3_s = exp(3 * lnrd);
      const real_t RH_c = RH[c] < A.RH_max ? RH[c] : A.RH_max;
      const size_t o = A.first + s;
      n[o] = 0;                                        // set by lcx_sd_set_n once the host has evaluated the spectrum
      rd3[o] = rd3_s;
      rw2[o] = pow(rw3_eq(rd3_s, A.kappa, RH_c, T[c]), real_t(2. / 3));
      kpa[o] = A.kappa;
      vt[o] = real_t(-1);
      if (rc2) rc2[o] = real_t(-1);
      sid[o] = idx_t(o);
      ijk[o] = c;
      const int nz1 = max(1, g.nz), ny1 = max(1, g.ny);
      const int ic = int(c);
      const int ii[3] = {(ic / nz1) / ny1, (ic / nz1) % ny1, ic % nz1};
      const real_t lo[3] = {g.x0, g.y0, g.z0}, hi[3] = {g.x1, g.y1, g.z1}, d[3] = {g.dx, g.dy, g.dz};
      real_t *out[3] = {g.nx ? x : nullptr, g.ny ? y : nullptr, g.nz ? z : nullptr};
      for (int ix = 0; ix < 3; ++ix)
      {
        if (!out[ix]) continue;
        const real_t u = u_p[ix];
        out[ix][o] = u * tmin(hi[ix], real_t(ii[ix] + 1) * d[ix]) + (real_t(1) - u) * tmax(lo[ix], real_t(ii[ix]) * d[ix]);
      }
    }
  }

  void sd_append_sd_conc(lcx_engine *e, size_t per_cell, real_t log_rd_min, real_t log_rd_max, real_t kappa, real_t RH_max,
                         uint64_t seed, uint32_t stream, uint64_t call, real_t *rd3_host)
  {
    const grid_t &g = e->grid;
    const size_t count = size_t(g.n_cell) * per_cell;
    if (count == 0) return;
    densify_sid(e);
    const size_t first = e->n_part;
    if (first + count > e->cap) throw error("n_sd_max (" + std::to_string(e->cap) + ") < n_part (" + std::to_string(first + count) + ")");
    sd_arrays &s = e->S();
    init_args A = {first, count, unsigned(per_cell), log_rd_min, log_rd_max, kappa, RH_max, uint32_t(seed), stream, uint32_t(call), uint32_t(call >> 32)};
    LCX_LAUNCH(e, k_init_sd_conc, div_up(count, TPB), TPB, 0, A, g, e->RH.p, e->T.p, s.n.p, s.rd3.p, s.rw2.p, s.kpa.p, s.vt.p, s.rc2.p,
               s.x.p, s.y.p, s.z.p, s.sid.p, s.ijk.p);
    LCX_CUDA(cudaMemcpyAsync(rd3_host, s.rd3.p + first, count * sizeof(real_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    e->n_part = first + count;
    e->sid_hi = e->n_part;
    e->grouped = false;
    e->keys_ready = 0;
  }
}
