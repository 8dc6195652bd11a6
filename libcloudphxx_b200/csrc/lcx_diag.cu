// Per-cell reductions over super-droplets (moments, counts, maxima) and the selectors that precede them.
// Reference: src/impl/diagnose_SD_attributes/particles_impl_moms.ipp:50-387 (thrust::reduce_by_key over
// permutation iterators), src/particles_diag.ipp:193-211,561-634.
//
// SDs are physically grouped by cell, so a cell's SDs are one contiguous segment [cell_off[c], cell_off[c+1]):
//   small cells (max population <= 2048): one warp per cell, lanes stride the segment, xor-shuffle tree;
//   big cells (0-D boxes, coarse grids): a CTA per 4096-SD chunk of a cell, then one thread per cell adds the
//   chunk partials in order.  Both are deterministic (no floating-point atomics).
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;
    constexpr int WARPS = TPB / 32;
    constexpr unsigned BIG_THRESHOLD = 2048;
    constexpr unsigned CHUNK = 4096;

    // ---- per-SD terms ---------------------------------------------------------------------------------
    struct term_moment   // weight * attr^power  (moment_counter, moms.ipp:240-274)
    {
      const real_t *w; const n_t *n; const real_t *x; real_t xp;
      __device__ __forceinline__ real_t operator()(size_t i) const
      {
        const real_t wi = w ? w[i] : real_t(n[i]);
        const real_t xi = x[i];
        return xi >= 0 ? wi * pow(xi, xp) : wi * pow(xi, real_t(int(xp)));
      }
    };
    struct term_positive   // 1 where the selector kept the SD (diag_sd_conc, particles_diag.ipp:26-35)
    {
      const real_t *w;
      __device__ __forceinline__ real_t operator()(size_t i) const { return w[i] > 0. ? real_t(1) : real_t(0); }
    };
    struct term_precip   // n_filtered * rw^3 * vt  (particles_diag.ipp:61-73,561-586)
    {
      const real_t *w, *rw2, *vt;
      __device__ __forceinline__ real_t operator()(size_t i) const { return w[i] * (pow(rw2[i], real_t(3. / 2)) * vt[i]); }
    };
    struct term_mass_dens   // Gaussian-kernel estimator of the mass density function (particles_impl_mass_dens.ipp:13-35)
    {
      const real_t *w, *x; const idx_t *ijk; const uint32_t *off; real_t rad, sig0, xp;
      __device__ __forceinline__ real_t operator()(size_t i) const
      {
        const idx_t c = ijk[i];
        const real_t sig = sig0 / pow(real_t(off[c + 1] - off[c]), real_t(0.2));
        const real_t xi = x[i];
        return w[i] / sig * pow(xi, 3 * xp) * exp(-pow((log(pow(xi, xp)) - log(rad)) / sig, 2) / 2.);
      }
    };
    struct term_plain    // a per-SD array as it is
    {
      const real_t *v;
      __device__ __forceinline__ real_t operator()(size_t i) const { return v[i]; }
    };
    struct term_sid      // storage index, reduced with max
    {
      const idx_t *sid;
      __device__ __forceinline__ real_t operator()(size_t i) const { return real_t(sid[i]); }
    };
    struct term_radius   // rw, reduced with max (particles_diag.ipp:606-634)
    {
      const real_t *rw2;
      __device__ __forceinline__ real_t operator()(size_t i) const { return sqrt(rw2[i]); }
    };

    template <bool IS_MAX> __device__ __forceinline__ real_t combine(real_t a, real_t b) { return IS_MAX ? (a < b ? b : a) : a + b; }
    template <bool IS_MAX> __device__ __forceinline__ real_t warp_reduce(real_t v)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = combine<IS_MAX>(v, __shfl_xor_sync(0xffffffffu, v, o));
      return v;
    }

    enum { NORM_NONE = 0, NORM_SPECIFIC = 1, NORM_MASS_DENS = 2 };
    __device__ __forceinline__ real_t normalise(real_t s, int mode, idx_t c, const real_t *dv, const real_t *rhod)
    {
      if (mode == NORM_SPECIFIC) { s = s / dv[c]; s = s / rhod[c]; }   // two successive divisions, as moms.ipp:322-350
      else if (mode == NORM_MASS_DENS)                                  // particles_impl_mass_dens.ipp:75-96
        s = (4. / 3. * cst<real_t>::rho_w() * sqrt(cst<real_t>::pi() / 2.)) * s / dv[c];
      return s;
    }

    template <class Term, bool IS_MAX>
    __global__ void __launch_bounds__(TPB) k_cell_reduce_small(idx_t n_cell, const uint32_t *__restrict__ off, Term term, int specific,
                                                              const real_t *__restrict__ dv, const real_t *__restrict__ rhod, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * WARPS + (threadIdx.x >> 5);
      if (c >= n_cell) return;
      const int lane = threadIdx.x & 31;
      const uint32_t b = off[c], en = off[c + 1];
      real_t acc = 0;
      for (uint32_t i = b + lane; i < en; i += 32) acc = combine<IS_MAX>(acc, term(i));
      acc = warp_reduce<IS_MAX>(acc);
      if (lane == 0) out[c] = normalise(acc, specific, c, dv, rhod);
    }

    template <class Term, bool IS_MAX>
    __global__ void __launch_bounds__(TPB) k_cell_reduce_chunks(const uint32_t *__restrict__ off, Term term, unsigned n_chunks, real_t *__restrict__ partial)
    {
      __shared__ real_t ws[WARPS];
      const idx_t c = blockIdx.y;
      const uint32_t b = off[c] + blockIdx.x * CHUNK;
      const uint32_t en = min(off[c + 1], b + CHUNK);
      real_t acc = 0;
      for (uint32_t i = b + threadIdx.x; i < en; i += TPB) acc = combine<IS_MAX>(acc, term(i));
      acc = warp_reduce<IS_MAX>(acc);
      if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
      __syncthreads();
      if (threadIdx.x == 0)
      {
        real_t s = ws[0];
        for (int w = 1; w < WARPS; ++w) s = combine<IS_MAX>(s, ws[w]);
        partial[size_t(c) * n_chunks + blockIdx.x] = s;
      }
    }

    template <bool IS_MAX>
    __global__ void __launch_bounds__(TPB) k_cell_reduce_final(idx_t n_cell, const uint32_t *__restrict__ off, unsigned n_chunks, const real_t *__restrict__ partial,
                                                              int specific, const real_t *__restrict__ dv, const real_t *__restrict__ rhod, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * TPB + threadIdx.x;
      if (c >= n_cell) return;
      const unsigned used = (off[c + 1] - off[c] + CHUNK - 1) / CHUNK;
      real_t s = 0;
      for (unsigned q = 0; q < used; ++q) s = combine<IS_MAX>(s, partial[size_t(c) * n_chunks + q]);
      out[c] = normalise(s, specific, c, dv, rhod);
    }

    template <class Term, bool IS_MAX>
    void cell_reduce(lcx_engine *e, Term term, int norm, real_t *out)
    {
      const grid_t &g = e->grid;
      if (!e->grouped) throw error("per-cell reduction requested while super-droplets are not grouped by cell");
      const int spec = (norm == NORM_SPECIFIC && g.n_dims == 0) ? int(NORM_NONE) : norm;   // a parcel is 1 kg of dry air
      if (e->max_count <= BIG_THRESHOLD)
      {
        LCX_LAUNCH(e, (k_cell_reduce_small<Term, IS_MAX>), div_up(g.n_cell, WARPS), TPB, 0, g.n_cell, e->cell_off.p, term, spec, e->dv.p, e->rhod.p, out);
        return;
      }
      const unsigned n_chunks = div_up(e->max_count, CHUNK);
      const size_t need = size_t(g.n_cell) * n_chunks;
      if (e->mom_partial.n < need) { LCX_CUDA(cudaStreamSynchronize(e->stream)); e->mom_partial.alloc(need); }
      LCX_LAUNCH(e, (k_cell_reduce_chunks<Term, IS_MAX>), dim3(n_chunks, g.n_cell), TPB, 0, e->cell_off.p, term, n_chunks, e->mom_partial.p);
      LCX_LAUNCH(e, (k_cell_reduce_final<IS_MAX>), div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->cell_off.p, n_chunks, e->mom_partial.p, spec, e->dv.p, e->rhod.p, out);
    }

    // ---- selectors ------------------------------------------------------------------------------------
    __global__ void __launch_bounds__(TPB) k_select(size_t n_part, int kind, int cons, real_t lo, real_t hi,
                                                   const n_t *__restrict__ n, const real_t *__restrict__ x,
                                                   const real_t *__restrict__ rw2, const real_t *__restrict__ rd3, const real_t *__restrict__ kpa,
                                                   const idx_t *__restrict__ ijk, const real_t *__restrict__ T, const real_t *__restrict__ RH,
                                                   real_t *__restrict__ nf)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i >= n_part) return;
      const real_t base = cons ? nf[i] : real_t(n[i]);
      real_t r;
      switch (kind)
      {
        case LCX_SEL_ALL:   r = real_t(n[i]); break;
        case LCX_SEL_RANGE: { const real_t v = x[i]; r = (v >= lo && v < hi) ? base : real_t(0); break; }     // moms.ipp:23-35
        case LCX_SEL_GT0:   r = base * (x[i] > 0); break;                                                     // moms.ipp:160-196
        case LCX_SEL_RW_GE_RC:                                                                                // particles_diag.ipp:37-59, 389-409
        {
          const real_t rc2 = pow(rw3_cr(rd3[i], kpa[i], T[ijk[i]]), real_t(2. / 3));
          r = rw2[i] >= rc2 ? real_t(n[i]) : real_t(0);
          break;
        }
        case LCX_SEL_RH_GE_SC:                                                                                // particles_diag.ipp:90-105, 361-387
        {
          const idx_t c = ijk[i];
          const real_t v = RH[c] - S_cr(rd3[i], kpa[i], T[c]);
          r = real_t(n[i]) * (v >= 0);
          break;
        }
        default: r = 0;
      }
      nf[i] = r;
    }
  }

  real_t *attr_ptr(lcx_engine *e, int attr)
  {
    sd_arrays &s = e->S();
    switch (attr)
    {
      case LCX_A_RD3: return s.rd3.p;
      case LCX_A_RW2: return s.rw2.p;
      case LCX_A_KPA: return s.kpa.p;
      case LCX_A_VT:  return s.vt.p;
      case LCX_A_X:   return s.x.p;
      case LCX_A_Y:   return s.y.p;
      case LCX_A_Z:   return s.z.p;
      default: throw error("unknown real-valued attribute id " + std::to_string(attr));
    }
  }

  void cell_moment(lcx_engine *e, const real_t *weight_or_null, const real_t *attr, real_t power, bool specific, real_t *out)
  {
    term_moment t = {weight_or_null, e->S().n.p, attr, power};
    cell_reduce<term_moment, false>(e, t, specific ? NORM_SPECIFIC : NORM_NONE, out);
  }

  void cell_sum(lcx_engine *e, const real_t *per_sd, real_t *out)
  {
    term_plain t = {per_sd};
    cell_reduce<term_plain, false>(e, t, NORM_NONE, out);
  }

  void cell_max_sid(lcx_engine *e, real_t *out)
  {
    term_sid t = {e->S().sid.p};
    cell_reduce<term_sid, true>(e, t, NORM_NONE, out);
  }

  void moms_select(lcx_engine *e, int kind, int attr, real_t lo, real_t hi, bool cons)
  {
    if (cons && !e->selected) throw error("consecutive selector called without a preceding selector");
    sd_arrays &s = e->S();
    const real_t *x = (kind == LCX_SEL_RANGE || kind == LCX_SEL_GT0) ? attr_ptr(e, attr) : nullptr;
    if (e->n_part)
      LCX_LAUNCH(e, k_select, div_up(e->n_part, TPB), TPB, 0, e->n_part, kind, int(cons), lo, hi, s.n.p, x, s.rw2.p, s.rd3.p, s.kpa.p, s.ijk.p,
                 e->T.p, e->RH.p, e->n_filtered.p);
    e->selected = true;
  }

  void diag_sd_conc(lcx_engine *e)
  {
    if (!e->selected) throw error("diag_sd_conc called before a selector");
    term_positive t = {e->n_filtered.p};
    cell_reduce<term_positive, false>(e, t, NORM_NONE, e->count_mom.p);
  }

  void diag_precip_rate(lcx_engine *e)
  {
    if (!e->selected) throw error("diag_precip_rate called before a selector");
    hskpng_vterm(e, false);   // side effect kept: the reference refreshes every vt here (particles_diag.ipp:565)
    term_precip t = {e->n_filtered.p, e->S().rw2.p, e->S().vt.p};
    cell_reduce<term_precip, false>(e, t, NORM_NONE, e->count_mom.p);
  }

  void diag_max_rw(lcx_engine *e)
  {
    term_radius t = {e->S().rw2.p};
    cell_reduce<term_radius, true>(e, t, NORM_NONE, e->count_mom.p);
  }

  void diag_mass_dens(lcx_engine *e, int attr, real_t rad, real_t sig0, real_t xp)
  {
    if (!e->selected) throw error("diag_*_mass_dens called before a selector");
    term_mass_dens t = {e->n_filtered.p, attr_ptr(e, attr), e->S().ijk.p, e->cell_off.p, rad, sig0, xp};
    cell_reduce<term_mass_dens, false>(e, t, NORM_MASS_DENS, e->count_mom.p);
  }

  namespace
  {
    // divergence of the Courant field, divided by dt (particles_diag.ipp:497-558; the reference's argument is named dx)
    __global__ void k_vel_div(grid_t g, real_t dt, const real_t *__restrict__ Cx, const real_t *__restrict__ Cy, const real_t *__restrict__ Cz, real_t *__restrict__ out)
    {
      const idx_t c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c >= g.n_cell) return;
      const idx_t cp = c + g.halo_x;
      real_t div = 0;
      if (g.n_dims == 3)
      {
        const idx_t col = idx_t(g.nz) * g.ny;
        const idx_t yl = cp + (cp / col) * g.nz;
        div = div + (Cy[yl + g.nz] - Cy[yl]) / dt;
      }
      if (g.n_dims >= 2)
      {
        const idx_t zl = g.n_dims == 3 ? cp + g.ny * (cp / (idx_t(g.nz) * g.ny)) + (cp - (cp / (idx_t(g.nz) * g.ny)) * (idx_t(g.nz) * g.ny)) / g.nz : cp + cp / g.nz;
        div = div + (Cz[zl + 1] - Cz[zl]) / dt;
      }
      const idx_t xr = cp + (g.n_dims == 3 ? idx_t(g.nz) * g.ny : idx_t(g.nz));
      div = div + (Cx[xr] - Cx[cp]) / dt;
      out[c] = div;
    }
  }

  void diag_vel_div(lcx_engine *e, real_t dt)
  {
    const grid_t &g = e->grid;
    if (g.n_dims == 0) return;
    wait_courant(e);
    LCX_LAUNCH(e, k_vel_div, div_up(g.n_cell, TPB), TPB, 0, g, dt, e->courant_x.p, e->courant_y.p, e->courant_z.p, e->count_mom.p);
  }
}
