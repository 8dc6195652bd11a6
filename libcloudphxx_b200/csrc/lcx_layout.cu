// End-of-step housekeeping: removal (or recycling) of zero-multiplicity SDs, new cell indices, and the physical
// re-grouping of the SoA arrays by cell, which is what lets every other kernel stream contiguous per-cell segments.
//
// Reference passes folded in here:
//   hskpng_remove_n0   src/impl/housekeeping/particles_impl_hskpng_remove.ipp:20-75   (9 remove_if sweeps)
//   rcyc               src/impl/housekeeping/particles_impl_rcyc.ipp:44-139
//   hskpng_ijk         src/impl/housekeeping/particles_impl_hskpng_ijk.ipp:159-200
//   hskpng_sort/count  src/impl/housekeeping/particles_impl_hskpng_sort.ipp:15-70, particles_impl_hskpng_count.ipp:16-48
//
// One key per SD (cell index, or n_cell for dead SDs) -> stable LSD radix sort of (key, physical index) on
// ceil(log2(n_cell+1)) bits -> cell offsets from the sorted keys -> one gather of every attribute into the alternate
// buffer set.  Dead SDs sort behind the last cell and are simply not copied.  The reference's storage index `sid`
// travels with each SD and is re-densified (rank among survivors) so that it keeps indexing the injected random streams.
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;

    __global__ void __launch_bounds__(TPB) k_make_keys(size_t first, size_t n, grid_t g, const n_t *__restrict__ ns,
                                                      const real_t *__restrict__ xs, const real_t *__restrict__ ys, const real_t *__restrict__ zs,
                                                      uint32_t *__restrict__ key, uint32_t *__restrict__ val)
    {
      const size_t t = first + size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t >= n) return;
      uint32_t k = g.n_cell;   // dead
      if (ns[t] != 0)
      {
        // i = size_t(double(x) / dx): the division is done in double whatever real_t is (hskpng_ijk.ipp:171)
        const idx_t i = g.nx ? idx_t(size_t(double(xs[t]) / double(g.dx))) : 0;
        const idx_t j = g.ny ? idx_t(size_t(double(ys[t]) / double(g.dy))) : 0;
        const idx_t kk = g.nz ? idx_t(size_t(double(zs[t]) / double(g.dz))) : 0;
        switch (g.n_dims)
        {
          case 0: k = 0; break;
          case 1: k = i; break;
          case 2: k = i * g.nz + kk; break;
          default: k = i * (idx_t(g.nz) * g.ny) + j * g.nz + kk; break;
        }
        if (k >= g.n_cell) k = g.n_cell - 1;   // never index outside the grid (the reference leaves this undefined)
      }
      key[t] = k;
      val[t] = uint32_t(t);
    }

    // off[c] = first sorted position whose key is >= c, for c in [0, n_cell+1]; off[n_cell+1] = n
    __global__ void __launch_bounds__(TPB) k_cell_offsets(size_t n, uint32_t n_cell, const uint32_t *__restrict__ key, uint32_t *__restrict__ off)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t > n) return;
      const uint32_t hi = (t == n) ? n_cell + 1 : key[t];
      const uint32_t lo = (t == 0) ? 0u : key[t - 1] + 1;
      for (uint32_t c = lo; c <= hi && c <= n_cell; ++c) off[c] = uint32_t(t);
      if (t == n) off[n_cell + 1] = uint32_t(n);
    }

    __global__ void __launch_bounds__(TPB) k_max_count(uint32_t n_cell, const uint32_t *__restrict__ off, dev_scalars *sc)
    {
      const uint32_t c = blockIdx.x * TPB + threadIdx.x;
      uint32_t m = (c < n_cell) ? off[c + 1] - off[c] : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0 && m) atomicMax(&sc->max_count, m);
      if (c == 0) sc->n_part = off[n_cell];
    }

    __global__ void __launch_bounds__(TPB) k_keys_from_ijk(size_t n, const idx_t *__restrict__ ijk, uint32_t *__restrict__ key, uint32_t *__restrict__ val)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) { key[t] = ijk[t]; val[t] = uint32_t(t); }
    }

    struct gather_set { const void *src[10]; void *dst[10]; int width[10]; int n; };

    __global__ void __launch_bounds__(TPB) k_gather(size_t n, const uint32_t *__restrict__ perm, gather_set G)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t >= n) return;
      const uint32_t s = perm[t];
#pragma unroll 1
      for (int a = 0; a < G.n; ++a)
      {
        if (G.width[a] == 8) static_cast<uint64_t *>(G.dst[a])[t] = static_cast<const uint64_t *>(G.src[a])[s];
        else                 static_cast<uint32_t *>(G.dst[a])[t] = static_cast<const uint32_t *>(G.src[a])[s];
      }
    }

    __global__ void __launch_bounds__(TPB) k_mark_sid(size_t n, const idx_t *__restrict__ sid, uint32_t *__restrict__ mark)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) mark[sid[t]] = 1u;
    }
    __global__ void __launch_bounds__(TPB) k_remap_sid(size_t n, idx_t *__restrict__ sid, const uint32_t *__restrict__ rank)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) sid[t] = rank[sid[t]];
    }

    template <class src_t>
    __global__ void __launch_bounds__(TPB) k_scatter_by_sid(size_t n, const idx_t *__restrict__ sid, const src_t *__restrict__ src, real_t *__restrict__ dst)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) dst[sid[t]] = real_t(src[t]);
    }

    int bit_length(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }

    void add(gather_set &G, const void *src, void *dst, int width)
    {
      if (!src) return;
      G.src[G.n] = src; G.dst[G.n] = dst; G.width[G.n] = width; ++G.n;
    }
  }

  void compute_cell_offsets(lcx_engine *e, const uint32_t *sorted_keys, size_t n_total)
  {
    const grid_t &g = e->grid;
    LCX_CUDA(cudaMemsetAsync(&e->scalars.p->max_count, 0, sizeof(unsigned int), e->stream));
    LCX_LAUNCH(e, k_cell_offsets, div_up(n_total + 1, TPB), TPB, 0, n_total, g.n_cell, sorted_keys, e->cell_off.p);
    LCX_LAUNCH(e, k_max_count, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->cell_off.p, e->scalars.p);
  }

  void post_copy(lcx_engine *e, bool rcyc, bool keep_all)
  {
    if (rcyc) throw error("opts.rcyc (recycling of super-droplets) is not implemented yet in the B200 back-end");
    const grid_t &g = e->grid;
    const size_t n_old = e->n_part;
    sd_arrays &s = e->S();
    sd_arrays &a = e->A();

    if (n_old == 0)
    {
      LCX_CUDA(cudaMemsetAsync(e->cell_off.p, 0, e->cell_off.bytes(), e->stream));
      e->max_count = 0; e->grouped = true;
      return;
    }

    if (keep_all)   // initial grouping: every SD stays, cells are the ones assigned at creation (init_ijk.ipp:36-52)
      LCX_LAUNCH(e, k_keys_from_ijk, div_up(n_old, TPB), TPB, 0, n_old, s.ijk.p, e->key[0].p, e->val[0].p);
    else
    {
      // k_transport already wrote the keys of the SDs it moved; only later arrivals (migration) are keyed here
      const size_t first = e->keys_ready <= n_old ? e->keys_ready : 0;
      if (first < n_old)
        LCX_LAUNCH(e, k_make_keys, div_up(n_old - first, TPB), TPB, 0, first, n_old, g, s.n.p, s.x.p, s.y.p, s.z.p, e->key[0].p, e->val[0].p);
    }
    e->keys_ready = 0;
    const int res = radix_sort_pairs(e, n_old, 0, bit_length(g.n_cell), 0);
    compute_cell_offsets(e, e->key[res].p, n_old);

    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    const size_t n_new = e->h_scalars->n_part;
    e->max_count = e->h_scalars->max_count;

    if (n_new)
    {
      gather_set G; G.n = 0;
      add(G, s.n.p, a.n.p, 8); add(G, s.rd3.p, a.rd3.p, 8); add(G, s.rw2.p, a.rw2.p, 8); add(G, s.kpa.p, a.kpa.p, 8);
      add(G, s.vt.p, a.vt.p, 8); add(G, s.x.p, a.x.p, 8); add(G, s.y.p, a.y.p, 8); add(G, s.z.p, a.z.p, 8);
      add(G, s.sid.p, a.sid.p, 4);
      LCX_LAUNCH(e, k_gather, div_up(n_new, TPB), TPB, 0, n_new, e->val[res].p, G);
      LCX_CUDA(cudaMemcpyAsync(a.ijk.p, e->key[res].p, n_new * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    }
    e->cur ^= 1;
    e->n_part = n_new;
    e->grouped = true;
    e->selected = false;

    if (n_new < n_old)
    {
      e->sid_dense = false;
      if (e->dense_always) densify_sid(e);
    }
    if (n_new == 0) { e->sid_hi = 0; e->sid_dense = true; }
  }

  // new sid = rank of the old sid among the survivors (the order a stable compaction of the reference's storage gives)
  void densify_sid(lcx_engine *e)
  {
    if (e->sid_dense) return;
    const size_t n = e->n_part, space = e->sid_hi;
    if (n)
    {
      if (space > e->flag.n) throw error("internal error: storage-index space exceeds the scratch buffer");
      sd_arrays &c = e->S();
      LCX_CUDA(cudaMemsetAsync(e->flag.p, 0, space * sizeof(uint32_t), e->stream));
      LCX_LAUNCH(e, k_mark_sid, div_up(n, TPB), TPB, 0, n, c.sid.p, e->flag.p);
      exclusive_scan_u32(e, e->flag.p, space);
      LCX_LAUNCH(e, k_remap_sid, div_up(n, TPB), TPB, 0, n, c.sid.p, e->flag.p);
    }
    e->sid_hi = n;
    e->sid_dense = true;
  }

  void scatter_attr_by_sid(lcx_engine *e, int attr, real_t *dst)
  {
    const size_t n = e->n_part;
    if (n == 0) return;
    densify_sid(e);
    sd_arrays &s = e->S();
    if (attr == LCX_A_N)
      LCX_LAUNCH(e, (k_scatter_by_sid<n_t>), div_up(n, TPB), TPB, 0, n, s.sid.p, s.n.p, dst);
    else if (attr == LCX_A_IJK)
      LCX_LAUNCH(e, (k_scatter_by_sid<idx_t>), div_up(n, TPB), TPB, 0, n, s.sid.p, s.ijk.p, dst);
    else
      LCX_LAUNCH(e, (k_scatter_by_sid<real_t>), div_up(n, TPB), TPB, 0, n, s.sid.p, attr_ptr(e, attr), dst);
  }
}
