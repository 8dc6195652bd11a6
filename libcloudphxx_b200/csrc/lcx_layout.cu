// End-of-step housekeeping: removal (or recycling) of zero-multiplicity SDs, new cell indices, and the physical
// re-grouping of the SoA arrays by cell, which is what lets every other kernel stream contiguous per-cell segments.
//
// Reference passes folded in here:
//   hskpng_remove_n0   src/impl/housekeeping/particles_impl_hskpng_remove.ipp:20-75   (9 remove_if sweeps)
//   rcyc               src/impl/housekeeping/particles_impl_rcyc.ipp:44-139
//   hskpng_ijk         src/impl/housekeeping/particles_impl_hskpng_ijk.ipp:159-200
//   hskpng_sort/count  src/impl/housekeeping/particles_impl_hskpng_sort.ipp:15-70, particles_impl_hskpng_count.ipp:16-48
//
// One key per SD (cell index, or n_cell for dead SDs).  Usual case (relayout_movers): most SDs stay in their cell, so only
// the movers are listed and radix-sorted by new cell; stayers keep their order at the front of their cell's new segment,
// arrivals follow; segment starts come from per-cell counts.  Otherwise (first grouping, > 40 % movers, huge cells): stable
// LSD radix sort of all (key, physical index) pairs on ceil(log2(n_cell+1)) bits and offsets from the sorted keys.  Either
// way one gather then moves every attribute into the alternate buffer set; dead SDs are simply not copied.  The reference's
// storage index `sid` travels with each SD and is re-densified lazily (rank among survivors) when the replayed random
// stream or get_attr need it dense.
#include "lcx_engine.cuh"

#include <cstdlib>
#include <string>

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;

    __global__ void __launch_bounds__(TPB) k_make_keys(size_t first, size_t n, grid_t g, const n_t *__restrict__ ns, const real_t *__restrict__ rw2,
                                                      const real_t *__restrict__ xs, const real_t *__restrict__ ys, const real_t *__restrict__ zs,
                                                      uint32_t *__restrict__ key, uint32_t *__restrict__ val)
    {
      const size_t t = first + size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t >= n) return;
      uint32_t k = g.n_cell;   // dead
      const bool live = ns[t] != 0;
      if (live)
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
      key[t] = live ? relayout_key(g, k, rw2[t]) : relayout_dead_key(g);
      val[t] = uint32_t(t);
    }

    // off[c] = first sorted position whose key is >= c, for c in [0, n_cell+1]; off[n_cell+1] = n
    __global__ void __launch_bounds__(TPB) k_cell_offsets(size_t n, uint32_t n_cell, int class_bits, const uint32_t *__restrict__ key, uint32_t *__restrict__ off)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t > n) return;
      const uint32_t hi = (t == n) ? n_cell + 1 : key[t] >> class_bits;
      const uint32_t lo = (t == 0) ? 0u : (key[t - 1] >> class_bits) + 1;
      for (uint32_t c = lo; c <= hi && c <= n_cell; ++c) off[c] = uint32_t(t);
      if (t == n) off[n_cell + 1] = uint32_t(n);
    }

    __global__ void __launch_bounds__(TPB) k_max_count(uint32_t n_cell, const uint32_t *__restrict__ off, dev_scalars *sc, uint32_t *__restrict__ big_cells)
    {
      const uint32_t c = blockIdx.x * TPB + threadIdx.x;
      uint32_t m = (c < n_cell) ? off[c + 1] - off[c] : 0u;
      if (m > BIG_CELL) big_cells[atomicAdd(&sc->n_big, 1u)] = c;      // rare: sedimenting drops piling up; coalescence treats these cells apart
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
      if ((threadIdx.x & 31) == 0 && m) atomicMax(&sc->max_count, m);
      if (c == 0) sc->n_part = off[n_cell];
    }

    __global__ void __launch_bounds__(TPB) k_keys_from_ijk(size_t n, grid_t g, const idx_t *__restrict__ ijk, const real_t *__restrict__ rw2,
                                                          uint32_t *__restrict__ key, uint32_t *__restrict__ val)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) { key[t] = relayout_key(g, ijk[t], rw2[t]); val[t] = uint32_t(t); }
    }

    struct gather_set { const void *src[15]; void *dst[15]; int width[15]; int n; };

    // every attribute of the survivors moves to its sorted position; the cell index comes from the sorted key
    __global__ void __launch_bounds__(TPB) k_gather(size_t n, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ key, int class_bits,
                                                   idx_t *__restrict__ ijk, gather_set G)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t >= n) return;
      const uint32_t s = perm[t];
      if (key) ijk[t] = key[t] >> class_bits;
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

    __global__ void __launch_bounds__(TPB) k_scatter_n_by_sid(size_t n, const idx_t *__restrict__ sid, const n_t *__restrict__ src, uint64_t *__restrict__ dst)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) dst[sid[t]] = uint64_t(src[t]);
    }
    __global__ void __launch_bounds__(TPB) k_iota(size_t n, uint32_t *__restrict__ val)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) val[t] = uint32_t(t);
    }

    // ---- recycling (rcyc.ipp:44-139) --------------------------------------------------------------------------------
    __global__ void __launch_bounds__(TPB) k_rcyc_stats(size_t n, const n_t *__restrict__ ns, dev_scalars *sc)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      const n_t v = (t < n) ? ns[t] : n_t(2);
      const unsigned zeros = __ballot_sync(0xffffffffu, v == 0), ones = __ballot_sync(0xffffffffu, v == 1);
      unsigned long long m = v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, m, o); m = w > m ? w : m; }
      if ((threadIdx.x & 31) == 0)
      {
        if (zeros) atomicAdd(&sc->rcyc_zero, (unsigned long long)__popc(zeros));
        if (ones) atomicAdd(&sc->rcyc_one, (unsigned long long)__popc(ones));
        atomicMax(&sc->rcyc_max, m);
      }
    }

    // order[sid] = physical index: the identity permutation of the reference's storage (thrust::sequence, rcyc.ipp:69)
    __global__ void __launch_bounds__(TPB) k_rcyc_order(size_t n, const idx_t *__restrict__ sid, uint32_t *__restrict__ order)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) order[sid[t]] = uint32_t(t);
    }

    __global__ void __launch_bounds__(TPB) k_rcyc_keys(size_t n, const n_t *__restrict__ ns, const uint32_t *__restrict__ order, int shift, uint32_t *__restrict__ key)
    {
      const size_t t = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t < n) key[t] = uint32_t(ns[order[t]] >> shift);
    }

    // q-th dead SD (ascending storage index) takes the attributes of the q-th largest multiplicity; the pair shares n
    __global__ void __launch_bounds__(TPB) k_rcyc_copy(size_t n_flagged, size_t n, const uint32_t *__restrict__ order, n_t *__restrict__ ns,
                                                      real_t *__restrict__ rd3, real_t *__restrict__ rw2, real_t *__restrict__ kpa, real_t *__restrict__ vt,
                                                      real_t *__restrict__ x, real_t *__restrict__ y, real_t *__restrict__ z,
                                                      real_t *__restrict__ p0, real_t *__restrict__ p1, real_t *__restrict__ p2, real_t *__restrict__ p3, real_t *__restrict__ rc2)
    {
      const size_t q = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (q >= n_flagged) return;
      const uint32_t dst = order[q], src = order[n - 1 - q];
      rd3[dst] = rd3[src]; rw2[dst] = rw2[src]; kpa[dst] = kpa[src]; vt[dst] = vt[src];
      if (x) x[dst] = x[src];
      if (y) y[dst] = y[src];
      if (z) z[dst] = z[src];
      if (p0) { p0[dst] = p0[src]; p1[dst] = p1[src]; p2[dst] = p2[src]; if (p3) p3[dst] = p3[src]; }      // sstp_tmp_* travel too (distmem_real_vctrs)
      if (rc2) rc2[dst] = rc2[src];
      const n_t m = ns[src];
      ns[dst] = m - m / 2;
      ns[src] = m / 2;
    }

    // ---- movers-only re-layout ------------------------------------------------------------------------------------------
    // Between two re-layouts most SDs stay in their cell (Courant numbers well below one): the stayers of a cell keep their
    // relative order, so only the SDs that changed cell (or arrived, or died) need sorting.  mv[t] = 1 for movers; after an
    // exclusive scan mv[t] = number of movers before t.
    // Everything below walks the OLD cell segments with 8 lanes per cell (four cells per warp): a lane knows its SD's old cell
    // from the segment it is in, movers are ranked inside the cell with a ballot, and only per-cell counts are scanned.
    constexpr int MVG = 8;                // lanes per cell of the counting / listing kernels
    constexpr int MVG_PLACE = 16;         // ... and of k_mv_place_stayers (measured: 16 lanes 0.57 ms, 8 lanes 0.69 ms; the other two lose with 16)
    __device__ __forceinline__ unsigned group_ballot(bool pred)
    { return (__ballot_sync(0xffffffffu, pred) >> ((threadIdx.x & 31) / MVG * MVG)) & ((1u << MVG) - 1u); }

    // movers (SDs whose new key names another cell, incl. the dead) per old cell; entry n_cell = arrivals appended since
    __global__ void __launch_bounds__(TPB) k_mv_count(uint32_t n_cell, size_t n_old, size_t n_grouped, int class_bits, const uint32_t *__restrict__ off,
                                                     const uint32_t *__restrict__ key, uint32_t *__restrict__ mvcnt)
    {
      const uint32_t c = (blockIdx.x * TPB + threadIdx.x) / MVG;
      const int l = threadIdx.x % MVG;
      const bool live = c < n_cell;
      const uint32_t b = live ? off[c] : 0u, en = live ? off[c + 1] : 0u;
      uint32_t cnt = 0;
      const uint32_t rounds = __reduce_max_sync(0xffffffffu, (en - b + MVG - 1) / MVG);
      for (uint32_t r = 0; r < rounds; ++r)
      {
        const uint32_t t = b + r * MVG + l;
        cnt += __popc(group_ballot(t < en && (key[t] >> class_bits) != c));
      }
      if (live && l == 0) mvcnt[c] = cnt;
      if (c == n_cell && l == 0) { mvcnt[n_cell] = uint32_t(n_old - n_grouped); mvcnt[n_cell + 1] = 0u; }
    }
    // compact list of the movers (new cell, old position), cell by cell in storage order; mvoff = exclusive scan of mvcnt
    __global__ void __launch_bounds__(TPB) k_mv_list(uint32_t n_cell, int class_bits, const uint32_t *__restrict__ off, const uint32_t *__restrict__ key,
                                                    const uint32_t *__restrict__ mvoff, uint32_t *__restrict__ mkey, uint32_t *__restrict__ mval)
    {
      const uint32_t c = (blockIdx.x * TPB + threadIdx.x) / MVG;
      const int l = threadIdx.x % MVG;
      const bool live = c < n_cell;
      const uint32_t b = live ? off[c] : 0u, en = live ? off[c + 1] : 0u;
      uint32_t base = live ? mvoff[c] : 0u;
      const uint32_t rounds = __reduce_max_sync(0xffffffffu, (en - b + MVG - 1) / MVG);
      for (uint32_t r = 0; r < rounds; ++r)
      {
        const uint32_t t = b + r * MVG + l;
        const uint32_t k = t < en ? key[t] >> class_bits : c;
        const bool mover = t < en && k != c;
        const unsigned m = group_ballot(mover);
        if (mover) { const uint32_t j = base + __popc(m & ((1u << l) - 1u)); mkey[j] = k; mval[j] = t; }
        base += __popc(m);
      }
    }
    __global__ void __launch_bounds__(TPB) k_mv_list_tail(size_t n_old, size_t n_grouped, int class_bits, const uint32_t *__restrict__ key, uint32_t first,
                                                         uint32_t *__restrict__ mkey, uint32_t *__restrict__ mval)
    {
      const size_t t = n_grouped + size_t(blockIdx.x) * TPB + threadIdx.x;
      if (t >= n_old) return;
      const size_t j = size_t(first) + (t - n_grouped);
      mkey[j] = key[t] >> class_bits; mval[j] = uint32_t(t);
    }
    // stayers move to the front of their cell's new segment in their old order
    __global__ void __launch_bounds__(TPB) k_mv_place_stayers(uint32_t n_cell, int class_bits, const uint32_t *__restrict__ off, const uint32_t *__restrict__ key,
                                                             const uint32_t *__restrict__ new_off, uint32_t *__restrict__ perm, idx_t *__restrict__ ijk_new,
                                                             const idx_t *__restrict__ sid_old, idx_t *__restrict__ sid_new)
    {
      const uint32_t c = (blockIdx.x * TPB + threadIdx.x) / MVG_PLACE;
      const int l = threadIdx.x % MVG_PLACE;
      const bool live = c < n_cell;
      const uint32_t b = live ? off[c] : 0u, en = live ? off[c + 1] : 0u;
      uint32_t base = live ? new_off[c] : 0u;
      const uint32_t rounds = __reduce_max_sync(0xffffffffu, (en - b + MVG_PLACE - 1) / MVG_PLACE);
      // (measured and dropped: issuing four rounds' loads before the first ballot - 0.725 ms against 0.694 ms)
      for (uint32_t r = 0; r < rounds; ++r)
      {
        const uint32_t t = b + r * MVG_PLACE + l;
        const bool stays = t < en && (key[t] >> class_bits) == c;
        const unsigned m = (__ballot_sync(0xffffffffu, stays) >> ((threadIdx.x & 31) / MVG_PLACE * MVG_PLACE)) & ((1u << MVG_PLACE) - 1u);
        if (stays)
        {
          const uint32_t d = base + __popc(m & ((1u << l) - 1u));
          perm[d] = t; ijk_new[d] = c;
          if (sid_new) sid_new[d] = sid_old[t];      // gather-on-read: the storage index moves right here, no gather kernel needed
        }
        base += __popc(m);
      }
    }
    // population of every cell after the move: stayers (old segment minus its movers) + arrivals
    __global__ void __launch_bounds__(TPB) k_mv_counts(uint32_t n_cell, const uint32_t *__restrict__ old_off, const uint32_t *__restrict__ mvoff,
                                                      const uint32_t *__restrict__ arr_off, uint32_t *__restrict__ cnt)
    {
      const uint32_t c = blockIdx.x * TPB + threadIdx.x;
      if (c > n_cell + 1) return;
      uint32_t v = 0;
      if (c < n_cell) v = (old_off[c + 1] - old_off[c]) - (mvoff[c + 1] - mvoff[c]) + (arr_off[c + 1] - arr_off[c]);
      cnt[c] = v;      // [n_cell], [n_cell + 1] = 0: after the scan both hold the number of survivors
    }
    __global__ void __launch_bounds__(TPB) k_mv_place_arrivals(uint32_t n_m, uint32_t n_cell, const uint32_t *__restrict__ mkey, const uint32_t *__restrict__ mval,
                                                              const uint32_t *__restrict__ new_off, const uint32_t *__restrict__ arr_off,
                                                              uint32_t *__restrict__ perm, idx_t *__restrict__ ijk_new,
                                                              const idx_t *__restrict__ sid_old, idx_t *__restrict__ sid_new)
    {
      const uint32_t j = blockIdx.x * TPB + threadIdx.x;
      if (j >= n_m) return;
      const uint32_t c = mkey[j];
      if (c >= n_cell) return;                               // dead
      const uint32_t a0 = arr_off[c];
      const uint32_t stay = (new_off[c + 1] - new_off[c]) - (arr_off[c + 1] - a0);
      const uint32_t d = new_off[c] + stay + (j - a0);
      perm[d] = mval[j];
      ijk_new[d] = c;
      if (sid_new) sid_new[d] = sid_old[mval[j]];
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
    LCX_CUDA(cudaMemsetAsync(&e->scalars.p->n_big, 0, sizeof(unsigned int), e->stream));
    LCX_LAUNCH(e, k_cell_offsets, div_up(n_total + 1, TPB), TPB, 0, n_total, g.n_cell, g.class_bits, sorted_keys, e->cell_off.p);
    LCX_LAUNCH(e, k_max_count, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->cell_off.p, e->scalars.p, e->big_cells.p);
  }

  // rcyc.ipp:44-139.  The reference sorts (stably) the whole storage by multiplicity; ties therefore resolve by storage
  // index, which is what the hidden sid carries here.  Afterwards every key is stale (hskpng_ijk follows in post_copy).
  static void recycle(lcx_engine *e)
  {
    const size_t n = e->n_part;
    if (n == 0) return;
    sd_arrays &s = e->S();
    dev_scalars *sc = e->scalars.p;
    LCX_CUDA(cudaMemsetAsync(&sc->rcyc_zero, 0, 3 * sizeof(unsigned long long), e->stream));
    LCX_LAUNCH(e, k_rcyc_stats, div_up(n, TPB), TPB, 0, n, s.n.p, sc);
    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, sc, sizeof(dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    const size_t n_zero = e->h_scalars->rcyc_zero, n_one = e->h_scalars->rcyc_one;
    const uint64_t n_max = e->h_scalars->rcyc_max;
    if (n_zero == 0) return;
    if (e->cfg.pure_const_multi) return;                     // plain removal (rcyc.ipp:58-62)
    // SDs counted from the large end of the sorted multiplicities up to the first n == 1; when nobody has n == 1 the
    // reference's find() runs off the end and returns the whole length (rcyc.ipp:87-90)
    const size_t n_splittable = n_one ? n - n_zero - n_one : n;
    if (n_splittable == 0) return;
    const size_t n_flagged = n_zero < n_splittable ? n_zero : n_splittable;

    densify_sid(e);
    uint32_t *key[2] = {e->key[0].p, e->key[1].p}, *val[2] = {e->val[0].p, e->val[1].p};
    LCX_LAUNCH(e, k_rcyc_order, div_up(n, TPB), TPB, 0, n, s.sid.p, val[0]);
    int cur = 0;
    const int bits = bit_length(n_max);
    for (int shift = 0; shift < bits; shift += 32)
    {
      LCX_LAUNCH(e, k_rcyc_keys, div_up(n, TPB), TPB, 0, n, s.n.p, val[cur], shift, key[cur]);
      cur = radix_sort_pairs(e, n, 0, bits - shift < 32 ? bits - shift : 32, key, val, cur);
    }
    LCX_LAUNCH(e, k_rcyc_copy, div_up(n_flagged, TPB), TPB, 0, n_flagged, n, val[cur], s.n.p, s.rd3.p, s.rw2.p, s.kpa.p, s.vt.p, s.x.p, s.y.p, s.z.p,
               s.pp_rv.p, s.pp_th.p, s.pp_rh.p, s.pp_p.p, s.rc2.p);
    e->keys_ready = 0;
  }

  // Re-layout that sorts only the SDs which changed cell.  Returns false (nothing done) when too many SDs moved for it to pay
  // off or when there is no previous grouping to start from; the caller then takes the full radix sort.
  // Result: perm in val[0], new cell indices in A().ijk, segment starts in cell_off, n_part / max_count in the device scalars.
  // perm_out: where the permutation goes (the sort scratch, or - gather-on-read - straight into pending_perm, in which case the
  // storage indices are moved on the fly as well)
  static bool relayout_movers(lcx_engine *e, size_t n_old, uint32_t *perm_out, bool move_sid)
  {
    static const bool enabled = [] { const char *v = std::getenv("LCX_RELAYOUT"); return !(v && std::string(v) == "sort"); }();
    const grid_t &g = e->grid;
    if (!enabled || e->n_grouped == 0 || e->n_grouped > n_old || e->max_count > 2048) return false;   // huge cells (0-D boxes): 8 lanes per cell would crawl
    // movers per old cell -> exclusive scan -> total (one readback)
    uint32_t *mvoff = e->mv_scan.p;
    const unsigned cell_blocks = div_up((size_t(g.n_cell) + 1) * MVG, TPB);
    LCX_LAUNCH(e, k_mv_count, cell_blocks, TPB, 0, g.n_cell, n_old, e->n_grouped, g.class_bits, e->cell_off.p, e->key[0].p, mvoff);
    exclusive_scan_u32(e, mvoff, size_t(g.n_cell) + 2);
    uint32_t n_m = 0;
    LCX_CUDA(cudaMemcpyAsync(&n_m, mvoff + g.n_cell + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    if (size_t(n_m) * 5 > n_old * 2 || size_t(n_m) * 2 > e->cap) return false;        // > 40 % movers: sort everything

    uint32_t *mk[2] = {e->key[1].p, e->key[1].p + n_m}, *mv[2] = {e->val[1].p, e->val[1].p + n_m};
    int res = 0;
    if (n_m)
    {
      LCX_LAUNCH(e, k_mv_list, cell_blocks, TPB, 0, g.n_cell, g.class_bits, e->cell_off.p, e->key[0].p, mvoff, mk[0], mv[0]);
      if (n_old > e->n_grouped)
      {
        const uint32_t first = n_m - uint32_t(n_old - e->n_grouped);
        LCX_LAUNCH(e, k_mv_list_tail, div_up(n_old - e->n_grouped, TPB), TPB, 0, n_old, e->n_grouped, g.class_bits, e->key[0].p, first, mk[0], mv[0]);
      }
      res = radix_sort_pairs(e, n_m, 0, bit_length(g.n_cell), mk, mv, 0);
    }
    LCX_LAUNCH(e, k_cell_offsets, div_up(size_t(n_m) + 1, TPB), TPB, 0, size_t(n_m), g.n_cell, 0, mk[res], e->arr_off.p);
    LCX_LAUNCH(e, k_mv_counts, div_up(g.n_cell + 2, TPB), TPB, 0, g.n_cell, e->cell_off.p, mvoff, e->arr_off.p, e->cell_off_new.p);
    exclusive_scan_u32(e, e->cell_off_new.p, size_t(g.n_cell) + 2);
    const idx_t *sid_old = move_sid ? e->S().sid.p : nullptr;
    idx_t *sid_new = move_sid ? e->A().sid.p : nullptr;
    LCX_LAUNCH(e, k_mv_place_stayers, div_up((size_t(g.n_cell) + 1) * MVG_PLACE, TPB), TPB, 0, g.n_cell, g.class_bits, e->cell_off.p, e->key[0].p, e->cell_off_new.p,
               perm_out, e->A().ijk.p, sid_old, sid_new);
    if (n_m)
      LCX_LAUNCH(e, k_mv_place_arrivals, div_up(n_m, TPB), TPB, 0, n_m, g.n_cell, mk[res], mv[res], e->cell_off_new.p, e->arr_off.p,
                 perm_out, e->A().ijk.p, sid_old, sid_new);
    LCX_CUDA(cudaMemcpyAsync(e->cell_off.p, e->cell_off_new.p, (size_t(g.n_cell) + 2) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    LCX_CUDA(cudaMemsetAsync(&e->scalars.p->max_count, 0, sizeof(unsigned int), e->stream));
    LCX_CUDA(cudaMemsetAsync(&e->scalars.p->n_big, 0, sizeof(unsigned int), e->stream));
    LCX_LAUNCH(e, k_max_count, div_up(g.n_cell, TPB), TPB, 0, g.n_cell, e->cell_off.p, e->scalars.p, e->big_cells.p);
    return true;
  }

  void post_copy(lcx_engine *e, bool rcyc, bool keep_all)
  {
    if (rcyc && !keep_all) recycle(e);
    const grid_t &g = e->grid;
    const size_t n_old = e->n_part;
    sd_arrays &s = e->S();
    sd_arrays &a = e->A();
    const uint32_t *sorted_keys = nullptr;     // full-sort path only: the gather takes the cell index from them
    bool gather_queued = false;
    size_t keyed_by_transport = 0;             // k_transport writes keys only: the identity permutation is made when the full sort needs it

    if (n_old == 0)
    {
      LCX_CUDA(cudaMemsetAsync(e->cell_off.p, 0, e->cell_off.bytes(), e->stream));
      e->max_count = 0; e->n_big = 0; e->grouped = true; e->n_grouped = 0;
      return;
    }

    if (keep_all)   // initial grouping: every SD stays, cells are the ones assigned at creation (init_ijk.ipp:36-52)
      LCX_LAUNCH(e, k_keys_from_ijk, div_up(n_old, TPB), TPB, 0, n_old, g, s.ijk.p, s.rw2.p, e->key[0].p, e->val[0].p);
    else
    {
      // k_transport already wrote the keys of the SDs it moved; only later arrivals (migration) are keyed here
      const size_t first = e->keys_ready <= n_old ? e->keys_ready : 0;
      if (first < n_old)
        LCX_LAUNCH(e, k_make_keys, div_up(n_old - first, TPB), TPB, 0, first, n_old, g, s.n.p, s.rw2.p, s.x.p, s.y.p, s.z.p, e->key[0].p, e->val[0].p);
      keyed_by_transport = first;
    }
    e->keys_ready = 0;

    // where every survivor goes: perm[new position] = old position, segment starts in cell_off, new cell index in a.ijk
    const uint32_t *perm = nullptr;
    const bool lazy_wanted = e->lazy_gather && !keep_all;
    if (lazy_wanted && e->pending_perm.n < e->cap) e->pending_perm.alloc(e->cap);
    bool sid_moved = false;
    if (!keep_all && relayout_movers(e, n_old, lazy_wanted ? e->pending_perm.p : e->val[0].p, lazy_wanted))
    {
      perm = lazy_wanted ? e->pending_perm.p : e->val[0].p;
      sid_moved = lazy_wanted;
    }
    else
    {
      if (keyed_by_transport) LCX_LAUNCH(e, k_iota, div_up(keyed_by_transport, TPB), TPB, 0, keyed_by_transport, e->val[0].p);
      const int res = radix_sort_pairs(e, n_old, 0, bit_length(g.n_cell) + g.class_bits, 0);
      compute_cell_offsets(e, e->key[res].p, n_old);
      perm = e->val[res].p;
      sorted_keys = e->key[res].p;
    }

    LCX_CUDA(cudaMemcpyAsync(e->h_scalars, e->scalars.p, sizeof(dev_scalars), cudaMemcpyDeviceToHost, e->stream));
    LCX_CUDA(cudaStreamSynchronize(e->stream));
    const size_t n_new = e->h_scalars->n_part;
    e->max_count = e->h_scalars->max_count;
    e->n_big = e->h_scalars->n_big;
    e->n_large = e->h_scalars->n_large;          // counted by this step's condensation; the next step's choice of kernel variant
    LCX_CUDA(cudaMemsetAsync(&e->scalars.p->n_large, 0, sizeof(unsigned int), e->stream));

    const bool lazy = e->lazy_gather && !keep_all && n_new > 0;
    if (n_new)
    {
      gather_set G; G.n = 0;
      if (!lazy)
      {
        add(G, s.n.p, a.n.p, 8); add(G, s.rd3.p, a.rd3.p, int(sizeof(real_t))); add(G, s.rw2.p, a.rw2.p, int(sizeof(real_t))); add(G, s.kpa.p, a.kpa.p, int(sizeof(real_t)));
        add(G, s.vt.p, a.vt.p, int(sizeof(real_t)));
      }
      if (!lazy) { add(G, s.x.p, a.x.p, int(sizeof(real_t))); add(G, s.y.p, a.y.p, int(sizeof(real_t))); add(G, s.z.p, a.z.p, int(sizeof(real_t))); }
      if (!sid_moved) add(G, s.sid.p, a.sid.p, 4);
      add(G, s.pp_rv.p, a.pp_rv.p, int(sizeof(real_t))); add(G, s.pp_th.p, a.pp_th.p, int(sizeof(real_t))); add(G, s.pp_rh.p, a.pp_rh.p, int(sizeof(real_t))); add(G, s.pp_p.p, a.pp_p.p, int(sizeof(real_t)));
      add(G, s.rc2.p, a.rc2.p, int(sizeof(real_t)));
      LCX_CUDA(cudaEventRecord(e->pre_gather, e->stream));      // uploads of the next step's fields may overtake the gather
      if (G.n || sorted_keys)
      {
        LCX_LAUNCH(e, k_gather, div_up(n_new, TPB), TPB, 0, n_new, perm, sorted_keys, g.class_bits, a.ijk.p, G);
        gather_queued = true;
      }
      if (lazy && perm != e->pending_perm.p)      // full-sort path: the permutation lives in sort scratch, keep a copy for whoever consumes the pending attributes
        LCX_CUDA(cudaMemcpyAsync(e->pending_perm.p, perm, n_new * sizeof(uint32_t), cudaMemcpyDeviceToDevice, e->stream));
    }
    e->pending = lazy ? (lcx_engine::PENDING_ATTR | (g.n_dims > 0 ? lcx_engine::PENDING_XYZ : 0u)) : 0u;
    e->cur ^= 1;
    e->n_part = n_new;
    e->n_grouped = n_new;
    e->grouped = true;
    e->selected = false;

    if (n_new < n_old)
    {
      e->sid_dense = false;
      if (e->dense_always) densify_sid(e);
    }
    if (n_new == 0) { e->sid_hi = 0; e->sid_dense = true; }
    e->tail_is_gather = gather_queued;           // cleared by the next entry point that works on cell fields
  }

  // what is still pending of a gather-on-read re-layout: from the old buffer set into the current one (see lcx_engine::pending)
  void finish_pending(lcx_engine *e, unsigned what)
  {
    what &= e->pending;
    if (!what) return;
    e->pending &= ~what;
    if (e->n_part == 0) return;
    sd_arrays &dst = e->S(), &src = e->A();
    gather_set G; G.n = 0;
    if (what & lcx_engine::PENDING_ATTR)
    {
      add(G, src.n.p, dst.n.p, 8); add(G, src.rd3.p, dst.rd3.p, int(sizeof(real_t))); add(G, src.rw2.p, dst.rw2.p, int(sizeof(real_t))); add(G, src.kpa.p, dst.kpa.p, int(sizeof(real_t)));
      add(G, src.vt.p, dst.vt.p, int(sizeof(real_t)));
    }
    if (what & lcx_engine::PENDING_XYZ) { add(G, src.x.p, dst.x.p, int(sizeof(real_t))); add(G, src.y.p, dst.y.p, int(sizeof(real_t))); add(G, src.z.p, dst.z.p, int(sizeof(real_t))); }
    if (G.n) LCX_LAUNCH(e, k_gather, div_up(e->n_part, TPB), TPB, 0, e->n_part, e->pending_perm.p, nullptr, 0, dst.ijk.p, G);
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

  void scatter_n_by_sid(lcx_engine *e, uint64_t *dst)
  {
    const size_t n = e->n_part;
    if (n == 0) return;
    densify_sid(e);
    LCX_LAUNCH(e, k_scatter_n_by_sid, div_up(n, TPB), TPB, 0, n, e->S().sid.p, e->S().n.p, dst);
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
