// Device-wide exclusive scan and stable LSD radix sort of (u32 key, u32 value) pairs, hand-written
// (no Thrust/CUB).  They replace the reference's thrust::sort_by_key / exclusive_scan call sites on the
// hot path (src/impl/housekeeping/particles_impl_hskpng_sort.ipp:18-53, coalescence/particles_impl_coal.ipp:
// 301-327, housekeeping/particles_impl_hskpng_remove.ipp:49-64).
//
// Radix sort: 8-bit digits, one (histogram, scan, scatter) round per digit.
//   tile  = 256 threads x 8 keys, laid out warp-blocked (each warp owns 256 consecutive keys and walks
//           them in 8 coalesced rounds of 32) so that in-tile order = memory order -> stable.
//   rank  = __match_any_sync groups equal digits inside a warp round; the lowest lane of a group bumps the
//           warp's per-digit counter in shared memory, the others derive their rank from the lane mask.
//   bytes = 4 (histogram read) + 8 (scatter read) + 8 (scatter write) per pair and digit.
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int SORT_THREADS = 256;
    constexpr int SORT_ITEMS = 8;
    constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 2048 pairs per CTA
    constexpr int SORT_WARPS = SORT_THREADS / 32;
    constexpr int RADIX = 256;

    constexpr int SCAN_THREADS = 256;
    constexpr int SCAN_ITEMS = 4;
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 1024 values per CTA

    __device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

    // ---- scan ---------------------------------------------------------------------------------------
    // block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix and the total
    __device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *warp_sums /*8*/)
    {
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      uint32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) warp_sums[w] = inc;
      __syncthreads();
      uint32_t base = 0, tot = 0;
#pragma unroll
      for (int i = 0; i < SCAN_THREADS / 32; ++i)
      {
        const uint32_t s = warp_sums[i];
        if (i < w) base += s;
        tot += s;
      }
      __syncthreads();
      *total = tot;
      return base + inc - v;
    }

    // phase 1: per-tile totals
    __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ data, size_t n, uint32_t *__restrict__ sums)
    {
      __shared__ uint32_t ws[SCAN_THREADS / 32];
      const size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
      uint32_t s = 0;
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += data[base + i];
      uint32_t tot;
      block_exclusive_scan(s, &tot, ws);
      if (threadIdx.x == 0) sums[blockIdx.x] = tot;
    }

    // phase 3 (and the single-tile case): scan each tile, adding the tile's global offset
    __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(uint32_t *__restrict__ data, size_t n, const uint32_t *__restrict__ tile_offsets)
    {
      __shared__ uint32_t ws[SCAN_THREADS / 32];
      const size_t base = size_t(blockIdx.x) * SCAN_TILE + size_t(threadIdx.x) * SCAN_ITEMS;
      uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? data[base + i] : 0u; s += v[i]; }
      uint32_t tot;
      uint32_t run = block_exclusive_scan(s, &tot, ws) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
#pragma unroll
      for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) data[base + i] = run; run += v[i]; }
    }

    void scan_recursive(lcx_engine *e, uint32_t *data, size_t n, uint32_t *scratch, size_t scratch_n)
    {
      const unsigned tiles = div_up(n, SCAN_TILE);
      if (tiles <= 1)
      {
        LCX_LAUNCH(e, k_scan_tiles, 1, SCAN_THREADS, 0, data, n, (const uint32_t *)nullptr);
        return;
      }
      if (scratch_n < tiles) throw error("exclusive_scan_u32: scratch too small");
      LCX_LAUNCH(e, k_scan_tile_sums, tiles, SCAN_THREADS, 0, data, n, scratch);
      scan_recursive(e, scratch, tiles, scratch + tiles, scratch_n - tiles);
      LCX_LAUNCH(e, k_scan_tiles, tiles, SCAN_THREADS, 0, data, n, (const uint32_t *)scratch);
    }

    // ---- radix sort ---------------------------------------------------------------------------------
    __global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint32_t *__restrict__ keys, size_t n, int shift, uint32_t mask,
                                                                uint32_t *__restrict__ ghist, unsigned n_tiles)
    {
      __shared__ uint32_t h[RADIX];
      h[threadIdx.x] = 0;
      __syncthreads();
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      const size_t wbase = size_t(blockIdx.x) * SORT_TILE + size_t(w) * (32 * SORT_ITEMS);
#pragma unroll
      for (int r = 0; r < SORT_ITEMS; ++r)
      {
        const size_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid)
        {
          const uint32_t d = (keys[i] >> shift) & mask;
          const unsigned peers = __match_any_sync(vmask, d);
          if ((__ffs(peers) - 1) == lane) atomicAdd(&h[d], (uint32_t)__popc(peers));
        }
      }
      __syncthreads();
      ghist[size_t(threadIdx.x) * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major for the global scan
    }

    __global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                                                   uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                                                                   size_t n, int shift, uint32_t mask,
                                                                   const uint32_t *__restrict__ gscan, unsigned n_tiles)
    {
      __shared__ uint32_t cnt[SORT_WARPS][RADIX];     // per-warp digit counters, then per-warp output bases
      for (int i = threadIdx.x; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
      __syncthreads();

      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      const unsigned lt = lanemask_lt();
      const size_t wbase = size_t(blockIdx.x) * SORT_TILE + size_t(w) * (32 * SORT_ITEMS);
      uint32_t k[SORT_ITEMS], v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
      for (int r = 0; r < SORT_ITEMS; ++r)
      {
        const size_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        k[r] = 0; v[r] = 0; rank[r] = 0;
        if (valid)
        {
          k[r] = keys_in[i];
          v[r] = vals_in[i];
          const uint32_t d = (k[r] >> shift) & mask;
          const unsigned peers = __match_any_sync(vmask, d);
          const int leader = __ffs(peers) - 1;
          uint32_t old = 0;
          if (lane == leader) { old = cnt[w][d]; cnt[w][d] = old + __popc(peers); }
          old = __shfl_sync(peers, old, leader);
          rank[r] = old + __popc(peers & lt);
        }
        __syncwarp();
      }
      __syncthreads();

      {   // thread d: turn per-warp counts of digit d into output bases (tile offset + earlier warps)
        const int d = threadIdx.x;
        uint32_t run = gscan[size_t(d) * n_tiles + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < SORT_WARPS; ++ww) { const uint32_t c = cnt[ww][d]; cnt[ww][d] = run; run += c; }
      }
      __syncthreads();

#pragma unroll
      for (int r = 0; r < SORT_ITEMS; ++r)
      {
        const size_t i = wbase + r * 32 + lane;
        if (i < n)
        {
          const uint32_t d = (k[r] >> shift) & mask;
          const uint32_t pos = cnt[w][d] + rank[r];
          keys_out[pos] = k[r];
          vals_out[pos] = v[r];
        }
      }
    }
  }

  void exclusive_scan_u32(lcx_engine *e, uint32_t *data, size_t n)
  {
    if (n == 0) return;
    scan_recursive(e, data, n, e->scan_tmp.p, e->scan_tmp.n);
  }

  int radix_sort_pairs(lcx_engine *e, size_t n, int bit_lo, int bit_hi, int in)
  {
    uint32_t *k[2] = {e->key[0].p, e->key[1].p}, *v[2] = {e->val[0].p, e->val[1].p};
    return radix_sort_pairs(e, n, bit_lo, bit_hi, k, v, in);
  }

  int radix_sort_pairs(lcx_engine *e, size_t n, int bit_lo, int bit_hi, uint32_t *const key[2], uint32_t *const val[2], int in)
  {
    if (n == 0) return in;
    const unsigned tiles = div_up(n, SORT_TILE);
    if (size_t(tiles) * RADIX > e->hist.n) throw error("radix_sort_pairs: histogram scratch too small");
    for (int shift = bit_lo; shift < bit_hi; shift += 8)
    {
      const int bits = (bit_hi - shift) < 8 ? (bit_hi - shift) : 8;
      const uint32_t mask = (1u << bits) - 1u;
      const int out = in ^ 1;
      LCX_LAUNCH(e, k_radix_hist, tiles, SORT_THREADS, 0, key[in], n, shift, mask, e->hist.p, tiles);
      exclusive_scan_u32(e, e->hist.p, size_t(tiles) * RADIX);
      LCX_LAUNCH(e, k_radix_scatter, tiles, SORT_THREADS, 0, key[in], val[in], key[out], val[out],
                 n, shift, mask, e->hist.p, tiles);
      in = out;
    }
    return in;
  }
}
