// SDM Monte-Carlo coalescence (Shima et al. 2009) for one sub-step.
// Reference: src/impl/coalescence/particles_impl_coal.ipp:273-546 (driver), :146-269 (collider), :118-143 (collide),
//            :59-96 (kappa mixing), :99-107 (scale factor); pairing order from
//            src/impl/housekeeping/particles_impl_hskpng_sort.ipp:31-53 (shuffle = stable sort by a fresh random key,
//            then stable sort by cell).
//
// The reference's candidate pairs are (2k, 2k+1) of each cell's SDs ordered by (random key un, storage index).
// SDs are physically grouped by cell here, so that order is a per-cell sort of the composite key
// (un[sid] << 32 | sid), which is unique - no dependence on the physical order inside the cell:
//   * small cells (max population <= 256): 16 lanes per cell (two cells per warp); random keys staged in shared memory,
//     rank of each SD by counting smaller keys (four per load), ties re-ranked with the storage index, pairs processed
//     by the lanes;
//   * big cells (0-D boxes, coarse grids): global LSD radix sort by sid, un, cell (lcx_sort.cu), then one thread
//     per candidate pair.
// Random numbers come from an injected stream (parity with the reference's mt19937 draw order: un by storage
// index, u01 by sorted position of the first SD of the pair) or from Philox4x32-10 evaluated at the same indices.
#include "lcx_engine.cuh"

namespace lcx
{
  namespace
  {
    constexpr int TPB = 256;
    constexpr int CELL_LANES = 16;           // lanes that share one cell in k_coal_small (two cells per warp)
#ifndef LCX_COAL_TPB
#define LCX_COAL_TPB 256                     // threads per CTA of k_coal_small
#endif
    constexpr int CTPB = LCX_COAL_TPB;
    constexpr int GROUPS = CTPB / CELL_LANES;
    constexpr unsigned SMALL_MAX = BIG_CELL; // largest cell population handled by k_coal_small in its usual configuration
    constexpr unsigned MEDIUM_MAX = 1024;    // ... and in the big-shared-memory configuration (rain piling up in a few cells)
    constexpr int KAPPA_ITER_MAX = 64;       // collisions of one pair up to which kappa is mixed event by event like the reference

    struct rng_src
    {
      int mode;
      const uint32_t *un;     // device copy of the injected integer stream (by storage index)
      const real_t *u01;      // device copy of the injected [0,1) stream (by sorted position)
      uint64_t seed, call;
      uint32_t cell_base, stream;   // global index of the slab's first cell; second key word (slab rank): slabs draw different streams

      __device__ __forceinline__ uint32_t get_un(uint32_t sid) const
      {
        if (mode == LCX_RNG_INJECT) return un[sid];
        uint32_t c[4] = {sid, 0u, uint32_t(call), uint32_t(call >> 32)};
        philox4x32_10(c, philox_key{uint32_t(seed), stream});
        return c[0];
      }
      __device__ __forceinline__ real_t get_u01(uint32_t pos) const
      {
        if (mode == LCX_RNG_INJECT) return u01[pos];
        uint32_t c[4] = {pos, 1u, uint32_t(call), uint32_t(call >> 32)};
        philox4x32_10(c, philox_key{uint32_t(seed), stream});
        const uint64_t bits = (uint64_t(c[0]) << 21) ^ (uint64_t(c[1]) >> 11);      // 53-bit mantissa from two words, [0,1)
        return real_t(bits & ((1ull << 53) - 1)) * real_t(1.0 / 9007199254740992.0);
      }
    };

    struct coal_ctx
    {
      n_t *n; real_t *rw2, *rd3, *vt, *kpa;
      real_t *rc2;            // nullptr unless activation sub-stepping keeps critical radii (coal.ipp:527-541)
      const real_t *dv;
      coal_kernel_params<real_t> kp;
      real_t dt;
      int multi_kappa, pure_const_multi;
      dev_scalars *sc;
    };

    // multiplicity / radius update of one collision event; `hi` keeps the larger multiplicity: coal.ipp:118-143
    __device__ __forceinline__ void collide(const coal_ctx &cx, uint32_t hi, uint32_t lo, n_t n_hi, n_t n_lo,
                                           real_t rw2_hi, real_t rw2_lo, real_t rd3_hi, real_t rd3_lo, n_t col_no)
    {
      cx.n[hi] = n_hi - col_no * n_lo;
      const real_t rw_lo = cbrt(col_no * rw2_hi * sqrt(rw2_hi) + rw2_lo * sqrt(rw2_lo));
      cx.rw2[lo] = rw_lo * rw_lo;
      const real_t rd3_new = col_no * rd3_hi + rd3_lo;
      cx.rd3[lo] = rd3_new;
      cx.vt[lo] = real_t(-1);
      if (cx.rc2) cx.rc2[lo] = real_t(-1);
      if (cx.multi_kappa)
      {
        // rd3-weighted mean of kappa, applied once per collision: weighted_summator, coal.ipp:59-96.  The reference iterates
        // col_no times (with an int counter); a rain drop sweeping up thousands of droplets would keep one lane - hence its whole
        // CTA - busy for that long, so beyond KAPPA_ITER_MAX collisions the recurrence's closed form is used: the same weighted
        // mean, summed in one step (differs from the iterated value by rounding only, <= col_no * 2^-53 relative)
        const real_t kpa_hi = cx.kpa[hi];
        real_t kpa_lo = cx.kpa[lo];
        real_t rd3_old = rd3_new - col_no * rd3_hi;
        if (col_no <= n_t(KAPPA_ITER_MAX))
          for (int ci = 0; ci < int(col_no); ++ci)
          {
            kpa_lo = (kpa_hi * rd3_hi + kpa_lo * rd3_old) / (rd3_hi + rd3_old);
            rd3_old += rd3_hi;
          }
        else
          kpa_lo = (real_t(col_no) * (kpa_hi * rd3_hi) + kpa_lo * rd3_old) / (real_t(col_no) * rd3_hi + rd3_old);
        cx.kpa[lo] = kpa_lo;
      }
    }

    // one candidate pair: a = physical index of the SD at even in-cell position, b = the next one: coal.ipp:181-268
    // pref = dt / dv * scale factor of the cell (the first two operations of the reference's product, the same for all its pairs)
    __device__ __forceinline__ void try_pair(const coal_ctx &cx, uint32_t a, uint32_t b, real_t u01, real_t pref,
                                            unsigned long long &n_coll, unsigned long long &n_pairs)
    {
      const n_t n_a = cx.n[a], n_b = cx.n[b];
      const real_t rw2_a = cx.rw2[a], rw2_b = cx.rw2[b];
      const real_t vt_a = cx.vt[a], vt_b = cx.vt[b];
      const real_t prob = pref * coal_kernel(cx.kp, n_a, n_b, rw2_a, rw2_b, vt_a, vt_b);
      n_t col_no = n_t(prob);
      if (cx.pure_const_multi && col_no >= 1) cx.sc->increase_sstp_coal = 1u;
      if (u01 < prob - col_no) ++col_no;
      if (col_no == 0) return;
      const real_t rd3_a = cx.rd3[a], rd3_b = cx.rd3[b];
      if (n_a >= n_b)
      {
        if (n_b > 0) col_no = tmin(col_no, n_t(n_a / n_b));
        collide(cx, a, b, n_a, n_b, rw2_a, rw2_b, rd3_a, rd3_b, col_no);
      }
      else
      {
        if (n_a > 0) col_no = tmin(col_no, n_t(n_b / n_a));
        collide(cx, b, a, n_b, n_a, rw2_b, rw2_a, rd3_b, rd3_a, col_no);
      }
      n_coll += col_no;
      n_pairs += 1;
    }

    // collision statistics: one pair of atomics per warp that saw a collision
    __device__ __forceinline__ void flush_stats(const coal_ctx &cx, unsigned long long n_coll, unsigned long long n_pairs)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        n_coll += __shfl_xor_sync(0xffffffffu, n_coll, o);
        n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o);
      }
      if ((threadIdx.x & 31) == 0 && n_pairs)
      {
        atomicAdd(&cx.sc->n_collisions, n_coll);
        atomicAdd(&cx.sc->n_pairs_collided, n_pairs);
      }
    }

    // ---- small cells: 16 lanes per cell -------------------------------------------------------------------
    // Pair order = rank of (un, sid) inside the cell.  Ranks are counted on the 32-bit random key alone, four keys per
    // shared-memory load; two SDs of a cell drawing the same key (about once per 10^7 cells) leave a slot of the
    // permutation unfilled, which is detected and that cell re-ranked with the storage index as tie-break.
    // Philox mode draws per cell: block q of cell c gives four keys (q < 2^31) or the u01 of two pairs (q >= 2^31).
    constexpr int KEY_PAD = 4;               // shifts the two cells of a warp to different banks for the 16-byte loads

    __device__ __forceinline__ void philox_cell_block(const rng_src &rng, uint32_t c, uint32_t q, uint32_t (&w)[4])
    {
      w[0] = rng.cell_base + c; w[1] = q; w[2] = uint32_t(rng.call); w[3] = uint32_t(rng.call >> 32);
      philox4x32_10(w, philox_key{uint32_t(rng.seed), rng.stream});
    }

    __device__ __forceinline__ real_t u01_from_words(uint32_t w0, uint32_t w1)      // 53-bit mantissa from two words, [0,1)
    {
      const uint64_t bits = (uint64_t(w0) << 21) ^ (uint64_t(w1) >> 11);
      return real_t(bits & ((1ull << 53) - 1)) * real_t(1.0 / 9007199254740992.0);
    }

    // ranks of the lane's SDs e0, e0 + 16, ... (N of them) among the m keys of the cell; returns the sum of the ranks it wrote
    template <int N>
    __device__ __forceinline__ uint32_t rank_sweep(const uint32_t *key, unsigned short *perm, uint32_t e0, uint32_t m, uint32_t m4)
    {
      uint32_t mine[N], r[N], sum = 0;
#pragma unroll
      for (int i = 0; i < N; ++i) { const uint32_t e = e0 + i * CELL_LANES; mine[i] = e < m ? key[e] : 0u; r[i] = 0; }
      for (uint32_t j = 0; j < m4; j += 4)
      {
        const uint4 k = *reinterpret_cast<const uint4 *>(key + j);
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] += (k.x < mine[i]) + (k.y < mine[i]) + (k.z < mine[i]) + (k.w < mine[i]);
      }
#pragma unroll
      for (int i = 0; i < N; ++i) { const uint32_t e = e0 + i * CELL_LANES; if (e < m) { perm[r[i]] = (unsigned short)e; sum += r[i]; } }
      return sum;
    }

#ifndef LCX_COAL_MINB
#define LCX_COAL_MINB 4      // 64 registers, 4 CTAs per SM: measured best of {1,3,4,5}
#endif
    // CAP = largest cell population: 256 (static shared memory, 4 CTAs per SM: the usual case) or MEDIUM_MAX (dynamic shared
    // memory, 2 CTAs per SM) for grids where sedimenting drops pile up in a few cells - far cheaper than the global-sort path
    template <int CAP>
    __global__ void __launch_bounds__(CTPB, (CAP == int(SMALL_MAX) ? LCX_COAL_MINB : 2) * 256 / CTPB) k_coal_small(idx_t n_cell, const uint32_t *__restrict__ off, const idx_t *__restrict__ sid,
                                                       rng_src rng, coal_ctx cx, const uint32_t *__restrict__ cell_list, uint32_t n_list)
    {
      extern __shared__ __align__(16) unsigned char coal_dyn_smem[];
      __shared__ __align__(16) uint32_t skey_static[CAP == int(SMALL_MAX) ? GROUPS : 1][CAP == int(SMALL_MAX) ? CAP + KEY_PAD : 1];
      __shared__ unsigned short sperm_static[CAP == int(SMALL_MAX) ? GROUPS : 1][CAP == int(SMALL_MAX) ? CAP : 1];
      uint32_t (*skey)[CAP + KEY_PAD] = CAP == int(SMALL_MAX) ? reinterpret_cast<uint32_t (*)[CAP + KEY_PAD]>(&skey_static[0][0])
                                                               : reinterpret_cast<uint32_t (*)[CAP + KEY_PAD]>(coal_dyn_smem);
      unsigned short (*sperm)[CAP] = CAP == int(SMALL_MAX) ? reinterpret_cast<unsigned short (*)[CAP]>(&sperm_static[0][0])
                                                            : reinterpret_cast<unsigned short (*)[CAP]>(coal_dyn_smem + sizeof(uint32_t) * GROUPS * (CAP + KEY_PAD));
      const int grp = threadIdx.x / CELL_LANES, l = threadIdx.x % CELL_LANES;
      // all cells of the grid, or (cell_list) only the listed ones - the populous cells the usual configuration leaves out
      const uint32_t slot = blockIdx.x * GROUPS + grp;
      const idx_t c = cell_list ? (slot < n_list ? cell_list[slot] : n_cell) : slot;
      uint32_t b = 0, m = 0;
      if (c < n_cell) { b = off[c]; m = off[c + 1] - b; }
      if (m < 2 || m > uint32_t(CAP)) m = 0;               // nothing to pair (or not this launch's cell); keep the lanes for the warp-wide syncs
      const uint32_t m4 = (m + 3u) & ~3u;
      uint32_t *const key = skey[grp];
      unsigned short *const perm = sperm[grp];
      const bool philox = rng.mode != LCX_RNG_INJECT;

      // random sort keys, padded to a multiple of four with the largest value (never "smaller than" anything)
      if (philox)
        for (uint32_t q = l; 4 * q < m; q += CELL_LANES)
        {
          uint32_t w[4];
          philox_cell_block(rng, c, q, w);
#pragma unroll
          for (int i = 0; i < 4; ++i) if (4 * q + i >= m) w[i] = 0xffffffffu;
          *reinterpret_cast<uint4 *>(key + 4 * q) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      else
        for (uint32_t e = l; e < m4; e += CELL_LANES) key[e] = e < m ? rng.un[sid[b + e]] : 0xffffffffu;
      __syncwarp();

      // rank = number of smaller keys; a lane ranks up to four of its SDs per sweep over the keys - as many as the more populous
      // of the warp's two cells needs (three at 40 SDs per cell: a fourth would be a quarter more comparisons for nothing).
      // Equal keys inside the cell share a rank, which shows in the sum of the ranks: m (m - 1) / 2 exactly when all differ.
      const uint32_t m_warp = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
      uint32_t rank_sum = 0;
      for (uint32_t e0 = l; e0 - l < m_warp; e0 += 4 * CELL_LANES)
      {
        const uint32_t left = m_warp - (e0 - l);           // warp-uniform
        if (left > 3 * CELL_LANES)      rank_sum += rank_sweep<4>(key, perm, e0, m, m4);
        else if (left > 2 * CELL_LANES) rank_sum += rank_sweep<3>(key, perm, e0, m, m4);
        else if (left > CELL_LANES)     rank_sum += rank_sweep<2>(key, perm, e0, m, m4);
        else                            rank_sum += rank_sweep<1>(key, perm, e0, m, m4);
      }
#pragma unroll
      for (int o = CELL_LANES / 2; o > 0; o >>= 1) rank_sum += __shfl_xor_sync(0xffffffffu, rank_sum, o);
      __syncwarp();

      // equal keys inside the cell: re-rank with the storage index as tie-break
      const bool hole = rank_sum != m * (m - 1) / 2 && m != 0;
      const unsigned group_lanes = 0xffffu << (threadIdx.x & 16);
      if (__ballot_sync(0xffffffffu, hole) & group_lanes)
      {
        for (uint32_t e = l; e < m; e += CELL_LANES)
        {
          const uint32_t mine = key[e], ms = sid[b + e];
          uint32_t r = 0;
          for (uint32_t j = 0; j < m; ++j) r += (key[j] < mine) || (key[j] == mine && sid[b + j] < ms);
          perm[r] = (unsigned short)e;
        }
      }
      __syncwarp();

      // u01 of the candidate pairs: two words each, written over the keys
      if (philox)
        for (uint32_t q = l; 4 * q < m; q += CELL_LANES)     // m/2 pairs x 2 words = m words
        {
          uint32_t w[4];
          philox_cell_block(rng, c, q | 0x80000000u, w);
          *reinterpret_cast<uint4 *>(key + 4 * q) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      __syncwarp();

      unsigned long long n_coll = 0, n_pairs = 0;
      if (m)
      {
        const real_t pref = cx.dt / cx.dv[c] * coal_scale_factor<real_t>(n_t(m));
        for (uint32_t k = l; 2 * k + 1 < m; k += CELL_LANES)
        {
          const real_t u01 = philox ? u01_from_words(key[2 * k], key[2 * k + 1]) : rng.u01[b + 2 * k];
          try_pair(cx, b + perm[2 * k], b + perm[2 * k + 1], u01, pref, n_coll, n_pairs);
        }
      }
      flush_stats(cx, n_coll, n_pairs);
    }

    // ---- big cells: global sort, then thread per pair ----------------------------------------------------
    __global__ void __launch_bounds__(TPB) k_fill_sid_keys(size_t n, const idx_t *__restrict__ sid, uint32_t *__restrict__ key, uint32_t *__restrict__ val)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i < n) { key[i] = sid[i]; val[i] = uint32_t(i); }
    }
    __global__ void __launch_bounds__(TPB) k_keys_from_un(size_t n, const idx_t *__restrict__ sid, const uint32_t *__restrict__ val, rng_src rng, uint32_t *__restrict__ key)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i < n) key[i] = rng.get_un(sid[val[i]]);
    }
    __global__ void __launch_bounds__(TPB) k_keys_from_ijk(size_t n, const idx_t *__restrict__ ijk, const uint32_t *__restrict__ val, uint32_t *__restrict__ key)
    {
      const size_t i = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (i < n) key[i] = ijk[val[i]];
    }
    __global__ void __launch_bounds__(TPB) k_coal_big(size_t n_part, const uint32_t *__restrict__ off, const idx_t *__restrict__ ijk,
                                                     const uint32_t *__restrict__ perm, rng_src rng, coal_ctx cx)
    {
      const size_t pos = size_t(blockIdx.x) * TPB + threadIdx.x;
      if (pos + 1 >= n_part) return;
      const idx_t c = ijk[pos];               // sorted position and physical position share the cell segments
      const uint32_t b = off[c], en = off[c + 1];
      if (((pos - b) & 1u) != 0u) return;     // only every second SD of a cell starts a pair
      if (pos + 1 >= en) return;              // the last SD of an odd-sized cell stays unpaired
      unsigned long long n_coll = 0, n_pairs = 0;
      try_pair(cx, perm[pos], perm[pos + 1], rng.get_u01(uint32_t(pos)), cx.dt / cx.dv[c] * coal_scale_factor<real_t>(n_t(en - b)), n_coll, n_pairs);
      if (n_pairs)
      {
        atomicAdd(&cx.sc->n_collisions, n_coll);
        atomicAdd(&cx.sc->n_pairs_collided, n_pairs);
      }
    }

    int bit_length(uint64_t v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }
  }

  void coal(lcx_engine *e, real_t dt_sub, const lcx_rng *r)
  {
    if (!e->grouped) throw error("coalescence requested while super-droplets are not grouped by cell");
    const size_t n = e->n_part;
    if (n < 2) return;
    const grid_t &g = e->grid;
    sd_arrays &s = e->S();

    rng_src rng;
    rng.mode = r->mode; rng.un = nullptr; rng.u01 = nullptr; rng.seed = r->seed; rng.call = r->call;
    rng.cell_base = r->cell_base; rng.stream = r->stream;
    if (r->mode == LCX_RNG_INJECT)
    {
      if (!r->un || !r->u01) throw error("lcx_coal: injected random streams are missing");
      densify_sid(e);      // un[] is indexed by the dense storage index
      LCX_CUDA(cudaMemcpyAsync(e->un.p, r->un, n * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
      LCX_CUDA(cudaMemcpyAsync(e->u01.p, r->u01, n * sizeof(real_t), cudaMemcpyHostToDevice, e->stream));
      rng.un = e->un.p; rng.u01 = e->u01.p;
    }

    coal_ctx cx;
    cx.n = s.n.p; cx.rw2 = s.rw2.p; cx.rd3 = s.rd3.p; cx.vt = s.vt.p; cx.kpa = s.kpa.p; cx.rc2 = s.rc2.p;
    cx.dv = e->dv.p;
    cx.kp.kind = e->cfg.kernel;
    cx.kp.has_multiplier = (e->cfg.kernel == KERNEL_GEOMETRIC && e->cfg.n_kernel_user_params == 1);
    cx.kp.user0 = e->cfg.n_kernel_user_params > 0 ? real_t(e->cfg.kernel_user_params[0]) : real_t(0);
    cx.kp.r_max = real_t(e->cfg.kernel_r_max);
    cx.kp.eff = e->eff.p;
    cx.dt = dt_sub;
    cx.multi_kappa = e->cfg.multi_kappa;
    cx.pure_const_multi = e->cfg.pure_const_multi;
    cx.sc = e->scalars.p;

    if (e->max_count <= SMALL_MAX)
    {
      LCX_LAUNCH(e, k_coal_small<int(SMALL_MAX)>, div_up(g.n_cell, GROUPS), CTPB, 0, g.n_cell, e->cell_off.p, s.sid.p, rng, cx, (const uint32_t *)nullptr, 0u);
      return;
    }
    if (e->max_count <= MEDIUM_MAX && g.n_cell > 1)
    {
      constexpr size_t smem = sizeof(uint32_t) * GROUPS * (MEDIUM_MAX + KEY_PAD) + sizeof(unsigned short) * GROUPS * MEDIUM_MAX;
      LCX_CUDA(cudaFuncSetAttribute(k_coal_small<int(MEDIUM_MAX)>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));      // per device; a cheap host call
      // the grid as usual (cells above SMALL_MAX skip themselves), then the few populous cells from the list post_copy made
      LCX_LAUNCH(e, k_coal_small<int(SMALL_MAX)>, div_up(g.n_cell, GROUPS), CTPB, 0, g.n_cell, e->cell_off.p, s.sid.p, rng, cx, (const uint32_t *)nullptr, 0u);
      if (e->n_big)
        LCX_LAUNCH(e, k_coal_small<int(MEDIUM_MAX)>, div_up(e->n_big, GROUPS), CTPB, smem, g.n_cell, e->cell_off.p, s.sid.p, rng, cx,
                   (const uint32_t *)e->big_cells.p, uint32_t(e->n_big));
      return;
    }

    // global path: stable LSD sort by storage index, then random key, then cell
    int in = 0;
    LCX_LAUNCH(e, k_fill_sid_keys, div_up(n, TPB), TPB, 0, n, s.sid.p, e->key[in].p, e->val[in].p);
    in = radix_sort_pairs(e, n, 0, bit_length(e->sid_hi ? e->sid_hi - 1 : 0), in);
    LCX_LAUNCH(e, k_keys_from_un, div_up(n, TPB), TPB, 0, n, s.sid.p, e->val[in].p, rng, e->key[in].p);
    in = radix_sort_pairs(e, n, 0, 32, in);
    if (g.n_cell > 1)
    {
      LCX_LAUNCH(e, k_keys_from_ijk, div_up(n, TPB), TPB, 0, n, s.ijk.p, e->val[in].p, e->key[in].p);
      in = radix_sort_pairs(e, n, 0, bit_length(g.n_cell - 1), in);
    }
    LCX_LAUNCH(e, k_coal_big, div_up(n, TPB), TPB, 0, n, e->cell_off.p, s.ijk.p, e->val[in].p, rng, cx);
  }
}
