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
//   k_cond_range  - the same work with the lanes balanced over SDs instead of cells: a warp owns a run of up to 16
//                   consecutive cells, i.e. ONE contiguous range of SDs, and walks it 32 SDs at a time; the per-cell
//                   constants of the run sit in shared memory, the per-cell sums come from a segmented warp reduction
//                   (the cell index is non-decreasing along the range).  k_cond_cells leaves lanes idle whenever the
//                   four cells of a warp differ in population (measured: 24.7 of 32 lanes in the kernel body at
//                   40 +- 6 SDs per cell); here only the last round of a warp is ragged (31.3 lanes).  Chosen on large
//                   grids: 7.8 -> 7.0 ms at the bench size - less than the lane count suggests, because the 32
//                   consecutive SDs of a round need less alike numbers of trial points than the old 4 x 8.
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
#ifndef LCX_COND_RANGE_TPB
#define LCX_COND_RANGE_TPB 32               // ONE warp per CTA of the run-per-warp kernel (32 CTAs per SM at 64 registers): warps of a CTA finish at different times and a 4-warp CTA holds its slot until the last one has (measured 7.02 -> 6.75 ms; 64 and 256 threads: 7.05, 7.07)
#endif
#ifndef LCX_COND_RANGE_MINB
#define LCX_COND_RANGE_MINB (LCX_COND_MINB * 128 / LCX_COND_RANGE_TPB)
#endif
    constexpr int RTPB = LCX_COND_RANGE_TPB;
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
    // gather-on-read (lcx_engine::pending): where the five attributes still sit, and where the copies go
    struct lazy_args
    {
      const uint32_t *perm; const real_t *rw2, *rd3, *kpa, *vt; const n_t *n;
      real_t *rd3_out, *kpa_out, *vt_out; n_t *n_out;
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
    // Measured and rejected as well (round 2): the straight-line TOMS 748 with the growth law as a __noinline__ function
    // reached from its eight call sites (no phase bookkeeping, 60 KB of SASS): 7.13 ms against 6.70 ms for the machine.
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

    // a warp per run of `run` consecutive cells (run <= RANGE_MAX): lanes balanced over the SDs of the run.
    // Per-cell sums: the cell index is non-decreasing along the run, so a segmented shuffle reduction after every round
    // leaves each cell's share of the round in the first lane of its segment, which adds it to the cell's accumulator in
    // shared memory (one writer per cell and round: the order of summation is fixed by the layout alone, hence reproducible).
    // Measured alternative: per-lane accumulator columns in shared memory instead of the shuffles - no faster, and its
    // 4 KB per warp limit the run to 8 cells (7.2-7.5 ms against 7.0-7.2 ms per launch for runs of 16).
#ifndef LCX_COND_RANGE_MAX
#define LCX_COND_RANGE_MAX 16
#endif
    constexpr int RANGE_MAX = LCX_COND_RANGE_MAX;
    // CLASSED (k_cond_classed): the droplets of a run are first split into two classes by size - drizzle / rain drops (rw > 40 um:
    // Re > 1, the branch of the ventilation factors with a log / exp pair and full cube roots) and everything else - and each class
    // is walked in rounds of its own (a 2-byte order list per run in shared memory, filled from both ends by ballot compaction,
    // ascending inside each class so that the cell index stays non-decreasing along a round).  In the plain order a round of 32
    // holds a few large drops among aerosol and cloud droplets: the expensive branches then run at 5-6 of 32 lanes and everybody
    // waits (ncu on the rain-laden cfg5 slab: 13.6 of 32 lanes per instruction, a quarter of the instructions in Re^0.077).  Same
    // arithmetic per droplet: wet radii are bit-identical to the plain order; a cell's droplets are summed in another order (th, rv
    // agree to rounding, like between the other work distributions).  The host picks the variant from the number of large drops
    // the fall-speed pass of the previous step counted (dev_scalars::n_large), so a given run always takes the same sequence of variants.
    constexpr int ORD_CAP = 1024;                       // longest run (SDs) the order list holds; longer runs keep the plain order
#define LCX_LARGE_RW2 real_t(1.6e-9)                 // (40 um)^2

    template <int MODE, bool LAZY, bool CLASSED>
    __device__ __forceinline__ void cond_range_body(idx_t c_begin, idx_t n_cell, int run, const uint32_t *__restrict__ off, real_t dt, real_t RH_max, const cond_args &a,
                                                    int n_dims, const real_t *__restrict__ dv, int first_step, int keep_after,
                                                    real_t *__restrict__ rw_mom3, real_t *__restrict__ drw_mom3,
                                                    real_t *__restrict__ th, real_t *__restrict__ rv, const lazy_args &z)
    {
      constexpr int WARPS = RTPB / 32;
      __shared__ cond_cell_consts<real_t> s_k[WARPS][RANGE_MAX];
      __shared__ real_t s_m[WARPS][RANGE_MAX][2];
      __shared__ unsigned short s_ord[WARPS][CLASSED ? ORD_CAP : 1];
      const int w = threadIdx.x / 32, l = threadIdx.x % 32;
      // cells [c_begin, n_cell): the whole grid, or one chunk of it whose first cell is a multiple of `run` (same runs either way)
      const size_t c0_ = size_t(c_begin) + (size_t(blockIdx.x) * WARPS + w) * size_t(run);
      if (c0_ >= n_cell) return;                      // whole warps leave; nothing below synchronises across warps
      const idx_t c0 = idx_t(c0_);
      const int nc = int(n_cell - c0 < idx_t(run) ? n_cell - c0 : idx_t(run));
      if (l < nc)
      {
        if (MODE != COND_EXACT) s_k[w][l] = make_cond_consts(load_cell(a, c0 + l), RH_max);
        s_m[w][l][0] = 0; s_m[w][l][1] = 0;
      }
      __syncwarp();
      const uint32_t b = off[c0], en = off[c0 + nc];
      // the two classes: [0, n0) of the order list ascending, the large drops from its far end downwards
      uint32_t n0 = 0, n1 = 0, rounds0 = 0, rounds = 0;
      bool ordered = false;
      if (CLASSED)
      {
        const uint32_t len = en - b;
        const unsigned lt = (1u << l) - 1u;
        ordered = len <= uint32_t(ORD_CAP);
        n0 = len;
        if (ordered)
        {
          n0 = 0;
          for (uint32_t base = 0; base < len; base += 32)
          {
            const uint32_t q = base + l;
            bool big = false, small = false;
            if (q < len)
            {
              const real_t r2 = LAZY ? z.rw2[z.perm[b + q]] : a.rw2[b + q];
              big = r2 > LCX_LARGE_RW2;
              small = !big;
            }
            const unsigned m_big = __ballot_sync(0xffffffffu, big), m_small = __ballot_sync(0xffffffffu, small);
            if (big) s_ord[w][len - 1 - (n1 + __popc(m_big & lt))] = (unsigned short)q;
            if (small) s_ord[w][n0 + __popc(m_small & lt)] = (unsigned short)q;
            n0 += __popc(m_small); n1 += __popc(m_big);
          }
          __syncwarp();
        }
        rounds0 = (n0 + 31) / 32; rounds = rounds0 + (n1 + 31) / 32;
      }
      // storage order: rounds of 32 consecutive SDs; class order: the rounds of the first class, then those of the large drops
      for (uint32_t r = 0, base = b; CLASSED ? r < rounds : base < en; ++r, base += 32)
      {
        uint32_t i = base + l;
        bool act = i < en;
        if (CLASSED)
        {
          const bool second = r >= rounds0;
          const uint32_t q = (second ? r - rounds0 : r) * 32 + l;
          act = q < (second ? n1 : n0);
          i = b + (ordered ? (act ? uint32_t(s_ord[w][second ? (en - b) - 1 - q : q]) : 0u) : q);
        }
        int ci = RANGE_MAX;                           // idle lanes sit behind the last SD: keys stay non-decreasing
        real_t mb = 0, ma = 0;
        if (act)
        {
          const idx_t cc = a.ijk[i] - c0;
          ci = cc < idx_t(nc) ? int(cc) : nc - 1;
          real_t r2, rd3_i, kpa_i, vt_i;
          n_t n_i;
          if (LAZY)      // the attributes still lie in the previous layout; their copies into the new one ride along
          {
            const uint32_t j = z.perm[i];
            r2 = z.rw2[j]; rd3_i = z.rd3[j]; kpa_i = z.kpa[j]; vt_i = z.vt[j]; n_i = z.n[j];
            z.rd3_out[i] = rd3_i; z.kpa_out[i] = kpa_i; z.vt_out[i] = vt_i; z.n_out[i] = n_i;
          }
          else { r2 = a.rw2[i]; rd3_i = a.rd3[i]; kpa_i = a.kpa[i]; vt_i = a.vt[i]; n_i = a.n[i]; }
          const real_t nn = real_t(n_i);
          mb = nn * (r2 * sqrt(r2));
          real_t r2n = r2;
          if (r2 > 0)
            r2n = MODE == COND_EXACT ? advance_rw2(r2, rd3_i, kpa_i, vt_i, load_cell(a, c0 + ci), dt, RH_max)
                                        : advance_rw2_fast<MODE == COND_TOMS748>(r2, rd3_i, kpa_i, vt_i, s_k[w][ci], dt);
          if (LAZY || r2 > 0) a.rw2[i] = r2n;
          ma = nn * (r2n * sqrt(r2n));
        }
        // segmented sums over lanes with equal cell index; afterwards the first lane of every segment holds its total
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const int ko = __shfl_down_sync(0xffffffffu, ci, o);
          const real_t vb = __shfl_down_sync(0xffffffffu, mb, o);
          const real_t va = __shfl_down_sync(0xffffffffu, ma, o);
          if (l + o < 32 && ko == ci) { mb += vb; ma += va; }
        }
        const int prev = __shfl_up_sync(0xffffffffu, ci, 1);
        if (act && (l == 0 || prev != ci)) { s_m[w][ci][0] += mb; s_m[w][ci][1] += ma; }     // one writer per cell and round
        __syncwarp();
      }
      if (l < nc)
      {
        const idx_t c = c0 + l;
        const cond_cell<real_t> cl = load_cell(a, c);      // read again rather than kept in registers across the solve; nobody wrote rv[c] yet
        real_t m3_before = s_m[w][l][0], m3_after = s_m[w][l][1];
        if (n_dims > 0) { m3_before = m3_before / dv[c] / cl.rhod; m3_after = m3_after / dv[c] / cl.rhod; }
        if (!first_step) m3_before = rw_mom3[c];
        if (keep_after) rw_mom3[c] = m3_after;
        const real_t drv = (-m3_before + m3_after) * (cst<real_t>::rho_w() * real_t(4. / 3) * cst<real_t>::pi());
        drw_mom3[c] = drv;
        rv[c] = cl.rv - drv;
        const real_t th_c = th[c];
        th[c] = th_c - drv * d_th_d_rv(cl.T, th_c);
      }
    }

    template <int MODE, bool LAZY>
    __global__ void __launch_bounds__(RTPB, LCX_COND_RANGE_MINB) k_cond_range(idx_t c_begin, idx_t n_cell, int run, const uint32_t *__restrict__ off, real_t dt, real_t RH_max, cond_args a,
                                                       int n_dims, const real_t *__restrict__ dv, int first_step, int keep_after,
                                                       real_t *__restrict__ rw_mom3, real_t *__restrict__ drw_mom3,
                                                       real_t *__restrict__ th, real_t *__restrict__ rv, lazy_args z)
    { cond_range_body<MODE, LAZY, false>(c_begin, n_cell, run, off, dt, RH_max, a, n_dims, dv, first_step, keep_after, rw_mom3, drw_mom3, th, rv, z); }

    template <int MODE, bool LAZY>
    __global__ void __launch_bounds__(RTPB, LCX_COND_RANGE_MINB) k_cond_classed(idx_t c_begin, idx_t n_cell, int run, const uint32_t *__restrict__ off, real_t dt, real_t RH_max, cond_args a,
                                                       int n_dims, const real_t *__restrict__ dv, int first_step, int keep_after,
                                                       real_t *__restrict__ rw_mom3, real_t *__restrict__ drw_mom3,
                                                       real_t *__restrict__ th, real_t *__restrict__ rv, lazy_args z)
    { cond_range_body<MODE, LAZY, true>(c_begin, n_cell, run, off, dt, RH_max, a, n_dims, dv, first_step, keep_after, rw_mom3, drw_mom3, th, rv, z); }

    // ---- staged variant of k_cond_range (TOMS 748 only) --------------------------------------------------------------------------------
    // A round of k_cond_range costs as many growth-law evaluations as its slowest droplet needs (5.1 on the bench workload) while the
    // droplets need 3.65 on average (2: 8 %, 3: 33 %, 4: 45 %, 5: 13 %): a third of the FP64 issue slots belong to lanes that are
    // already done.  Letting a lane start its next droplet at once was measured and rejected: the lanes then sit
    // in different phases of TOMS 748 and the division-heavy interpolation code of each phase runs serially.  The number of
    // evaluations done IS the phase, though, so droplets are re-grouped by it: the same warp-per-run-of-cells decomposition, but
    //   stage 0  takes 32 new droplets through the three evaluations everybody needs (old point, far bracket end, secant point);
    //            whoever is not finished parks its solver state in a per-warp queue in shared memory (ballot-compacted);
    //   stage 1  runs the 4th evaluation on 32 parked droplets at a time - all in the same phase, all lanes busy;
    //            the few still unfinished park again;
    //   stage 2  runs those 32 at a time to completion (5th evaluation for nearly all of them).
    // A stage runs as soon as its queue holds a full warp (later stages first, so no queue exceeds 63 entries); the partial rest is
    // drained at the end of the run.  Same trial points, same arithmetic per droplet: rw2 is bit-identical to k_cond_range.  The
    // per-cell sums of n r^3 are taken in a final sweep over the run in exactly the order k_cond_range uses, so th and rv are
    // bit-identical as well.  Queue entry: 13 doubles + 2 words; 13 KB of shared memory per warp -> 2 warps per CTA, 7 CTAs per SM.
    // MEASURED (B200, cfg4 slab): slower than k_cond_range despite the better lane use - see cond_staged() below; kept as an opt-in.
    constexpr int ST_TPB = 64, ST_WARPS = ST_TPB / 32, QCAP = 64, QF0 = 11, QF1 = 13;
#ifndef LCX_COND_ST_MINB
#define LCX_COND_ST_MINB 7
#endif
    template <int NF>
    struct cond_queue
    {
      real_t f[NF][QCAP];      // a b fa fb d fd e fe c rw2_old rd2 [a0 b0]
      uint32_t idx[QCAP], meta[QCAP];
    };
    struct staged_smem
    {
      cond_cell_consts<real_t> k[ST_WARPS][RANGE_MAX];
      real_t m[ST_WARPS][RANGE_MAX][2];
      cond_queue<QF0> q0[ST_WARPS];
      cond_queue<QF1> q1[ST_WARPS];
    };

    __device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

    template <int NF>
    __device__ __forceinline__ void queue_push(cond_queue<NF> &q, int &count, bool busy, const euler_rw2_solver<real_t> &sv, uint32_t i, int ci)
    {
      const unsigned m = __ballot_sync(0xffffffffu, busy);
      if (busy)
      {
        const int pos = count + __popc(m & lanemask_lt());
        q.f[0][pos] = sv.s.a; q.f[1][pos] = sv.s.b; q.f[2][pos] = sv.s.fa; q.f[3][pos] = sv.s.fb; q.f[4][pos] = sv.s.d; q.f[5][pos] = sv.s.fd;
        q.f[6][pos] = sv.s.e; q.f[7][pos] = sv.s.fe; q.f[8][pos] = sv.c; q.f[9][pos] = sv.rw2_old; q.f[10][pos] = sv.rd2;
        if (NF > 11) { q.f[11][pos] = sv.a0; q.f[12][pos] = sv.b0; }
        q.idx[pos] = i;
        q.meta[pos] = uint32_t(ci) | (uint32_t(sv.phase) << 8) | (uint32_t(sv.bracketing) << 12) | (sv.left << 16);
      }
      count += __popc(m);
      __syncwarp();
    }
    template <int NF>
    __device__ __forceinline__ bool queue_pop(cond_queue<NF> &q, int &count, int lane, euler_rw2_solver<real_t> &sv, uint32_t &i, int &ci)
    {
      const int take = count < 32 ? count : 32;
      const bool busy = lane < take;
      if (busy)
      {
        const int pos = count - take + lane;
        sv.s.a = q.f[0][pos]; sv.s.b = q.f[1][pos]; sv.s.fa = q.f[2][pos]; sv.s.fb = q.f[3][pos]; sv.s.d = q.f[4][pos]; sv.s.fd = q.f[5][pos];
        sv.s.e = q.f[6][pos]; sv.s.fe = q.f[7][pos]; sv.c = q.f[8][pos]; sv.rw2_old = q.f[9][pos]; sv.rd2 = q.f[10][pos];
        if (NF > 11) { sv.a0 = q.f[11][pos]; sv.b0 = q.f[12][pos]; }
        i = q.idx[pos];
        const uint32_t mt = q.meta[pos];
        ci = int(mt & 0xffu); sv.phase = int((mt >> 8) & 0xfu); sv.bracketing = ((mt >> 12) & 1u) != 0u; sv.left = mt >> 16;
      }
      count -= take;
      __syncwarp();
      return busy;
    }

    template <bool LAZY>
    __global__ void __launch_bounds__(ST_TPB, LCX_COND_ST_MINB) k_cond_staged(idx_t n_cell, int run, const uint32_t *__restrict__ off, real_t dt, real_t RH_max, cond_args a,
                                                       int n_dims, const real_t *__restrict__ dv, int first_step, int keep_after,
                                                       real_t *__restrict__ rw_mom3, real_t *__restrict__ drw_mom3,
                                                       real_t *__restrict__ th, real_t *__restrict__ rv, lazy_args z)
    {
      __shared__ staged_smem sm;
      const int w = threadIdx.x / 32, l = threadIdx.x % 32;
      const size_t c0_ = (size_t(blockIdx.x) * ST_WARPS + w) * size_t(run);
      if (c0_ >= n_cell) return;                      // whole warps leave; nothing below synchronises across warps
      const idx_t c0 = idx_t(c0_);
      const int nc = int(n_cell - c0 < idx_t(run) ? n_cell - c0 : idx_t(run));
      if (l < nc)
      {
        sm.k[w][l] = make_cond_consts(load_cell(a, c0 + l), RH_max);
        sm.m[w][l][0] = 0; sm.m[w][l][1] = 0;
      }
      __syncwarp();
      const uint32_t b = off[c0], en = off[c0 + nc];
      cond_queue<QF0> &q0 = sm.q0[w];
      cond_queue<QF1> &q1 = sm.q1[w];
      int n0 = 0, n1 = 0;                              // queue fills (warp-uniform)
      uint32_t next = b;                               // first droplet not yet started

      for (;;)
      {
        // pick the work of this round (warp-uniform): the latest stage that can fill the warp, else new droplets, else leftovers
        int src;
        if (n1 >= 32) src = 2; else if (n0 >= 32) src = 1; else if (next < en) src = 0; else if (n1 > 0) src = 2; else if (n0 > 0) src = 1; else break;

        euler_rw2_solver<real_t> sv;
        uint32_t i = 0;
        int ci = 0;
        bool busy = false;
        real_t rd3_i = 0, rd3_dry = 0, vt_cRe = 0;
        if (src == 0)
        {
          i = next + l;
          const bool act = i < en;
          ci = RANGE_MAX;                               // idle lanes sit behind the last SD: keys stay non-decreasing
          real_t mb = 0;
          sv.start(real_t(0), dt);
          if (act)
          {
            const idx_t cc = a.ijk[i] - c0;
            ci = cc < idx_t(nc) ? int(cc) : nc - 1;
            real_t r2, kpa_i, vt_i;
            n_t n_i;
            if (LAZY)      // the attributes still lie in the previous layout; their copies into the new one ride along
            {
              const uint32_t j = z.perm[i];
              r2 = z.rw2[j]; rd3_i = z.rd3[j]; kpa_i = z.kpa[j]; vt_i = z.vt[j]; n_i = z.n[j];
              z.rd3_out[i] = rd3_i; z.kpa_out[i] = kpa_i; z.vt_out[i] = vt_i; z.n_out[i] = n_i;
              if (!(r2 > 0)) a.rw2[i] = r2;
            }
            else { r2 = a.rw2[i]; rd3_i = a.rd3[i]; kpa_i = a.kpa[i]; vt_i = a.vt[i]; n_i = a.n[i]; }
            mb = real_t(n_i) * (r2 * sqrt(r2));
            if (r2 > 0)
            {
              busy = true;
              sv.start(r2, dt);
              rd3_dry = rd3_i * (real_t(1) - kpa_i); vt_cRe = vt_i * sm.k[w][ci].c_Re;
            }
          }
          // the "before" sums, segment by segment exactly as k_cond_range takes them
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
          {
            const int ko = __shfl_down_sync(0xffffffffu, ci, o);
            const real_t vb = __shfl_down_sync(0xffffffffu, mb, o);
            if (l + o < 32 && ko == ci) mb += vb;
          }
          const int prev = __shfl_up_sync(0xffffffffu, ci, 1);
          if (act && (l == 0 || prev != ci)) sm.m[w][ci][0] += mb;
          __syncwarp();
          next += 32;
        }
        else
        {
          busy = src == 1 ? queue_pop(q0, n0, l, sv, i, ci) : queue_pop(q1, n1, l, sv, i, ci);
          sv.dt = dt;
          if (busy)
          {
            rd3_i = a.rd3[i];                           // written in place (or, gather-on-read, by stage 0 of this warp)
            rd3_dry = rd3_i * (real_t(1) - a.kpa[i]); vt_cRe = a.vt[i] * sm.k[w][ci].c_Re;
          }
        }

        // ONE site for the growth law and the solver step, whatever the stage: 3 evaluations for new droplets, 1 for stage 1,
        // to completion for stage 2
        const int limit = src == 0 ? 3 : src == 1 ? 1 : 1000;
        for (int it = 0; it < limit && __any_sync(0xffffffffu, busy); ++it)
        {
          if (busy)
          {
            real_t result;
            const real_t g = drw2_dt_fast(sv.point(), rd3_i, rd3_dry, vt_cRe, sm.k[w][ci < RANGE_MAX ? ci : 0]);
            if (sv.feed(g, rd3_i, result)) { a.rw2[i] = result; busy = false; }
          }
        }
        if (src == 0) queue_push(q0, n0, busy, sv, i, ci);
        else if (src == 1) queue_push(q1, n1, busy, sv, i, ci);
      }
      __syncwarp();

      // the "after" sums from the radii just written, in the order of k_cond_range: round by round, segment by segment
      for (uint32_t base = b; base < en; base += 32)
      {
        const uint32_t i = base + l;
        const bool act = i < en;
        int ci = RANGE_MAX;
        real_t ma = 0;
        if (act)
        {
          const idx_t cc = a.ijk[i] - c0;
          ci = cc < idx_t(nc) ? int(cc) : nc - 1;
          const real_t r2n = a.rw2[i];
          ma = real_t(a.n[i]) * (r2n * sqrt(r2n));
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const int ko = __shfl_down_sync(0xffffffffu, ci, o);
          const real_t va = __shfl_down_sync(0xffffffffu, ma, o);
          if (l + o < 32 && ko == ci) ma += va;
        }
        const int prev = __shfl_up_sync(0xffffffffu, ci, 1);
        if (act && (l == 0 || prev != ci)) sm.m[w][ci][1] += ma;
        __syncwarp();
      }
      if (l < nc)
      {
        const idx_t c = c0 + l;
        const cond_cell<real_t> cl = load_cell(a, c);
        real_t m3_before = sm.m[w][l][0], m3_after = sm.m[w][l][1];
        if (n_dims > 0) { m3_before = m3_before / dv[c] / cl.rhod; m3_after = m3_after / dv[c] / cl.rhod; }
        if (!first_step) m3_before = rw_mom3[c];
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
    int g_classed = -2;        // -2: not read yet; -1 automatic, 0 never, 1 always
    int g_staged = -1;         // -1: not read yet ($LCX_COND_STAGED, default on); 0 / 1
    int g_layout = -2;         // -2: not read yet; -1: 8 lanes per cell; 0: automatic; 1..RANGE_MAX: cells per warp of k_cond_range
  }

  int cond_layout()
  {
    if (g_layout == -2)
    {
      const char *v = std::getenv("LCX_COND_LAYOUT");
      g_layout = v ? std::atoi(v) : 0;
      if (g_layout < -1 || g_layout > RANGE_MAX) g_layout = 0;
    }
    return g_layout;
  }
  void set_cond_layout(int cells_per_warp) { g_layout = (cells_per_warp < -1 || cells_per_warp > RANGE_MAX) ? 0 : cells_per_warp; }

  // cells per warp of k_cond_range, or 0 for k_cond_cells.  Automatic rule: runs as long as possible (the ragged last round
  // of a warp costs 16 idle lanes on average, so a run should hold a few hundred SDs) while the grid still fills the GPU
  // four times over; small grids keep the 8-lanes-per-cell kernel, which has more parallelism to offer there.
  static int range_run(const lcx_engine *e)
  {
    const int lay = cond_layout();
    if (lay < 0) return 0;
    if (lay > 0) return lay;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
    const size_t warps_wanted = size_t(sms) * 32 * 4;
    size_t run = e->grid.n_cell / warps_wanted;
    if (run > size_t(RANGE_MAX)) run = RANGE_MAX;
    if (run < 4) return 0;
    return int(run);
  }

  // cells per CTA of the run-per-warp kernel if the next lcx_cond would take it (and not its staged form), else 0: chunks of a
  // windowed step must start at multiples of it
  int cond_granule(lcx_engine *e)
  {
    if (cond_staged() && cond_solver() == COND_TOMS748) return 0;
    const int run = e->max_count <= FUSED_MAX ? range_run(e) : 0;
    return run > 0 ? run * (RTPB / 32) : 0;
  }

  int cond_classed()
  {
    if (g_classed == -2) { const char *v = std::getenv("LCX_COND_CLASSED"); g_classed = !v || !v[0] ? -1 : v[0] == '1' ? 1 : v[0] == '0' ? 0 : -1; }
    return g_classed;
  }
  void set_cond_classed(int mode) { g_classed = mode > 0 ? 1 : mode == 0 ? 0 : -1; }

  int cond_staged()
  {
    // off by default: measured on the cfg4 slab it raises the active lanes from 20.3 to 25.8 of 32 and executes 12 % fewer warp
    // instructions, but its 29 KB of queues per CTA leave 14 warps per SM instead of 29 and the issue rate drops from 66 % to 42 %:
    // 9.75 ms against 7.12 ms under ncu (profiles/r02_cond_staged_vs_range_ncu.md)
    // Round 2 also measured a lighter form that parks only the stragglers (droplets needing a fifth evaluation, 13.6 %: 11 doubles
    // each, one queue, 24-28 warps per SM): 7.3-8.0 ms against 6.7 ms - the generic loop around the one evaluation site costs more
    // issue slots than the nearly empty fifth pass it removes.  Not kept.
    if (g_staged < 0) { const char *v = std::getenv("LCX_COND_STAGED"); g_staged = (v && v[0] == '1') ? 1 : 0; }
    return g_staged;
  }
  void set_cond_staged(int on) { g_staged = on ? 1 : 0; }

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

    const int run = e->max_count <= FUSED_MAX ? range_run(e) : 0;
    if (run > 0 && mode == COND_TOMS748 && cond_staged() == 1 && e->win_end == 0)
    {
      const unsigned blocks = div_up(div_up(g.n_cell, run), ST_WARPS);
      lazy_args z = {};
      if (e->pending & lcx_engine::PENDING_ATTR)
      {
        sd_arrays &o = e->A();
        z = {e->pending_perm.p, o.rw2.p, o.rd3.p, o.kpa.p, o.vt.p, o.n.p, s.rd3.p, s.kpa.p, s.vt.p, s.n.p};
        e->pending &= ~unsigned(lcx_engine::PENDING_ATTR);
        LCX_LAUNCH(e, k_cond_staged<true>, blocks, ST_TPB, 0, g.n_cell, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                   int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z);
        return;
      }
      LCX_LAUNCH(e, k_cond_staged<false>, blocks, ST_TPB, 0, g.n_cell, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                 int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z);
      return;
    }
    const bool windowed = e->win_end != 0;
    if (run > 0)
    {
      // chunked step_sync: only the window's cells; the chunks arrive in order and the last one (ending at n_cell) closes a
      // pending gather-on-read re-layout
      const idx_t c_begin = windowed ? e->win_begin : 0, c_end = windowed ? e->win_end : g.n_cell;
      if (windowed && (c_begin % idx_t(run) != 0 || c_end <= c_begin || c_end > g.n_cell)) throw error("lcx_cond: cell window not aligned to the kernel's runs");
      const unsigned blocks = div_up(div_up(c_end - c_begin, run), RTPB / 32);
      // droplets walked class by class when the previous call counted enough large drops (more than one in 64; cond_classed())
      const int cls = cond_classed();
      const bool classed = cls == 1 || (cls < 0 && e->n_large * 64 * 64 > e->n_part);      // n_large: counted in one CTA of 64 (lcx_cells.cu)
      lazy_args z = {};
      if (e->pending & lcx_engine::PENDING_ATTR)      // consume the pending re-layout: read through the permutation from the old buffer set, write everything into the new one
      {
        sd_arrays &o = e->A();
        z = {e->pending_perm.p, o.rw2.p, o.rd3.p, o.kpa.p, o.vt.p, o.n.p, s.rd3.p, s.kpa.p, s.vt.p, s.n.p};
        if (c_end == g.n_cell) e->pending &= ~unsigned(lcx_engine::PENDING_ATTR);
        if (classed)
          LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond_classed<M, true>), blocks, RTPB, 0, c_begin, c_end, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                                            int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z))
        else
          LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond_range<M, true>), blocks, RTPB, 0, c_begin, c_end, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                                            int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z))
        return;
      }
      if (classed)
        LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond_classed<M, false>), blocks, RTPB, 0, c_begin, c_end, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                                          int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z))
      else
        LCX_BY_COND_MODE(mode, LCX_LAUNCH(e, (k_cond_range<M, false>), blocks, RTPB, 0, c_begin, c_end, run, e->cell_off.p, dt_sub, RH_max, a, g.n_dims, e->dv.p,
                                          int(step == 0), keep_after, e->rw_mom3.p, e->drw_mom3.p, e->th.p, e->rv.p, z))
      return;
    }
    if (windowed) throw error("lcx_cond: a cell window needs the run-per-warp kernel (lcx_cond_granule says when)");
    finish_pending(e, lcx_engine::PENDING_ATTR);      // the other condensation kernels work on the current layout only
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
