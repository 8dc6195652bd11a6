// Internal definition of the engine behind include/lcx_b200.h: device buffers, launch helpers and the
// declarations shared by the translation units (lcx_api.cu, lcx_sort.cu, lcx_cells.cu, lcx_cond.cu,
// lcx_coal.cu, lcx_transport.cu, lcx_layout.cu, lcx_diag.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lcx_b200.h"
#include "lcx_physics.h"

namespace lcx
{
#ifdef LCX_F32
  typedef float real_t;           // the single-precision engine: same sources, liblcx_b200_f32.so (build.py)
#else
  typedef double real_t;
#endif
  typedef uint32_t idx_t;         // SD indices / cell indices / storage indices (n_sd_max < 2^32 per slab)

  struct error : std::runtime_error { explicit error(const std::string &s) : std::runtime_error(s) {} };

  inline void cuda_check(cudaError_t st, const char *what, const char *file, int line)
  {
    if (st != cudaSuccess)
      throw error(std::string(what) + ": " + cudaGetErrorString(st) + " (" + file + ":" + std::to_string(line) + ")");
  }
#define LCX_CUDA(call) ::lcx::cuda_check((call), #call, __FILE__, __LINE__)

  // raw device allocation with a size, freed by the engine
  template <class T>
  struct dbuf
  {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count)
    {
      release();
      n = count;
      if (count) LCX_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    size_t bytes() const { return n * sizeof(T); }
  };

  // one set of per-SD arrays (structure of arrays)
  struct sd_arrays
  {
    dbuf<n_t> n;
    dbuf<real_t> rd3, rw2, kpa, vt, x, y, z;
    dbuf<idx_t> sid, ijk;
    dbuf<real_t> pp_rv, pp_th, pp_rh, pp_p;      // per-particle sub-stepping: the SD's own record of rv, th, rhod, p (sstp_tmp_*)
    dbuf<real_t> rc2;                            // critical radius squared at rc2_T (only with sstp_cond_act > 1)
    void alloc_pp(size_t cap, bool with_p) { pp_rv.alloc(cap); pp_th.alloc(cap); pp_rh.alloc(cap); if (with_p) pp_p.alloc(cap); }
    void alloc(size_t cap, bool has_x, bool has_y, bool has_z)
    {
      n.alloc(cap); rd3.alloc(cap); rw2.alloc(cap); kpa.alloc(cap); vt.alloc(cap);
      if (has_x) x.alloc(cap);
      if (has_y) y.alloc(cap);
      if (has_z) z.alloc(cap);
      sid.alloc(cap); ijk.alloc(cap);
    }
    void release()
    { n.release(); rd3.release(); rw2.release(); kpa.release(); vt.release(); x.release(); y.release(); z.release(); sid.release(); ijk.release();
      pp_rv.release(); pp_th.release(); pp_rh.release(); pp_p.release(); rc2.release(); }
  };

  // grid description handed to kernels by value
  struct grid_t
  {
    int n_dims, nx, ny, nz;
    real_t dx, dy, dz, x0, y0, z0, x1, y1, z1;
    idx_t n_cell;
    int halo_size;          // cells of x-halo on each side of the Courant fields (2 for pred_corr, else 0)
    idx_t halo_x;           // halo_size * (cells per x-column): offset of the first real cell in halo-extended numbering
    int class_bits;         // low bits of the re-layout sort key that order the SDs of a cell by size class (0..3)
  };

  // Re-layout sort key: cell index in the high bits, a coarse size class (factor 4 in radius per class) in the bits that
  // the last 8-bit radix digit leaves unused anyway.  Inside a cell the SDs then lie ordered by size, so the lanes that
  // run the condensation root solve in lock-step work on similar droplets (similar iteration counts).  Any order inside
  // a cell is valid: pairing for coalescence is by random key, sums over a cell carry a tolerance.
  __host__ __device__ inline uint32_t relayout_key(const grid_t &g, idx_t cell, real_t rw2)
  {
    if (g.class_bits == 0) return cell;
#if defined(__CUDA_ARCH__)
    const int e = int((__double_as_longlong(double(rw2)) >> 52) & 0x7ff) - 1023;
#else
    int e = 0; if (rw2 > 0) { (void)frexp(double(rw2), &e); e -= 1; } else e = -1023;
#endif
    int cls = (e + 50) >> 2;            // r < 0.1 um -> 0, 0.1-0.4 -> 1, 0.4-1.6 -> 2, 1.6-6.4 -> 3, 6.4-25 -> 4, 25-100 -> 5, ...
    cls = cls < 0 ? 0 : cls > 7 ? 7 : cls;
    return (cell << g.class_bits) | uint32_t(cls >> (3 - g.class_bits));
  }
  __host__ __device__ inline uint32_t relayout_dead_key(const grid_t &g) { return g.n_cell << g.class_bits; }

  // device-side scalars that kernels produce and the host occasionally reads back
  struct dev_scalars
  {
    unsigned int n_part;            // live SDs after the last lcx_post_copy
    unsigned int max_count;         // largest per-cell SD count
    unsigned int n_lft, n_rgt;      // SDs that left through the left / right face in the last lcx_transport
    unsigned int increase_sstp_coal;
    unsigned int mig_timeout;       // a neighbour's delivery did not arrive (cross-process wait gave up)
    unsigned long long n_collisions, n_pairs_collided;
    double puddle[8];               // liq_vol, dry_vol, liq_num, prtcl_num, accumulated; [4], [5]: dry volume / SDs lost through the lid
    unsigned long long rcyc_zero, rcyc_one, rcyc_max;   // SDs with n == 0, with n == 1, largest n (recycling)
    unsigned int n_big;             // cells holding more than BIG_CELL super-droplets (listed in lcx_engine::big_cells)
    unsigned int n_large;           // drops with rw > 40 um counted by the condensation kernels since the last lcx_post_copy
  };
  constexpr unsigned BIG_CELL = 256;  // largest cell population the per-cell coalescence kernel takes in its usual configuration
}

struct lcx_engine
{
  lcx_config cfg;
  lcx::grid_t grid;
  int device = 0;
  cudaStream_t stream = nullptr;
  // Courant fields are needed late in the step (transport): their host->device copies run on a second stream, so that
  // they overlap the condensation kernel; consumers call lcx::wait_courant first
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t courant_ready = nullptr, main_mark = nullptr;
  bool courant_pending = false;
  // th, rv, rhod, p uploads use the same copy stream: right after a re-layout the engine's stream is still busy with the
  // gather (which touches no cell field), and the next step's fields can already travel (pre_gather marks the point the
  // copy stream has to wait for); every entry point that consumes cell fields first orders itself after scalars_ready
  cudaEvent_t scalars_ready = nullptr, pre_gather = nullptr;
  bool scalars_pending = false, upload_batch_open = false, tail_is_gather = false;
  // Chunked step_sync (lcx_set_cell_window): the host layer uploads th / rv / rhod of one x-chunk, runs hskpng_Tpr and the
  // condensation kernel on that chunk's cells only, and reads th / rv of the chunk back on a third stream while the next chunk
  // computes - the host<->device copies of step_sync then hide behind the condensation kernel instead of bracketing it
  lcx::idx_t win_begin = 0, win_end = 0;          // [begin, end) in cells; 0, 0 = the whole grid
  cudaStream_t d2h_stream = nullptr;
  cudaStream_t win_stream = nullptr, win_home = nullptr;   // odd chunks run on win_stream (`stream` points at it meanwhile), win_home keeps the engine's own
  cudaEvent_t win_join = nullptr;
  int win_count = 0;
  cudaEvent_t d2h_mark = nullptr;
  bool d2h_open = false;
  uint64_t launches = 0;

  size_t cap = 0;                // n_sd_max
  size_t n_part = 0;             // live SDs (host copy)
  size_t n_tail = 0;             // SDs appended (migration) since the last post_copy, already counted in n_part
  unsigned max_count = 0;        // host copy of the largest cell population
  unsigned n_big = 0;            // host copy: number of cells with more than BIG_CELL super-droplets
  lcx::dbuf<uint32_t> big_cells; // their indices (any order)
  bool grouped = false;          // SD arrays physically grouped by cell, cell_off valid
  // storage indices: every sid is < sid_hi; dense means {sid} = [0, n_part).  With injected random streams (and for
  // get_attr) they must be dense, so removals are followed by a re-numbering; with Philox any unique, order-preserving
  // index will do and the re-numbering is postponed until something needs it (dense_always = false)
  size_t sid_hi = 0;
  bool sid_dense = true;
  bool dense_always = true;

  lcx::sd_arrays sd[2];          // current / alternate (re-layout target)
  int cur = 0;
  lcx::sd_arrays &S() { return sd[cur]; }
  lcx::sd_arrays &A() { return sd[cur ^ 1]; }

  // per-SD scratch
  lcx::dbuf<uint32_t> key[2], val[2], un, flag;
  lcx::dbuf<lcx::real_t> u01, n_filtered, tmp_real;
  lcx::dbuf<uint32_t> perm;      // sorted position -> physical index (big-cell coalescence)

  // per-cell
  lcx::dbuf<lcx::real_t> th, rv, rhod, p, T, RH, eta, dv, lambda_D, lambda_K;
  lcx::dbuf<lcx::real_t> sstp_tmp_rv, sstp_tmp_th, sstp_tmp_rh;
  lcx::dbuf<lcx::real_t> drw_mom3, rw_mom3, count_mom, mom_partial;
  lcx::dbuf<lcx::real_t> cell_tmp4;        // 4 reals per cell: per-cell parts of the Beard (1977) fall-speed correction
  lcx::dbuf<lcx::real_t> courant_x, courant_y, courant_z, w_LS;
  // Gather-on-read (default; LCX_LAZY_GATHER=0 restores the eager gather): after a re-layout only sid and the per-particle records are moved at once.  The
  // other attributes of the live SDs stay in the OLD buffer set A() with pending_perm[new position] = old position until their
  // first consumer, which reads them through the permutation and writes them into S():
  //   PENDING_ATTR  n, rd3, rw2, kpa, vt  - the condensation kernel (FP64-bound, 7 % of the DRAM bandwidth: free of charge);
  //   PENDING_XYZ   x, y, z               - the transport kernel (nothing reads positions before it).
  // Every other consumer first runs finish_pending() for what it needs.
  enum { PENDING_ATTR = 1u, PENDING_XYZ = 2u };
  bool lazy_gather = true;
  unsigned pending = 0;
  lcx::dbuf<uint32_t> pending_perm;
  size_t n_grouped = 0;          // SDs [0, n_grouped) still lie in the segments described by cell_off / ijk of the last re-layout
  lcx::dbuf<uint32_t> mv_scan;   // n_cell + 2: movers per old cell, then their exclusive scan (movers-only re-layout)
  lcx::dbuf<uint32_t> arr_off, cell_off_new;   // n_cell + 2 each: arrivals per cell, segment starts being built
  lcx::dbuf<uint32_t> cell_off;  // n_cell + 2 entries: start of each cell's segment; [n_cell] = n_part, [n_cell+1] = total incl. dead

  // tables
  lcx::dbuf<lcx::real_t> vt0, eff;

  // radix-sort / scan scratch
  lcx::dbuf<uint32_t> hist, scan_tmp;

  // x-slab migration (include/lcx_b200.h): lists of leavers per side (storage index, physical index; ping-pong for the sort),
  // this engine's inboxes and the neighbours' inboxes it delivers into
  lcx::dbuf<uint32_t> mig_key[2][2], mig_val[2][2];
  size_t mig_cap = 0;
  int mig_n_real = 0;
  lcx::dbuf<unsigned char> inbox[2];
  struct mig_remote { unsigned char *base = nullptr; size_t cap = 0; bool ipc = false; } remote[2];
  cudaEvent_t ev_put = nullptr;
  unsigned mig_seq = 0, halo_seq = 0;   // halo_seq: deliveries of Courant halo planes (lcx_halo_put / lcx_halo_take)
  size_t n_large = 0;            // dev_scalars::n_large as of the last lcx_post_copy
  size_t keys_ready = 0;         // key[0] / val[0] already hold the re-layout sort keys of SDs [0, keys_ready) (written by k_transport)

  lcx::dbuf<lcx::dev_scalars> scalars;
  lcx::dev_scalars *h_scalars = nullptr;   // pinned host mirror

  lcx::dbuf<double> red_partial;           // deterministic two-pass reductions

  bool selected = false;                   // a selector filled n_filtered

  // optional per-kernel timing (CUDA events around every launch on the engine's stream), used by bench.py
  struct prof_rec { const char *name; cudaEvent_t t0, t1; };
  bool profiling = false;
  std::vector<prof_rec> prof;
  cudaEvent_t timer0 = nullptr, timer1 = nullptr;

  ~lcx_engine();
};

namespace lcx
{
  // LCX_TRACE=1: every launch is announced on stderr and waited for (finds the kernel that hangs or faults)
  inline bool trace_launches() { static const bool on = [] { const char *v = std::getenv("LCX_TRACE"); return v && v[0] == '1'; }(); return on; }

  inline unsigned div_up(size_t a, size_t b) { return unsigned((a + b - 1) / b); }

  // ---- lcx_sort.cu -----------------------------------------------------------------------------------
  // exclusive prefix sum of n uint32 values, in place
  void exclusive_scan_u32(lcx_engine *e, uint32_t *data, size_t n);
  // stable LSD radix sort of (key, value) pairs on bits [bit_lo, bit_hi); result ends in key[*out]/val[*out]
  // (ping-pong between the two buffers of e->key / e->val; `in` is the buffer that holds the input)
  int radix_sort_pairs(lcx_engine *e, size_t n, int bit_lo, int bit_hi, int in);
  // the same on caller-supplied ping-pong buffers (migration lists keep the engine's key/val buffers untouched)
  int radix_sort_pairs(lcx_engine *e, size_t n, int bit_lo, int bit_hi, uint32_t *const key[2], uint32_t *const val[2], int in);

  // ---- lcx_layout.cu ---------------------------------------------------------------------------------
  void compute_cell_offsets(lcx_engine *e, const uint32_t *sorted_keys, size_t n_total);
  void post_copy(lcx_engine *e, bool rcyc, bool keep_all);
  void finish_pending(lcx_engine *e, unsigned what = 3u);      // completes (the named parts of) a gather-on-read re-layout
  void scatter_attr_by_sid(lcx_engine *e, int attr, real_t *dst);
  void scatter_n_by_sid(lcx_engine *e, uint64_t *dst);
  void densify_sid(lcx_engine *e);

  // ---- lcx_cells.cu ----------------------------------------------------------------------------------
  void hskpng_Tpr(lcx_engine *e);
  void hskpng_mfp(lcx_engine *e);
  void hskpng_vterm(lcx_engine *e, bool only_invalid);
  void sstp_percell_step(lcx_engine *e, int step, int sstp, bool var_rho);
  void sstp_save(lcx_engine *e);
  void update_th_rv(lcx_engine *e);

  // ---- lcx_diag.cu -----------------------------------------------------------------------------------
  // per-cell sum over the cell's SDs of weight(i) * pow(attr(i), power); weight = n (as real) or n_filtered
  void cell_moment(lcx_engine *e, const real_t *weight_or_null, const real_t *attr, real_t power, bool specific, real_t *out);
  void moms_select(lcx_engine *e, int kind, int attr, real_t lo, real_t hi, bool cons);
  void diag_sd_conc(lcx_engine *e);
  void diag_precip_rate(lcx_engine *e);
  void diag_max_rw(lcx_engine *e);
  void diag_mass_dens(lcx_engine *e, int attr, real_t rad, real_t sig0, real_t xp);
  void diag_vel_div(lcx_engine *e, real_t dt);

  // ---- lcx_cond.cu -----------------------------------------------------------------------------------
  int cond_solver();                      // COND_SECANT / COND_TOMS748 / COND_EXACT: lcx_set_cond_solver, else $LCX_COND_SOLVER, else TOMS 748
  void set_cond_solver(int mode);
  int cond_layout();                      // lcx_set_cond_layout, else $LCX_COND_LAYOUT, else 0 (automatic)
  void set_cond_layout(int cells_per_warp);
  int cond_classed();                     // lcx_set_cond_classed, else $LCX_COND_CLASSED, else -1 (automatic): droplets walked class by class
  void set_cond_classed(int mode);
  int cond_staged();                      // lcx_set_cond_staged, else $LCX_COND_STAGED, else on: the phase-grouped variant of the range kernel
  void set_cond_staged(int on);
  void cond(lcx_engine *e, real_t dt_sub, real_t RH_max, int step, int sstp);
  int cond_granule(lcx_engine *e);       // cells per CTA of the run-per-warp condensation kernel, 0 if it would not be chosen
  void cond_perparticle(lcx_engine *e, real_t dt, real_t RH_max, int sstp, bool mix);
  void cond_perparticle_adaptive(lcx_engine *e, real_t dt, real_t RH_max, int sstp_max, int sstp_act, real_t drw2_eps, real_t drw2_max);
  void hskpng_rc2(lcx_engine *e);
  void pp_save(lcx_engine *e, size_t first);                 // sstp_tmp_x[i] = x[ijk[i]] for i >= first (sstp_save.ipp, init_perparticle_sstp.ipp)
  void cell_sum(lcx_engine *e, const real_t *per_sd, real_t *out);       // plain per-cell sums of a per-SD array
  void cell_max_sid(lcx_engine *e, real_t *out);             // largest storage index in every cell (as real)
  void wait_courant(lcx_engine *e);      // orders the engine's stream after pending Courant-field uploads

  // ---- lcx_coal.cu -----------------------------------------------------------------------------------
  void coal(lcx_engine *e, real_t dt_sub, const lcx_rng *rng);

  // ---- lcx_transport.cu ------------------------------------------------------------------------------
  void transport(lcx_engine *e, const lcx_transport_opts *o);
  void migr_put(lcx_engine *e, int64_t *n_lft, int64_t *n_rgt);
  void migr_take(lcx_engine *e, lcx_engine *rgt, lcx_engine *lft, int64_t *n_from_rgt, int64_t *n_from_lft);
  size_t halo_values(const grid_t &g);      // Courant values one neighbour delivers per step: halo_x + halo_y + halo_z
  void halo_put(lcx_engine *e);
  void halo_take(lcx_engine *e);

  // layout of an inbox allocation: 256 bytes of headers (one per parity), then n[2][cap], then real[2][n_real][cap], then the
  // Courant halo planes of the neighbour halo[2][halo_n] (predictor-corrector advection only, else halo_n = 0)
  struct mig_hdr { unsigned int count, seq, halo_seq, pad1; };
  constexpr size_t MIG_HDR_BYTES = 256;
  inline size_t inbox_sd_bytes(size_t cap, int n_real) { return MIG_HDR_BYTES + 2 * cap * (sizeof(n_t) + size_t(n_real) * sizeof(real_t)); }
  inline size_t inbox_bytes(size_t cap, int n_real, size_t halo_n) { return inbox_sd_bytes(cap, n_real) + 2 * halo_n * sizeof(real_t); }
  inline real_t *box_halo(unsigned char *b, size_t cap, int n_real, size_t halo_n, int parity)
  { return reinterpret_cast<real_t *>(b + inbox_sd_bytes(cap, n_real)) + size_t(parity) * halo_n; }
  inline mig_hdr *box_hdr(unsigned char *b, int parity) { return reinterpret_cast<mig_hdr *>(b) + parity; }
  inline n_t *box_n(unsigned char *b, size_t cap, int parity) { return reinterpret_cast<n_t *>(b + MIG_HDR_BYTES) + size_t(parity) * cap; }
  inline real_t *box_real(unsigned char *b, size_t cap, int n_real, int parity)
  { return reinterpret_cast<real_t *>(b + MIG_HDR_BYTES + 2 * cap * sizeof(n_t)) + size_t(parity) * cap * size_t(n_real); }

  real_t *attr_ptr(lcx_engine *e, int attr);

  // ---- lcx_init.cu -----------------------------------------------------------------------------------
  void sd_append_sd_conc(lcx_engine *e, size_t per_cell, real_t log_rd_min, real_t log_rd_max, real_t kappa, real_t RH_max,
                         uint64_t seed, uint32_t stream, uint64_t call, real_t *rd3_host);

  // ---- Philox4x32-10 (Salmon, Moraes, Dror & Shaw, SC'11) --------------------------------------------
  struct philox_key { uint32_t k0, k1; };
  __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
  {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], philox_key k)
  {
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
      philox_round(c, k.k0, k.k1);
      k.k0 += 0x9E3779B9u; k.k1 += 0xBB67AE85u;
    }
  }
  // 53-bit mantissa from two Philox words: uniform in [0, 1)
  __device__ __forceinline__ double philox_u01(uint32_t w0, uint32_t w1)
  {
    const uint64_t bits = (uint64_t(w0) << 21) ^ (uint64_t(w1) >> 11);
    return double(bits & ((1ull << 53) - 1)) * (1.0 / 9007199254740992.0);
  }
}

// runs CALL once with the compile-time constant M set to the run-time solver mode
#define LCX_BY_COND_MODE(mode, CALL)                                          \
  switch (mode)                                                               \
  {                                                                           \
    case lcx::COND_EXACT:   { constexpr int M = lcx::COND_EXACT; CALL; } break;     \
    case lcx::COND_SECANT:  { constexpr int M = lcx::COND_SECANT; CALL; } break;    \
    default:                { constexpr int M = lcx::COND_TOMS748; CALL; } break;   \
  }

#define LCX_LAUNCH(e, kernel, grid, block, smem, ...)                          \
  do {                                                                         \
    lcx_engine::prof_rec lcx_pr_ = {#kernel, nullptr, nullptr};                \
    if ((e)->profiling) {                                                      \
      LCX_CUDA(cudaEventCreate(&lcx_pr_.t0));                                  \
      LCX_CUDA(cudaEventCreate(&lcx_pr_.t1));                                  \
      LCX_CUDA(cudaEventRecord(lcx_pr_.t0, (e)->stream));                      \
    }                                                                          \
    if (::lcx::trace_launches()) std::fprintf(stderr, "[lcx dev %d] %s grid %u\n", (e)->device, #kernel, dim3(grid).x); \
    kernel<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__);             \
    LCX_CUDA(cudaGetLastError());                                              \
    if (::lcx::trace_launches()) LCX_CUDA(cudaStreamSynchronize((e)->stream)); \
    if ((e)->profiling) {                                                      \
      LCX_CUDA(cudaEventRecord(lcx_pr_.t1, (e)->stream));                      \
      (e)->prof.push_back(lcx_pr_);                                            \
    }                                                                          \
    ++(e)->launches;                                                           \
  } while (0)
