/* lcx_b200.h - C ABI of the B200-native super-droplet engine (liblcx_b200.so).
 *
 * One `lcx_engine` owns the super-droplets (SDs) and Eulerian cell fields of ONE x-slab on ONE GPU and
 * exposes the passes of libcloudph++'s Lagrangian hot path as plain C calls: POD arguments, raw
 * pointers and sizes, `int` status (0 = ok, message via lcx_last_error()).  No C++ types, exceptions,
 * Thrust or torch types cross this boundary.  The C++ host layer (libcloudphxx_b200/host) implements
 * the reference's `lgrngn::particles_t<real_t, CUDA | multi_CUDA>` state machine on top of it; any other
 * host language can bind the same symbols.
 *
 * Every entry point names the reference routine whose effect it reproduces (paths relative to the
 * reference repository root).  All work is enqueued on the engine's own CUDA stream; calls that return
 * data to the host synchronise that stream.  There is no CPU fallback: every pass is a hand-written
 * sm_100a kernel and creation fails if no CUDA device is usable.
 *
 * Layout in HBM: structure-of-arrays, physically grouped by grid cell after every lcx_post_copy
 * (so per-cell segments are contiguous); the reference's storage index of each SD is carried as `sid`
 * and is what orders ties, random-stream look-ups and lcx_get_attr output.
 */
#ifndef LCX_B200_H
#define LCX_B200_H

#include <stddef.h>
#include <stdint.h>

/* The single-precision engine is the same source compiled with LCX_F32: real = float, every symbol below carries the suffix  */
/* _f32 (lcx_create_f32, ...), `void *` array arguments then point to floats; scalars stay double in the signatures.           */
#ifdef LCX_F32
#include "lcx_b200_f32_names.h"
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lcx_engine lcx_engine;

/* x-boundary handling of this slab (reference src/detail/bcond.hpp:9-15) */
enum lcx_bcond_t { LCX_BCOND_SHAREDMEM = 0, LCX_BCOND_DISTMEM = 1, LCX_BCOND_OPEN = 3 };

/* per-cell fields addressable through lcx_cells_set / lcx_cells_get */
enum lcx_field_t
{
  LCX_F_TH = 0, LCX_F_RV, LCX_F_RHOD, LCX_F_P, LCX_F_COURANT_X, LCX_F_COURANT_Y, LCX_F_COURANT_Z,
  LCX_F_T, LCX_F_RH, LCX_F_ETA, LCX_F_DV, LCX_F_W_LS, LCX_F_MOM, LCX_F_COUNT
};

/* per-SD attributes */
enum lcx_attr_t
{
  LCX_A_RD3 = 0, LCX_A_RW2, LCX_A_KPA, LCX_A_VT, LCX_A_X, LCX_A_Y, LCX_A_Z, LCX_A_N, LCX_A_SID, LCX_A_IJK, LCX_A_COUNT
};

/* selectors (reference src/impl/diagnose_SD_attributes/particles_impl_moms.ipp:50-234) */
enum lcx_select_t
{
  LCX_SEL_ALL = 0,      /* moms_all                                  */
  LCX_SEL_RANGE,        /* moms_rng: lo <= attr < hi                 */
  LCX_SEL_GT0,          /* moms_gt0: attr > 0                        */
  LCX_SEL_RW_GE_RC,     /* moms_cmp(rw2, rc2): activated droplets    */
  LCX_SEL_RH_GE_SC      /* moms_ge0(RH - S_crit)                     */
};

/* source of the uniform random numbers consumed by coalescence */
enum lcx_rng_mode_t
{
  LCX_RNG_INJECT = 0,   /* caller supplies un[n_part] (by storage index) and u01[n_part] (by sorted position) */
  LCX_RNG_PHILOX = 1    /* counter-based Philox4x32-10 evaluated in the kernels                               */
};

typedef struct
{
  int device;                 /* CUDA device ordinal (-1: current)                                            */
  int real_bytes;             /* 8 = double (4 = float reserved)                                              */
  int nx, ny, nz;             /* cells of THIS slab (0 = dimension absent)                                    */
  double dx, dy, dz;
  double x0, y0, z0, x1, y1, z1;      /* Lagrangian domain of this slab, slab-local coordinates                */
  uint64_t n_sd_max;          /* capacity                                                                     */
  int kernel;                 /* kernel_t ordinal                                                             */
  int terminal_velocity;      /* vt_t ordinal                                                                 */
  int adve_scheme;            /* as_t ordinal                                                                 */
  int RH_formula;             /* RH_formula_t ordinal                                                         */
  int th_dry, const_p;
  int n_kernel_user_params;
  double kernel_user_params[4];
  double kernel_r_max;        /* largest tabulated radius [um] of an efficiency table                         */
  int open_side_walls, periodic_topbot_walls;
  int bcond_lft, bcond_rgt;   /* lcx_bcond_t                                                                  */
  double lft_x1, rgt_x0;      /* x1 of the left neighbour / x0 of the right neighbour (distmem)               */
  int multi_kappa;            /* more than one (kappa) species: kappa is mixed on coalescence                 */
  int pure_const_multi;       /* constant-multiplicity run: report probabilities >= 1                         */
  int allow_sstp_cond;        /* keep old rv/th/rhod for per-cell condensation sub-stepping                   */
  int exact_sstp_cond;        /* per-particle sub-stepping: every SD carries its own rv, th, rhod (, p) history */
  int sstp_cond_act;          /* > 1: SDs crossing their critical radius sub-step this many times; keeps rc2 per SD */
  double rc2_T;               /* temperature [deg C] at which that critical radius is evaluated (opts_init.rc2_T) */
} lcx_config;

typedef struct
{
  int mode;                   /* lcx_rng_mode_t                                                               */
  const uint32_t *un;         /* INJECT: host array, n_part entries, indexed by storage index (sid)           */
  const void *u01;            /* INJECT: host array of real, n_part entries, indexed by sorted position       */
  uint64_t seed, call;        /* PHILOX: key (low 32 bits used) and per-call counter                          */
  uint32_t cell_base;         /* PHILOX: global index of this slab's first cell (streams differ between slabs) */
  uint32_t stream;            /* PHILOX: second key word, e.g. the slab's rank                                 */
} lcx_rng;

typedef struct
{
  int adve, sedi, subs;       /* processes to apply in lcx_transport                                          */
  int adve_scheme;            /* as_t actually used this step (pred_corr may fall back to euler)              */
  double dt;
} lcx_transport_opts;

/* ---- life cycle ---------------------------------------------------------------------------------- */
const char *lcx_last_error(void);
const char *lcx_version(void);
int  lcx_device_count(void);
int  lcx_create(const lcx_config *cfg, lcx_engine **out);       /* src/impl/particles_impl.ipp:327-544       */
int  lcx_destroy(lcx_engine *e);
int  lcx_sync(lcx_engine *e);                                   /* wait for the engine's stream              */
void *lcx_stream(lcx_engine *e);                                /* cudaStream_t of the engine                 */

/* ---- Eulerian cell fields (consecutive, z fastest; Courant fields staggered + x-halo) -------------- */
int  lcx_field_size(lcx_engine *e, int field, int64_t *count);  /* init_sync.ipp:13-52                       */
int  lcx_cells_set(lcx_engine *e, int field, const void *src, int64_t count, int src_on_device);  /* impl_sync.ipp:15-40 */
int  lcx_cells_get(lcx_engine *e, int field, void *dst, int64_t count);                            /* impl_sync.ipp:42-68 */
/* contiguous pieces of a field, queued on the engine's stream without waiting (the caller ends a batch with lcx_sync);  */
/* host memory may be pageable or pinned - pinned memory (lcx_host_alloc or the caller's own) moves at full PCIe rate      */
int  lcx_cells_set_part(lcx_engine *e, int field, int64_t offset, const void *src, int64_t count);
int  lcx_cells_get_part(lcx_engine *e, int field, int64_t offset, void *dst, int64_t count);
/* Both take HOST or DEVICE memory on the caller's side (unified addressing tells them apart): a host model whose Eulerian   */
/* fields already live on the GPU passes device pointers through arrinfo_t and no PCIe traffic happens.                       */
/* Chunked step_sync: while a cell window [c_begin, c_end) is set, lcx_hskpng_Tpr and lcx_cond (run-per-warp kernel only) work on */
/* those cells alone, and lcx_cells_get_part read-backs do not hold up the next chunk's kernels.  The host layer uploads chunk    */
/* k + 1 and reads chunk k back while chunk k + 1 computes: the copies of step_sync hide behind the condensation kernel.          */
/* Windows must start at multiples of lcx_cond_granule() cells (0: the kernel in use cannot be windowed), arrive in order, cover   */
/* the grid, and be cleared with (0, 0).  Results are bit-identical to the un-chunked step.  Consecutive windows run on two         */
/* alternating streams (they touch disjoint cells and SDs); clearing the window joins them.  No reference counterpart: the        */
/* reference's step_sync is upload -> step_cond -> read-back in sequence (src/particles_step.ipp:32-336).                          */
int  lcx_set_cell_window(lcx_engine *e, int64_t c_begin, int64_t c_end);
int  lcx_cond_granule(lcx_engine *e, int64_t *cells);
int  lcx_pointer_on_device(const void *p, int *on_device);      /* 1: device (or managed) memory, 0: host memory           */
int  lcx_copy_to_host(void *dst_host, const void *src_any, size_t bytes);   /* blocking; used by init() for device-resident fields */
int  lcx_host_alloc(size_t bytes, void **out);                  /* page-locked staging memory for the host layer          */
int  lcx_host_free(void *p);
int  lcx_set_vt0_table(lcx_engine *e, const void *table, int n);            /* init_vterm.ipp:36-59          */
int  lcx_set_efficiencies(lcx_engine *e, const void *table, int64_t n);     /* init_kernel.ipp:60-145        */

/* ---- super-droplets -------------------------------------------------------------------------------- */
/* append `count` SDs (host arrays; absent dimensions may be NULL); vt := invalid; sid continues      */
int  lcx_sd_append(lcx_engine *e, int64_t count, const uint64_t *n, const void *rd3, const void *rw2,
                   const void *kpa, const void *x, const void *y, const void *z, const uint32_t *ijk);
/* Device-side creation of the `sd_conc` flavour (per_cell super-droplets in every cell; reference: init_dry_sd_conc.ipp:25-66,    */
/* init_wet.ipp:18-74, init_xyz.ipp:16-73, init_ijk.ipp:36-52) from the counter-based random stream (key = seed, stream; counter  */
/* = SD index, call): dry radii stratified in ln(rd) over [log_rd_min, log_rd_max], equilibrium wet radii at min(RH, RH_max) of    */
/* the cell (lcx_hskpng_Tpr must have run), positions uniform in the cell.  Multiplicities are left 0: the dry radii cubed come    */
/* back in rd3_host (n_cell * per_cell reals, host memory), the caller evaluates its spectrum and sends n with lcx_sd_set_n       */
/* (physical order == order of creation until the first lcx_post_copy).                                                          */
int  lcx_sd_append_sd_conc(lcx_engine *e, int64_t per_cell, double log_rd_min, double log_rd_max, double kappa, double RH_max,
                           uint64_t seed, uint32_t stream, uint64_t call, void *rd3_host);
int  lcx_sd_set_n(lcx_engine *e, int64_t first, int64_t count, const uint64_t *n);
int  lcx_n_part(lcx_engine *e, int64_t *n_part);
/* always != 0 (default): storage indices are re-numbered after every removal, as injected random streams and          */
/* lcx_get_attr need; 0: the re-numbering is postponed until something needs it (enough for the Philox stream)         */
int  lcx_set_dense_storage_index(lcx_engine *e, int always);
/* copy one attribute to the host in reference storage order (sid); n is converted to real           */
int  lcx_get_attr(lcx_engine *e, int attr, void *dst, int64_t cap, int64_t *n_out);   /* fill_outbuf.ipp:40-79 */
int  lcx_get_attr_u64(lcx_engine *e, int attr, uint64_t *dst, int64_t cap, int64_t *n_out);   /* exact for n >= 2^53 */
/* storage index and cell of every SD in PHYSICAL order (introspection for tests: which in-cell slot an SD occupies) */
int  lcx_get_layout(lcx_engine *e, uint32_t *sid, uint32_t *ijk, int64_t cap, int64_t *n_out);

/* ---- housekeeping passes ------------------------------------------------------------------------------ */
int  lcx_hskpng_Tpr(lcx_engine *e);                             /* hskpng_Tpr.ipp:219-305                    */
int  lcx_hskpng_mfp(lcx_engine *e);                             /* hskpng_mfp.ipp:42-52                      */
int  lcx_hskpng_vterm(lcx_engine *e, int only_invalid);         /* hskpng_vterm.ipp:185-342                  */
int  lcx_sstp_percell_step(lcx_engine *e, int step, int sstp_cond, int var_rho);   /* sstp_percell_step.ipp:7-47 */
int  lcx_sstp_save(lcx_engine *e);                              /* sstp_save.ipp:7-29                        */

/* ---- condensation (per-cell sub-stepping path) ---------------------------------------------------------- */
/* one sub-step: 3rd wet moment before (step 0; later sub-steps reuse the previous "after"), implicit-Euler growth */
/* of every liquid SD, 3rd wet moment after, and the vapour / heat feedback rv -= drv, th -= drv dth/drv              */
/* root search used by every condensation entry point below, process-wide: 1 = the reference's TOMS 748 with identical trial */
/* points (default; toms748.hpp:291-454), 2 = the same with the growth law transcribed operation by operation                */
/* (cond_common.ipp:79-174), 0 = opt-in fast mode: safeguarded secant that stops when the root is known to the reference's    */
/* tolerance (half the evaluations; results within 2^-15 per step of the reference but on a different trajectory - the rows  */
/* of the reference's fixture that count threshold crossings are not reproduced in this mode)                               */
int  lcx_set_cond_solver(int mode);
int  lcx_get_cond_solver(void);
/* work distribution of the fused per-cell condensation kernel, process-wide; results are the same up to the order in   */
/* which the droplets of a cell are summed: 0 = automatic (default), -1 = eight lanes per cell, k in 1..16 = a warp per   */
/* run of k consecutive cells with its lanes balanced over the run's super-droplets                                       */
int  lcx_set_cond_layout(int cells_per_warp);
int  lcx_get_cond_layout(void);
/* the run-per-warp kernel in its phase-grouped form (opt-in, default 0; measured slower on B200: DESIGN.md section 8): droplets that need a 4th / 5th growth-law evaluation are parked   */
/* in shared memory and processed a full warp at a time; results are bit-identical to the plain form (0), which stays for A/B runs */
/* Order in which the run-per-warp kernel walks a run's droplets: -1 automatic (default; $LCX_COND_CLASSED): class by class -    */
/* drizzle / rain drops (rw > 40 um) apart from the rest - once the previous step counted more than one large drop in 64, else  */
/* in storage order; 0 never; 1 always.  Wet radii are bit-identical either way, a cell's sums are taken in another order.      */
int  lcx_set_cond_classed(int mode);
int  lcx_get_cond_classed(void);
int  lcx_set_cond_staged(int on);
int  lcx_get_cond_staged(void);
/* per-particle condensation sub-stepping, all sub-steps of one time step (particles_step.ipp:199-236,                 */
/* condensation/perparticle/ *.ipp); mix != 0: the vapour / heat exchanged by the SDs of a cell is shared after each sub-step */
int  lcx_cond_perparticle(lcx_engine *e, double dt, double RH_max, int sstp_cond, int mix);
/* the same with the number of sub-steps chosen per SD (perparticle_nomixing_adaptive_sstp_cond.ipp:8-335); no mixing      */
int  lcx_cond_perparticle_adaptive(lcx_engine *e, double dt, double RH_max, int sstp_cond_max, int sstp_cond_act,
                                   double drw2_eps, double drw2_max);
int  lcx_hskpng_rc2(lcx_engine *e);                             /* hskpng_rc2.ipp:13-32: critical radii flagged invalid */
int  lcx_cond(lcx_engine *e, double dt_sub, double RH_max, int step, int sstp_cond);   /* percell/particles_impl_cond.ipp:13-139, common/particles_impl_update_th_rv.ipp:74-191 */

/* ---- coalescence --------------------------------------------------------------------------------------- */
/* per-cell random pairing + SDM Monte-Carlo collisions for one sub-step                              */
int  lcx_coal(lcx_engine *e, double dt_sub, const lcx_rng *rng);          /* coalescence/particles_impl_coal.ipp:273-546 */
int  lcx_coal_flag(lcx_engine *e, int *increase_sstp_coal);               /* particles_step.ipp:396-400      */
int  lcx_coal_stats(lcx_engine *e, uint64_t *n_collisions, uint64_t *n_pairs_collided);  /* since creation   */

/* ---- transport: advection + sedimentation + subsidence + boundary conditions ---------------------------- */
int  lcx_transport(lcx_engine *e, const lcx_transport_opts *o);           /* adve.ipp:98-304, sedi.ipp:13-24, subs.ipp:13-25, bcnd.ipp:114-368 */
int  lcx_puddle(lcx_engine *e, double out[14]);                           /* accumulated precipitation       */
/* what left through the LID (z >= z1), which the reference removes without accounting (bcnd.ipp:330-336): accumulated dry   */
/* volume (same convention as the puddle) and number of super-droplets; lets callers close the dry-volume budget            */
int  lcx_top_loss(lcx_engine *e, double out[2]);

/* ---- x-slab migration (distributed memory) ------------------------------------------------------------- */
/* Every engine owns one inbox per side (0: left-movers arriving from the RIGHT neighbour, 1: right-movers arriving */
/* from the LEFT neighbour): a single device allocation [2 headers | n x 2 | reals x 2] - two parities, so a sender may  */
/* already deliver step s+1 while step s is being unpacked.  Senders write their packed migrants STRAIGHT into the        */
/* neighbour's inbox (peer memory over NVLink: no staging buffer, no copy engine, no host in the data path) and then      */
/* publish {count, sequence number} in its header.                                                                        */
/*   lcx_migr_connect      neighbour engine in the same process (enables peer access when on another device)              */
/*   lcx_migr_ipc_export / lcx_migr_ipc_connect   neighbour in another process: 64-byte CUDA IPC handle + capacity          */
/*   lcx_migr_put          sorts the leavers found by lcx_transport by storage index (bcnd.ipp:160-172), packs them with x   */
/*                         shifted into the neighbour's coordinates (pack.ipp:15-121), zeroes their local multiplicity       */
/*                         (unpack.ipp:122-145) and publishes the headers; one host read-back (the two counts)               */
/*   lcx_migr_take         waits for both neighbours' deliveries - same process: their stream event (pass the engines; the   */
/*                         caller guarantees their lcx_migr_put has RETURNED, e.g. by a thread barrier); other process: pass  */
/*                         NULL and the sequence number is awaited on the device - then appends the arrivals: right            */
/*                         neighbour's first, then the left's (unpack.ipp:50-120, step_async_and_copy.ipp:100-190)              */
enum { LCX_IPC_BLOB_BYTES = 96 };
int  lcx_migr_connect(lcx_engine *e, int side, lcx_engine *neighbour);
int  lcx_migr_ipc_export(lcx_engine *e, int side, void *blob);           /* inbox `side` of e                */
int  lcx_migr_ipc_connect(lcx_engine *e, int side, const void *blob);    /* e's movers of `side` go there    */
int  lcx_migr_put(lcx_engine *e, int64_t *n_lft, int64_t *n_rgt);
int  lcx_migr_take(lcx_engine *e, lcx_engine *rgt_neighbour, lcx_engine *lft_neighbour, int64_t *n_from_rgt, int64_t *n_from_lft);
int  lcx_migr_real_attrs(lcx_engine *e, int *count);                      /* number of real attributes sent  */
/* Courant halo of predictor-corrector advection between process-distributed slabs (replaces xchng_courants,                   */
/* src/impl/distributed_memory/particles_impl_xchng_courants.ipp:15-153: MPI_Isend / MPI_Recv of halo_size x-planes of Cx, Cy, */
/* Cz per side): lcx_halo_put copies this slab's outermost interior planes straight into the neighbours' inboxes (peer memory)  */
/* and publishes them; lcx_halo_take waits on the device for both neighbours' planes and copies them into this slab's halo.     */
/* Call both, in this order, on every rank, after the Courant fields of the step were set and before lcx_transport.            */
int  lcx_halo_put(lcx_engine *e);
int  lcx_halo_take(lcx_engine *e);

/* ---- end of step: removal / recycling, cell index, per-cell grouping ------------------------------------ */
/* keep_all != 0: initial grouping - nothing is removed and the cell indices given to lcx_sd_append are used */
int  lcx_post_copy(lcx_engine *e, int rcyc, int keep_all);                              /* post_copy.ipp:18-35, hskpng_remove.ipp:20-75, rcyc.ipp:44-139, hskpng_ijk.ipp:159-200, hskpng_sort.ipp:15-70, hskpng_count.ipp:16-48 */

/* ---- diagnostics ----------------------------------------------------------------------------------------- */
int  lcx_moms_select(lcx_engine *e, int kind, int attr, double lo, double hi, int cons);
int  lcx_moms_calc(lcx_engine *e, int attr, double power, int specific);  /* moms.ipp:277-387                */
int  lcx_diag_sd_conc(lcx_engine *e);                                     /* particles_diag.ipp:193-211      */
int  lcx_diag_cell_field(lcx_engine *e, int field);                       /* particles_diag.ipp:148-190      */
int  lcx_diag_precip_rate(lcx_engine *e);                                 /* particles_diag.ipp:561-586      */
int  lcx_diag_mass_dens(lcx_engine *e, int attr, double rad, double sig0, double xp);   /* particles_impl_mass_dens.ipp:13-98 */
int  lcx_diag_vel_div(lcx_engine *e, double dt);                          /* particles_diag.ipp:497-558      */
int  lcx_diag_max_rw(lcx_engine *e);                                      /* particles_diag.ipp:606-634      */
int  lcx_outbuf(lcx_engine *e, void *dst, int64_t count);                 /* fill_outbuf.ipp:13-37           */

/* ---- timing / introspection (used by bench.py) ----------------------------------------------------------- */
/* device timer on the engine's stream (CUDA events): start, then stop returns the elapsed milliseconds           */
int  lcx_timer_start(lcx_engine *e);
int  lcx_timer_stop(lcx_engine *e, float *ms);
/* per-kernel profile: while enabled every launch is bracketed by CUDA events; the report is a text table        */
/* "kernel launches total_ms" (one line per kernel), written into buf (truncated to size)                        */
int  lcx_profile_enable(lcx_engine *e, int on);
int  lcx_profile_report(lcx_engine *e, char *buf, int64_t size);
int  lcx_launch_count(lcx_engine *e, uint64_t *launches);                 /* kernels launched since creation */
int  lcx_cell_stats(lcx_engine *e, int64_t *n_cell, int64_t *max_count);

#ifdef __cplusplus
}
#endif
#endif
