/* Flat C view of libcloudphxx::lgrngn::particles_proto_t<real>, real = double (symbols lgc_*) or, compiled with
 * -DLGC_FLOAT, float (symbols lgcf_*; the reference instantiates both: src/lib.cpp:43-44).  Arrays are `lgc_real`,
 * scalar options stay double in the structs and are converted on the way in.
 *
 * Role: what the reference's Boost.Python module (reference bindings/python/lgrngn.hpp:41-149,
 * bindings/python/lib.cpp:217-434) does for Python callers, done here as a plain C surface that
 * ctypes can load.  The implementation (lgrngn_capi.cpp) touches nothing but the public lgrngn API
 * (factory / particles_proto_t / opts_init_t / opts_t / arrinfo_t), so the very same source is
 * compiled twice: against this library's headers (the B200 back-end) and, for the parity oracle,
 * against the reference's own headers and CPU back-ends (oracle/_ref).
 */
#ifndef LGRNGN_CAPI_H
#define LGRNGN_CAPI_H
#ifdef __cplusplus
extern "C" {
#endif

#ifdef LGC_FLOAT
typedef float lgc_real;
#define lgc_create lgcf_create
#define lgc_destroy lgcf_destroy
#define lgc_diag lgcf_diag
#define lgc_get_attr lgcf_get_attr
#define lgc_get_n lgcf_get_n
#define lgc_impl_name lgcf_impl_name
#define lgc_init lgcf_init
#define lgc_last_error lgcf_last_error
#define lgc_n_cell lgcf_n_cell
#define lgc_opts_defaults lgcf_opts_defaults
#define lgc_opts_init_defaults lgcf_opts_init_defaults
#define lgc_outbuf lgcf_outbuf
#define lgc_proto lgcf_proto
#define lgc_puddle lgcf_puddle
#define lgc_step_async lgcf_step_async
#define lgc_step_cond lgcf_step_cond
#define lgc_step_sync lgcf_step_sync
#define lgc_sync_in lgcf_sync_in
#else
typedef double lgc_real;
#endif

typedef struct lgc_handle lgc_handle;

enum { LGC_MAX_MODES = 4, LGC_MAX_DISTROS = 4, LGC_MAX_KPARAMS = 4, LGC_MAX_SIZES = 16 };

/* dry spectrum n(ln r) [m^-3 per unit ln r, at STP] */
typedef struct
{
  int    kind;                     /* 0: sum of lognormal modes; 1: exponential in volume; 2: caller's function */
  double kappa, rd_insol;
  int    n_modes;
  double mean_r[LGC_MAX_MODES], stdev[LGC_MAX_MODES], n_tot[LGC_MAX_MODES];
  double r0, n0;                   /* kind 1: n0 * 3 (r/r0)^3 exp(-(r/r0)^3)                       */
  double (*fn)(double lnr, void *ctx);   /* kind 2: n(ln r), called from the host-side initialisation (what the  */
  void  *ctx;                            /*         reference's Python binding does with a Python callable)      */
} lgc_distro;

/* one entry of opts_init_t::dry_sizes: (kappa, rd_insol) -> radius -> (STP concentration, SDs per cell) */
typedef struct { double kappa, rd_insol, radius, conc; int count; } lgc_dry_size;

typedef struct
{
  int backend;                     /* backend_t: 1 serial, 2 OpenMP, 3 CUDA, 4 multi_CUDA          */
  int nx, ny, nz;
  double dx, dy, dz, dt;
  int sstp_cond, sstp_coal;
  double x0, y0, z0, x1, y1, z1;
  unsigned long long sd_conc, sd_const_multi, n_sd_max;
  int kernel, terminal_velocity, adve_scheme, RH_formula;     /* enum ordinals                      */
  int n_kernel_parameters;
  double kernel_parameters[LGC_MAX_KPARAMS];
  int coal_switch, sedi_switch, subs_switch, exact_sstp_cond;
  int turb_adve_switch, turb_cond_switch, turb_coal_switch, ice_switch, chem_switch;
  double RH_max;
  int rng_seed, rng_seed_init, rng_seed_init_switch;
  int dev_count, dev_id;
  double rd_min, rd_max;
  int open_side_walls, periodic_topbot_walls, variable_dt_switch, th_dry, const_p;
  int aerosol_independent_of_rhod;
  int n_distros;
  lgc_distro distros[LGC_MAX_DISTROS];
  int n_w_LS;
  const double *w_LS;
  int sd_conc_large_tail, no_ccn_at_init;
  int n_dry_sizes;
  lgc_dry_size dry_sizes[LGC_MAX_SIZES];
  int n_aerosol_conc_factor;
  const double *aerosol_conc_factor;
  int sstp_cond_mix;               /* per-particle sub-stepping: share vapour / heat inside a cell after each sub-step */
  int adaptive_sstp_cond, sstp_cond_act;
  double sstp_cond_adapt_drw2_eps, sstp_cond_adapt_drw2_max, rc2_T;
} lgc_opts_init;

typedef struct
{
  int adve, sedi, subs, cond, coal, rcyc;
  double RH_max, dt;
} lgc_opts;

/* data == NULL means "not supplied" (null arrinfo_t); strides are in elements */
typedef struct { lgc_real *data; long strides[3]; } lgc_arr;

enum lgc_diag_t
{
  LGC_DIAG_ALL = 0, LGC_DIAG_RW_GE_RC, LGC_DIAG_RH_GE_SC,
  LGC_DIAG_DRY_RNG, LGC_DIAG_WET_RNG, LGC_DIAG_KAPPA_RNG,
  LGC_DIAG_DRY_RNG_CONS, LGC_DIAG_WET_RNG_CONS, LGC_DIAG_KAPPA_RNG_CONS,
  LGC_DIAG_WATER, LGC_DIAG_WATER_CONS,
  LGC_DIAG_SD_CONC, LGC_DIAG_PRESSURE, LGC_DIAG_TEMPERATURE, LGC_DIAG_RH,
  LGC_DIAG_DRY_MOM, LGC_DIAG_WET_MOM, LGC_DIAG_KAPPA_MOM,
  LGC_DIAG_PRECIP_RATE, LGC_DIAG_MAX_RW, LGC_DIAG_VEL_DIV, LGC_DIAG_WET_MASS_DENS
};

void        lgc_opts_init_defaults(lgc_opts_init *);
void        lgc_opts_defaults(lgc_opts *);
const char *lgc_last_error(void);
const char *lgc_impl_name(void);          /* "b200" or "reference" */

int  lgc_create(const lgc_opts_init *, lgc_handle **out);
void lgc_destroy(lgc_handle *);
int  lgc_init(lgc_handle *, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod, const lgc_arr *p,
              const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz);
int  lgc_step_sync(lgc_handle *, const lgc_opts *, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod,
                   const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz);
int  lgc_sync_in(lgc_handle *, const lgc_arr *th, const lgc_arr *rv, const lgc_arr *rhod,
                 const lgc_arr *cx, const lgc_arr *cy, const lgc_arr *cz);
int  lgc_step_cond(lgc_handle *, const lgc_opts *, const lgc_arr *th, const lgc_arr *rv);
int  lgc_step_async(lgc_handle *, const lgc_opts *);
int  lgc_diag(lgc_handle *, int what, double a, double b);
long lgc_n_cell(lgc_handle *);
int  lgc_outbuf(lgc_handle *, lgc_real *dst, long n);
/* copies a per-SD attribute (storage order); *n_out = number of SDs, even if cap is too small */
int  lgc_get_attr(lgc_handle *, const char *name, lgc_real *dst, long cap, long *n_out);
int  lgc_get_n(lgc_handle *, unsigned long long *dst, long cap, long *n_out);
int  lgc_puddle(lgc_handle *, double *out14);
void *lgc_proto(lgc_handle *);            /* the underlying particles_proto_t<real>* (for back-end specific extras) */

#ifdef __cplusplus
}
#endif
#endif
