/* Extra C entry points of liblgrngn_b200.so that have no counterpart in the reference's public API:
 *   - choice of the random stream feeding coalescence (replay of the reference's mt19937 draw order for parity
 *     runs, or Philox evaluated inside the kernels);
 *   - process-distributed runs (one process per GPU, e.g. under torchrun): rank / size / neighbour geometry are
 *     set before `factory(CUDA, opts_init)` is called with the rank-LOCAL opts_init, the way the reference's
 *     MPI mode works (reference src/particles_ctor.ipp:22-75, tests/mpi/mpi_adve_test.cpp:88-141); step_async then
 *     stops before migration and the caller drives pack -> exchange -> unpack -> post_copy through the engine ABI.
 */
#ifndef LGRNGN_B200_EXTRAS_H
#define LGRNGN_B200_EXTRAS_H
#ifdef __cplusplus
extern "C" {
#endif

enum { LGRNGN_B200_RNG_PHILOX = 0, LGRNGN_B200_RNG_MT19937 = 1 };

typedef struct
{
  int rank, size;          /* size <= 1: single process                                     */
  double lft_x1, rgt_x0;   /* x1 of the left neighbour's slab, x0 of the right neighbour's   */
  int n_x_tot;             /* total number of x cells over all ranks (0: unknown)            */
} lgrngn_b200_distmem;

void  lgrngn_b200_set_rng_mode(int mode);       /* applies to particle systems created afterwards (default: $LCX_RNG or Philox) */
int   lgrngn_b200_get_rng_mode(void);
void  lgrngn_b200_set_distmem(const lgrngn_b200_distmem *d);
void *lgrngn_b200_engine(void *particles_proto); /* lcx_engine* of the first (or only) slab */
int   lgrngn_b200_n_slabs(void *particles_proto);
void *lgrngn_b200_engine_of_slab(void *particles_proto, int slab);
int   lgrngn_b200_post_copy(void *particles_proto, int rcyc);
/* Process-distributed runs: after init() every rank exports two 96-byte blobs (CUDA IPC handles of its two migration     */
/* inboxes: [0] filled by its RIGHT neighbour, [1] by its LEFT neighbour), the caller moves them to the neighbours with    */
/* whatever transport it has (MPI, torch.distributed, a file), and each rank connects with the blob pairs of its left and   */
/* right neighbours.  From then on step_async() itself migrates the super-droplets: packed straight into the neighbour's    */
/* inbox over NVLink, ordered by sequence numbers in device memory; no host communication in the step.                      */
int   lgrngn_b200_distmem_export(void *particles_proto, void *blobs /* 2 x 96 bytes */);
int   lgrngn_b200_distmem_connect(void *particles_proto, const void *lft_neighbour_blobs, const void *rgt_neighbour_blobs);
/* migrants of the last step, summed over the slabs: sent left, sent right, received from the right, received from the left */
int   lgrngn_b200_migr_stats(void *particles_proto, long long *out4);
/* test hooks: storage indices kept dense (1) / re-numbered lazily (0) / chosen by the random-stream mode (-1, default);     */
/* random streams for the next coalescence sub-step (un by storage index, u01 by sorted position; one call per sub-step);   */
/* number of Philox calls made so far; storage index and cell of every super-droplet in physical order                       */
void  lgrngn_b200_set_dense_sid(int mode);
int   lgrngn_b200_inject_rng(void *particles_proto, const unsigned int *un, const double *u01, long long n);
long long lgrngn_b200_philox_call(void *particles_proto);
int   lgrngn_b200_get_layout(void *particles_proto, unsigned int *sid, unsigned int *ijk, long long cap, long long *n_out);
/* one full model step (condensation + th/rv feedback, then coalescence / transport / housekeeping) on the Eulerian  */
/* fields ALREADY RESIDENT in device memory: no host<->device field traffic.  flags: bit0 adve, bit1 sedi, bit2 cond, */
/* bit3 coal.  Used by bench.py for the device-resident throughput figure.                                            */
int   lgrngn_b200_step_resident(void *particles_proto, int flags);
/* Single-precision particle systems (factory<float>: the engine liblcx_b200_f32.so, entry points lcx_*_f32): the same handles */
int   lgrngn_b200_n_slabs_f32(void *particles_proto_float);
void *lgrngn_b200_engine_of_slab_f32(void *particles_proto_float, int slab);
int   lgrngn_b200_step_resident_f32(void *particles_proto_float, int flags);
/* multiplicities as 64-bit integers in storage order, either precision (real_bytes = 8 or 4 says which proto it is)          */
int   lgrngn_b200_get_n(void *particles_proto, int real_bytes, unsigned long long *dst, long long cap, long long *n_out);
/* Binary-compatibility guard: the layout of opts_init_t / opts_t / arrinfo_t and the v-table slots of particles_proto_t as THIS  */
/* library was compiled (text, see host/include/lgrngn_abi_probe.hpp); returns the length.  A host model built with other headers */
/* (the reference's) compiles the same probe against them and compares the two strings before calling factory().                */
long  lgrngn_b200_abi_layout(char *buf, long size);
/* scalar thermodynamic helpers and constants of libcloudphxx::common as the reference's Python module exposes them  */
/* (bindings/python/common.hpp:20-170, lib.cpp:40-120); unknown names give NaN                                       */
double lgrngn_b200_common(const char *name, double a, double b, double c, double d, double e);

#ifdef __cplusplus
}
#endif
#endif
