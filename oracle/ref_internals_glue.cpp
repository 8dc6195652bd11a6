// Oracle-only glue: routes the flat binding's internal-state queries to the per-back-end readers in
// ref_internals.cpp (compiled once for serial, once for OpenMP).  Test infrastructure.
#include <cstring>
extern "C" {
long lgc_ref_dump_u64_serial(void *, const char *, unsigned long long *, long);
long lgc_ref_dump_f64_serial(void *, const char *, double *, long);
long lgc_ref_dump_u64_omp(void *, const char *, unsigned long long *, long);
long lgc_ref_dump_f64_omp(void *, const char *, double *, long);

long lgc_ref_dump_u64_serial_f32(void *, const char *, unsigned long long *, long);
long lgc_ref_dump_f64_serial_f32(void *, const char *, double *, long);
long lgc_ref_dump_u64_omp_f32(void *, const char *, unsigned long long *, long);
long lgc_ref_dump_f64_omp_f32(void *, const char *, double *, long);

// single-precision particle systems (the lgcf_* binding)
long lgc_ref_internal_get_n_f32(void *proto, int backend, unsigned long long *dst, long cap)
{
  if (backend == 1) return lgc_ref_dump_u64_serial_f32(proto, "n", dst, cap);
  if (backend == 2) return lgc_ref_dump_u64_omp_f32(proto, "n", dst, cap);
  return -1;
}
long lgc_ref_dump_u64_f32(void *handle, const char *name, unsigned long long *dst, long cap);
long lgc_ref_dump_f64_f32(void *handle, const char *name, double *dst, long cap);

long lgc_ref_internal_get_n(void *proto, int backend, unsigned long long *dst, long cap)
{
  if (backend == 1) return lgc_ref_dump_u64_serial(proto, "n", dst, cap);
  if (backend == 2) return lgc_ref_dump_u64_omp(proto, "n", dst, cap);
  return -1;
}

// lgc_handle layout is { unique_ptr<proto> p; int backend; long n_cell; } - see lgrngn_capi.cpp
struct lgc_handle_view { void *p; int backend; long n_cell; };

long lgc_ref_dump_u64(void *handle, const char *name, unsigned long long *dst, long cap)
{
  lgc_handle_view *h = static_cast<lgc_handle_view *>(handle);
  return h->backend == 2 ? lgc_ref_dump_u64_omp(h->p, name, dst, cap) : lgc_ref_dump_u64_serial(h->p, name, dst, cap);
}
long lgc_ref_dump_f64(void *handle, const char *name, double *dst, long cap)
{
  lgc_handle_view *h = static_cast<lgc_handle_view *>(handle);
  return h->backend == 2 ? lgc_ref_dump_f64_omp(h->p, name, dst, cap) : lgc_ref_dump_f64_serial(h->p, name, dst, cap);
}
long lgc_ref_dump_u64_f32(void *handle, const char *name, unsigned long long *dst, long cap)
{
  lgc_handle_view *h = static_cast<lgc_handle_view *>(handle);
  return h->backend == 2 ? lgc_ref_dump_u64_omp_f32(h->p, name, dst, cap) : lgc_ref_dump_u64_serial_f32(h->p, name, dst, cap);
}
long lgc_ref_dump_f64_f32(void *handle, const char *name, double *dst, long cap)      // values widened to double
{
  lgc_handle_view *h = static_cast<lgc_handle_view *>(handle);
  return h->backend == 2 ? lgc_ref_dump_f64_omp_f32(h->p, name, dst, cap) : lgc_ref_dump_f64_serial_f32(h->p, name, dst, cap);
}
}
