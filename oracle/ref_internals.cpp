// Oracle-only helper: reads private state of the *reference's* particles_t<double, serial|OpenMP>
// (multiplicities n, cell indices, sort permutation, per-cell counts, terminal velocities) so that the
// integer parts of the hot path can be compared element by element.  The reference exposes none of
// these through its public API (reference src/impl/diagnose_SD_attributes/
// particles_impl_fill_outbuf.ipp:42 - "TODO implement n").
//
// Built only by oracle/build_ref.py, from the reference sources where they lie; this file contains
// no reference code - it includes the reference's own translation-unit body (src/particles.tpp) the
// same way reference src/lib_cpp.cpp:1-9 / src/lib_omp.cpp:1-9 do.
// Test infrastructure: never linked into the product.
#include "lib.hpp"

// -DLGC_INTERNALS_F32: the same readers for the reference's single-precision instantiation (src/lib.cpp:43)
#if defined(LGC_INTERNALS_F32)
   typedef float lgc_ref_real;
#  define LGC_SFX(name, be) name##_##be##_f32
#else
   typedef double lgc_ref_real;
#  define LGC_SFX(name, be) name##_##be
#endif
#if defined(LGC_INTERNALS_OMP)
#  include <thrust/system/omp/execution_policy.h>
#  include <thrust/system/omp/vector.h>
   namespace thrust_device = ::thrust::omp;
#  define LGC_BACKEND OpenMP
#  define LGC_FN(name) LGC_SFX(name, omp)
#else
#  include <thrust/system/cpp/vector.h>
   namespace thrust_device = ::thrust::cpp;
#  define LGC_BACKEND serial
#  define LGC_FN(name) LGC_SFX(name, serial)
#endif

#include "particles.tpp"

#include <cstring>
#include <string>

namespace
{
  using namespace libcloudphxx::lgrngn;
  typedef particles_t<lgc_ref_real, LGC_BACKEND> prt_t;

  template <class vec_t, class out_t>
  long copy_out(const vec_t &v, std::size_t n, out_t *dst, long cap)
  {
    for (std::size_t i = 0; i < n && long(i) < cap; ++i) dst[i] = out_t(v[i]);
    return long(n);
  }
}

// returns the number of elements of the named array (copies at most cap of them), -1 if unknown
extern "C" long LGC_FN(lgc_ref_dump_u64)(void *proto, const char *name, unsigned long long *dst, long cap)
{
  prt_t *p = static_cast<prt_t *>(static_cast<particles_proto_t<lgc_ref_real> *>(proto));
  auto &s = *p->pimpl;
  const std::string nm(name);
  if (nm == "n")          return copy_out(s.n, s.n_part, dst, cap);
  if (nm == "ijk")        return copy_out(s.ijk, s.n_part, dst, cap);
  if (nm == "sorted_id")  return copy_out(s.sorted_id, s.n_part, dst, cap);
  if (nm == "sorted_ijk") return copy_out(s.sorted_ijk, s.n_part, dst, cap);
  if (nm == "count_ijk")  return copy_out(s.count_ijk, s.count_n, dst, cap);
  if (nm == "count_num")  return copy_out(s.count_num, s.count_n, dst, cap);
  if (nm == "n_part")     { if (cap > 0) dst[0] = s.n_part; return 1; }
  if (nm == "sorted")     { if (cap > 0) dst[0] = s.sorted; return 1; }
  if (nm == "sstp_coal")  { if (cap > 0) dst[0] = s.sstp_coal; return 1; }
  return -1;
}

extern "C" long LGC_FN(lgc_ref_dump_f64)(void *proto, const char *name, double *dst, long cap)
{
  prt_t *p = static_cast<prt_t *>(static_cast<particles_proto_t<lgc_ref_real> *>(proto));
  auto &s = *p->pimpl;
  const std::string nm(name);
  if (nm == "vt")   return copy_out(s.vt, s.n_part, dst, cap);
  if (nm == "rw2")  return copy_out(s.rw2, s.n_part, dst, cap);
  if (nm == "rd3")  return copy_out(s.rd3, s.n_part, dst, cap);
  if (nm == "kpa")  return copy_out(s.kpa, s.n_part, dst, cap);
  if (nm == "x")    return copy_out(s.x, s.x.size() ? s.n_part : 0, dst, cap);
  if (nm == "y")    return copy_out(s.y, s.y.size() ? s.n_part : 0, dst, cap);
  if (nm == "z")    return copy_out(s.z, s.z.size() ? s.n_part : 0, dst, cap);
  if (nm == "T")    return copy_out(s.T, s.T.size(), dst, cap);
  if (nm == "p")    return copy_out(s.p, s.p.size(), dst, cap);
  if (nm == "RH")   return copy_out(s.RH, s.RH.size(), dst, cap);
  if (nm == "eta")  return copy_out(s.eta, s.eta.size(), dst, cap);
  if (nm == "th")   return copy_out(s.th, s.th.size(), dst, cap);
  if (nm == "rv")   return copy_out(s.rv, s.rv.size(), dst, cap);
  if (nm == "rhod") return copy_out(s.rhod, s.rhod.size(), dst, cap);
  if (nm == "dv")   return copy_out(s.dv, s.dv.size(), dst, cap);
  return -1;
}
