// Host-side functor interface used for dry-aerosol spectra n(ln r)
// (counterpart of reference include/libcloudph++/common/unary_function.hpp:11-19).
#pragma once
namespace libcloudphxx { namespace common {
  template <typename real_t>
  struct unary_function
  {
    virtual real_t funval(const real_t) const = 0;
    real_t operator()(const real_t arg) const { return funval(arg); }
    virtual ~unary_function() {}
  };
}}
