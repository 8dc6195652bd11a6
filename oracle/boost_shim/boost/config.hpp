// Minimal stand-in for <boost/config.hpp>, used ONLY to compile the reference's
// own serial/OpenMP back-ends as the parity oracle (oracle/_ref).  Boost is not
// installed in this image; the reference needs just a handful of names from it.
// Test infrastructure - never included by the product (libcloudphxx_b200/).
#pragma once
#include <sstream>
#include <functional>
#include <iostream>
#include <cmath>
#include <algorithm>
#include <stdexcept>
#include <cstdint>

#if defined(__CUDACC__)
#  define BOOST_GPU_ENABLED __host__ __device__
#else
#  define BOOST_GPU_ENABLED
#endif

namespace boost { namespace math {
  template <class T> inline bool isfinite(T x) { return std::isfinite(x); }
}}
