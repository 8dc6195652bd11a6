// stand-in for <boost/math/constants/constants.hpp> (oracle build only)
#pragma once
namespace boost { namespace math { namespace constants {
  // same correctly-rounded literal Boost ships for double/float
  template <class T> constexpr T pi() { return T(3.141592653589793238462643383279502884e+00L); }
}}}
