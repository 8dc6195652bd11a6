// stand-in for <boost/units/systems/si.hpp>: Boost.units is a zero-overhead
// compile-time dimension checker, so the reference's own unit-less substitute
// (fake_units.hpp, which it uses under nvcc) performs the identical arithmetic.
#pragma once
#include <libcloudph++/common/detail/fake_units.hpp>
namespace boost { namespace units {
  namespace si = libcloudphxx::common::detail::fake_units::si;
  using libcloudphxx::common::detail::fake_units::quantity;
  using libcloudphxx::common::detail::fake_units::divide_typeof_helper;
  using libcloudphxx::common::detail::fake_units::multiply_typeof_helper;
  using libcloudphxx::common::detail::fake_units::power_typeof_helper;
  using libcloudphxx::common::detail::fake_units::static_rational;
}}
