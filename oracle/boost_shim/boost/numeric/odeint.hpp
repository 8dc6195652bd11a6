// stand-in for <boost/numeric/odeint.hpp>: only the aqueous-chemistry code
// (out of scope, never executed by the oracle) names these types.
#pragma once
#include <stdexcept>
namespace boost { namespace numeric { namespace odeint {
  struct thrust_algebra {};
  struct thrust_operations {};
  struct never_resizer {};
  template <class S, class V, class D, class T, class A, class O, class R>
  struct runge_kutta4
  {
    template <class... Args> void adjust_size(Args&&...) {}
    template <class... Args> void do_step(Args&&...)
    { throw std::runtime_error("oracle shim: odeint (chemistry) is not available"); }
  };
}}}
