// stand-in for <boost/math/tools/minima.hpp>: Brent's one-dimensional minimiser
// (Brent 1973, "Algorithms for Minimization without Derivatives", ch. 5), used by
// the reference only in the const-multiplicity auto-range initialisation.
#pragma once
#include <cmath>
#include <utility>
#include <limits>
#include <cstdint>
namespace boost { namespace math { namespace tools {
  template <class F, class T>
  std::pair<T, T> brent_find_minima(F f, T lo, T hi, int bits, std::uintmax_t &max_iter)
  {
    const int digits = std::numeric_limits<T>::digits;
    if (bits > digits / 2) bits = digits / 2;
    const T tol = std::ldexp(T(1), 1 - bits);
    const T golden = T(0.3819660f);   // single-precision literal, as in Boost
    T x = hi, w = hi, v = hi, fx = f(x), fw = fx, fv = fx, step = 0, prev_step = 0;
    std::uintmax_t left = max_iter;
    while (left)
    {
      const T mid = (lo + hi) / 2;
      const T t1 = tol * std::fabs(x) + tol / 4, t2 = 2 * t1;
      if (std::fabs(x - mid) <= t2 - (hi - lo) / 2) break;
      bool use_golden = true;
      if (std::fabs(prev_step) > t1)
      {
        T r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), p = (x - v) * q - (x - w) * r;
        q = 2 * (q - r);
        if (q > 0) p = -p;
        q = std::fabs(q);
        const T old = prev_step;
        prev_step = step;
        if (!(std::fabs(p) >= std::fabs(q * old / 2) || p <= q * (lo - x) || p >= q * (hi - x)))
        {
          step = p / q;
          const T u = x + step;
          if ((u - lo) < t2 || (hi - u) < t2) step = (mid - x) < 0 ? -std::fabs(t1) : std::fabs(t1);
          use_golden = false;
        }
      }
      if (use_golden)
      {
        prev_step = (x >= mid) ? lo - x : hi - x;
        step = golden * prev_step;
      }
      const T u = (std::fabs(step) >= t1) ? x + step : (step > 0 ? x + std::fabs(t1) : x - std::fabs(t1));
      const T fu = f(u);
      if (fu <= fx)
      {
        if (u >= x) lo = x; else hi = x;
        v = w; w = x; x = u; fv = fw; fw = fx; fx = fu;
      }
      else
      {
        if (u < x) lo = u; else hi = u;
        if (fu <= fw || w == x) { v = w; w = u; fv = fw; fw = fu; }
        else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
      }
      --left;
    }
    max_iter -= left;
    return std::make_pair(x, fx);
  }
}}}
