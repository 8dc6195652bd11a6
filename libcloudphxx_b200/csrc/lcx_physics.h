// Cloud-microphysics formulae of the super-droplet hot path, written once for host and device.
//
// Every function states which reference formula it evaluates (file:line under /root/reference) and
// keeps the reference's ORDER OF FLOATING-POINT OPERATIONS, because parity with the reference's CPU
// back-end is judged at the level of last-bit agreement wherever only + - * / sqrt are involved
// (compile with -fmad=false / -ffp-contract=off).  No Boost.units, no Thrust: plain templates.
//
// Used by: the CUDA kernels (csrc/*.cu) and the C++ host layer (host/particles_b200.cpp, for the
// host-side initialisation that must reproduce the reference's libm results bit for bit).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#  define LCX_HD __host__ __device__ __forceinline__
#else
#  define LCX_HD inline
#endif

namespace lcx
{
  typedef unsigned long long n_t;

  // enum ordinals are those of the public API (lgrngn_b200_api.hpp)
  enum { KERNEL_UNDEFINED = 0, KERNEL_GEOMETRIC, KERNEL_GOLOVIN, KERNEL_HALL, KERNEL_HALL_DAVIS_NO_WAALS, KERNEL_LONG,
         KERNEL_ONISHI_HALL, KERNEL_ONISHI_HALL_DAVIS_NO_WAALS, KERNEL_HALL_PINSKY_1000MB_GRAV,
         KERNEL_HALL_PINSKY_CUMULONIMBUS, KERNEL_HALL_PINSKY_STRATOCUMULUS, KERNEL_VOHL_DAVIS_NO_WAALS };
  enum { VT_UNDEFINED = 0, VT_BEARD76, VT_BEARD77, VT_BEARD77FAST, VT_KHVOROSTYANOV_SPHERICAL, VT_KHVOROSTYANOV_NONSPHERICAL };
  enum { AS_UNDEFINED = 0, AS_IMPLICIT, AS_EULER, AS_PRED_CORR };
  enum { RH_PV_CC = 0, RH_RV_CC, RH_PV_TET, RH_RV_TET };

  // Arithmetic inside the condensation root solve.  There every quotient is either the next trial abscissa or part of the
  // residual whose root is wanted to 2^-15 only, so a correctly rounded result buys nothing: a translation unit may define
  // LCX_FAST_MATH to get reciprocal-approximation + two Newton steps + one residual correction on the device (error <= 1 ulp,
  // no special-case branch), about 1/3 of the instructions of the IEEE division sequence.  Everywhere else "/" is IEEE.
  // FP64 instructions cannot carry a full 64-bit immediate: a literal such as 1.71 or 1/5040 is put together by two UMOVs
  // at every use, which made up 8 % of the executed instructions of the condensation kernel (ncu source view, profiles/).
  // Operands read from the constant bank cost no instruction, so the literals of the root solve's inner loop live in a table.
  // Same values, same operations: results are bit-identical to the literal form (the host build keeps the literals).
#if defined(LCX_FAST_MATH) && defined(__CUDACC__) && !defined(LCX_NO_KC)
  static __constant__ double lcx_kc[] = {
    1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,   // 0..8: Taylor series of exp
    1e-4, 1. / 3, 1. / 9, 5. / 81,                                                                                                  // 9..12: cbrt(1 + x) series
    1.71, 1.33                                                                                                                      // 13, 14: transition-regime correction
  };
#endif
#if defined(LCX_FAST_MATH) && defined(__CUDA_ARCH__) && !defined(LCX_NO_KC)
#  define LCX_KC(i, literal) (::lcx::lcx_kc[i])
#else
#  define LCX_KC(i, literal) (literal)
#endif
#if defined(LCX_FAST_MATH) && defined(__CUDA_ARCH__)
  __device__ __forceinline__ double lcx_div(double a, double b)
  {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
  }
  __device__ __forceinline__ float lcx_div(float a, float b) { return a / b; }
  // 1/sqrt(x), x normal and positive: hardware seed + two Newton steps (<= 1 ulp; the library routine spends ~50
  // instructions on the same thing because it also serves subnormals, zeros and infinities)
  __device__ __forceinline__ double lcx_rsqrt(double x)
  {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = fma(r, fma(-hx * r, r, 0.5), r);
    r = fma(r, fma(-hx * r, r, 0.5), r);
    return r;
  }
  __device__ __forceinline__ float lcx_rsqrt(float x) { return rsqrtf(x); }
  // exp(x) for the Kelvin term (x = A / rw, A ~ 1e-9 m).  0 <= x < 1/8 (rw > 8 nm): Taylor polynomial of degree 11, remainder
  // < 3e-20.  1/8 <= x < 1 (the smallest aerosol, a few lanes of a warp at a time: the library routine there was a divergent
  // 45-instruction detour taking 4 % of the kernel's instructions at 4 of 32 lanes): exp(x) = exp(x/8)^8, three squarings,
  // <= 7 ulp.  Anything else goes to the library.
  __device__ __forceinline__ double lcx_exp_small(double x)
  {
    if (!(x >= 0 && x < 1.0)) return exp(x);
    const bool scaled = x >= 0.125;
    const double y = scaled ? x * 0.125 : x;
    double p = LCX_KC(0, 1.0 / 39916800.0);
    p = fma(p, y, LCX_KC(1, 1.0 / 3628800.0)); p = fma(p, y, LCX_KC(2, 1.0 / 362880.0)); p = fma(p, y, LCX_KC(3, 1.0 / 40320.0));
    p = fma(p, y, LCX_KC(4, 1.0 / 5040.0)); p = fma(p, y, LCX_KC(5, 1.0 / 720.0)); p = fma(p, y, LCX_KC(6, 1.0 / 120.0));
    p = fma(p, y, LCX_KC(7, 1.0 / 24.0)); p = fma(p, y, LCX_KC(8, 1.0 / 6.0));
    p = fma(p, y, 0.5); p = fma(p, y, 1.0); p = fma(p, y, 1.0);
    if (scaled) { p *= p; p *= p; p *= p; }
    return p;
  }
  __device__ __forceinline__ float lcx_exp_small(float x) { return exp(x); }
  // cbrt(y), 1 <= y < 1e30: single-precision seed, two Newton steps with the seed's slope (error 1e-7 -> 1e-14 -> 1e-21)
  __device__ __forceinline__ double lcx_cbrt_ge1(double y)
  {
    const float c0 = cbrtf(float(y));
    const double s = double(1.0f / (3.0f * c0 * c0));
    double c = double(c0);
    c = fma(-fma(c * c, c, -y), s, c);
    c = fma(-fma(c * c, c, -y), s, c);
    return c;
  }
  __device__ __forceinline__ float lcx_cbrt_ge1(float y) { return cbrtf(y); }
  // cbrt(1 + x), 0 <= x <= 1/2 (ventilation factors of cloud and drizzle drops: x = Re Sc <= 0.5 up to r ~ 40 um): the seed is a
  // degree-5 polynomial in single precision (Chebyshev fit on [0, 1/2], error 1.8e-7 with rounding) instead of cbrtf with its
  // range reduction, then the same two fixed-slope Newton steps in double (1.8e-7 -> 5e-14 -> 1.8e-16, checked over the whole
  // interval by tools/check_fastmath.cu).  Half the instructions of lcx_cbrt_ge1, which made up 13 % of the kernel's.
  __device__ __forceinline__ double lcx_cbrt1p_mid(double x)
  {
    const float xf = float(x);
    float c0 = 0.011139679700136185f;
    c0 = fmaf(c0, xf, -0.03272373229265213f); c0 = fmaf(c0, xf, 0.059791404753923416f); c0 = fmaf(c0, xf, -0.1108996570110321f);
    c0 = fmaf(c0, xf, 0.333324670791626f); c0 = fmaf(c0, xf, 1.0f);
    float sf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sf) : "f"(3.0f * c0 * c0));
    const double s = double(sf), y = 1.0 + x;
    double c = double(c0);
    c = fma(-fma(c * c, c, -y), s, c);
    c = fma(-fma(c * c, c, -y), s, c);
    return c;
  }
  __device__ __forceinline__ float lcx_cbrt1p_mid(float x) { return cbrtf(1.0f + x); }
  __device__ __forceinline__ double lcx_pow077(double x) { return exp(0.077 * log(x)); }      // x > 1
  __device__ __forceinline__ float lcx_pow077(float x) { return powf(x, 0.077f); }
#else
  template <class T> LCX_HD T lcx_div(T a, T b) { return a / b; }
  template <class T> LCX_HD T lcx_rsqrt(T x) { return T(1) / sqrt(x); }
  template <class T> LCX_HD T lcx_exp_small(T x) { return exp(x); }
  template <class T> LCX_HD T lcx_cbrt_ge1(T y) { return cbrt(y); }
  template <class T> LCX_HD T lcx_cbrt1p_mid(T x) { return cbrt(T(1) + x); }
  template <class T> LCX_HD T lcx_pow077(T x) { return pow(x, T(.077)); }
#endif

  template <class T> LCX_HD T tmin(T a, T b) { return (b < a) ? b : a; }   // std::min semantics
  template <class T> LCX_HD T tmax(T a, T b) { return (a < b) ? b : a; }   // std::max semantics (NaN in b is dropped)

  // ------------------------------------------------------------------------------------------------
  // constants: reference include/libcloudph++/common/{moist_air.hpp:26-45,104,110; const_cp.hpp:27-31;
  // theta_std.hpp:20; earth.hpp:17-22; molar_mass.hpp:23-24}.  Derived constants are evaluated in
  // real_t in the same order as the reference's libcloudphxx_const_derived expressions.
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  struct cst
  {
    static LCX_HD real_t c_pd()   { return real_t(1005); }
    static LCX_HD real_t c_pv()   { return real_t(1850); }
    static LCX_HD real_t c_pw()   { return real_t(4218); }
    static LCX_HD real_t M_d()    { return real_t(0.02897); }
    static LCX_HD real_t M_v()    { return real_t(1 * 1e-3) + real_t(17 * 1e-3); }
    static LCX_HD real_t eps()    { return M_v() / M_d(); }
    static LCX_HD real_t kaBoNA() { return real_t(8.3144621); }
    static LCX_HD real_t R_d()    { return kaBoNA() / M_d(); }
    static LCX_HD real_t R_v()    { return kaBoNA() / M_v(); }
    static LCX_HD real_t rho_w()  { return real_t(1e3); }
    static LCX_HD real_t D_0()    { return real_t(2.26e-5); }
    static LCX_HD real_t K_0()    { return real_t(2.4e-2); }
    static LCX_HD real_t p_1000() { return real_t(100000); }
    static LCX_HD real_t p_tri()  { return real_t(611.73); }
    static LCX_HD real_t T_tri()  { return real_t(273.16); }
    static LCX_HD real_t l_tri()  { return real_t(2.5e6); }
    static LCX_HD real_t g()      { return real_t(9.81); }
    static LCX_HD real_t p_stp()  { return real_t(101325); }
    static LCX_HD real_t T_stp()  { return real_t(273.15 + 15); }
    static LCX_HD real_t rho_stp(){ return p_stp() / T_stp() / R_d(); }
    static LCX_HD real_t pi()     { return real_t(3.141592653589793238462643383279502884L); }
  };

  // ------------------------------------------------------------------------------------------------
  // per-cell thermodynamics  (reference src/impl/housekeeping/particles_impl_hskpng_Tpr.ipp:219-305)
  // ------------------------------------------------------------------------------------------------
  // temperature from dry potential temperature and dry-air density: common/theta_dry.hpp:24-35
  template <class real_t>
  LCX_HD real_t T_of_th_dry(real_t th, real_t rhod)
  {
    typedef cst<real_t> c;
    return pow(th * pow(rhod * c::R_d() / c::p_1000(), c::R_d() / c::c_pd()), c::c_pd() / (c::c_pd() - c::R_d()));
  }
  // Exner function: common/theta_std.hpp:36-41
  template <class real_t>
  LCX_HD real_t exner(real_t p) { typedef cst<real_t> c; return pow(p / c::p_1000(), c::R_d() / c::c_pd()); }
  // total pressure from the gas law: common/theta_dry.hpp:47-55
  template <class real_t>
  LCX_HD real_t p_of_rhod_rv_T(real_t rhod, real_t rv, real_t T)
  { typedef cst<real_t> c; return rhod * (c::R_d() + rv * c::R_v()) * T; }
  // vapour partial pressure: common/moist_air.hpp:88-95
  template <class real_t>
  LCX_HD real_t p_v(real_t p, real_t r) { return p * r / (r + cst<real_t>::eps()); }
  // saturation vapour pressure, Clausius-Clapeyron with constant c_p: common/const_cp.hpp:34-43
  template <class real_t>
  LCX_HD real_t p_vs_cc(real_t T)
  {
    typedef cst<real_t> c;
    return c::p_tri() * exp(
      (c::l_tri() + (c::c_pw() - c::c_pv()) * c::T_tri()) / c::R_v() * (real_t(1) / c::T_tri() - real_t(1) / T)
      - (c::c_pw() - c::c_pv()) / c::R_v() * log(T / c::T_tri()));
  }
  // saturation mixing ratio: common/const_cp.hpp:57-63
  template <class real_t>
  LCX_HD real_t r_vs_cc(real_t T, real_t p) { return cst<real_t>::eps() / (p / p_vs_cc(T) - 1); }
  // Tetens: common/tetens.hpp:15-35
  template <class real_t>
  LCX_HD real_t p_vs_tet(real_t T)
  {
    const real_t Tc(T - 273.15);
    return real_t(real_t(6.1078e2) * exp((real_t(17.27) * Tc) / (Tc + real_t(237.3))));
  }
  template <class real_t>
  LCX_HD real_t r_vs_tet(real_t T, real_t p)
  {
    const real_t Tc(T - real_t(273.15));
    return real_t(real_t(380) / (p * exp(real_t(-17.2693882) * (Tc) / (T - real_t(35.86))) - real_t(610.9)));
  }
  // relative humidity by the four formulae: hskpng_Tpr.ipp:71-103,141-161
  template <class real_t>
  LCX_HD real_t RH_of(int formula, real_t p, real_t rv, real_t T)
  {
    switch (formula)
    {
      case RH_PV_CC:  return real_t(p_v(p, rv) / p_vs_cc(T));
      case RH_RV_CC:  return real_t(rv / r_vs_cc(T, p));
      case RH_PV_TET: return real_t(p_v(p, rv) / p_vs_tet(T));
      case RH_RV_TET: return real_t(rv / r_vs_tet(T, p));
      default:        return real_t(0);
    }
  }
  // dynamic viscosity of air: common/vterm.hpp:22-31
  template <class real_t>
  LCX_HD real_t visc(real_t T)
  {
    const real_t q = T / cst<real_t>::T_tri();
    return real_t(1.72 * 1e-5) * (real_t(393) / (T + real_t(120))) * real_t(q * sqrt(q));
  }
  // mean free paths: common/mean_free_path.hpp:16-51
  template <class real_t>
  LCX_HD real_t lambda_D(real_t T)
  { typedef cst<real_t> c; return real_t(2) * c::D_0() / (real_t(sqrt(real_t(2) * real_t(c::R_v() * T)))); }
  template <class real_t>
  LCX_HD real_t lambda_K(real_t T, real_t p)
  { typedef cst<real_t> c; return real_t(.8) * (c::K_0() * T / p) / (real_t(sqrt(real_t(2) * real_t(c::R_d() * T)))); }
  // latent heat of evaporation: common/const_cp.hpp:82-87
  template <class real_t>
  LCX_HD real_t l_v(real_t T)
  { typedef cst<real_t> c; return c::l_tri() + (c::c_pv() - c::c_pw()) * (T - c::T_tri()); }
  // d(theta)/d(rv) at constant ... : common/theta_dry.hpp:60-66
  template <class real_t>
  LCX_HD real_t d_th_d_rv(real_t T, real_t th) { return -th / T * l_v(T) / cst<real_t>::c_pd(); }

  // ------------------------------------------------------------------------------------------------
  // kappa-Koehler / Kelvin  (common/kelvin_term.hpp:23-50, common/kappa_koehler.hpp:31-54)
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  LCX_HD real_t sg_surf(real_t T) { return real_t(0.07275) * (real_t(1.) - real_t(0.002) * (T - real_t(291.))); }
  template <class real_t>
  LCX_HD real_t kelvin_A(real_t T)
  { typedef cst<real_t> c; return real_t(2) * sg_surf(T) / c::R_v() / T / c::rho_w(); }
  template <class real_t>
  LCX_HD real_t klvntrm(real_t r, real_t T) { return exp(kelvin_A(T) / r); }
  template <class real_t>
  LCX_HD real_t a_w(real_t rw3, real_t rd3, real_t kappa) { return (rw3 - rd3) / (rw3 - rd3 * (real_t(1) - kappa)); }
  template <class real_t>
  LCX_HD real_t rw3_eq_nokelvin(real_t rd3, real_t kappa, real_t RH) { return rd3 * (1 - RH * (1 - kappa)) / (1 - RH); }

  // ------------------------------------------------------------------------------------------------
  // TOMS 748 bracketing root finder (Alefeld, Potra & Shi 1995, ACM TOMS 21, 327-344), in the variant the
  // reference ships (common/detail/toms748.hpp:291-454, after Boost.Math): same interpolation choices, the
  // same guard constants and the same termination test, so that the sequence of evaluated abscissae - and
  // hence the returned mid-point - coincides with the reference's whenever f itself agrees.
  // ------------------------------------------------------------------------------------------------
  template <class T> struct fpl;
  template <> struct fpl<double> { static LCX_HD double eps() { return DBL_EPSILON; } static LCX_HD double tiny() { return DBL_MIN; } static LCX_HD double huge_() { return DBL_MAX; } };
  template <> struct fpl<float>  { static LCX_HD float  eps() { return FLT_EPSILON; } static LCX_HD float  tiny() { return FLT_MIN; } static LCX_HD float  huge_() { return FLT_MAX; } };

  // relative-width stopping rule: toms748.hpp:262-286 (bits = sizeof(real_t)*8/4, src/detail/config.hpp:39)
  template <class T>
  struct width_tol
  {
    T eps;
    LCX_HD explicit width_tol(unsigned bits) { eps = tmax(T(ldexp(1.0F, 1 - int(bits))), T(4 * fpl<T>::eps())); }
    LCX_HD bool operator()(const T &a, const T &b) const { return fabs(a - b) <= (eps * tmin(fabs(a), fabs(b))); }
  };

  namespace root748
  {
    template <class T>
    struct state { T a, b, fa, fb, d, fd, e, fe; };

    template <class T>
    LCX_HD T guarded_div(T num, T den, T fallback)
    {
      if (fabs(den) < 1 && fabs(den * fpl<T>::huge_()) <= fabs(num)) return fallback;
      return lcx_div(num, den);
    }

    // shrink [a,b] around the sign change using the trial point c; the discarded end goes to (d,fd)
    template <class F, class T>
    LCX_HD void rebracket(const F &f, state<T> &s, T c)
    {
      const T tol = fpl<T>::eps() * 2;
      if ((s.b - s.a) < 2 * tol * s.a)            c = s.a + (s.b - s.a) / 2;
      else if (c <= s.a + fabs(s.a) * tol)        c = s.a + fabs(s.a) * tol;
      else if (c >= s.b - fabs(s.b) * tol)        c = s.b - fabs(s.a) * tol;
      const T fc = f(c);
      if (fc == 0) { s.a = c; s.fa = 0; s.d = 0; s.fd = 0; return; }
      if (copysign(T(1), s.fa * fc) < 0) { s.d = s.b; s.fd = s.fb; s.b = c; s.fb = fc; }
      else                               { s.d = s.a; s.fd = s.fa; s.a = c; s.fa = fc; }
    }

    template <class T>
    LCX_HD T secant(const T &a, const T &b, const T &fa, const T &fb)
    {
      const T tol = fpl<T>::eps() * 5;
      const T c = a - lcx_div(fa, fb - fa) * (b - a);
      if ((c <= a + fabs(a) * tol) || (c >= b - fabs(b) * tol)) return (a + b) / 2;
      return c;
    }

    // root of the parabola through (a,fa),(b,fb),(d,fd) by `count` Newton steps
    template <class T>
    LCX_HD T quadratic(const T &a, const T &b, const T &d, const T &fa, const T &fb, const T &fd, unsigned count)
    {
      const T B = guarded_div(T(fb - fa), T(b - a), fpl<T>::huge_());
      T A = guarded_div(T(fd - fb), T(d - b), fpl<T>::huge_());
      A = guarded_div(T(A - B), T(d - a), T(0));
      if (A == 0) return secant(a, b, fa, fb);
      T c = (copysign(T(1), A * fa) > 0) ? a : b;
      for (unsigned i = 1; i <= count; ++i)
        c -= guarded_div(T(fa + (B + A * (c - b)) * (c - a)), T(B + A * (2 * c - a - b)), T(1 + c - a));
      if ((c <= a) || (c >= b)) c = secant(a, b, fa, fb);
      return c;
    }

    // inverse cubic interpolation through the four best points
    template <class T>
    LCX_HD T cubic(const T &a, const T &b, const T &d, const T &e, const T &fa, const T &fb, const T &fd, const T &fe)
    {
      // lcx_div is the plain quotient except inside the condensation kernels' fast-math root solve
      const T q11 = lcx_div((d - e) * fd, fe - fd);
      const T q21 = lcx_div((b - d) * fb, fd - fb);
      const T q31 = lcx_div((a - b) * fa, fb - fa);
      const T d21 = lcx_div((b - d) * fd, fd - fb);
      const T d31 = lcx_div((a - b) * fb, fb - fa);
      const T q22 = lcx_div((d21 - q11) * fb, fe - fb);
      const T q32 = lcx_div((d31 - q21) * fa, fd - fa);
      const T d32 = lcx_div((d31 - q21) * fd, fd - fa);
      const T q33 = lcx_div((d32 - q22) * fa, fe - fa);
      T c = q31 + q32 + q33 + a;
      if ((c <= a) || (c >= b)) c = quadratic(a, b, d, fa, fb, fd, 3);
      return c;
    }

    template <class T>
    LCX_HD bool values_coincide(const state<T> &s)
    {
      // "any of the six differences below the threshold" as one comparison of their minimum: no chain of six branches
      // (fmin drops NaNs like the comparisons do; all six NaN -> NaN < m -> false)
      const T m = fpl<T>::tiny() * 32;
      const T d1 = fmin(fabs(s.fa - s.fb), fabs(s.fa - s.fd)), d2 = fmin(fabs(s.fa - s.fe), fabs(s.fb - s.fd)),
              d3 = fmin(fabs(s.fb - s.fe), fabs(s.fd - s.fe));
      return fmin(fmin(d1, d2), d3) < m;
    }
  }

  // returns the midpoint of the final bracket; max_iter is in/out (evaluations allowed / used)
  template <class F, class T, class Tol>
  LCX_HD T toms748(const F &f, const T &ax, const T &bx, const T &fax, const T &fbx, Tol tol, uintmax_t &max_iter)
  {
    using namespace root748;
    uintmax_t left = max_iter;
    state<T> s;
    s.a = ax; s.b = bx; s.fa = fax; s.fb = fbx;
    const T mu = 0.5f;

    if (tol(s.a, s.b) || (s.fa == 0) || (s.fb == 0))
    {
      max_iter = 0;
      if (s.fa == 0) s.b = s.a; else if (s.fb == 0) s.a = s.b;
      return (s.a + s.b) / 2;
    }

    s.fe = s.e = s.fd = 1e5F;
    s.d = 0;

    if (s.fa != 0)
    {
      rebracket(f, s, secant(s.a, s.b, s.fa, s.fb));
      --left;
      if (left && (s.fa != 0) && !tol(s.a, s.b))
      {
        const T c = quadratic(s.a, s.b, s.d, s.fa, s.fb, s.fd, 2);
        s.e = s.d; s.fe = s.fd;
        rebracket(f, s, c);
        --left;
      }
    }

    while (left && (s.fa != 0) && !tol(s.a, s.b))
    {
      const T a0 = s.a, b0 = s.b;
      T c = values_coincide(s) ? quadratic(s.a, s.b, s.d, s.fa, s.fb, s.fd, 2)
                               : cubic(s.a, s.b, s.d, s.e, s.fa, s.fb, s.fd, s.fe);
      s.e = s.d; s.fe = s.fd;
      rebracket(f, s, c);
      if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) break;

      c = values_coincide(s) ? quadratic(s.a, s.b, s.d, s.fa, s.fb, s.fd, 3)
                             : cubic(s.a, s.b, s.d, s.e, s.fa, s.fb, s.fd, s.fe);
      rebracket(f, s, c);
      if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) break;

      // double-length secant step from the end with the smaller residual
      T u, fu;
      if (fabs(s.fa) < fabs(s.fb)) { u = s.a; fu = s.fa; } else { u = s.b; fu = s.fb; }
      c = u - 2 * lcx_div(fu, s.fb - s.fa) * (s.b - s.a);
      if (fabs(c - u) > (s.b - s.a) / 2) c = s.a + (s.b - s.a) / 2;
      s.e = s.d; s.fe = s.fd;
      rebracket(f, s, c);
      if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) break;

      // bisect when the bracket did not halve
      if ((s.b - s.a) < mu * (b0 - a0)) continue;
      s.e = s.d; s.fe = s.fd;
      rebracket(f, s, T(s.a + (s.b - s.a) / 2));
      --left;
    }

    max_iter -= left;
    if (s.fa == 0) s.b = s.a; else if (s.fb == 0) s.a = s.b;
    return (s.a + s.b) / 2;
  }

  // convenience form with the default tolerance / 100 evaluations (toms748.hpp:431-452)
  template <class F, class T>
  LCX_HD T toms748(const F &f, const T &ax, const T &bx)
  {
    uintmax_t it = 100;
    return toms748(f, ax, bx, f(ax), f(bx), width_tol<T>(sizeof(T) * 8 / 4), it);
  }

  // ------------------------------------------------------------------------------------------------
  // equilibrium and critical wet radii (common/kappa_koehler.hpp:58-192)
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  struct rw3_eq_resid
  {
    real_t RH, rd3, kappa, T;
    LCX_HD real_t operator()(real_t rw3) const
    { return RH - a_w(rw3, rd3, kappa) * klvntrm(real_t(cbrt(rw3)), T); }
  };
  template <class real_t>
  LCX_HD real_t rw3_eq(real_t rd3, real_t kappa, real_t RH, real_t T)
  {
    if (kappa == 0) return rd3;
    rw3_eq_resid<real_t> f = {RH, rd3, kappa, T};
    return toms748(f, rd3, rw3_eq_nokelvin(rd3, kappa, RH));
  }
  struct rw3_cr_resid   // always evaluated in double: kappa_koehler.hpp:157-165
  {
    double rd3, kappa, T;
    LCX_HD double operator()(double rw3) const
    {
      return (kelvin_A(T) * (rd3 - rw3) * ((kappa - 1) * rd3 + rw3) + 3 * kappa * rd3 * rw3 * cbrt(rw3));
    }
  };
  template <class real_t>
  LCX_HD real_t rw3_cr(real_t rd3, real_t kappa, real_t T)
  {
    rw3_cr_resid f = {double(rd3), double(kappa), double(T)};
    return real_t(toms748(f, double(1e0 * rd3), double(1e8 * rd3)));
  }
  template <class real_t>
  LCX_HD real_t S_cr(real_t rd3, real_t kappa, real_t T)
  {
    const real_t rw3 = rw3_cr(rd3, kappa, T);
    return a_w(rw3, rd3, kappa) * klvntrm(real_t(cbrt(real_t(rw3))), T);
  }

  // ------------------------------------------------------------------------------------------------
  // condensational growth  (src/impl/condensation/common/particles_impl_cond_common.ipp:79-338,
  // common/maxwell-mason.hpp:15-47, common/ventil.hpp:17-79, common/transition_regime.hpp:15-19)
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  LCX_HD real_t beta_tr(real_t Kn) { return (1 + Kn) / (1 + real_t(1.71) * Kn + real_t(1.33) * Kn * Kn); }

  // Nu = 1 + cbrt(1 + Re Pr) max(1, Re^0.077).  pow() is only evaluated when it can exceed 1:
  // x^0.077 <= 1 for x <= 1 and is NaN for x < 0 (a -1 "invalid" fall speed), which max() drops.
  template <class real_t>
  LCX_HD real_t nusselt(real_t Pr, real_t Re)
  {
    const real_t boost = (Re > real_t(1)) ? tmax(real_t(1), real_t(pow(Re, real_t(.077)))) : real_t(1);
    return real_t(1) + cbrt(real_t(1) + Re * Pr) * boost;
  }

  // everything drw2/dt needs that is shared by all super-droplets of one cell
  template <class real_t>
  struct cond_cell
  {
    real_t rhod, rv, T, p, RH, eta, lambda_D, lambda_K;
  };

  template <class real_t>
  struct growth_fn   // f(x) = rw2_old + dt * drw2_dt(x) - x, the backward-Euler residual
  {
    real_t rw2_old, dt, rhod, rv, T, p, RH_eff, eta, rd3, kpa, vt, lam_D, lam_K;

    LCX_HD real_t drw2_dt(real_t rw2) const
    {
      typedef cst<real_t> c;
      const real_t rw = sqrt(rw2);
      const real_t rw3 = rw * rw * rw;
      const real_t Re = vt * (real_t(2) * rw) * rhod / eta;
      const real_t Sc = eta / rhod / c::D_0();
      const real_t Pr = c::c_pd() * eta / c::K_0();
      const real_t D = c::D_0() * beta_tr(lam_D / rw) * (nusselt(Sc, Re) / 2);
      const real_t K = c::K_0() * beta_tr(lam_K / rw) * (nusselt(Pr, Re) / 2);
      const real_t lv = l_v(T);
      const real_t rho_v = rhod * rv;
      const real_t rdrdt = (real_t(1) - a_w(rw3, rd3, kpa) * klvntrm(rw, T) / RH_eff)
        / c::rho_w()
        / (real_t(1) / D / rho_v + lv / K / RH_eff / T * (lv / c::R_v() / T - real_t(1)));
      return real_t(2) * rdrdt;
    }
    LCX_HD real_t operator()(const real_t &x) const { return (rw2_old + dt * drw2_dt(x) - x); }
  };

  // One implicit-Euler step of rw^2: control flow of advance_rw2::operator() (cond_common.ipp:187-337) with its
  // TOMS 748 root solve (toms748 above) turned inside out into a resumable state machine: the caller evaluates the
  // growth law at point() and hands the value to feed().  Same abscissae, same arithmetic, same result as calling
  // toms748(f, a, b, fa, fb, tol, n_iter).  Why: on the GPU the growth law then has ONE evaluation site, so (i) the
  // kernel fits the instruction cache (straightforward inlining of toms748 puts ~6 copies of f into 100 KB of SASS and
  // stalls on instruction fetch) and (ii) a lane that finishes its droplet early can start the next one while its
  // neighbours keep iterating - all lanes stay in the same loop body whatever phase or droplet they are at.
  template <class real_t>
  struct euler_rw2_solver
  {
    enum { INIT0, INIT1, FIRST, SECOND, LOOP1, LOOP2, LOOP3, LOOP4 };
    root748::state<real_t> s;
    real_t rw2_old, dt, drw2, rd2, a0, b0, c;
    unsigned left;
    int phase;
    bool bracketing;      // the pending evaluation is a rebracket() step: c was guarded, [a,b] is updated afterwards

    LCX_HD void start(real_t rw2_old_, real_t dt_, unsigned n_iter = 100)
    {
      rw2_old = rw2_old_; dt = dt_; c = rw2_old_; left = n_iter; phase = INIT0; bracketing = false;
      s.a = s.b = s.fa = s.fb = s.d = s.fd = s.e = s.fe = 0; drw2 = rd2 = a0 = b0 = 0;
    }

    // abscissa at which drw2/dt is wanted next
    LCX_HD real_t point()
    {
      if (bracketing)
      {
        const real_t t2 = fpl<real_t>::eps() * 2;
        if ((s.b - s.a) < 2 * t2 * s.a)          c = s.a + (s.b - s.a) / 2;
        else if (c <= s.a + fabs(s.a) * t2)      c = s.a + fabs(s.a) * t2;
        else if (c >= s.b - fabs(s.b) * t2)      c = s.b - fabs(s.a) * t2;
      }
      return c;
    }

    LCX_HD real_t clamp(real_t r) const { return r < rd2 ? rd2 : r; }

    // consumes g = drw2_dt(point()); returns true when the step is complete (result = new rw2)
    LCX_HD bool feed(real_t g, real_t rd3, real_t &result, real_t cond_mlt = 2)
    {
      using namespace root748;
      const width_tol<real_t> tol(sizeof(real_t) * 8 / 4);
      const real_t mu = 0.5f;
      const real_t fc = (rw2_old + dt * g - c);
      if (bracketing)
      {
        if (fc == 0) { s.a = c; s.fa = 0; s.d = 0; s.fd = 0; }
        else if (copysign(real_t(1), s.fa * fc) < 0) { s.d = s.b; s.fd = s.fb; s.b = c; s.fb = fc; }
        else                                         { s.d = s.a; s.fd = s.fa; s.a = c; s.fa = fc; }
      }

      bool finish = false, loop_head = false;
      int newton_steps = 0;          // > 0: interpolate (cubic when the four function values are distinct, else quadratic)
      bool quadratic_only = false;
      switch (phase)
      {
        case INIT0:
        {
          drw2 = dt * g;
          if (drw2 == 0) { result = rw2_old; return true; }
          const real_t rd = cbrt(rd3);
          rd2 = rd * rd;
          s.a = tmax(rd2, rw2_old + tmin(real_t(0), cond_mlt * drw2));
          s.b = rw2_old + tmax(real_t(0), cond_mlt * drw2);
          if (s.a == s.b) { result = rw2_old; return true; }
          c = (drw2 > 0) ? s.b : s.a;
          phase = INIT1;
          return false;
        }
        case INIT1:
        {
          if (drw2 > 0) { s.fa = drw2; s.fb = fc; } else { s.fa = fc; s.fb = drw2; }
          if (s.fa * s.fb > 0) { result = clamp(rw2_old + drw2); return true; }      // not bracketed: explicit Euler
          if (tol(s.a, s.b) || (s.fa == 0) || (s.fb == 0)) { finish = true; break; }
          s.fe = s.e = s.fd = 1e5F;
          c = secant(s.a, s.b, s.fa, s.fb);
          bracketing = true;
          phase = FIRST;
          return false;
        }
        case FIRST:
          --left;
          if (left && (s.fa != 0) && !tol(s.a, s.b)) { newton_steps = 2; quadratic_only = true; phase = SECOND; }
          else loop_head = true;
          break;
        case SECOND:
          --left;
          loop_head = true;
          break;
        case LOOP1:
          if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) { finish = true; break; }
          newton_steps = 3; phase = LOOP2;
          break;
        case LOOP2:
        {
          if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) { finish = true; break; }
          real_t u, fu;
          if (fabs(s.fa) < fabs(s.fb)) { u = s.a; fu = s.fa; } else { u = s.b; fu = s.fb; }
          c = u - 2 * lcx_div(fu, s.fb - s.fa) * (s.b - s.a);
          if (fabs(c - u) > (s.b - s.a) / 2) c = s.a + (s.b - s.a) / 2;
          s.e = s.d; s.fe = s.fd;
          phase = LOOP3;
          return false;
        }
        case LOOP3:
          if ((0 == --left) || (s.fa == 0) || tol(s.a, s.b)) { finish = true; break; }
          if ((s.b - s.a) < mu * (b0 - a0)) { loop_head = true; break; }
          s.e = s.d; s.fe = s.fd;
          c = real_t(s.a + (s.b - s.a) / 2);
          phase = LOOP4;
          return false;
        case LOOP4:
          --left;
          loop_head = true;
          break;
      }

      if (loop_head)
      {
        if (left && (s.fa != 0) && !tol(s.a, s.b)) { a0 = s.a; b0 = s.b; newton_steps = 2; phase = LOOP1; }
        else finish = true;
      }
      if (finish)
      {
        if (s.fa == 0) s.b = s.a; else if (s.fb == 0) s.a = s.b;
        result = clamp((s.a + s.b) / 2);
        return true;
      }
      // newton_steps > 0 here: next trial point by interpolation (one shared site for the cubic / quadratic formulae)
      bool have = false;
      if (!quadratic_only && !values_coincide(s))
      {
        // inverse cubic interpolation; out-of-bracket results fall back to three Newton steps on the parabola
        const real_t q11 = lcx_div((s.d - s.e) * s.fd, s.fe - s.fd);
        const real_t q21 = lcx_div((s.b - s.d) * s.fb, s.fd - s.fb);
        const real_t q31 = lcx_div((s.a - s.b) * s.fa, s.fb - s.fa);
        const real_t d21 = lcx_div((s.b - s.d) * s.fd, s.fd - s.fb);
        const real_t d31 = lcx_div((s.a - s.b) * s.fb, s.fb - s.fa);
        const real_t q22 = lcx_div((d21 - q11) * s.fb, s.fe - s.fb);
        const real_t q32 = lcx_div((d31 - q21) * s.fa, s.fd - s.fa);
        const real_t d32 = lcx_div((d31 - q21) * s.fd, s.fd - s.fa);
        const real_t q33 = lcx_div((d32 - q22) * s.fa, s.fe - s.fa);
        c = q31 + q32 + q33 + s.a;
        have = !((c <= s.a) || (c >= s.b));
        if (!have) newton_steps = 3;
      }
      if (!have) c = quadratic(s.a, s.b, s.d, s.fa, s.fb, s.fd, unsigned(newton_steps));
      // the point removed by the coming rebracket becomes (e,fe) - except after the second step of the loop body,
      // where the reference does not refresh it (toms748.hpp:381-392)
      if (phase != LOOP2) { s.e = s.d; s.fe = s.fd; }
      return false;
    }
  };

  template <class F, class real_t>
  LCX_HD real_t implicit_euler_rw2(const F &f, real_t rw2_old, real_t rd3, real_t dt)
  {
    euler_rw2_solver<real_t> sv;
    sv.start(rw2_old, dt);
    real_t result = rw2_old;
    for (;;)
      if (sv.feed(f.drw2_dt(sv.point()), rd3, result)) return result;
  }


  // The same implicit-Euler step with a root search that stops as soon as the root is KNOWN to the reference's tolerance,
  // instead of shrinking a bracket to it from both sides.  Everything up to and including the choice between "explicit
  // Euler", "unchanged" and "root search" is the reference's (cond_common.ipp:187-303); the search itself is a safeguarded
  // secant iteration: f(x) = rw2_old + dt g(x) - x is smooth and, over one (sub-)step, close to linear for every droplet that
  // is not a haze particle in equilibrium, so the first secant point is usually already 1e-6 of the bracket from the root.
  // An iterate c is accepted when |f(c) / f'| < 2^-19 c with f' from the last two points (and then corrected by that last
  // secant step): error <= 2^-18 relative, the reference's own midpoint-of-bracket answer carries 2^-16, so the two differ
  // by less than the 2^-15 that two valid TOMS 748 answers may differ by.  The sign-changing bracket is kept throughout and
  // bisected when the secant point leaves it; if nothing is accepted within 40 evaluations the bracket midpoint is returned
  // once it is narrower than 2^-15 relative (never observed).  North star: "toms748/Newton ... within a stated tolerance".
  template <class F, class real_t>
  LCX_HD real_t implicit_euler_rw2_secant(const F &f, real_t rw2_old, real_t rd3, real_t dt, real_t cond_mlt = 2)
  {
    const real_t drw2 = dt * f.drw2_dt(rw2_old);
    if (drw2 == 0) return rw2_old;
    const real_t rd = cbrt(rd3), rd2 = rd * rd;
    real_t a = tmax(rd2, rw2_old + tmin(real_t(0), cond_mlt * drw2));
    real_t b = rw2_old + tmax(real_t(0), cond_mlt * drw2);
    if (a == b) return rw2_old;
    real_t fa, fb;
    if (drw2 > 0) { fa = drw2; fb = rw2_old + dt * f.drw2_dt(b) - b; }
    else          { fb = drw2; fa = rw2_old + dt * f.drw2_dt(a) - a; }
    if (fa * fb > 0) { const real_t r = rw2_old + drw2; return r < rd2 ? rd2 : r; }      // not bracketed: explicit Euler
    if (fa == 0) return a < rd2 ? rd2 : a;
    if (fb == 0) return b < rd2 ? rd2 : b;

    const real_t accept = real_t(1) / real_t(524288), width = real_t(1) / real_t(32768);        // 2^-19, 2^-15
    // the two most recent points (p0 older) drive the secant; [a, b] always brackets the root
    real_t p0 = a, f0 = fa, p1 = b, f1 = fb;
    if (fabs(fa) < fabs(fb)) { p0 = b; f0 = fb; p1 = a; f1 = fa; }
    for (int it = 0; it < 40; ++it)
    {
      const real_t slope = lcx_div(f1 - f0, p1 - p0);
      real_t c = p1 - lcx_div(f1, slope);
      if (!(c > a && c < b)) c = a + (b - a) / 2;
      const real_t fc = rw2_old + dt * f.drw2_dt(c) - c;
      if (fc == 0) return c < rd2 ? rd2 : c;
      if ((fc > 0) == (fa > 0)) { a = c; fa = fc; } else { b = c; fb = fc; }
      // slope through the newest two points; a usable estimate needs it to point the way f does (f decreases across the root)
      const real_t s2 = lcx_div(fc - f1, c - p1);
      if (s2 < 0)
      {
        const real_t step = lcx_div(fc, s2);
        if (fabs(step) < accept * c)
        {
          real_t r = c - step;
          if (!(r > a && r < b)) r = c;
          return r < rd2 ? rd2 : r;
        }
      }
      if ((b - a) <= width * tmin(fabs(a), fabs(b))) break;
      p0 = p1; f0 = f1; p1 = c; f1 = fc;
      if (f1 == f0) { p0 = (fc > 0) == (fa > 0) ? b : a; f0 = (fc > 0) == (fa > 0) ? fb : fa; }     // flat spot: fall back to the bracket end
    }
    const real_t r = a + (b - a) / 2;
    return r < rd2 ? rd2 : r;
  }

  template <class real_t>
  LCX_HD real_t advance_rw2(real_t rw2_old, real_t rd3, real_t kpa, real_t vt, const cond_cell<real_t> &cl,
                            real_t dt, real_t RH_max)
  {
    if (rw2_old <= 0) return rw2_old;
    growth_fn<real_t> f;
    f.rw2_old = rw2_old; f.dt = dt; f.rhod = cl.rhod; f.rv = cl.rv; f.T = cl.T; f.p = cl.p;
    f.RH_eff = cl.RH > RH_max ? RH_max : cl.RH;
    f.eta = cl.eta; f.rd3 = rd3; f.kpa = kpa; f.vt = vt; f.lam_D = cl.lambda_D; f.lam_K = cl.lambda_K;
    return implicit_euler_rw2(f, rw2_old, rd3, dt);
  }

  // ---- the same growth law, arranged for throughput ---------------------------------------------------
  // drw2/dt written as ONE quotient: every per-cell factor is precomputed once per cell (cond_cell_consts), the two
  // transition-regime factors, the water activity and the Maxwell-Mason denominator share a single division, 1/rw
  // comes from rsqrt, and cbrt(1 + x) uses its series for x < 1e-4 (cloud droplets: Re Sc ~ 1e-6).  Algebraically
  // identical to growth_fn::drw2_dt; numerically within a few ulp of it (the same size as the libm differences between
  // host and device), so results stay inside the root solver's own 2^-15 bracket tolerance.
  template <class real_t>
  struct cond_cell_consts
  {
    real_t c_Re;      // 2 rhod / eta              -> Re = vt * rw * c_Re
    real_t Sc, Pr;
    real_t lam_D, lam_K;
    real_t A;         // Kelvin length 2 sigma / (R_v T rho_w)
    real_t inv_RH;    // 1 / min(RH, RH_max)
    real_t X;         // 1 / (D_0 rho_v)
    real_t Y;         // l_v (l_v / (R_v T) - 1) / (K_0 RH T)
  };

  template <class real_t>
  LCX_HD cond_cell_consts<real_t> make_cond_consts(const cond_cell<real_t> &cl, real_t RH_max)
  {
    typedef cst<real_t> c;
    cond_cell_consts<real_t> k;
    const real_t RH_eff = cl.RH > RH_max ? RH_max : cl.RH;
    const real_t lv = l_v(cl.T);
    k.c_Re = real_t(2) * cl.rhod / cl.eta;
    k.Sc = cl.eta / cl.rhod / c::D_0();
    k.Pr = c::c_pd() * cl.eta / c::K_0();
    k.lam_D = cl.lambda_D; k.lam_K = cl.lambda_K;
    k.A = kelvin_A(cl.T);
    k.inv_RH = real_t(1) / RH_eff;
    k.X = real_t(1) / (c::D_0() * (cl.rhod * cl.rv));
    k.Y = lv * (lv / c::R_v() / cl.T - real_t(1)) / (c::K_0() * RH_eff * cl.T);
    return k;
  }

  template <class real_t>
  LCX_HD real_t nusselt_fast(real_t P, real_t Re)
  {
    const real_t x = Re * P;
    const real_t cb = (fabs(x) < real_t(1e-4))
      ? real_t(1) + x * (real_t(1. / 3) - x * (real_t(1. / 9) - x * real_t(5. / 81)))
      : real_t(cbrt(real_t(1) + x));
    const real_t boost = (Re > real_t(1)) ? tmax(real_t(1), real_t(pow(Re, real_t(.077)))) : real_t(1);
    return real_t(1) + cb * boost;
  }

  // drw2/dt in the single-quotient form; k may live in shared memory (the staged condensation kernel reads it from there)
  template <class real_t>
  LCX_HD real_t drw2_dt_fast(real_t rw2, real_t rd3, real_t rd3_dry, real_t vt_cRe, const cond_cell_consts<real_t> &k)
  {
    const real_t inv_rw = lcx_rsqrt(rw2);
    const real_t rw = rw2 * inv_rw;
    const real_t rw3 = rw2 * rw;
    const real_t Re = vt_cRe * rw;
    // Sh = Nu(Sc, Re), Nu = Nu(Pr, Re): the Re^0.077 factor is shared, the two cube roots go through one code site.
    // Re > 1 means drizzle and rain drops; for them Re^0.077 = exp(0.077 ln Re) (relative error ~ 0.077 ln Re * 2^-53, a fraction
    // of an ulp) at a third of the instructions of the library pow(), which works to full precision for any exponent.
    const real_t boost = (Re > real_t(1)) ? tmax(real_t(1), real_t(lcx_pow077(Re))) : real_t(1);
    real_t nu[2] = {k.Sc, k.Pr};
    // (both rounds unrolled on purpose: the two cube-root chains are independent and interleave; measured 5 % on the kernel)
    for (int q = 0; q < 2; ++q)
    {
      const real_t x = Re * nu[q];
      const real_t cb = (fabs(x) < real_t(LCX_KC(9, 1e-4)))
        ? real_t(1) + x * (real_t(LCX_KC(10, 1. / 3)) - x * (real_t(LCX_KC(11, 1. / 9)) - x * real_t(LCX_KC(12, 5. / 81))))
        : (x <= real_t(0.5)) ? real_t(lcx_cbrt1p_mid(x)) : real_t(lcx_cbrt_ge1(real_t(1) + x));
      nu[q] = real_t(1) + cb * boost;
    }
    const real_t Sh = nu[0], Nu = nu[1];
    const real_t KnD = k.lam_D * inv_rw, KnK = k.lam_K * inv_rw;
    const real_t c171 = real_t(LCX_KC(13, 1.71)), c133 = real_t(LCX_KC(14, 1.33));
    const real_t bDn = real_t(1) + KnD, bDd = real_t(1) + KnD * (c171 + c133 * KnD);
    const real_t bKn = real_t(1) + KnK, bKd = real_t(1) + KnK * (c171 + c133 * KnK);
    const real_t awn = rw3 - rd3, awd = rw3 - rd3_dry;
    const real_t klv = lcx_exp_small(k.A * inv_rw);
    const real_t tD = bDn * Sh, tK = bKn * Nu;
    const real_t num = (awd - awn * klv * k.inv_RH) * (tD * tK);
    const real_t den = awd * (k.X * bDd * tK + k.Y * bKd * tD);
    return lcx_div(num, cst<real_t>::rho_w() * den);
  }

  template <class real_t>
  struct growth_fast
  {
    real_t rw2_old, dt, rd3, rd3_dry, vt_cRe;   // rd3_dry = rd3 (1 - kappa), vt_cRe = vt * c_Re
    cond_cell_consts<real_t> k;

    LCX_HD real_t drw2_dt(real_t rw2) const { return drw2_dt_fast(rw2, rd3, rd3_dry, vt_cRe, k); }
    LCX_HD real_t operator()(const real_t &x) const { return (rw2_old + dt * drw2_dt(x) - x); }
  };

  // root search of the condensation step: COND_TOMS748 (default) - the reference's TOMS 748 with the same trial points (fast
  // growth law); COND_EXACT - TOMS 748 and the growth law transcribed operation by operation; COND_SECANT (opt-in) - stop as
  // soon as the root is known to the reference's tolerance
  enum { COND_SECANT = 0, COND_TOMS748 = 1, COND_EXACT = 2 };

  template <bool TOMS, class real_t>
  LCX_HD real_t advance_rw2_fast(real_t rw2_old, real_t rd3, real_t kpa, real_t vt, const cond_cell_consts<real_t> &k, real_t dt)
  {
    if (rw2_old <= 0) return rw2_old;
    growth_fast<real_t> f;
    f.rw2_old = rw2_old; f.dt = dt; f.rd3 = rd3; f.rd3_dry = rd3 * (real_t(1) - kpa); f.vt_cRe = vt * k.c_Re; f.k = k;
    if (TOMS) return implicit_euler_rw2(f, rw2_old, rd3, dt);          // the reference's root search, same trial points
    return implicit_euler_rw2_secant(f, rw2_old, rd3, dt);
  }

  // ------------------------------------------------------------------------------------------------
  // terminal velocities  (common/vterm.hpp:38-221; dispatch src/impl/housekeeping/
  // particles_impl_hskpng_vterm.ipp:38-121; LUT src/impl/initialization/particles_impl_init_vterm.ipp:36-59)
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  LCX_HD real_t vt_khvorostyanov(real_t r, real_t /*T*/, real_t rhoa, real_t eta, bool spherical)
  {
    const double rho_w = cst<double>::rho_w(), g = cst<double>::g();
    const double r_d(r), rhoa_d(rhoa), eta_d(eta);
    const double X = double(32. / 3) * (rho_w - rhoa_d) / rhoa_d * g * r_d * r_d * r_d / eta_d / eta_d * rhoa_d * rhoa_d;
    const double b = double(.0902 / 2) * sqrt(X) /
      ((sqrt(double(1) + double(.0902) * sqrt(X)) - double(1)) * (sqrt(double(1) + double(.0902) * sqrt(X))));
    const double pow_hlpr = sqrt(double(1) + double(.0902) * sqrt(X)) - double(1);
    const double a = double(9.06 * 9.06 / 4) * pow_hlpr * pow_hlpr / pow(X, b);
    double Av;
    if (spherical)
    {
      Av = a * pow(eta_d / rhoa_d * double(1e4), double(1) - double(2) * b)
             * pow(double(4. / 3) * rho_w / rhoa_d * g * double(1e2), b);
    }
    else
    {
      const double lambda_half = double(2.35e-3);
      const double ksi = exp(-r_d / lambda_half) + (double(1) - exp(-r_d / lambda_half)) / (double(1) + r_d / lambda_half);
      const double alfa = cst<double>::pi() / double(6) * rho_w * ksi;
      Av = a * pow(eta_d / rhoa_d * double(1e4), double(1) - double(2) * b)
             * pow(double(2.546479) * alfa / rhoa_d * g * double(1e2), b);
    }
    const double Bv = double(3) * b - double(1);
    return real_t((Av * double(pow(double(2 * 1e2) * r_d, Bv))) / double(1e2));
  }

  // sea-level fall speed of Beard (1977), always in double: vterm.hpp:112-134
  template <class real_t>
  LCX_HD real_t vt_beard77_v0(real_t r)
  {
    const double m_s[4] = {0.105035e2, 0.108750e1, -0.133245, -0.659969e-2};
    const double m_l[8] = {0.65639e1, -0.10391e1, -0.14001e1, -0.82736e0, -0.34277e0, -0.83072e-1, -0.10583e-1, -0.54208e-3};
    const double x = log(2 * 100 * r);
    double y = 0;
    if (r <= double(20e-6)) { for (int i = 0; i < 4; ++i) y += m_s[i] * pow(x, double(i)); }
    else                    { for (int i = 0; i < 8; ++i) y += m_l[i] * pow(x, double(i)); }
    return real_t(exp(y) / 100.);
  }
  // altitude correction factor of Beard (1977): vterm.hpp:137-164
  template <class real_t>
  LCX_HD real_t vt_beard77_fact(real_t r, real_t p, real_t rhoa, real_t eta)
  {
    typedef cst<real_t> c;
    const real_t eta_0(1.818e-5);
    if (r <= real_t(20e-6))
    {
      const real_t l_0(6.62e-8);
      const real_t l(l_0 * (eta / eta_0) * sqrt(c::p_stp() / p * c::rho_stp() / rhoa));
      return (eta_0 / eta) * (1 + real_t(1.255) * (l / r)) / (1 + real_t(1.255) * (l_0 / r));
    }
    const real_t eps_s = (eta_0 / eta) - 1;
    const real_t eps_c = sqrt(c::rho_stp() / rhoa) - 1;
    return real_t(1.104) * eps_s
      + ((real_t(1.058) * eps_c - real_t(1.104) * eps_s) * (real_t(5.52) + log(2 * 100 * r)) / real_t(5.01)) + 1;
  }
  // The same factor with its per-cell sub-expressions evaluated once per cell (identical operations in identical
  // order, so identical results): what depends on the droplet is l / r, l_0 / r and log(r) only.
  template <class real_t>
  struct beard77_cell { real_t l, eta_ratio, eps_s, eps_c; };
  template <class real_t>
  LCX_HD beard77_cell<real_t> vt_beard77_cell_consts(real_t p, real_t rhoa, real_t eta)
  {
    typedef cst<real_t> c;
    const real_t eta_0(1.818e-5), l_0(6.62e-8);
    beard77_cell<real_t> k;
    k.l = l_0 * (eta / eta_0) * sqrt(c::p_stp() / p * c::rho_stp() / rhoa);
    k.eta_ratio = eta_0 / eta;
    k.eps_s = (eta_0 / eta) - 1;
    k.eps_c = sqrt(c::rho_stp() / rhoa) - 1;
    return k;
  }
  template <class real_t>
  LCX_HD real_t vt_beard77_fact(real_t r, const beard77_cell<real_t> &k)
  {
    if (r <= real_t(20e-6))
      return k.eta_ratio * (1 + real_t(1.255) * (k.l / r)) / (1 + real_t(1.255) * (real_t(6.62e-8) / r));
    return real_t(1.104) * k.eps_s
      + ((real_t(1.058) * k.eps_c - real_t(1.104) * k.eps_s) * (real_t(5.52) + log(2 * 100 * r)) / real_t(5.01)) + 1;
  }
  // Beard (1976): vterm.hpp:168-221
  template <class real_t>
  LCX_HD real_t vt_beard76(real_t r, real_t T, real_t p, real_t rhoa, real_t eta)
  {
    typedef cst<real_t> c;
    if (r <= real_t(9.5e-6))
    {
      const real_t l = (real_t(6.62e-8) * (eta / real_t(1.818e-5)) * (c::p_stp() / p) * sqrt(real_t(T) / real_t(293.15)));
      const real_t C_ac = real_t(1.) + real_t(1.255) * l / r;
      return ((c::rho_w() - rhoa) * c::g() / (real_t(4.5) * eta) * C_ac * r * r);
    }
    else if (r <= real_t(5.035e-4))
    {
      const double b[7] = {-0.318657e1, 0.992696, -0.153193e-2, -0.987059e-3, -0.578878e-3, 0.855176e-4, -0.327815e-5};
      const real_t l = (real_t(6.62e-8) * (eta / real_t(1.818e-5)) * (c::p_stp() / p) * sqrt(real_t(T) / real_t(293.15)));
      const real_t C_ac = real_t(1.) + real_t(1.255) * l / r;
      const real_t log_N_Da = log(real_t(32. / 3.) * r * r * r * rhoa * (c::rho_w() - rhoa) * c::g() / eta / eta);
      real_t Y = 0.;
      for (int i = 0; i < 7; ++i) Y = double(Y) + b[i] * pow(double(log_N_Da), double(i));
      const real_t N_Re = C_ac * exp(double(Y));
      return (eta * N_Re / rhoa / real_t(2.) / r);
    }
    else
    {
      const real_t b[6] = {real_t(-0.500015e1), real_t(0.523778e1), real_t(-0.204914e1), real_t(0.475294), real_t(-0.542819e-1), real_t(0.238449e-2)};
      const real_t sg = sg_surf(T);
      const real_t Bo = real_t(16. / 3.) * r * r * (c::rho_w() - rhoa) * c::g() / sg;
      const real_t N_p = sg * sg * sg * rhoa * rhoa / eta / eta / eta / eta / c::g() / (c::rho_w() - rhoa);
      const real_t X = log(Bo * pow(N_p, real_t(1. / 6.)));
      real_t Y = 0.;
      for (int i = 0; i < 6; ++i) Y = Y + b[i] * pow(X, real_t(i));
      const real_t N_Re = pow(N_p, real_t(1. / 6.)) * exp(Y);
      return (eta * N_Re / rhoa / real_t(2.) / r);
    }
  }

  // log-spaced cache of vt_beard77_v0 (src/detail/config.hpp:27-44)
  enum { VT0_N_BIN = 10000 };
  template <class real_t>
  struct vt0_bins
  {
    real_t ln_r_min, ln_r_max, dlnr;
    LCX_HD vt0_bins() : ln_r_min(real_t(log(5e-7))), ln_r_max(real_t(log(3e-3))) { dlnr = (ln_r_max - ln_r_min) / VT0_N_BIN; }
    // bin index is round-tripped through real_t exactly as the reference stores it (hskpng_vterm.ipp:198-207)
    LCX_HD int bin_of(real_t rw2) const
    {
      const real_t lnr = real_t(.5) * log(rw2);
      const real_t as_real = lnr <= ln_r_min ? real_t(0) : lnr >= ln_r_max ? real_t(VT0_N_BIN - 1) : real_t(int((lnr - ln_r_min) / dlnr));
      return int(as_real);
    }
    LCX_HD real_t mid(int it) const { return real_t(exp(ln_r_min + (it + 0.5) * dlnr)); }
  };

  template <class real_t>
  LCX_HD real_t vt_of(int formula, real_t rw2, real_t T, real_t p, real_t rhod, real_t eta, const real_t *vt0_table)
  {
    switch (formula)
    {
      case VT_BEARD76:  return vt_beard76(real_t(sqrt(rw2)), T, p, rhod, eta);
      case VT_BEARD77:  return vt_beard77_fact(real_t(sqrt(rw2)), p, rhod, eta) * vt_beard77_v0(real_t(sqrt(rw2)));
      case VT_BEARD77FAST:
      {
        const vt0_bins<real_t> bins;
        return vt_beard77_fact(real_t(sqrt(rw2)), p, rhod, eta) * vt0_table[bins.bin_of(rw2)];
      }
      case VT_KHVOROSTYANOV_SPHERICAL:    return vt_khvorostyanov(real_t(sqrt(rw2)), T, rhod, eta, true);
      case VT_KHVOROSTYANOV_NONSPHERICAL: return vt_khvorostyanov(real_t(sqrt(rw2)), T, rhod, eta, false);
      default: return real_t(0);
    }
  }

  // ------------------------------------------------------------------------------------------------
  // collision kernels (src/detail/kernels.hpp:40-202, kernel_interpolation.hpp:9-64, kernel_utils.hpp:12-29)
  // ------------------------------------------------------------------------------------------------
  template <class real_t>
  struct coal_kernel_params
  {
    int kind;                 // KERNEL_*
    int has_multiplier;       // geometric kernel with one user parameter
    real_t user0;             // Golovin b / geometric multiplier
    real_t r_max;             // largest tabulated radius [um], efficiency kernels
    const real_t *eff;        // packed lower-triangular efficiency table
  };

  // kernel_utils.hpp:12-19.  The reference takes the integer radius, compares and divides it in double and truncates:
  // int(100 + (R - 100.) / 10.).  For an integer R > 100 the rounded quotient never reaches the next integer (its fractional
  // part is a multiple of 1/10), so integer arithmetic gives the same index without the 64-bit int -> double conversions.
  LCX_HD int kernel_index(n_t R)
  {
    const uint32_t r = uint32_t(R);            // radii in micrometres, below the table's r_max
    if (r <= 100u) return int(r);
    return int(100u + (r - 100u) / 10u);
  }
  // kernel_utils.hpp:21-28 (n_user_params = 0 for tables): size_t(0.5 * i * (i + 1) + j), exact in double, hence in integers
  LCX_HD size_t kernel_vector_index(int i, int j)
  {
    const uint32_t a = uint32_t(i >= j ? i : j), b = uint32_t(i >= j ? j : i);
    return size_t(a * (a + 1u) / 2u + b);
  }

  template <class real_t>
  LCX_HD real_t interpolated_efficiency(const coal_kernel_params<real_t> &kp, real_t r1, real_t r2)
  {
    r1 *= 1e6; r2 *= 1e6;
    if (r1 >= kp.r_max) r1 = kp.r_max - 1e-6;
    if (r2 >= kp.r_max) r2 = kp.r_max - 1e-6;
    n_t dx, dy, x[4];
    if (r1 >= 100.) { x[0] = n_t(floor(r1 / 10.) * 10); dx = 10; } else { x[0] = n_t(floor(r1)); dx = 1; }
    if (r2 >= 100.) { x[2] = n_t(floor(r2 / 10.) * 10); dy = 10; } else { x[2] = n_t(floor(r2)); dy = 1; }
    x[1] = x[0] + dx;
    x[3] = x[2] + dy;
    const size_t iv0 = kernel_vector_index(kernel_index(x[0]), kernel_index(x[2])),
                 iv1 = kernel_vector_index(kernel_index(x[1]), kernel_index(x[2])),
                 iv2 = kernel_vector_index(kernel_index(x[0]), kernel_index(x[3])),
                 iv3 = kernel_vector_index(kernel_index(x[1]), kernel_index(x[3]));
    real_t w[4];
    w[0] = r1 - x[0];
    w[1] = x[1] - r1;
    w[2] = r2 - x[2];
    w[3] = x[3] - r2;
    // ... / dx / dy with dx, dy = 1 below 100 um: dividing by one changes nothing, so those two divisions are skipped
    real_t res = kp.eff[iv0] * w[1] * w[3] + kp.eff[iv1] * w[0] * w[3] + kp.eff[iv2] * w[1] * w[2] + kp.eff[iv3] * w[0] * w[2];
    if (dx != 1) res = res / dx;
    if (dy != 1) res = res / dy;
    return res;
  }

  template <class real_t>
  LCX_HD real_t geometric_kernel(n_t n_a, n_t n_b, real_t rw2_a, real_t rw2_b, real_t vt_a, real_t vt_b)
  {
    return cst<real_t>::pi() * tmax(n_a, n_b) * fabs(vt_a - vt_b) * (rw2_a + rw2_b + 2. * sqrt(rw2_a * rw2_b));
  }

  template <class real_t>
  LCX_HD real_t coal_kernel(const coal_kernel_params<real_t> &kp, n_t n_a, n_t n_b, real_t rw2_a, real_t rw2_b, real_t vt_a, real_t vt_b)
  {
    switch (kp.kind)
    {
      case KERNEL_GOLOVIN:
        return cst<real_t>::pi() * 4. / 3. * kp.user0 * tmax(n_a, n_b) * (rw2_a * sqrt(rw2_a) + rw2_b * sqrt(rw2_b));
      case KERNEL_GEOMETRIC:
      {
        const real_t res = geometric_kernel(n_a, n_b, rw2_a, rw2_b, vt_a, vt_b);
        return kp.has_multiplier ? res * kp.user0 : res;
      }
      case KERNEL_LONG:
      {
        real_t res = geometric_kernel(n_a, n_b, rw2_a, rw2_b, vt_a, vt_b);
        const real_t r_L = tmax(real_t(sqrt(rw2_a)), real_t(sqrt(rw2_b)));
        if (r_L < 50.e-6)
        {
          const real_t r_s = tmin(real_t(sqrt(rw2_a)), real_t(sqrt(rw2_b)));
          if (r_s <= 3e-6) res = 0.;
          else res *= 4.5e8 * r_L * r_L * (1. - 3e-6 / r_s);
        }
        return res;
      }
      case KERNEL_HALL: case KERNEL_HALL_DAVIS_NO_WAALS: case KERNEL_VOHL_DAVIS_NO_WAALS:
      case KERNEL_HALL_PINSKY_1000MB_GRAV: case KERNEL_HALL_PINSKY_CUMULONIMBUS: case KERNEL_HALL_PINSKY_STRATOCUMULUS:
        return interpolated_efficiency(kp, real_t(sqrt(rw2_a)), real_t(sqrt(rw2_b)))
               * geometric_kernel(n_a, n_b, rw2_a, rw2_b, vt_a, vt_b);
      default: return real_t(0);
    }
  }

  // Shima et al. (2009) sec. 5.1.3 scaling of the pair-sampling probability: coal.ipp:99-107
  template <class real_t>
  LCX_HD real_t coal_scale_factor(n_t n) { return n > 1 ? (real_t(n * (n - 1)) / 2) / (n / 2) : 0; }

  // volume factor used by the precipitation accounting: bcnd.ipp:26-46
  template <class real_t>
  LCX_HD real_t count_vol(real_t n_filtered, real_t radius_pow, real_t exponent)
  { return 4. / 3. * cst<real_t>::pi() * n_filtered * pow(radius_pow, exponent); }
}
