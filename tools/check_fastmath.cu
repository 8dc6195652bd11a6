// Accuracy check of the device-only fast arithmetic used inside the condensation root solve (lcx_physics.h) against libm.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I libcloudphxx_b200/csrc -I include tools/check_fastmath.cu -o /tmp/check_fastmath
#define LCX_FAST_MATH 1
#include "lcx_physics.h"
#include <cstdio>
#include <cmath>
#include <vector>

__global__ void k(int n, const double *x, double *o_rs, double *o_seed_rs, double *o_div, double *o_seed_rcp, double *o_exp, double *o_cbrt, double *o_exp8, double *o_cbrtm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  o_rs[i] = lcx::lcx_rsqrt(v);
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); o_seed_rs[i] = r;
  o_div[i] = lcx::lcx_div(1.2345678901234567, v);
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); o_seed_rcp[i] = r;
  o_exp[i] = lcx::lcx_exp_small(fmod(v, 0.125));
  o_cbrt[i] = lcx::lcx_cbrt_ge1(1.0 + v);
  o_exp8[i] = lcx::lcx_exp_small(0.125 + fmod(v, 0.875));      // the scaled-and-squared branch
  o_cbrtm[i] = lcx::lcx_cbrt1p_mid(fmod(v, 0.5));
}

int main()
{
  const int n = 1 << 22;
  std::vector<double> x(n);
  for (int i = 0; i < n; ++i) x[i] = std::exp(-40.0 + 70.0 * (i + 0.5) / n);     // 4e-18 .. 1e13
  double *d[9];
  for (auto &p : d) cudaMalloc(&p, n * sizeof(double));
  cudaMemcpy(d[0], x.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]);
  std::vector<std::vector<double>> o(8, std::vector<double>(n));
  for (int q = 0; q < 8; ++q) cudaMemcpy(o[q].data(), d[q + 1], n * sizeof(double), cudaMemcpyDeviceToHost);
  double e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
  {
    const double ref[8] = {1.0 / std::sqrt(x[i]), 1.0 / std::sqrt(x[i]), 1.2345678901234567 / x[i], 1.0 / x[i], std::exp(std::fmod(x[i], 0.125)), std::cbrt(1.0 + x[i]), std::exp(0.125 + std::fmod(x[i], 0.875)), std::cbrt(1.0 + std::fmod(x[i], 0.5))};
    for (int q = 0; q < 8; ++q) e[q] = std::fmax(e[q], std::fabs(o[q][i] / ref[q] - 1.0));
  }
  std::printf("max rel err: rsqrt %.3g (seed %.3g)  div %.3g (rcp seed %.3g)  exp_small %.3g (x >= 1/8: %.3g)  cbrt_ge1 %.3g  cbrt1p_mid %.3g\n", e[0], e[1], e[2], e[3], e[4], e[6], e[5], e[7]);
  return (e[0] < 5e-16 && e[2] < 5e-16 && e[4] < 5e-16 && e[5] < 5e-16 && e[6] < 2.5e-15 && e[7] < 5e-16) ? 0 : 1;
}
