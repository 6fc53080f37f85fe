// Shared device/host helpers for the qd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>

#include "qd_b200.h"

namespace qd {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return QD_ERR_CUDA;
  }
  return QD_OK;
}

#define QD_CHECK_CUDA(expr)                                          \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) {                                         \
      qd::set_error("%s: %s", #expr, cudaGetErrorString(_e));        \
      return QD_ERR_CUDA;                                            \
    }                                                                \
  } while (0)

#define QD_REQUIRE(cond, msg)             \
  do {                                    \
    if (!(cond)) {                        \
      qd::set_error("%s", msg);           \
      return QD_ERR_INVALID_ARG;          \
    }                                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int grid_for(int64_t work_items, int threads, int max_ctas_per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)kNumSMs * max_ctas_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---- order-preserving 64-bit key of a finite double (-0.0 folded onto +0.0) ----------
__device__ __forceinline__ unsigned long long ord_key(double x) {
  x = __dadd_rn(x, 0.0);  // -0.0 -> +0.0 so that equal values have equal keys
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_unkey(unsigned long long k) {
  unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// ---- exact (order-independent) summation: 3-level binning ----------------------------
// A value x with |x| <= 2^e is split against fixed grids so that up to 2^b partial values
// per level add without rounding; any association order gives the same bits.
struct BinConst {
  double c0, c1, c2;
};
__host__ __device__ inline BinConst make_bins(double absmax, int64_t count) {
  int e;
  if (!(absmax > 0.0)) absmax = 1e-300;
  frexp(absmax, &e);
  e += 1;
  int b = 2;
  while (((int64_t)1 << b) < count) ++b;
  b += 1;
  BinConst k;
  int e0 = e + b - 1;
  if (e0 > 1020) {  // cannot bin without overflow: degrade to plain accumulation
    k.c0 = k.c1 = k.c2 = 0.0;
    return k;
  }
  int e1 = e0 - 54 + b;
  int e2 = e1 - 54 + b;
  k.c0 = ldexp(1.5, e0);
  k.c1 = e1 > -1000 ? ldexp(1.5, e1) : 0.0;
  k.c2 = e2 > -1000 ? ldexp(1.5, e2) : 0.0;
  return k;
}
__device__ __forceinline__ void bin_split(double x, const BinConst& k, double& h0, double& h1, double& h2) {
  if (k.c0 == 0.0) {
    h0 = x;
    h1 = 0.0;
    h2 = 0.0;
    return;
  }
  h0 = __dadd_rn(__dadd_rn(x, k.c0), -k.c0);
  double r = __dadd_rn(x, -h0);
  if (k.c1 == 0.0) {
    h1 = r;
    h2 = 0.0;
    return;
  }
  h1 = __dadd_rn(__dadd_rn(r, k.c1), -k.c1);
  double r2 = __dadd_rn(r, -h1);
  h2 = (k.c2 == 0.0) ? r2 : __dadd_rn(__dadd_rn(r2, k.c2), -k.c2);
}

__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }
__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= 3.402823466e38f; }

template <typename T>
__device__ __forceinline__ double load_as_double(const void* p, int64_t i) {
  return (double)reinterpret_cast<const T*>(p)[i];
}

}  // namespace qd
