// Variation operators of the MAP-Elites emitters:
//   GaussianEmitter.ask   ribs/emitters/_gaussian_emitter.py:139-166    x = clip(p + sigma z)
//   IsoLineEmitter.ask    ribs/emitters/_iso_line_emitter.py:156-201    x = clip(p1 + iso_sigma z + line_z (p2 - p1))
// One fused kernel: gather of the parents' rows from the archive's solution field (sample_elites picks the
// cells on the host with the archive's generator), noise scaling, line term and bounds clip, with the
// roundings of the reference: the noise is sigma * z evaluated in double and then cast to the solution dtype
// (numpy's normal(scale=...).astype), sums and products are rounded in the solution dtype, no contraction.
#include "qd_common.cuh"

namespace qd {

template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }

template <typename T>
__global__ void __launch_bounds__(256) emit_variation_kernel(const T* __restrict__ store, int64_t n, int64_t B,
                                                             const int32_t* __restrict__ pa,
                                                             const int32_t* __restrict__ pb, const T* __restrict__ x0,
                                                             const double* __restrict__ z,
                                                             const double* __restrict__ line_z,
                                                             const T* __restrict__ iso_sigma, int sigma_is_vector,
                                                             double line_sigma, const T* __restrict__ lower,
                                                             const T* __restrict__ upper, T* __restrict__ X) {
  const int64_t total = B * n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = t / n, j = t - k * n;
    const T p1 = pa ? store[(int64_t)pa[k] * n + j] : x0[j];
    const T sig = iso_sigma[sigma_is_vector ? j : 0];
    const T noise = (T)__dmul_rn((double)sig, z[t]);  // rng.normal(scale=sigma).astype(dtype)
    T x = add_rn<T>(p1, noise);
    if (line_z) {
      const T p2 = pb ? store[(int64_t)pb[k] * n + j] : x0[j];
      const T lg = (T)__dmul_rn(line_sigma, line_z[k]);
      x = add_rn<T>(x, mul_rn<T>(lg, add_rn<T>(p2, -p1)));  // elites + iso + line * (p2 - p1)
    }
    x = x < lower[j] ? lower[j] : x;  // np.clip
    x = x > upper[j] ? upper[j] : x;
    X[t] = x;
  }
}

}  // namespace qd

extern "C" int qd_emit_variation(const void* store_solution, int dtype, int64_t n, int64_t B, const int32_t* parent_a,
                                 const int32_t* parent_b, const void* x0, const double* z, const double* line_z,
                                 const void* iso_sigma, int sigma_is_vector, double line_sigma, const void* lower,
                                 const void* upper, void* X_out, void* stream) {
  using namespace qd;
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "solution dtype must be f32/f64");
  QD_REQUIRE(n >= 1 && B >= 0, "bad shape");
  if (B == 0) return QD_OK;
  QD_REQUIRE(z && iso_sigma && lower && upper && X_out, "null pointer");
  QD_REQUIRE((parent_a && store_solution) || x0, "either parents or x0 must be given");
  if (line_z) QD_REQUIRE((parent_a == nullptr) == (parent_b == nullptr), "the line operator needs two parent lists");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(B * n, 256, 8);
  if (dtype == QD_F64)
    emit_variation_kernel<double><<<grid, 256, 0, s>>>((const double*)store_solution, n, B, parent_a, parent_b,
                                                       (const double*)x0, z, line_z, (const double*)iso_sigma,
                                                       sigma_is_vector, line_sigma, (const double*)lower,
                                                       (const double*)upper, (double*)X_out);
  else
    emit_variation_kernel<float><<<grid, 256, 0, s>>>((const float*)store_solution, n, B, parent_a, parent_b,
                                                      (const float*)x0, z, line_z, (const float*)iso_sigma,
                                                      sigma_is_vector, line_sigma, (const float*)lower,
                                                      (const float*)upper, (float*)X_out);
  return check_launch("emit_variation_kernel");
}
