// A2: GridArchive.index_of as one streaming kernel (reference
// ribs/archives/_grid_archive.py:362-410). HBM-bound: B*(d*s + 4) bytes.
//
// Bit-exactness notes:
//  * (m - lb) is evaluated in the measures dtype, everything after it in fp64, because the
//    reference multiplies by an int32 `dims` array (NumPy promotes to float64);
//  * no FMA contraction: explicit __dmul_rn / __dadd_rn / __ddiv_rn;
//  * x86 cvttsd2si returns INT_MIN for out-of-range or NaN inputs, which the following
//    clip maps to cell 0 -- CUDA's saturating conversion is replaced by a range test.
#include "qd_common.cuh"

namespace qd {

struct GridParams {
  int d;
  int32_t dims[QD_MAX_MEASURE_DIM];
  int32_t stride[QD_MAX_MEASURE_DIM];
  double lower[QD_MAX_MEASURE_DIM];
  double interval[QD_MAX_MEASURE_DIM];
  double eps;
};

template <typename T>
__device__ __forceinline__ int32_t grid_coord(T m, int k, const GridParams& p) {
  T diff = m - (T)p.lower[k];  // measures dtype
  double v = __ddiv_rn(__dadd_rn(__dmul_rn((double)p.dims[k], (double)diff), p.eps), p.interval[k]);
  int32_t g;
  if (v > -2147483649.0 && v < 2147483648.0)
    g = (int32_t)v;  // truncation toward zero
  else
    g = INT32_MIN;  // x86 "integer indefinite"
  g = max(g, 0);
  g = min(g, p.dims[k] - 1);
  return g;
}

template <typename T, int D>
__global__ void __launch_bounds__(256) grid_index_kernel(const T* __restrict__ meas, int64_t B,
                                                         const __grid_constant__ GridParams p,
                                                         int32_t* __restrict__ idx, uint32_t* flags) {
  const int d = D > 0 ? D : p.d;
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t cell = 0;
    if constexpr (D == 2 && sizeof(T) == 8) {
      double2 m = reinterpret_cast<const double2*>(meas)[i];
      bad |= !(finite_d(m.x) && finite_d(m.y));
      cell = grid_coord<double>(m.x, 0, p) * p.stride[0] + grid_coord<double>(m.y, 1, p);
    } else if constexpr (D == 2 && sizeof(T) == 4) {
      float2 m = reinterpret_cast<const float2*>(meas)[i];
      bad |= !(finite_f(m.x) && finite_f(m.y));
      cell = grid_coord<float>(m.x, 0, p) * p.stride[0] + grid_coord<float>(m.y, 1, p);
    } else {
      for (int k = 0; k < d; ++k) {
        T m = meas[i * d + k];
        bad |= !(fabs((double)m) <= (sizeof(T) == 8 ? 1.7976931348623157e308 : 3.402823466e38));
        cell += grid_coord<T>(m, k, p) * p.stride[k];
      }
    }
    idx[i] = cell;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, QD_FLAG_NONFINITE_MEASURES);
}

int fill_grid_params(GridParams& p, int d, const int32_t* dims, const double* lower, const double* interval,
                     double epsilon) {
  QD_REQUIRE(d >= 1 && d <= QD_MAX_MEASURE_DIM, "measure_dim out of range [1, 32]");
  p.d = d;
  int64_t cells = 1;
  for (int k = 0; k < d; ++k) {
    QD_REQUIRE(dims[k] >= 1, "dims must be positive");
    p.dims[k] = dims[k];
    p.lower[k] = lower[k];
    p.interval[k] = interval[k];
    cells *= dims[k];
    QD_REQUIRE(cells < ((int64_t)1 << 31), "number of cells must fit int32");
  }
  int64_t s = 1;
  for (int k = d - 1; k >= 0; --k) {
    p.stride[k] = (int32_t)s;
    s *= dims[k];
  }
  p.eps = epsilon;
  return QD_OK;
}

}  // namespace qd

extern "C" int qd_grid_index_of(const void* measures, int dtype, int64_t B, int d, const int32_t* dims,
                                const double* lower, const double* interval, double epsilon, int32_t* idx_out,
                                uint32_t* flags, void* stream) {
  using namespace qd;
  GridParams p;
  int rc = fill_grid_params(p, d, dims, lower, interval, epsilon);
  if (rc != QD_OK) return rc;
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  if (B == 0) return QD_OK;
  QD_REQUIRE(measures && idx_out && flags, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  int grid = grid_for(B, 256, 8);
  if (dtype == QD_F64) {
    if (d == 2)
      grid_index_kernel<double, 2><<<grid, 256, 0, s>>>((const double*)measures, B, p, idx_out, flags);
    else
      grid_index_kernel<double, 0><<<grid, 256, 0, s>>>((const double*)measures, B, p, idx_out, flags);
  } else {
    if (d == 2)
      grid_index_kernel<float, 2><<<grid, 256, 0, s>>>((const float*)measures, B, p, idx_out, flags);
    else
      grid_index_kernel<float, 0><<<grid, 256, 0, s>>>((const float*)measures, B, p, idx_out, flags);
  }
  return check_launch("grid_index_kernel");
}
