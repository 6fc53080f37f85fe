// Grid cell arithmetic shared by the stand-alone index kernel and the fused grid insert.
// Reference: ribs/archives/_grid_archive.py:362-410 (see grid_index.cu for the bit-exactness notes).
#pragma once
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
__device__ __forceinline__ int32_t grid_cell_of(const T* __restrict__ meas, int64_t i, const GridParams& p, bool& bad) {
  const int d = D > 0 ? D : p.d;
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
  return cell;
}

int fill_grid_params(GridParams& p, int d, const int32_t* dims, const double* lower, const double* interval,
                     double epsilon);

}  // namespace qd
