// A2: GridArchive.index_of as one streaming kernel (reference
// ribs/archives/_grid_archive.py:362-410). HBM-bound: B*(d*s + 4) bytes.
//
// Bit-exactness notes:
//  * (m - lb) is evaluated in the measures dtype, everything after it in fp64, because the
//    reference multiplies by an int32 `dims` array (NumPy promotes to float64);
//  * no FMA contraction: explicit __dmul_rn / __dadd_rn / __ddiv_rn;
//  * x86 cvttsd2si returns INT_MIN for out-of-range or NaN inputs, which the following
//    clip maps to cell 0 -- CUDA's saturating conversion is replaced by a range test.
#include "grid_common.cuh"

namespace qd {

template <typename T, int D>
__global__ void __launch_bounds__(256) grid_index_kernel(const T* __restrict__ meas, int64_t B,
                                                         const __grid_constant__ GridParams p,
                                                         int32_t* __restrict__ idx, uint32_t* flags) {
  bool bad = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
    idx[i] = grid_cell_of<T, D>(meas, i, p, bad);
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
