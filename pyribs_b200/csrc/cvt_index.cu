// A3: CVTArchive.index_of (reference ribs/archives/_cvt_archive.py:579-644).
//
// method 1 (this file): exact fp64 scan, the in-library cross-check and small-problem path.
// method 0 (cvt_tc.cu): tcgen05 filter + fp64 rescoring for large B x C.
#include "cvt_common.cuh"

namespace qd {

constexpr int kScanThreads = 128;
constexpr int kScanTile = 512;  // centroids per shared-memory tile

template <typename T, int D>
__global__ void __launch_bounds__(kScanThreads) cvt_scan_fp64_kernel(const T* __restrict__ meas, int64_t B, int d_rt,
                                                                   const T* __restrict__ cent, int64_t C,
                                                                   int32_t base, int32_t* __restrict__ idx,
                                                                   double* __restrict__ dist_out, uint32_t* flags) {
  extern __shared__ double s_cent[];
  const int d = D > 0 ? D : d_rt;
  constexpr int XN = D > 0 ? D : QD_MAX_MEASURE_DIM;
  const int64_t n_qblocks = (B + kScanThreads - 1) / kScanThreads;
  bool bad = false;
  for (int64_t qb = blockIdx.x; qb < n_qblocks; qb += gridDim.x) {
    const int64_t q = qb * kScanThreads + threadIdx.x;
    double x[XN];
    if (q < B) {
#pragma unroll
      for (int k = 0; k < XN; ++k)
        if (k < d) {
          x[k] = (double)meas[q * d + k];
          bad |= !finite_d(x[k]);
        }
    }
    double best = INFINITY;
    int32_t bi = 0;
    for (int64_t t0 = 0; t0 < C; t0 += kScanTile) {
      const int tn = (int)min((int64_t)kScanTile, C - t0);
      __syncthreads();
      for (int e = threadIdx.x; e < tn * d; e += kScanThreads) s_cent[e] = (double)cent[t0 * d + e];
      __syncthreads();
      if (q < B) {
        for (int j = 0; j < tn; ++j) {
          double dd = sqdist_np<D>(x, &s_cent[j * d], d);
          if (dd < best) {  // strict: first minimum wins, as np.argmin
            best = dd;
            bi = (int32_t)(t0 + j);
          }
        }
      }
    }
    if (q < B) {
      idx[q] = bi + base;
      if (dist_out) dist_out[q] = best;
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, QD_FLAG_NONFINITE_MEASURES);
}

template <typename T>
int launch_scan(const T* meas, int64_t B, int d, const T* cent, int64_t C, int32_t base, int32_t* idx, double* dist,
                uint32_t* flags, cudaStream_t s) {
  int64_t qblocks = (B + kScanThreads - 1) / kScanThreads;
  int grid = (int)(qblocks < (int64_t)kNumSMs * 16 ? qblocks : (int64_t)kNumSMs * 16);
  size_t smem = (size_t)kScanTile * d * sizeof(double);
#define QD_SCAN(DD) \
  cvt_scan_fp64_kernel<T, DD><<<grid, kScanThreads, smem, s>>>(meas, B, d, cent, C, base, idx, dist, flags)
  switch (d) {
    case 1: QD_SCAN(1); break;
    case 2: QD_SCAN(2); break;
    case 3: QD_SCAN(3); break;
    case 4: QD_SCAN(4); break;
    case 8: QD_SCAN(8); break;
    default: {
      if (smem > 48 * 1024) {
        QD_CHECK_CUDA(cudaFuncSetAttribute(cvt_scan_fp64_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      }
      QD_SCAN(0);
    }
  }
#undef QD_SCAN
  return check_launch("cvt_scan_fp64_kernel");
}

__global__ void mask_losers_kernel(const double* __restrict__ local, const double* __restrict__ global,
                                   int32_t* __restrict__ idx, int64_t B) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
    if (local[i] != global[i]) idx[i] = INT32_MAX;
}

int cvt_tc_index_of(const void* measures, int dtype, int64_t B, int d, const void* centroids, int64_t C,
                    const void* prepared, int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags,
                    void* scratch, size_t scratch_bytes, cudaStream_t s);  // cvt_tc.cu
int cvt_bucket_index_of(const void* measures, int dtype, int64_t B, int d, int64_t C, const void* buckets,
                        int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags,
                        cudaStream_t s);  // cvt_bucket.cu
int cvt_tree_index_of(const void* measures, int dtype, int64_t B, int d, int64_t C, const void* tree,
                      int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags, void* scratch,
                      size_t scratch_bytes, cudaStream_t s);  // cvt_tree.cu

}  // namespace qd

extern "C" int qd_cvt_index_of(const void* measures, int dtype, int64_t B, int d, const void* centroids, int64_t C,
                               const void* prepared, int32_t index_base, int method, int32_t* idx_out,
                               double* dist_out, uint32_t* flags, void* scratch, size_t scratch_bytes, void* stream) {
  using namespace qd;
  QD_REQUIRE(d >= 1 && d <= QD_MAX_MEASURE_DIM, "measure_dim out of range [1, 32]");
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  QD_REQUIRE(C >= 1 && C + (int64_t)index_base < ((int64_t)1 << 31), "centroid count must be in [1, 2^31)");
  if (B == 0) return QD_OK;
  QD_REQUIRE(measures && centroids && idx_out && flags, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  if (method == 1) {
    if (dtype == QD_F64)
      return launch_scan<double>((const double*)measures, B, d, (const double*)centroids, C, index_base, idx_out,
                                 dist_out, flags, s);
    return launch_scan<float>((const float*)measures, B, d, (const float*)centroids, C, index_base, idx_out, dist_out,
                              flags, s);
  }
  if (method == 2) {
    QD_REQUIRE(prepared != nullptr, "method 2 needs the bucket grid built by qd_cvt_bucket_prepare");
    return cvt_bucket_index_of(measures, dtype, B, d, C, prepared, index_base, idx_out, dist_out, flags, s);
  }
  if (method == 3) {
    QD_REQUIRE(prepared != nullptr, "method 3 needs the box tree built by qd_cvt_tree_prepare");
    return cvt_tree_index_of(measures, dtype, B, d, C, prepared, index_base, idx_out, dist_out, flags, scratch,
                             scratch_bytes, s);
  }
  QD_REQUIRE(method == 0, "method must be 0 (tensor-core), 1 (fp64 scan), 2 (bucket grid) or 3 (box tree)");
  QD_REQUIRE(prepared != nullptr, "method 0 needs the image built by qd_cvt_prepare");
  return cvt_tc_index_of(measures, dtype, B, d, centroids, C, prepared, index_base, idx_out, dist_out, flags, scratch,
                         scratch_bytes, s);
}

extern "C" int qd_cvt_mask_losers(const double* local_dist, const double* global_min, int32_t* idx_inout, int64_t B,
                                  void* stream) {
  using namespace qd;
  if (B == 0) return QD_OK;
  mask_losers_kernel<<<grid_for(B, 256), 256, 0, (cudaStream_t)stream>>>(local_dist, global_min, idx_inout, B);
  return check_launch("mask_losers_kernel");
}
