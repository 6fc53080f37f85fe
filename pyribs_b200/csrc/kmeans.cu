// Lloyd update step of k_means_centroids (reference ribs/archives/_cvt_archive.py:36-134 hands the samples to
// sklearn.cluster.k_means). The assignment step is the archive's own exact
// nearest-centroid search (cvt_*.cu); this file turns (sample, nearest centroid) pairs into the new centroids.
// Sums use the three-bin exact split of qd_common.cuh: every partial value adds without rounding, so the
// centroids do not depend on the order the atomics land in and a seeded run is reproducible bit for bit.
#include "qd_common.cuh"

namespace qd {

template <typename T>
__global__ void __launch_bounds__(256) kmeans_accumulate_kernel(const T* __restrict__ X, const int32_t* __restrict__ idx,
                                                                int64_t n, int d, BinConst bc, double* __restrict__ bins,
                                                                uint32_t* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = idx[i];
    atomicAdd(&count[c], 1u);
    for (int j = 0; j < d; ++j) {
      double h0, h1, h2;
      bin_split((double)X[i * d + j], bc, h0, h1, h2);
      double* b = bins + (c * d + j) * 3;
      if (h0 != 0.0) atomicAdd(b + 0, h0);
      if (h1 != 0.0) atomicAdd(b + 1, h1);
      if (h2 != 0.0) atomicAdd(b + 2, h2);
    }
  }
}

// new centroid = sum / count (an empty cluster keeps its old centroid; the host relocates it);
// shift[c] = squared distance the centroid moved
template <typename T>
__global__ void __launch_bounds__(256) kmeans_finish_kernel(const double* __restrict__ bins,
                                                            const uint32_t* __restrict__ count, int64_t k, int d,
                                                            const T* __restrict__ old_c, T* __restrict__ new_c,
                                                            double* __restrict__ sums, double* __restrict__ shift) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < k; c += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t m = count[c];
    double s2 = 0.0;
    for (int j = 0; j < d; ++j) {
      const double* b = bins + (c * d + j) * 3;
      const double s = __dadd_rn(__dadd_rn(b[0], b[1]), b[2]);
      sums[c * d + j] = s;
      const T o = old_c[c * d + j];
      const T v = m ? (T)__ddiv_rn(s, (double)m) : o;
      new_c[c * d + j] = v;
      const double df = (double)v - (double)o;
      s2 += df * df;
    }
    shift[c] = s2;
  }
}

}  // namespace qd

extern "C" size_t qd_kmeans_workspace_doubles(int64_t k, int64_t d) { return (size_t)k * d * 3; }

extern "C" int qd_kmeans_update(const void* X, int dtype, int64_t n, int64_t d, const int32_t* idx, int64_t k,
                                double absmax, double* bins, uint32_t* count, const void* old_centroids,
                                void* new_centroids, double* sums, double* shift, void* stream) {
  using namespace qd;
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be f32/f64");
  QD_REQUIRE(X && idx && bins && count && old_centroids && new_centroids && sums && shift, "null pointer");
  QD_REQUIRE(n >= 1 && d >= 1 && d <= 1024 && k >= 1, "bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  QD_CHECK_CUDA(cudaMemsetAsync(bins, 0, (size_t)k * d * 3 * sizeof(double), s));
  QD_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)k * sizeof(uint32_t), s));
  const BinConst bc = make_bins(absmax, n);
  if (dtype == QD_F64) {
    kmeans_accumulate_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((const double*)X, idx, n, (int)d, bc, bins, count);
    int rc = check_launch("kmeans_accumulate_kernel");
    if (rc) return rc;
    kmeans_finish_kernel<double><<<grid_for(k, 256), 256, 0, s>>>(bins, count, k, (int)d, (const double*)old_centroids,
                                                                  (double*)new_centroids, sums, shift);
  } else {
    kmeans_accumulate_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((const float*)X, idx, n, (int)d, bc, bins, count);
    int rc = check_launch("kmeans_accumulate_kernel");
    if (rc) return rc;
    kmeans_finish_kernel<float><<<grid_for(k, 256), 256, 0, s>>>(bins, count, k, (int)d, (const float*)old_centroids,
                                                                 (float*)new_centroids, sums, shift);
  }
  return check_launch("kmeans_finish_kernel");
}
