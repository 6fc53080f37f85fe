// A12-A14: evolution-strategy ask / tell kernels (reference ribs/emitters/opt/_cma_es.py,
// _sep_cma_es.py, _lm_ma_es.py). All strategy state is fp64 and lives in HBM; sigma is a device
// scalar so ask -> tell -> ask chains never synchronise with the host.
//
// Rooflines: sep ask/tell, LM-MA tell, weighted parent reduction are HBM streams;
// the CMA sampling / rank-mu / C^-1/2 products are fp64 GEMMs (B200 runs fp64 FMA and fp64 MMA
// at the same 40 TFLOP/s, and tcgen05 has no fp64 kind, so they are tiled FMA kernels).
#include "qd_common.cuh"

namespace qd {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based normals
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

__global__ void __launch_bounds__(256) fill_normal_kernel(double* __restrict__ out, int64_t count, uint64_t seed,
                                                          uint64_t offset) {
  const int64_t pairs = (count + 1) / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + (uint64_t)p;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x243F6A88u, 0x85A308D3u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    // two 53-bit uniforms; u1 in (0,1]
    const double u1 = ((double)(((uint64_t)(c[0] >> 5) << 26) | (c[1] >> 6)) + 1.0) * 0x1p-53;
    const double u2 = (double)(((uint64_t)(c[2] >> 5) << 26) | (c[3] >> 6)) * 0x1p-53;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    out[2 * p] = rad * cs;
    if (2 * p + 1 < count) out[2 * p + 1] = rad * sn;
  }
}

// ---------------------------------------------------------------------------------------------
// bounds test shared by the three asks (_cma_es.py:190-194)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) oob_kernel(const double* __restrict__ X, int64_t rows, int64_t n,
                                                  const int32_t* __restrict__ row_ids,
                                                  const double* __restrict__ lower, const double* __restrict__ upper,
                                                  int32_t* __restrict__ oob) {
  // one warp per row
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = w0; k < rows; k += nw) {
    const int64_t row = row_ids ? row_ids[k] : k;
    bool bad = false;
    for (int64_t j = lane; j < n; j += 32) {
      const double x = X[row * n + j];
      bad |= (x < lower[j]) || (x > upper[j]);
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) oob[k] = bad ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// A13 ask: X[row] = sqrt(C) * (sigma z) + mean      (_sep_cma_es.py:140-179)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sep_ask_kernel(const double* __restrict__ z, int64_t rows, int64_t n,
                                                      const int32_t* __restrict__ row_ids,
                                                      const double* __restrict__ mean, const double* __restrict__ C,
                                                      const double* __restrict__ sigma_p, double* __restrict__ X) {
  const double sigma = *sigma_p;
  const int64_t total = rows * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / n, j = e - k * n;
    const int64_t row = row_ids ? row_ids[k] : k;
    const double u = __dmul_rn(sigma, z[e]);  // rng.normal(0, sigma): 0 + sigma * z
    X[row * n + j] = __dadd_rn(__dmul_rn(sqrt(C[j]), u), mean[j]);
  }
}

// ---------------------------------------------------------------------------------------------
// weighted parent reduction: mean_out[j] = sum_k w[k] X[rank[k]][j]   (_cma_es.py:297)
//                            sq_out[j]   = sum_k w[k] (X[rank[k]][j] - old[j])^2  (_sep_cma_es.py:303)
// pass 1: (column block, row split) partials; pass 2: fixed-order sum of the splits.
// ---------------------------------------------------------------------------------------------
constexpr int kWsumCols = 256;
constexpr int kWsumRowsPerSplit = 128;

__global__ void __launch_bounds__(kWsumCols) wsum_partial_kernel(const double* __restrict__ X, int64_t n,
                                                                 const int64_t* __restrict__ rank,
                                                                 const double* __restrict__ w, int64_t mu,
                                                                 const double* __restrict__ old_mean,
                                                                 double* __restrict__ partial, int nsplit) {
  const int64_t j = blockIdx.x * (int64_t)kWsumCols + threadIdx.x;
  const int split = blockIdx.y;
  const int64_t k0 = (int64_t)split * kWsumRowsPerSplit;
  const int64_t k1 = min(mu, k0 + kWsumRowsPerSplit);
  __shared__ int64_t s_rank[kWsumRowsPerSplit];
  __shared__ double s_w[kWsumRowsPerSplit];
  for (int t = threadIdx.x; t < k1 - k0; t += kWsumCols) {
    s_rank[t] = rank[k0 + t];
    s_w[t] = w[k0 + t];
  }
  __syncthreads();
  if (j >= n) return;
  const double om = old_mean ? old_mean[j] : 0.0;
  double a = 0.0, b = 0.0;
  for (int t = 0; t < k1 - k0; ++t) {
    const double x = X[s_rank[t] * n + j];
    a = __dadd_rn(a, __dmul_rn(x, s_w[t]));  // parents * weights, then sum (no contraction)
    const double y = __dsub_rn(x, om);
    b = __dadd_rn(b, __dmul_rn(__dmul_rn(y, s_w[t]), y));
  }
  partial[((int64_t)split * 2 + 0) * n + j] = a;
  partial[((int64_t)split * 2 + 1) * n + j] = b;
}

__global__ void __launch_bounds__(256) wsum_final_kernel(const double* __restrict__ partial, int64_t n, int nsplit,
                                                         double* __restrict__ mean_out, double* __restrict__ sq_out) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < nsplit; ++s) {
      a = __dadd_rn(a, partial[((int64_t)s * 2 + 0) * n + j]);
      b = __dadd_rn(b, partial[((int64_t)s * 2 + 1) * n + j]);
    }
    if (mean_out) mean_out[j] = a;
    if (sq_out) sq_out[j] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// deterministic block sums of squares, used for ||ps||^2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) t += sh[w];
  __syncthreads();
  return t;  // valid on thread 0
}

struct CmaScalars {
  double cc, cs, c1, cmu, mueff, damps, wsum, hsig_exp;
};

// status block layout (device doubles): [0] sigma, [1] ||ps||^2, [2] max(C diag / eig), [3] min
constexpr int kStatusSigma = 0, kStatusSsq = 1, kStatusMax = 2, kStatusMin = 3;

// A13 tell, pass 1 (_sep_cma_es.py:281-286): y, z = y / sqrt(C), ps; per-block sum(ps^2)
__global__ void __launch_bounds__(256) sep_tell_ps_kernel(int64_t n, const double* __restrict__ new_mean,
                                                          const double* __restrict__ mean, const double* __restrict__ C,
                                                          double* __restrict__ ps, const double* __restrict__ status,
                                                          CmaScalars sc, double* __restrict__ blocksum) {
  __shared__ double sh[8];
  const double sigma = status[kStatusSigma];
  const double coef = sqrt(sc.cs * (2.0 - sc.cs) * sc.mueff) / sigma;
  double acc = 0.0;
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) {
    const double y = __dsub_rn(new_mean[j], mean[j]);
    const double zz = __dmul_rn(1.0 / sqrt(C[j]), y);
    const double p = __dadd_rn(__dmul_rn(1.0 - sc.cs, ps[j]), __dmul_rn(coef, zz));
    ps[j] = p;
    acc = p * p;
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
}

// A13 tell, pass 2 (_sep_cma_es.py:288-311): hsig, pc, C, mean, sigma
__global__ void __launch_bounds__(256) sep_tell_update_kernel(int64_t n, const double* __restrict__ new_mean,
                                                              const double* __restrict__ rank_mu,
                                                              double* __restrict__ mean, double* __restrict__ pc,
                                                              double* __restrict__ C, double* __restrict__ status,
                                                              CmaScalars sc, const double* __restrict__ blocksum,
                                                              int nblocks, double* __restrict__ minmax) {
  __shared__ double sh[8];
  __shared__ double s_ssq;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += blocksum[b];  // same order in every block
    s_ssq = s;
  }
  __syncthreads();
  const double ssq = s_ssq;
  const double sigma = status[kStatusSigma];
  const double left = ssq / (double)n / (1.0 - pow(1.0 - sc.cs, sc.hsig_exp));
  const double right = 2.0 + 4.0 / ((double)n + 1.0);
  const double hsig = left < right ? 1.0 : 0.0;
  const double c1a = sc.c1 * (1.0 - (1.0 - hsig * hsig) * sc.cc * (2.0 - sc.cc));
  const double pc_coef = hsig * sqrt(sc.cc * (2.0 - sc.cc) * sc.mueff);
  const double decay = 1.0 - c1a - sc.cmu * sc.wsum;
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double cmax = -INFINITY, cmin = INFINITY;
  if (j < n) {
    const double y = __dsub_rn(new_mean[j], mean[j]);
    const double p = __dadd_rn(__dmul_rn(1.0 - sc.cc, pc[j]), __dmul_rn(pc_coef, y));
    pc[j] = p;
    // cov*(1-c1a-cmu*sum w) + (c1*pc^2)*c1 + rank_mu*cmu/sigma^2      (:246-255)
    const double t1 = __dmul_rn(C[j], decay);
    const double t2 = __dmul_rn(__dmul_rn(sc.c1, __dmul_rn(p, p)), sc.c1);
    const double t3 = __ddiv_rn(__dmul_rn(rank_mu[j], sc.cmu), __dmul_rn(sigma, sigma));
    const double cn = __dadd_rn(__dadd_rn(t1, t2), t3);
    C[j] = cn;
    mean[j] = new_mean[j];
    cmax = cn;
    cmin = cn;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cmax = fmax(cmax, __shfl_down_sync(0xffffffffu, cmax, o));
    cmin = fmin(cmin, __shfl_down_sync(0xffffffffu, cmin, o));
  }
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = cmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = -INFINITY;
    for (int w = 0; w < 8; ++w) m = fmax(m, sh[w]);
    minmax[2 * blockIdx.x] = m;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cmin;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = INFINITY;
    for (int w = 0; w < 8; ++w) m = fmin(m, sh[w]);
    minmax[2 * blockIdx.x + 1] = m;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double cn = sc.cs / sc.damps;
    status[kStatusSsq] = ssq;
    status[kStatusSigma] = sigma * exp(fmin(1.0, cn * (ssq / (double)n - 1.0) / 2.0));  // :310-311
  }
}

__global__ void minmax_final_kernel(const double* __restrict__ minmax, int nblocks, double* __restrict__ status) {
  double mx = -INFINITY, mn = INFINITY;
  for (int b = 0; b < nblocks; ++b) {
    mx = fmax(mx, minmax[2 * b]);
    mn = fmin(mn, minmax[2 * b + 1]);
  }
  status[kStatusMax] = mx;
  status[kStatusMin] = mn;
}

// ---------------------------------------------------------------------------------------------
// fp64 GEMM building block: out[i][j] = sum_k A(i,k) * B(j,k)  ("NT"), 64x64x16 tiles, 4x4/thread
// A and B are supplied as functors so gathers / scalings fuse into the tile loads.
// ---------------------------------------------------------------------------------------------
template <class AF, class BF, class EF>
__global__ void __launch_bounds__(256) gemm_nt_kernel(int64_t M, int64_t N, int64_t K, AF a_of, BF b_of, EF epilogue) {
  __shared__ double sA[16][64 + 1];
  __shared__ double sB[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  double acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, kk = e & 15;
      sA[kk][r] = (i0 + r < M && k0 + kk < K) ? a_of(i0 + r, k0 + kk) : 0.0;
      sB[kk][r] = (j0 + r < N && k0 + kk < K) ? b_of(j0 + r, k0 + kk) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = sA[kk][ty * 4 + u];
        b[u] = sB[kk][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < M && j < N) epilogue(i, j, acc[u][v]);
    }
}

template <class AF, class BF, class EF>
int launch_gemm_nt(int64_t M, int64_t N, int64_t K, AF a, BF b, EF e, cudaStream_t s, const char* what) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  gemm_nt_kernel<<<grid, 256, 0, s>>>(M, N, K, a, b, e);
  return check_launch(what);
}

// T = B * sqrt(eig)   (_cma_es.py:206);  also eig max/min into the status block
__global__ void cma_transform_kernel(int64_t n, const double* __restrict__ Bm, const double* __restrict__ eig,
                                     double* __restrict__ T) {
  const int64_t total = n * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    T[e] = __dmul_rn(Bm[e], sqrt(eig[e % n]));
}

// A12 tell vector part 1: y, z = invsqrt @ y (GEMV, one warp per row), ps, block sums
__global__ void __launch_bounds__(256) cma_tell_ps_kernel(int64_t n, const double* __restrict__ new_mean,
                                                          const double* __restrict__ mean,
                                                          const double* __restrict__ invsqrt, double* __restrict__ ps,
                                                          const double* __restrict__ status, CmaScalars sc,
                                                          double* __restrict__ rowsq) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double acc = 0.0;
  for (int64_t k = lane; k < n; k += 32) acc = fma(invsqrt[row * n + k], __dsub_rn(new_mean[k], mean[k]), acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double sigma = status[kStatusSigma];
    const double coef = sqrt(sc.cs * (2.0 - sc.cs) * sc.mueff) / sigma;
    const double p = (1.0 - sc.cs) * ps[row] + coef * acc;
    ps[row] = p;
    rowsq[row] = p * p;
  }
}

// A12 tell vector part 2: ssq, hsig, pc; writes the scalars the GEMM epilogue needs
// scal[0] = decay, scal[1] = c1*c1, scal[2] = cmu / sigma^2
__global__ void __launch_bounds__(256) cma_tell_pc_kernel(int64_t n, const double* __restrict__ new_mean,
                                                          const double* __restrict__ mean, double* __restrict__ pc,
                                                          double* __restrict__ status, CmaScalars sc,
                                                          const double* __restrict__ rowsq, double* __restrict__ scal) {
  __shared__ double s_ssq;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int64_t r = 0; r < n; ++r) s += rowsq[r];
    s_ssq = s;
  }
  __syncthreads();
  const double ssq = s_ssq;
  const double sigma = status[kStatusSigma];
  const double left = ssq / (double)n / (1.0 - pow(1.0 - sc.cs, sc.hsig_exp));
  const double hsig = left < 2.0 + 4.0 / ((double)n + 1.0) ? 1.0 : 0.0;
  const double c1a = sc.c1 * (1.0 - (1.0 - hsig * hsig) * sc.cc * (2.0 - sc.cc));
  const double pc_coef = hsig * sqrt(sc.cc * (2.0 - sc.cc) * sc.mueff);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    pc[j] = (1.0 - sc.cc) * pc[j] + pc_coef * __dsub_rn(new_mean[j], mean[j]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[0] = 1.0 - c1a - sc.cmu * sc.wsum;
    scal[1] = sc.c1 * sc.c1;
    scal[2] = sc.cmu / (sigma * sigma);
    scal[3] = sigma * exp(fmin(1.0, (sc.cs / sc.damps) * (ssq / (double)n - 1.0) / 2.0));  // new sigma
    status[kStatusSsq] = ssq;
  }
}

__global__ void cma_tell_finish_kernel(int64_t n, const double* __restrict__ new_mean, double* __restrict__ mean,
                                       double* __restrict__ status, const double* __restrict__ scal) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    mean[j] = new_mean[j];
  if (blockIdx.x == 0 && threadIdx.x == 0) status[kStatusSigma] = scal[3];
}

__global__ void sym_max_kernel(int64_t n, double* __restrict__ A) {  // A = max(A, A^T)  (_cma_es.py:67,80)
  const int64_t total = n * n;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n, j = e - i * n;
    if (i < j) {
      const double m = fmax(A[i * n + j], A[j * n + i]);
      A[i * n + j] = m;
      A[j * n + i] = m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// A14: LM-MA-ES
// ---------------------------------------------------------------------------------------------
// ask (_lm_ma_es.py:126-152): es_stream.cu (tall-skinny products + forward substitution)

// tell (_lm_ma_es.py:224-239): ps, M rows, mean, sigma
__global__ void __launch_bounds__(256) lmma_tell_ps_kernel(int64_t n, const double* __restrict__ z_mean,
                                                           double* __restrict__ ps, double csigma, double mueff,
                                                           double* __restrict__ blocksum) {
  __shared__ double sh[8];
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double acc = 0.0;
  if (j < n) {
    const double p = __dadd_rn(__dmul_rn(1.0 - csigma, ps[j]), __dmul_rn(sqrt(mueff * csigma * (2.0 - csigma)), z_mean[j]));
    ps[j] = p;
    acc = p * p;
  }
  const double t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
}

// m_i <- (1 - cc_i) m_i + sqrt(mueff cc_i (2 - cc_i)) z_mean for every direction vector (_lm_ma_es.py:217-225): a
// stream over n_vectors x n (654 MB read + written at the default n_vectors = batch size). A CTA owns a column strip
// of 256 threads (VEC2: two columns each, 16-byte accesses) and a range of vectors whose two coefficients it
// computes once, into shared memory; z_mean stays in registers.
constexpr int kLmTellVecs = 128;  // vectors per CTA at most

template <bool VEC2>
__global__ void __launch_bounds__(256) lmma_tell_m_kernel(int64_t n, int64_t n_vectors, int64_t vecs_per_cta,
                                                          const double* __restrict__ z_mean,
                                                          const double* __restrict__ new_mean, double* __restrict__ mean,
                                                          double* __restrict__ Mv, const double* __restrict__ cc,
                                                          double csigma, double mueff, double* __restrict__ status,
                                                          const double* __restrict__ blocksum, int nblocks) {
  __shared__ double keep[kLmTellVecs], gain[kLmTellVecs];
  const int64_t i0 = (int64_t)blockIdx.y * vecs_per_cta;
  const int nv = (int)(n_vectors - i0 < vecs_per_cta ? n_vectors - i0 : vecs_per_cta);
  for (int v = threadIdx.x; v < nv; v += 256) {
    const double c = cc[i0 + v];
    keep[v] = 1.0 - c;
    gain[v] = sqrt(mueff * c * (2.0 - c));
  }
  __syncthreads();
  const int64_t j = (blockIdx.x * (int64_t)256 + threadIdx.x) * (VEC2 ? 2 : 1);
  if (j < n) {
    if (VEC2) {  // n even: every pair is whole and 16-byte aligned
      const double2 zm = *reinterpret_cast<const double2*>(z_mean + j);
      double2* col = reinterpret_cast<double2*>(Mv + i0 * n + j);
#pragma unroll 4
      for (int v = 0; v < nv; ++v) {
        double2 mv = col[(int64_t)v * (n / 2)];
        mv.x = __dadd_rn(__dmul_rn(keep[v], mv.x), __dmul_rn(gain[v], zm.x));
        mv.y = __dadd_rn(__dmul_rn(keep[v], mv.y), __dmul_rn(gain[v], zm.y));
        col[(int64_t)v * (n / 2)] = mv;
      }
      if (blockIdx.y == 0) *reinterpret_cast<double2*>(mean + j) = *reinterpret_cast<const double2*>(new_mean + j);
    } else {
      const double zm = z_mean[j];
      double* col = Mv + i0 * n + j;
#pragma unroll 4
      for (int v = 0; v < nv; ++v) col[(int64_t)v * n] = __dadd_rn(__dmul_rn(keep[v], col[(int64_t)v * n]), __dmul_rn(gain[v], zm));
      if (blockIdx.y == 0) mean[j] = new_mean[j];
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    double ssq = 0.0;
    for (int b = 0; b < nblocks; ++b) ssq += blocksum[b];
    status[kStatusSsq] = ssq;
    status[kStatusSigma] = status[kStatusSigma] * exp(csigma / 2.0 * (ssq / (double)n - 1.0));
  }
}

static CmaScalars to_dev(const qd_cma_scalars* s) {
  CmaScalars c;
  c.cc = s->cc;
  c.cs = s->cs;
  c.c1 = s->c1;
  c.cmu = s->cmu;
  c.mueff = s->mueff;
  c.damps = s->damps;
  c.wsum = s->wsum;
  c.hsig_exp = s->hsig_exp;
  return c;
}


// ---------------------------------------------------------------------------------------------
// CMA-ES pool: E strategies of identical (n, lambda) stacked in HBM, every kernel batched over the
// emitters through blockIdx.z (the many-emitter regime of configs[0] and
// configs[2], where a launch sequence per emitter is pure launch latency). The arithmetic of each
// kernel repeats its single-strategy counterpart above operation for operation, so a pooled strategy
// and a stand-alone one fed the same normals stay bit-identical (tests/test_gpu_es.py).
// ---------------------------------------------------------------------------------------------
struct PoolView {
  int64_t E, n, lam;
  double *mean, *pc, *ps, *C, *B, *eig, *T, *invsqrt, *status, *new_mean, *ws, *X, *z;
  int64_t ws_stride;
};

__device__ __forceinline__ int pool_slot(const int32_t* ids) { return ids ? ids[blockIdx.z] : (int)blockIdx.z; }

__global__ void __launch_bounds__(256) pool_fill_normal_kernel(PoolView pv, const uint64_t* __restrict__ seeds,
                                                               const uint64_t* __restrict__ offsets) {
  const int e = blockIdx.z;
  const int64_t count = pv.lam * pv.n;
  double* out = pv.z + (int64_t)e * count;
  const uint64_t seed = seeds[e], offset = offsets[e];
  const int64_t pairs = (count + 1) / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + (uint64_t)p;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x243F6A88u, 0x85A308D3u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    const double u1 = ((double)(((uint64_t)(c[0] >> 5) << 26) | (c[1] >> 6)) + 1.0) * 0x1p-53;
    const double u2 = (double)(((uint64_t)(c[2] >> 5) << 26) | (c[3] >> 6)) * 0x1p-53;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    out[2 * p] = rad * cs;
    if (2 * p + 1 < count) out[2 * p + 1] = rad * sn;
  }
}

// batched "NT" GEMM: the tile code of gemm_nt_kernel with an emitter index threaded through the functors
template <class KF, class AF, class BF, class EF>
__global__ void __launch_bounds__(256) gemm_nt_pool_kernel(int64_t M, int64_t N, const int32_t* __restrict__ ids,
                                                           KF k_of, AF a_of, BF b_of, EF epilogue) {
  __shared__ double sA[16][64 + 1];
  __shared__ double sB[16][64 + 1];
  const int e = pool_slot(ids);
  const int64_t K = k_of(e);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
  double acc[4][4] = {};
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    for (int t = threadIdx.x; t < 64 * 16; t += 256) {
      const int r = t >> 4, kk = t & 15;
      sA[kk][r] = (i0 + r < M && k0 + kk < K) ? a_of(e, i0 + r, k0 + kk) : 0.0;
      sB[kk][r] = (j0 + r < N && k0 + kk < K) ? b_of(e, j0 + r, k0 + kk) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = sA[kk][ty * 4 + u];
        b[u] = sB[kk][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int64_t i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < M && j < N) epilogue(e, i, j, acc[u][v]);
    }
}

template <class KF, class AF, class BF, class EF>
int launch_gemm_nt_pool(int64_t M, int64_t N, int batch, const int32_t* ids, KF kf, AF a, BF b, EF ep, cudaStream_t s,
                        const char* what) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)batch);
  gemm_nt_pool_kernel<<<grid, 256, 0, s>>>(M, N, ids, kf, a, b, ep);
  return check_launch(what);
}

__global__ void pool_sym_max_kernel(PoolView pv, const int32_t* __restrict__ ids, int which) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n;
  double* A = (which == 0 ? pv.C : pv.invsqrt) + (int64_t)e * n * n;
  const int64_t total = n * n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / n, j = t - i * n;
    if (i < j) {
      const double m = fmax(A[i * n + j], A[j * n + i]);
      A[i * n + j] = m;
      A[j * n + i] = m;
    }
  }
}

// |eig|, T = B * sqrt(eig), eigenvalue extrema into the status block
__global__ void __launch_bounds__(256) pool_transform_kernel(PoolView pv, const int32_t* __restrict__ ids) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n;
  double* eig = pv.eig + (int64_t)e * n;
  const double* Bm = pv.B + (int64_t)e * n * n;
  double* T = pv.T + (int64_t)e * n * n;
  __shared__ double sh_mx[8], sh_mn[8];
  double mx = -INFINITY, mn = INFINITY;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = fabs(eig[j]);
    mx = fmax(mx, v);
    mn = fmin(mn, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    sh_mx[threadIdx.x >> 5] = mx;
    sh_mn[threadIdx.x >> 5] = mn;
  }
  __syncthreads();
  const int64_t total = n * n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
    T[t] = __dmul_rn(Bm[t], sqrt(fabs(eig[t % n])));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      mx = fmax(mx, sh_mx[w]);
      mn = fmin(mn, sh_mn[w]);
    }
    pv.status[(int64_t)e * 8 + kStatusMax] = mx;
    pv.status[(int64_t)e * 8 + kStatusMin] = mn;
  }
}

__global__ void pool_abs_eig_kernel(PoolView pv, const int32_t* __restrict__ ids) {
  const int e = pool_slot(ids);
  double* eig = pv.eig + (int64_t)e * pv.n;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < pv.n; j += (int64_t)gridDim.x * blockDim.x)
    eig[j] = fabs(eig[j]);
}

// weighted parent mean: the split-then-sum order of wsum_partial_kernel + wsum_final_kernel in one kernel
__global__ void __launch_bounds__(kWsumCols) pool_wsum_kernel(PoolView pv, const int32_t* __restrict__ ids,
                                                              const int64_t* __restrict__ rank,
                                                              const double* __restrict__ w,
                                                              const int64_t* __restrict__ mu_arr) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n, lam = pv.lam, mu = mu_arr[e];
  const double* X = pv.X + (int64_t)e * lam * n;
  const int64_t* rk = rank + (int64_t)e * lam;
  const double* we = w + (int64_t)e * lam;
  const int64_t j = blockIdx.x * (int64_t)kWsumCols + threadIdx.x;
  __shared__ int64_t s_rank[kWsumRowsPerSplit];
  __shared__ double s_w[kWsumRowsPerSplit];
  double total = 0.0;
  for (int64_t k0 = 0; k0 < mu; k0 += kWsumRowsPerSplit) {
    const int64_t k1 = min(mu, k0 + kWsumRowsPerSplit);
    __syncthreads();
    for (int t = threadIdx.x; t < k1 - k0; t += kWsumCols) {
      s_rank[t] = rk[k0 + t];
      s_w[t] = we[k0 + t];
    }
    __syncthreads();
    if (j < n) {
      double a = 0.0;
      for (int t = 0; t < k1 - k0; ++t) a = __dadd_rn(a, __dmul_rn(X[s_rank[t] * n + j], s_w[t]));
      total = __dadd_rn(total, a);
    }
  }
  if (j < n) pv.new_mean[(int64_t)e * n + j] = total;
}

__global__ void __launch_bounds__(256) pool_tell_ps_kernel(PoolView pv, const int32_t* __restrict__ ids,
                                                           const CmaScalars* __restrict__ sc_arr) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n;
  const CmaScalars sc = sc_arr[e];
  const double* new_mean = pv.new_mean + (int64_t)e * n;
  const double* mean = pv.mean + (int64_t)e * n;
  const double* invsqrt = pv.invsqrt + (int64_t)e * n * n;
  double* ps = pv.ps + (int64_t)e * n;
  double* rowsq = pv.ws + (int64_t)e * pv.ws_stride;
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= n) return;
  double acc = 0.0;
  for (int64_t k = lane; k < n; k += 32) acc = fma(invsqrt[row * n + k], __dsub_rn(new_mean[k], mean[k]), acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const double sigma = pv.status[(int64_t)e * 8 + kStatusSigma];
    const double coef = sqrt(sc.cs * (2.0 - sc.cs) * sc.mueff) / sigma;
    const double p = (1.0 - sc.cs) * ps[row] + coef * acc;
    ps[row] = p;
    rowsq[row] = p * p;
  }
}

__global__ void __launch_bounds__(256) pool_tell_pc_kernel(PoolView pv, const int32_t* __restrict__ ids,
                                                           const CmaScalars* __restrict__ sc_arr) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n;
  const CmaScalars sc = sc_arr[e];
  const double* new_mean = pv.new_mean + (int64_t)e * n;
  const double* mean = pv.mean + (int64_t)e * n;
  double* pc = pv.pc + (int64_t)e * n;
  double* status = pv.status + (int64_t)e * 8;
  const double* rowsq = pv.ws + (int64_t)e * pv.ws_stride;
  double* scal = pv.ws + (int64_t)e * pv.ws_stride + n;
  __shared__ double s_ssq;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int64_t r = 0; r < n; ++r) s += rowsq[r];
    s_ssq = s;
  }
  __syncthreads();
  const double ssq = s_ssq;
  const double sigma = status[kStatusSigma];
  const double left = ssq / (double)n / (1.0 - pow(1.0 - sc.cs, sc.hsig_exp));
  const double hsig = left < 2.0 + 4.0 / ((double)n + 1.0) ? 1.0 : 0.0;
  const double c1a = sc.c1 * (1.0 - (1.0 - hsig * hsig) * sc.cc * (2.0 - sc.cc));
  const double pc_coef = hsig * sqrt(sc.cc * (2.0 - sc.cc) * sc.mueff);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    pc[j] = (1.0 - sc.cc) * pc[j] + pc_coef * __dsub_rn(new_mean[j], mean[j]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal[0] = 1.0 - c1a - sc.cmu * sc.wsum;
    scal[1] = sc.c1 * sc.c1;
    scal[2] = sc.cmu / (sigma * sigma);
    scal[3] = sigma * exp(fmin(1.0, (sc.cs / sc.damps) * (ssq / (double)n - 1.0) / 2.0));  // new sigma
    status[kStatusSsq] = ssq;
  }
}

__global__ void pool_tell_finish_kernel(PoolView pv, const int32_t* __restrict__ ids) {
  const int e = pool_slot(ids);
  const int64_t n = pv.n;
  const double* new_mean = pv.new_mean + (int64_t)e * n;
  double* mean = pv.mean + (int64_t)e * n;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    mean[j] = new_mean[j];
  if (blockIdx.x == 0 && threadIdx.x == 0)
    pv.status[(int64_t)e * 8 + kStatusSigma] = pv.ws[(int64_t)e * pv.ws_stride + n + 3];
}

static int to_view(const qd_cma_pool* p, PoolView& v) {
  QD_REQUIRE(p, "null pool");
  QD_REQUIRE(p->E >= 1 && p->n >= 1 && p->lam >= 1, "bad pool shape");
  QD_REQUIRE(p->mean && p->pc && p->ps && p->C && p->B && p->eig && p->T && p->invsqrt && p->status && p->new_mean &&
                 p->ws && p->X && p->z,
             "null pool pointer");
  QD_REQUIRE(p->ws_stride >= p->n + 8, "pool workspace stride must be >= n + 8 doubles");
  v.E = p->E;
  v.n = p->n;
  v.lam = p->lam;
  v.mean = p->mean;
  v.pc = p->pc;
  v.ps = p->ps;
  v.C = p->C;
  v.B = p->B;
  v.eig = p->eig;
  v.T = p->T;
  v.invsqrt = p->invsqrt;
  v.status = p->status;
  v.new_mean = p->new_mean;
  v.ws = p->ws;
  v.X = p->X;
  v.z = p->z;
  v.ws_stride = p->ws_stride;
  return QD_OK;
}

}  // namespace qd

using namespace qd;

int qd_sym_eigh_indexed(int64_t n, int64_t batch, const int32_t* ids, const double* A, double* eig, double* V,
                        cudaStream_t stream);  // eigh.cu

extern "C" int qd_cma_pool_refresh(const qd_cma_pool* pool, const int32_t* ids, int n_ids, void* stream) {
  PoolView pv;
  int rc = to_view(pool, pv);
  if (rc) return rc;
  if (n_ids == 0) return QD_OK;
  QD_REQUIRE(ids && n_ids > 0 && n_ids <= pool->E, "bad ids");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = pv.n;
  const dim3 gnn((unsigned)grid_for(n * n, 256), 1, (unsigned)n_ids);
  pool_sym_max_kernel<<<gnn, 256, 0, s>>>(pv, ids, 0);  // C = max(C, C^T)   (_cma_es.py:67)
  if ((rc = check_launch("pool_sym_max_kernel"))) return rc;
  if ((rc = qd_sym_eigh_indexed(n, n_ids, ids, pv.C, pv.eig, pv.B, s))) return rc;
  pool_transform_kernel<<<gnn, 256, 0, s>>>(pv, ids);
  if ((rc = check_launch("pool_transform_kernel"))) return rc;
  pool_abs_eig_kernel<<<dim3((unsigned)grid_for(n, 256), 1, (unsigned)n_ids), 256, 0, s>>>(pv, ids);
  if ((rc = check_launch("pool_abs_eig_kernel"))) return rc;
  // invsqrt = (B * (1/sqrt(eig))) @ B^T, then max with its transpose   (_cma_es.py:76-80)
  const double* Bm = pv.B;
  const double* eig = pv.eig;
  double* inv = pv.invsqrt;
  auto kf = [=] __device__(int) { return n; };
  auto a_of = [=] __device__(int e, int64_t i, int64_t k) {
    return __dmul_rn(Bm[(int64_t)e * n * n + i * n + k], 1.0 / sqrt(eig[(int64_t)e * n + k]));
  };
  auto b_of = [=] __device__(int e, int64_t j, int64_t k) { return Bm[(int64_t)e * n * n + j * n + k]; };
  auto epi = [=] __device__(int e, int64_t i, int64_t j, double v) { inv[(int64_t)e * n * n + i * n + j] = v; };
  if ((rc = launch_gemm_nt_pool(n, n, n_ids, ids, kf, a_of, b_of, epi, s, "pool_invsqrt_gemm"))) return rc;
  pool_sym_max_kernel<<<gnn, 256, 0, s>>>(pv, ids, 1);
  return check_launch("pool_sym_max_kernel");
}

extern "C" int qd_cma_pool_ask(const qd_cma_pool* pool, const uint64_t* seeds, const uint64_t* offsets, void* stream) {
  PoolView pv;
  int rc = to_view(pool, pv);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = pv.n, lam = pv.lam;
  if (seeds) {  // NULL: the caller filled pool->z itself (parity runs)
    QD_REQUIRE(offsets, "null offsets");
    const dim3 g((unsigned)grid_for((lam * n + 1) / 2, 256, 2), 1, (unsigned)pv.E);
    pool_fill_normal_kernel<<<g, 256, 0, s>>>(pv, seeds, offsets);
    if ((rc = check_launch("pool_fill_normal_kernel"))) return rc;
  }
  // X[e][k][i] = mean[e][i] + sum_j T[e][i][j] * (sigma_e z[e][k][j])      (_cma_es.py:181-189)
  const double *z = pv.z, *T = pv.T, *status = pv.status, *mean = pv.mean;
  double* X = pv.X;
  auto kf = [=] __device__(int) { return n; };
  auto a_of = [=] __device__(int e, int64_t k, int64_t j) {
    return __dmul_rn(status[(int64_t)e * 8 + kStatusSigma], z[((int64_t)e * lam + k) * n + j]);
  };
  auto b_of = [=] __device__(int e, int64_t i, int64_t j) { return T[(int64_t)e * n * n + i * n + j]; };
  auto epi = [=] __device__(int e, int64_t k, int64_t i, double v) {
    X[((int64_t)e * lam + k) * n + i] = v + mean[(int64_t)e * n + i];
  };
  return launch_gemm_nt_pool(lam, n, (int)pv.E, nullptr, kf, a_of, b_of, epi, s, "pool_ask_gemm");
}

extern "C" int qd_cma_pool_tell(const qd_cma_pool* pool, const int32_t* ids, int n_ids, const int64_t* rank,
                                const double* w, const int64_t* mu, const qd_cma_scalars* sc, void* stream) {
  PoolView pv;
  int rc = to_view(pool, pv);
  if (rc) return rc;
  if (n_ids == 0) return QD_OK;
  QD_REQUIRE(ids && n_ids > 0 && n_ids <= pool->E && rank && w && mu && sc, "null pointer");
  static_assert(sizeof(qd_cma_scalars) == sizeof(CmaScalars), "scalar block layouts differ");
  const CmaScalars* sca = reinterpret_cast<const CmaScalars*>(sc);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = pv.n, lam = pv.lam;
  pool_wsum_kernel<<<dim3((unsigned)((n + kWsumCols - 1) / kWsumCols), 1, (unsigned)n_ids), kWsumCols, 0, s>>>(
      pv, ids, rank, w, mu);
  if ((rc = check_launch("pool_wsum_kernel"))) return rc;
  pool_tell_ps_kernel<<<dim3((unsigned)((n * 32 + 255) / 256), 1, (unsigned)n_ids), 256, 0, s>>>(pv, ids, sca);
  if ((rc = check_launch("pool_tell_ps_kernel"))) return rc;
  pool_tell_pc_kernel<<<dim3((unsigned)grid_for(n, 256), 1, (unsigned)n_ids), 256, 0, s>>>(pv, ids, sca);
  if ((rc = check_launch("pool_tell_pc_kernel"))) return rc;
  // rank-mu, in place: C[i][j] = C[i][j]*decay + c1^2 pc_i pc_j + (cmu/sigma^2) sum_k w_k ys_ki ys_kj   (:261-270, :318-323)
  const double *X = pv.X, *old_mean = pv.mean, *pc = pv.pc, *ws = pv.ws;
  double* Cm = pv.C;
  const int64_t wss = pv.ws_stride;
  auto kf = [=] __device__(int e) { return mu[e]; };
  auto a_of = [=] __device__(int e, int64_t i, int64_t k) {
    return (X[((int64_t)e * lam + rank[(int64_t)e * lam + k]) * n + i] - old_mean[(int64_t)e * n + i]) *
           w[(int64_t)e * lam + k];
  };
  auto b_of = [=] __device__(int e, int64_t j, int64_t k) {
    return X[((int64_t)e * lam + rank[(int64_t)e * lam + k]) * n + j] - old_mean[(int64_t)e * n + j];
  };
  auto epi = [=] __device__(int e, int64_t i, int64_t j, double v) {
    const double* scal = ws + (int64_t)e * wss + n;
    double* c = Cm + (int64_t)e * n * n + i * n + j;
    *c = *c * scal[0] + (pc[(int64_t)e * n + i] * pc[(int64_t)e * n + j]) * scal[1] + v * scal[2];
  };
  if ((rc = launch_gemm_nt_pool(n, n, n_ids, ids, kf, a_of, b_of, epi, s, "pool_rank_mu_gemm"))) return rc;
  pool_tell_finish_kernel<<<dim3((unsigned)grid_for(n, 256), 1, (unsigned)n_ids), 256, 0, s>>>(pv, ids);
  return check_launch("pool_tell_finish_kernel");
}

namespace qd {
}  // namespace qd

using namespace qd;

extern "C" int qd_fill_normal(double* out, int64_t count, uint64_t seed, uint64_t offset, void* stream) {
  if (count == 0) return QD_OK;
  QD_REQUIRE(out, "null pointer");
  fill_normal_kernel<<<grid_for((count + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(out, count, seed, offset);
  return check_launch("fill_normal_kernel");
}

extern "C" int qd_es_out_of_bounds(const double* X, int64_t rows, int64_t n, const int32_t* row_ids,
                                   const double* lower, const double* upper, int32_t* oob, void* stream) {
  if (rows == 0) return QD_OK;
  QD_REQUIRE(X && lower && upper && oob, "null pointer");
  oob_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(X, rows, n, row_ids, lower, upper, oob);
  return check_launch("oob_kernel");
}

extern "C" int qd_sep_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                          const double* C, const double* status, double* X, void* stream) {
  if (rows == 0) return QD_OK;
  QD_REQUIRE(z && mean && C && status && X, "null pointer");
  sep_ask_kernel<<<grid_for(rows * n, 256), 256, 0, (cudaStream_t)stream>>>(z, rows, n, row_ids, mean, C, status, X);
  return check_launch("sep_ask_kernel");
}

extern "C" size_t qd_wsum_partial_bytes(int64_t mu, int64_t n) {
  const int64_t nsplit = (mu + kWsumRowsPerSplit - 1) / kWsumRowsPerSplit;
  return (size_t)(nsplit > 0 ? nsplit : 1) * 2 * n * sizeof(double);
}

extern "C" int qd_weighted_rows(const double* X, int64_t n, const int64_t* rank, const double* w, int64_t mu,
                                const double* old_mean, double* out_mean, double* out_sq, void* partial, void* stream) {
  QD_REQUIRE(X && rank && w && partial && mu >= 1 && n >= 1, "bad arguments");
  const int nsplit = (int)((mu + kWsumRowsPerSplit - 1) / kWsumRowsPerSplit);
  dim3 grid((unsigned)((n + kWsumCols - 1) / kWsumCols), (unsigned)nsplit);
  cudaStream_t s = (cudaStream_t)stream;
  wsum_partial_kernel<<<grid, kWsumCols, 0, s>>>(X, n, rank, w, mu, old_mean, (double*)partial, nsplit);
  int rc = check_launch("wsum_partial_kernel");
  if (rc) return rc;
  wsum_final_kernel<<<grid_for(n, 256), 256, 0, s>>>((const double*)partial, n, nsplit, out_mean, out_sq);
  return check_launch("wsum_final_kernel");
}

extern "C" size_t qd_es_workspace_doubles(int64_t n) { return (size_t)(4 * ((n + 255) / 256) + n + 64); }

extern "C" int qd_sep_tell(int64_t n, const double* new_mean, const double* rank_mu, double* mean, double* pc,
                           double* ps, double* C, double* status, const qd_cma_scalars* sc, double* ws, void* stream) {
  QD_REQUIRE(new_mean && rank_mu && mean && pc && ps && C && status && sc && ws, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (int)((n + 255) / 256);
  CmaScalars c = to_dev(sc);
  sep_tell_ps_kernel<<<nb, 256, 0, s>>>(n, new_mean, mean, C, ps, status, c, ws);
  int rc = check_launch("sep_tell_ps_kernel");
  if (rc) return rc;
  sep_tell_update_kernel<<<nb, 256, 0, s>>>(n, new_mean, rank_mu, mean, pc, C, status, c, ws, nb, ws + nb);
  if ((rc = check_launch("sep_tell_update_kernel"))) return rc;
  minmax_final_kernel<<<1, 1, 0, s>>>(ws + nb, nb, status);
  return check_launch("minmax_final_kernel");
}

extern "C" int qd_cma_transform(int64_t n, const double* Bmat, const double* eigvals, double* T, void* stream) {
  QD_REQUIRE(Bmat && eigvals && T, "null pointer");
  cma_transform_kernel<<<grid_for(n * n, 256), 256, 0, (cudaStream_t)stream>>>(n, Bmat, eigvals, T);
  return check_launch("cma_transform_kernel");
}

extern "C" int qd_cma_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                          const double* T, const double* status, double* X, void* stream) {
  if (rows == 0) return QD_OK;
  QD_REQUIRE(z && mean && T && status && X, "null pointer");
  // X[row][i] = mean[i] + sum_j T[i][j] * (sigma z[k][j])      (_cma_es.py:181-189)
  auto a_of = [=] __device__(int64_t k, int64_t j) { return __dmul_rn(status[kStatusSigma], z[k * n + j]); };
  auto b_of = [=] __device__(int64_t i, int64_t j) { return T[i * n + j]; };
  auto epi = [=] __device__(int64_t k, int64_t i, double v) {
    const int64_t row = row_ids ? row_ids[k] : k;
    X[row * n + i] = v + mean[i];
  };
  return launch_gemm_nt(rows, n, n, a_of, b_of, epi, (cudaStream_t)stream, "cma_ask_gemm");
}

extern "C" int qd_cma_invsqrt(int64_t n, const double* Bmat, const double* eigvals, double* invsqrt, void* stream) {
  QD_REQUIRE(Bmat && eigvals && invsqrt, "null pointer");
  // (B * (1/sqrt(eig))) @ B^T, then max with its transpose   (_cma_es.py:76-80)
  auto a_of = [=] __device__(int64_t i, int64_t k) { return __dmul_rn(Bmat[i * n + k], 1.0 / sqrt(eigvals[k])); };
  auto b_of = [=] __device__(int64_t j, int64_t k) { return Bmat[j * n + k]; };
  auto epi = [=] __device__(int64_t i, int64_t j, double v) { invsqrt[i * n + j] = v; };
  int rc = launch_gemm_nt(n, n, n, a_of, b_of, epi, (cudaStream_t)stream, "cma_invsqrt_gemm");
  if (rc) return rc;
  sym_max_kernel<<<grid_for(n * n, 256), 256, 0, (cudaStream_t)stream>>>(n, invsqrt);
  return check_launch("sym_max_kernel");
}

extern "C" int qd_sym_max(int64_t n, double* A, void* stream) {
  QD_REQUIRE(A, "null pointer");
  sym_max_kernel<<<grid_for(n * n, 256), 256, 0, (cudaStream_t)stream>>>(n, A);
  return check_launch("sym_max_kernel");
}

extern "C" int qd_cma_tell(int64_t n, const double* X, const int64_t* rank, const double* w, int64_t mu,
                           const double* new_mean, const double* invsqrt, double* mean, double* pc, double* ps,
                           const double* Cmat, double* Cnext, double* status, const qd_cma_scalars* sc, double* ws,
                           void* stream) {
  QD_REQUIRE(X && rank && w && new_mean && invsqrt && mean && pc && ps && Cmat && Cnext && status && sc && ws,
             "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  CmaScalars c = to_dev(sc);
  double* rowsq = ws;       // n
  double* scal = ws + n;    // 8
  cma_tell_ps_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(n, new_mean, mean, invsqrt, ps, status, c, rowsq);
  int rc = check_launch("cma_tell_ps_kernel");
  if (rc) return rc;
  cma_tell_pc_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, new_mean, mean, pc, status, c, rowsq, scal);
  if ((rc = check_launch("cma_tell_pc_kernel"))) return rc;
  // rank-mu: Cnext[i][j] = C[i][j]*decay + c1^2 pc_i pc_j + (cmu/sigma^2) sum_k w_k ys_ki ys_kj   (:261-270, :318-323)
  const double* old_mean = mean;
  auto a_of = [=] __device__(int64_t i, int64_t k) { return (X[rank[k] * n + i] - old_mean[i]) * w[k]; };
  auto b_of = [=] __device__(int64_t j, int64_t k) { return X[rank[k] * n + j] - old_mean[j]; };
  auto epi = [=] __device__(int64_t i, int64_t j, double v) {
    Cnext[i * n + j] = Cmat[i * n + j] * scal[0] + (pc[i] * pc[j]) * scal[1] + v * scal[2];
  };
  if ((rc = launch_gemm_nt(n, n, mu, a_of, b_of, epi, s, "cma_rank_mu_gemm"))) return rc;
  cma_tell_finish_kernel<<<grid_for(n, 256), 256, 0, s>>>(n, new_mean, mean, status, scal);
  return check_launch("cma_tell_finish_kernel");
}

extern "C" int qd_lmma_tell(int64_t n, int64_t n_vectors, const double* new_mean, const double* z_mean, double* mean,
                            double* ps, double* M, const double* cc, double csigma, double mueff, double* status,
                            double* ws, void* stream) {
  QD_REQUIRE(new_mean && z_mean && mean && ps && M && cc && status && ws, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = (int)((n + 255) / 256);
  lmma_tell_ps_kernel<<<nb, 256, 0, s>>>(n, z_mean, ps, csigma, mueff, ws);
  int rc = check_launch("lmma_tell_ps_kernel");
  if (rc) return rc;
  // column strips x vector ranges: enough CTAs for eight per SM, at most kLmTellVecs vectors each; with no vectors
  // one row of CTAs still copies the mean and updates sigma
  const bool vec2 = n % 2 == 0 && (((uintptr_t)M | (uintptr_t)z_mean | (uintptr_t)mean | (uintptr_t)new_mean) & 15u) == 0;
  const int64_t gx = vec2 ? (n / 2 + 255) / 256 : (n + 255) / 256;
  int64_t gy = (8 * (int64_t)kNumSMs + gx - 1) / gx;
  if (gy > n_vectors) gy = n_vectors;
  if (gy < 1) gy = 1;
  int64_t vpc = (n_vectors + gy - 1) / gy;
  if (vpc > kLmTellVecs) vpc = kLmTellVecs;
  if (vpc < 1) vpc = 1;
  gy = n_vectors > 0 ? (n_vectors + vpc - 1) / vpc : 1;
  QD_REQUIRE(gx <= 0x7fffffff && gy <= 65535, "sizes out of range");
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (vec2)
    lmma_tell_m_kernel<true><<<grid, 256, 0, s>>>(n, n_vectors, vpc, z_mean, new_mean, mean, M, cc, csigma, mueff, status,
                                                  ws, nb);
  else
    lmma_tell_m_kernel<false><<<grid, 256, 0, s>>>(n, n_vectors, vpc, z_mean, new_mean, mean, M, cc, csigma, mueff, status,
                                                   ws, nb);
  return check_launch("lmma_tell_m_kernel");
}
