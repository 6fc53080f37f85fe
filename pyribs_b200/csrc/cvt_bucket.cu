// A3, low measure_dim: exact nearest centroid through a uniform bucket grid (method 2).
//
// Reference: ribs/archives/_cvt_archive.py:579-609. The reference's default is a KD-tree
// (scipy.spatial.KDTree.query, k=1): a spatial index, not a B x C contraction. For
// measure_dim <= 3 this file is its GPU analogue. The centroids of a shard are binned once
// (qd_cvt_bucket_prepare, the moment the reference builds its tree, :410-429) into G^d
// equal buckets over their bounding box, stored bucket-major (CSR: start[], points, ids).
// A query walks Chebyshev rings of buckets around its own bucket, evaluating the exact
// fp64 squared distance (cvt_common.cuh, NumPy's summation order) of every centroid it
// meets, and stops when the best distance found is below a conservative lower bound on
// anything outside the rings already visited. The (distance, index) minimum is taken
// lexicographically, so the result is bit-identical to the fp64 scan / np.argmin
// (_cvt_archive.py:632) -- lowest centroid index on exact ties.
//
// Cost per query: ~9 (2-D) / ~27 (3-D) buckets of ~2 centroids, against C evaluations
// for the scan and C accumulator reads for the tensor-core filter; the whole structure of
// the 10k-centroid 2-D case (220 KB) lives in L1/L2.
#include "cvt_common.cuh"

namespace qd {

constexpr int kBucketMaxDim = 3;
constexpr int kBucketHeaderBytes = 256;

struct BucketHeader {
  double lo[kBucketMaxDim];
  double w[kBucketMaxDim];
  double inv_w[kBucketMaxDim];
  double margin;  // absolute slack on the ring lower bound (covers face / binning rounding)
  int32_t G[kBucketMaxDim];
  int32_t nb;
  int32_t C;
  int32_t d;
};
static_assert(sizeof(BucketHeader) <= kBucketHeaderBytes, "header too large");

struct BucketView {
  BucketHeader* h;
  double* mm;       // 2*d running min / max of the centroid coordinates
  int32_t* start;   // nb + 1
  int32_t* cursor;  // nb (build only)
  int32_t* ids;     // C
  double* pts;      // C * d, bucket-major
};

static int64_t bucket_count(int64_t C, int d) {
  // ~2 centroids per bucket
  double g = ceil(pow((double)C / 2.0, 1.0 / d));
  int64_t G = g < 1 ? 1 : (int64_t)g;
  int64_t nb = 1;
  for (int k = 0; k < d; ++k) nb *= G;
  return nb;
}
static int grid_side(int64_t C, int d) {
  double g = ceil(pow((double)C / 2.0, 1.0 / d));
  return g < 1 ? 1 : (int)g;
}

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

static BucketView carve_bucket(void* base, int64_t C, int d) {
  const int64_t nb = bucket_count(C, d);
  BucketView v;
  char* p = (char*)base;
  v.h = (BucketHeader*)p;
  p += kBucketHeaderBytes;
  v.mm = (double*)p;
  p += 256;
  v.start = (int32_t*)p;
  p += a256((size_t)(nb + 1) * 4);
  v.cursor = (int32_t*)p;
  p += a256((size_t)nb * 4);
  v.ids = (int32_t*)p;
  p += a256((size_t)C * 4);
  v.pts = (double*)p;
  return v;
}

static size_t bucket_bytes(int64_t C, int d) {
  const int64_t nb = bucket_count(C, d);
  return kBucketHeaderBytes + 256 + a256((size_t)(nb + 1) * 4) + a256((size_t)nb * 4) + a256((size_t)C * 4) +
         a256((size_t)C * d * 8);
}

// ---- build -------------------------------------------------------------------------------
__global__ void bucket_init_kernel(BucketView v, int d, int64_t nb) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < d) {
    v.mm[i] = INFINITY;
    v.mm[d + i] = -INFINITY;
  }
  for (int64_t b = i; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    v.start[b] = 0;
    v.cursor[b] = 0;
  }
  if (i == 0) v.start[nb] = 0;
}

__device__ __forceinline__ void atomic_min_double(double* a, double x) {
  // order-preserving key on the bit pattern; coordinates are finite here
  unsigned long long* p = (unsigned long long*)a;
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) > x) {
    unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(x));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_double(double* a, double x) {
  unsigned long long* p = (unsigned long long*)a;
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) < x) {
    unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(x));
    if (prev == old) break;
    old = prev;
  }
}

template <typename T>
__global__ void bucket_minmax_kernel(const T* __restrict__ cent, int64_t C, int d, BucketView v) {
  double lo[kBucketMaxDim], hi[kBucketMaxDim];
  for (int k = 0; k < d; ++k) {
    lo[k] = INFINITY;
    hi[k] = -INFINITY;
  }
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x)
    for (int k = 0; k < d; ++k) {
      const double x = (double)cent[c * d + k];
      if (finite_d(x)) {
        lo[k] = fmin(lo[k], x);
        hi[k] = fmax(hi[k], x);
      }
    }
  for (int k = 0; k < d; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fmin(lo[k], __shfl_down_sync(0xffffffffu, lo[k], o));
      hi[k] = fmax(hi[k], __shfl_down_sync(0xffffffffu, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) {
      if (lo[k] != INFINITY) atomic_min_double(&v.mm[k], lo[k]);
      if (hi[k] != -INFINITY) atomic_max_double(&v.mm[d + k], hi[k]);
    }
  }
}

__global__ void bucket_header_kernel(BucketView v, int64_t C, int d, int G) {
  if (threadIdx.x || blockIdx.x) return;
  BucketHeader h;
  double scale = 0.0;
  for (int k = 0; k < kBucketMaxDim; ++k) {
    h.lo[k] = 0.0;
    h.w[k] = 1.0;
    h.inv_w[k] = 0.0;
    h.G[k] = 1;
  }
  int64_t nb = 1;
  for (int k = 0; k < d; ++k) {
    double lo = v.mm[k], hi = v.mm[d + k];
    if (!(lo <= hi)) lo = hi = 0.0;  // no finite coordinate at all
    const double ext = hi - lo;
    h.lo[k] = lo;
    if (ext > 0.0 && finite_d(ext)) {
      h.G[k] = G;
      h.w[k] = ext / (double)G;
      h.inv_w[k] = (double)G / ext;
    } else {  // degenerate axis: a single slice (inv_w = 0 sends everything to slice 0)
      h.G[k] = 1;
      h.w[k] = 1.0;
      h.inv_w[k] = 0.0;
    }
    scale = fmax(scale, fmax(fabs(lo), fabs(hi)));
    nb *= h.G[k];
  }
  h.margin = scale * 1e-11 + 1e-300;
  h.nb = (int32_t)nb;
  h.C = (int32_t)C;
  h.d = d;
  *v.h = h;
}

template <int D>
__device__ __forceinline__ int bucket_coord(double x, int k, const BucketHeader& h) {
  double t = (x - h.lo[k]) * h.inv_w[k];
  int g = (t >= 0.0) ? ((t < 2147483000.0) ? (int)t : 2147483000) : 0;  // NaN -> 0
  return min(g, h.G[k] - 1);
}

template <typename T>
__global__ void bucket_count_kernel(const T* __restrict__ cent, int64_t C, int d, BucketView v, bool fill) {
  const BucketHeader h = *v.h;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < C; c += (int64_t)gridDim.x * blockDim.x) {
    int b = 0;
    for (int k = 0; k < d; ++k) b = b * h.G[k] + bucket_coord<0>((double)cent[c * d + k], k, h);
    if (!fill) {
      atomicAdd(&v.start[b + 1], 1);
    } else {
      const int pos = v.start[b] + atomicAdd(&v.cursor[b], 1);
      v.ids[pos] = (int32_t)c;
      for (int k = 0; k < d; ++k) v.pts[(int64_t)pos * d + k] = (double)cent[c * d + k];
    }
  }
}

// single-block inclusive scan over start[1..nb] (set-up path, once per archive)
__global__ void __launch_bounds__(1024) bucket_scan_kernel(BucketView v, int64_t nb) {
  __shared__ int32_t sh[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 1; base <= nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int32_t x = i <= nb ? v.start[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i <= nb) v.start[i] = sh[threadIdx.x] + carry;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
}

// ---- search ------------------------------------------------------------------------------
// Where the search reads the structure from: global memory through the read-only path, or a copy the CTA
// staged in shared memory (random 16-byte gathers cost ~2 L1 wavefront-cycles per lane from global memory,
// a few bank conflicts from shared memory). ids stay in global memory: they are fetched for improving
// candidates only.
template <bool SMEM>
struct BSrc {
  const double* pts;
  const int32_t* start;
  const int32_t* ids;
  __device__ __forceinline__ int32_t st(int i) const { return SMEM ? start[i] : __ldg(start + i); }
  __device__ __forceinline__ int32_t id(int p) const { return __ldg(ids + p); }
  __device__ __forceinline__ double pt(int64_t i) const { return SMEM ? pts[i] : __ldg(pts + i); }
  __device__ __forceinline__ double2 pt2(int p) const {
    return SMEM ? reinterpret_cast<const double2*>(pts)[p] : __ldg(reinterpret_cast<const double2*>(pts) + p);
  }
};

template <int D, class S>
__device__ __forceinline__ void consider(const double* __restrict__ x, const S& v, int p, const double (&c)[D],
                                         double& best, int32_t& bi) {
  const double dd = sqdist_np<D>(x, c, D);
  if (dd < best) {  // the id is only fetched for a candidate that matters
    best = dd;
    bi = v.id(p);
  } else if (dd == best) {
    const int32_t id = v.id(p);
    if (id < bi) bi = id;
  }
}

// one point of the bucket-major array; 2-D points are 16-byte aligned: one vector load (scattered loads cost
// L1 wavefronts per lane and per instruction, and this kernel is bound by them)
template <int D, class S>
__device__ __forceinline__ void load_point(const S& v, int p, double (&c)[D]) {
  if constexpr (D == 2) {
    const double2 t = v.pt2(p);
    c[0] = t.x;
    c[1] = t.y;
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = v.pt((int64_t)p * D + k);
  }
}

// points [p0, p1) of the bucket-major array, two at a time so that their loads overlap
template <int D, class S>
__device__ __forceinline__ void scan_range(const double* __restrict__ x, const S& v, int p0, int p1,
                                           double& best, int32_t& bi) {
  int p = p0;
  for (; p + 1 < p1; p += 2) {
    double c0[D], c1[D];
    load_point<D>(v, p, c0);
    load_point<D>(v, p + 1, c1);
    consider<D>(x, v, p, c0, best, bi);
    consider<D>(x, v, p + 1, c1, best, bi);
  }
  if (p < p1) {
    double c0[D];
    load_point<D>(v, p, c0);
    consider<D>(x, v, p, c0, best, bi);
  }
}

// conservative lower bound on the distance to anything outside the (2r+1)-box around the query's bucket
template <int D>
__device__ __forceinline__ double ring_bound(const double* x, const int* bx, int r, const BucketHeader& h) {
  double lb = INFINITY;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (bx[k] - r > 0) lb = fmin(lb, x[k] - (h.lo[k] + (double)(bx[k] - r) * h.w[k]));
    if (bx[k] + r < h.G[k] - 1) lb = fmin(lb, (h.lo[k] + (double)(bx[k] + r + 1) * h.w[k]) - x[k]);
  }
  if (lb == INFINITY) return lb;                  // the box covers the whole grid
  return lb * (1.0 - 1e-11) - h.margin;           // far larger than any rounding of the faces / the binning
}

template <typename T, int D, bool VEC, bool SMEM>
__global__ void __launch_bounds__(SMEM ? 1024 : 256) cvt_bucket_kernel(const T* __restrict__ meas, int64_t B,
                                                                     BucketView bv, int32_t base,
                                                                     int32_t* __restrict__ idx,
                                                                     double* __restrict__ dist_out, uint32_t* flags,
                                                                     int64_t C) {
  __shared__ BucketHeader h;
  extern __shared__ __align__(16) unsigned char bucket_smem[];
  if (threadIdx.x == 0) h = *bv.h;
  __syncthreads();
  BSrc<SMEM> v;
  v.ids = bv.ids;
  if constexpr (SMEM) {  // stage points + CSR offsets once per CTA (both 16-byte aligned, sizes padded by the host)
    const int64_t pts16 = (C * D * 8 + 15) / 16, st16 = ((int64_t)(h.nb + 1) * 4 + 15) / 16;
    uint4* sp = reinterpret_cast<uint4*>(bucket_smem);
    uint4* ss = sp + pts16;
    const uint4* gp = reinterpret_cast<const uint4*>(bv.pts);
    const uint4* gs = reinterpret_cast<const uint4*>(bv.start);
    for (int64_t t = threadIdx.x; t < pts16; t += blockDim.x) sp[t] = __ldg(gp + t);
    for (int64_t t = threadIdx.x; t < st16; t += blockDim.x) ss[t] = __ldg(gs + t);
    __syncthreads();
    v.pts = reinterpret_cast<const double*>(sp);
    v.start = reinterpret_cast<const int32_t*>(ss);
  } else {
    v.pts = bv.pts;
    v.start = bv.start;
  }
  bool bad = false;
  constexpr int kRows = D == 1 ? 1 : (D == 2 ? 3 : 9);  // rows of the 3^D start box (last axis contiguous)
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < B; q += (int64_t)gridDim.x * blockDim.x) {
    double x[D];
    int bx[D];
    bool qbad = false;
    if constexpr (VEC) {  // 2-D fp64 rows on a 16-byte aligned base
      const double2 t = reinterpret_cast<const double2*>(meas)[q];
      x[0] = t.x;
      x[1] = t.y;
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) x[k] = (double)meas[q * D + k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
      qbad |= !finite_d(x[k]);
      bx[k] = bucket_coord<D>(x[k], k, h);
    }
    bad |= qbad;
    double best = INFINITY;
    int32_t bi = INT32_MAX;
    int maxG = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) maxG = max(maxG, h.G[k]);
    if (!qbad) {
      // rings 0 and 1 together: the whole 3^D box. All row extents are fetched before any point is touched.
      const int GL = h.G[D - 1];
      const int s0 = max(bx[D - 1] - 1, 0), s1 = min(bx[D - 1] + 1, GL - 1);
      int p0[kRows], p1[kRows];
#pragma unroll
      for (int t = 0; t < kRows; ++t) {
        int row = 0;
        bool in = true;
        if (D == 2) {
          row = bx[0] + t - 1;
          in = row >= 0 && row < h.G[0];
        } else if (D == 3) {
          const int a = bx[0] + t / 3 - 1, b2 = bx[1] + t % 3 - 1;
          in = a >= 0 && a < h.G[0] && b2 >= 0 && b2 < h.G[1];
          row = a * h.G[1] + b2;
        }
        p0[t] = in ? v.st(row * GL + s0) : 0;
        p1[t] = in ? v.st(row * GL + s1 + 1) : 0;
      }
#pragma unroll
      for (int t = 0; t < kRows; ++t) scan_range<D>(x, v, p0[t], p1[t], best, bi);
      double lb = ring_bound<D>(x, bx, 1, h);
      for (int r = 2; r < maxG && !(lb == INFINITY) && !(lb > 0.0 && best < lb * lb); ++r) {
        // shell of Chebyshev radius r: rows over the leading D-1 axes; the last axis is contiguous in the CSR layout
        const int lo1 = D >= 2 ? max(bx[0] - r, 0) : 0, hi1 = D >= 2 ? min(bx[0] + r, h.G[0] - 1) : 0;
        for (int a = lo1; a <= hi1; ++a) {
          const int lo2 = D >= 3 ? max(bx[1] - r, 0) : 0, hi2 = D >= 3 ? min(bx[1] + r, h.G[1] - 1) : 0;
          for (int b2 = lo2; b2 <= hi2; ++b2) {
            int row = 0;
            bool shell = false;  // some leading axis already sits at Chebyshev distance r
            if (D >= 2) {
              row = a;
              shell |= (abs(a - bx[0]) == r);
            }
            if (D >= 3) {
              row = row * h.G[1] + b2;
              shell |= (abs(b2 - bx[1]) == r);
            }
            const int c0 = bx[D - 1] - r, c1 = bx[D - 1] + r;
            const int rowbase = row * GL;
            if (shell) {
              const int t0 = max(c0, 0), t1 = min(c1, GL - 1);
              scan_range<D>(x, v, v.st(rowbase + t0), v.st(rowbase + t1 + 1), best, bi);
            } else {
              if (c0 >= 0) scan_range<D>(x, v, v.st(rowbase + c0), v.st(rowbase + c0 + 1), best, bi);
              if (c1 < GL) scan_range<D>(x, v, v.st(rowbase + c1), v.st(rowbase + c1 + 1), best, bi);
            }
          }
        }
        lb = ring_bound<D>(x, bx, r, h);
      }
    }
    idx[q] = qbad ? base : ((bi == INT32_MAX ? 0 : bi) + base);
    if (dist_out) dist_out[q] = qbad ? INFINITY : best;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, QD_FLAG_NONFINITE_MEASURES);
}

constexpr size_t kBucketSmemMax = 200 * 1024;

template <typename T, int D>
static int launch_bucket_d(const T* meas, int64_t B, int64_t C, BucketView v, int32_t base, int32_t* idx, double* dist,
                           uint32_t* flags, cudaStream_t s) {
  const bool vec = D == 2 && sizeof(T) == 8 && ((uintptr_t)meas & 15) == 0;
  // points + CSR offsets of small shards fit in shared memory: one CTA per SM stages them and serves a slice
  // of the queries from there
  const int64_t nb = bucket_count(C, D);
  const size_t smem = (size_t)((C * D * 8 + 15) / 16 + ((nb + 1) * 4 + 15) / 16) * 16;
  if (smem <= kBucketSmemMax && B >= 4096) {
    const int grid = (int)((B + 1023) / 1024 < kNumSMs ? (B + 1023) / 1024 : kNumSMs);
    if (vec) {
      QD_CHECK_CUDA(cudaFuncSetAttribute(cvt_bucket_kernel<T, D, D == 2 && sizeof(T) == 8, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBucketSmemMax));
      cvt_bucket_kernel<T, D, D == 2 && sizeof(T) == 8, true><<<grid, 1024, smem, s>>>(meas, B, v, base, idx, dist,
                                                                                     flags, C);
    } else {
      QD_CHECK_CUDA(cudaFuncSetAttribute(cvt_bucket_kernel<T, D, false, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBucketSmemMax));
      cvt_bucket_kernel<T, D, false, true><<<grid, 1024, smem, s>>>(meas, B, v, base, idx, dist, flags, C);
    }
    return check_launch("cvt_bucket_kernel(smem)");
  }
  const int grid = grid_for(B, 256, 8);
  if (vec)
    cvt_bucket_kernel<T, D, D == 2 && sizeof(T) == 8, false><<<grid, 256, 0, s>>>(meas, B, v, base, idx, dist, flags, C);
  else
    cvt_bucket_kernel<T, D, false, false><<<grid, 256, 0, s>>>(meas, B, v, base, idx, dist, flags, C);
  return check_launch("cvt_bucket_kernel");
}

template <typename T>
static int launch_bucket(const T* meas, int64_t B, int d, int64_t C, BucketView v, int32_t base, int32_t* idx,
                         double* dist, uint32_t* flags, cudaStream_t s) {
  switch (d) {
    case 1: return launch_bucket_d<T, 1>(meas, B, C, v, base, idx, dist, flags, s);
    case 2: return launch_bucket_d<T, 2>(meas, B, C, v, base, idx, dist, flags, s);
    case 3: return launch_bucket_d<T, 3>(meas, B, C, v, base, idx, dist, flags, s);
    default: set_error("bucket search supports measure_dim <= 3"); return QD_ERR_UNSUPPORTED;
  }
}

int cvt_bucket_index_of(const void* measures, int dtype, int64_t B, int d, int64_t C, const void* buckets,
                        int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags, cudaStream_t s) {
  BucketView v = carve_bucket(const_cast<void*>(buckets), C, d);
  if (dtype == QD_F64)
    return launch_bucket<double>((const double*)measures, B, d, C, v, index_base, idx_out, dist_out, flags, s);
  return launch_bucket<float>((const float*)measures, B, d, C, v, index_base, idx_out, dist_out, flags, s);
}

}  // namespace qd

extern "C" size_t qd_cvt_bucket_bytes(int64_t C, int d) {
  if (C < 1 || d < 1 || d > qd::kBucketMaxDim) return 0;
  return qd::bucket_bytes(C, d);
}

extern "C" int qd_cvt_bucket_prepare(const void* centroids, int dtype, int64_t C, int d, void* buckets, void* stream) {
  using namespace qd;
  QD_REQUIRE(d >= 1 && d <= kBucketMaxDim, "bucket search supports measure_dim in [1, 3]");
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  QD_REQUIRE(C >= 1 && C < ((int64_t)1 << 31), "centroid count must be in [1, 2^31)");
  QD_REQUIRE(centroids && buckets, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  BucketView v = carve_bucket(buckets, C, d);
  const int64_t nb = bucket_count(C, d);
  const int G = grid_side(C, d);
  bucket_init_kernel<<<grid_for(nb > 64 ? nb : 64, 256), 256, 0, s>>>(v, d, nb);
  int rc = check_launch("bucket_init_kernel");
  if (rc) return rc;
  const int grid = grid_for(C, 256, 4);
  if (dtype == QD_F64)
    bucket_minmax_kernel<double><<<grid, 256, 0, s>>>((const double*)centroids, C, d, v);
  else
    bucket_minmax_kernel<float><<<grid, 256, 0, s>>>((const float*)centroids, C, d, v);
  if ((rc = check_launch("bucket_minmax_kernel"))) return rc;
  bucket_header_kernel<<<1, 32, 0, s>>>(v, C, d, G);
  if ((rc = check_launch("bucket_header_kernel"))) return rc;
  for (int pass = 0; pass < 2; ++pass) {
    if (dtype == QD_F64)
      bucket_count_kernel<double><<<grid, 256, 0, s>>>((const double*)centroids, C, d, v, pass == 1);
    else
      bucket_count_kernel<float><<<grid, 256, 0, s>>>((const float*)centroids, C, d, v, pass == 1);
    if ((rc = check_launch("bucket_count_kernel"))) return rc;
    if (pass == 0) {
      bucket_scan_kernel<<<1, 1024, 0, s>>>(v, nb);
      if ((rc = check_launch("bucket_scan_kernel"))) return rc;
    }
  }
  return QD_OK;
}
