// A3, method 0: CVTArchive.index_of as a tensor-core filter + exact fp64 rescoring
// (reference ribs/archives/_cvt_archive.py:579-644; brute-force semantics :626-632).
//
//   score(q, c) = ||c||^2 - 2 x_q . c      (argmin_c score == argmin_c ||x_q - c||^2)
//
// is evaluated for every (query, centroid) pair by tcgen05.mma (kind::f16, fp32 accumulate in
// TMEM) on fp16 operand images whose K dimension carries the coordinate products AND the
// norm:   A row (query)    = [-2u_hi | -2u_hi | -2u_lo | t t t | 0..]      (P = 3 split)
//         B row (centroid) = [ c_hi  |  c_lo  |  c_hi  | n_hi n_mid n_lo | 0..]
// with u = t * s * x, c = s * centroid, s = a power of two that brings all centroid
// coordinates into [-1, 1], t = a per-query power of two that does the same for the query
// (a positive per-row factor does not change that row's argmin), n = ||c||^2 split three
// ways. P = 1 keeps only the hi*hi block. The epilogue warps read the accumulator straight
// out of TMEM (tcgen05.ld) and keep a running minimum per query; every pair whose score is
// within eps_q of the running minimum is appended to a candidate list. eps_q is a rigorous
// bound on twice the approximation error (derivation in DESIGN.md), so the exact nearest
// centroid -- and every exact tie -- is always a candidate. Candidates are rescored in fp64
// in NumPy's summation order and the lowest index among exact minima wins, which makes the
// result bit-identical to the fp64 scan (method 1).
//
// CTA = 128 queries x a contiguous range of 256-centroid tiles; 18 warps:
//   warp 0      bulk-copy (TMA, cp.async.bulk) producer of B tiles, 4-stage mbarrier ring
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue, 2 accumulator buffers
//   warps 2-17  epilogue: warp w reads TMEM lanes 32*(w%4).., columns 64*((w-2)/4)..
#include <cuda_fp16.h>

#include "cvt_common.cuh"

namespace qd {

constexpr int kTileM = 128;         // queries per CTA (TMEM lanes)
constexpr int kTileN = 256;         // centroids per MMA / accumulator buffer (TMEM columns)
constexpr int kStages = 4;          // B-tile ring depth
constexpr int kMaxKSteps = 4;       // K = 16 * ksteps
constexpr int kEpiWarps = 16;       // 4 per SM sub-partition
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiWarp0 = 2;
constexpr int kWarpBuf = 192;       // candidates staged per epilogue warp before one global reservation
constexpr int kRegions = 1024;      // candidate list is split into independently filled regions

struct PreparedHeader {  // lives at the start of the prepared image (device memory)
  double scale;                      // s (power of two)
  unsigned long long cmax2_sq_bits;  // max ||s c||_2^2 (double bits, for atomicMax)
  unsigned long long cmax1_bits;     // max ||s c||_1
  unsigned long long absmax_bits;    // max |coordinate|
  int32_t P, ksteps, ntiles, d;
  int64_t C;
};
constexpr size_t kHeaderBytes = 256;

struct SearchScratch {
  uint32_t* region_count;  // kRegions
  uint32_t* n_outliers;
  unsigned long long* bestd;  // B
  int32_t* outliers;          // B
  int2* cand;                 // kRegions * region_cap: (query, first column of a 64-column slice)
  float* cand_score;          // kRegions * region_cap: tensor-core minimum of that slice
  unsigned long long* cand_bits;  // kRegions * region_cap: exact minimum distance of the slice (pass 0 -> 1)
  uint32_t* best_approx;      // B: ordered-uint image of the smallest tensor-core score per query
  float* eps;                 // B: per-query margin
  uint32_t region_cap;
  float* dbg_scores;  // diagnostics only: B x (ntiles*256) raw accumulator values, or NULL
  float* dbg_eps;     // diagnostics only: B margins, or NULL
};

// monotone float -> uint map (so that atomicMin on the image orders like the floats)
__device__ __forceinline__ uint32_t ford(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unford(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__host__ __device__ inline void pick_split(int d, int& P, int& ksteps) {
  P = (3 * d + 3 <= 32) ? 3 : 1;
  ksteps = (P * d + 3 + 15) / 16;
}

// byte offset of element (row r, k in [0,16)) inside one K-step block of a K-major,
// no-swizzle UMMA operand: core matrices of 8 rows x 16 bytes; SBO = 256 B between 8-row
// groups, LBO = 128 B between the two K chunks.
__host__ __device__ inline uint32_t canon_off(int r, int k) {
  return (uint32_t)((r >> 3) * 256 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// ------------------------------------------------------------------------------------------
// prepare: centroid image
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void prep_absmax_kernel(const T* __restrict__ cent, int64_t n, PreparedHeader* h) {
  double m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmax(m, fabs((double)cent[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(&h->absmax_bits, (unsigned long long)__double_as_longlong(m));
}

__global__ void prep_scale_kernel(PreparedHeader* h) {
  double m = __longlong_as_double((long long)h->absmax_bits);
  int e = 0;
  if (m > 0.0 && m < 1e300) frexp(m, &e);  // m = f * 2^e, f in [0.5, 1)  ->  m * 2^-e in [0.5, 1)
  h->scale = ldexp(1.0, -e);
}

__device__ __forceinline__ void split3(double v, __half& a, __half& b, __half& c) {
  a = __double2half(v);
  double r = v - (double)__half2float(a);
  b = __double2half(r);
  r -= (double)__half2float(b);
  c = __double2half(r);
}

template <typename T>
__global__ void prep_image_kernel(const T* __restrict__ cent, int64_t C, int d, PreparedHeader* h,
                                  unsigned char* __restrict__ image) {
  const double s = h->scale;
  const int P = h->P, ksteps = h->ksteps;
  const int64_t rows = (int64_t)h->ntiles * kTileN;
  double m2 = 0.0, m1 = 0.0;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < rows; c += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = c / kTileN;
    const int r = (int)(c % kTileN);
    unsigned char* tbase = image + (size_t)tile * ksteps * (kTileN * 32);
    auto put = [&](int k, __half v) {
      *reinterpret_cast<__half*>(tbase + (size_t)(k >> 4) * (kTileN * 32) + canon_off(r, k & 15)) = v;
    };
    for (int k = 0; k < ksteps * 16; ++k) put(k, __float2half(0.f));
    if (c < C) {
      double n2 = 0.0, n1 = 0.0;
      for (int k = 0; k < d; ++k) {
        const double v = (double)cent[c * d + k] * s;
        n2 += v * v;
        n1 += fabs(v);
        const __half hi = __double2half(v);
        put(k, hi);
        if (P == 3) {
          const __half lo = __double2half(v - (double)__half2float(hi));
          put(d + k, lo);
          put(2 * d + k, hi);
        }
      }
      __half a, b, e;
      split3(n2, a, b, e);
      put(P * d, a);
      put(P * d + 1, b);
      put(P * d + 2, e);
      m2 = fmax(m2, n2);
      m1 = fmax(m1, n1);
    }  // padding rows stay zero; the epilogue masks their columns
  }
  for (int o = 16; o > 0; o >>= 1) {
    m2 = fmax(m2, __shfl_down_sync(0xffffffffu, m2, o));
    m1 = fmax(m1, __shfl_down_sync(0xffffffffu, m1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (m2 > 0.0) atomicMax(&h->cmax2_sq_bits, (unsigned long long)__double_as_longlong(m2));
    if (m1 > 0.0) atomicMax(&h->cmax1_bits, (unsigned long long)__double_as_longlong(m1));
  }
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// K-major, no-swizzle smem matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;  // LBO: next K chunk (8 elements)
  d |= (uint64_t)(256 >> 4) << 32;  // SBO: next 8-row group
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc() {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

struct SmemLayout {
  unsigned char a[kMaxKSteps][kTileM * 32];              // 16 KB
  unsigned char b[kStages][kMaxKSteps][kTileN * 32];     // 128 KB
  uint64_t full[kStages], empty[kStages], tfull[2], tempty[2];
  uint32_t tmem_base;
  float eps[kTileM];
  int2 wbuf[kEpiWarps][kWarpBuf];  // per-epilogue-warp candidate staging (flushed in bulk)
  float wscore[kEpiWarps][kWarpBuf];
  uint32_t rowbest[kTileM];  // ordered-uint running minimum shared by the 4 warps of a row
  uint32_t wcnt[kEpiWarps];
};

// ------------------------------------------------------------------------------------------
// main search kernel
// ------------------------------------------------------------------------------------------
template <typename T, bool DBG>
__global__ void __launch_bounds__(kThreads, 1)
    cvt_tc_kernel(const T* __restrict__ meas, int64_t B, int d, const unsigned char* __restrict__ prepared,
                  int tiles_per_cta, SearchScratch ws, uint32_t* flags) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
  const PreparedHeader* hdr = reinterpret_cast<const PreparedHeader*>(prepared);
  const unsigned char* image = prepared + kHeaderBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = hdr->P, ksteps = hdr->ksteps, ntiles = hdr->ntiles;
  const int64_t q0 = (int64_t)blockIdx.x * kTileM;
  const int tile_begin = blockIdx.y * tiles_per_cta;
  const int tile_end = min(ntiles, tile_begin + tiles_per_cta);
  const int my_tiles = tile_end - tile_begin;
  const uint32_t tile_bytes = (uint32_t)ksteps * kTileN * 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&sm.tfull[a], 1);
      mbar_init(&sm.tempty[a], kEpiWarps);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&sm.tmem_base, 512);

  // ---- A tile: 128 query rows built in place by the first four epilogue warps ------------
  if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
    const int r = (warp & 3) * 32 + lane;
    const int64_t q = q0 + r;
    const double s = hdr->scale;
    double xs[QD_MAX_MEASURE_DIM];
    double umax = 0.0;
    bool finite = true;
    if (q < B) {
      for (int k = 0; k < d; ++k) {
        const double x = (double)meas[q * d + k];
        finite &= finite_d(x);
        xs[k] = x * s;
        umax = fmax(umax, fabs(xs[k]));
      }
    }
    double t = 1.0;
    bool outlier = false;
    if (q < B) {
      if (!finite) {
        if (blockIdx.y == 0) atomicOr(flags, QD_FLAG_NONFINITE_MEASURES);
        outlier = true;
      } else if (umax > 1.0) {
        int e;
        frexp(umax, &e);  // umax = f 2^e, f in [0.5,1)
        if (e > 24 || !finite_d(umax)) outlier = true;
        t = ldexp(1.0, -e);
      }
    }
    const bool live = (q < B) && !outlier;
    for (int k = 0; k < ksteps * 16; ++k)
      *reinterpret_cast<__half*>(&sm.a[k >> 4][canon_off(r, k & 15)]) = __float2half(0.f);
    float eps = 0.f;
    if (live) {
      double u2 = 0.0, u1 = 0.0;
      for (int k = 0; k < d; ++k) {
        const double u = xs[k] * t;
        u2 += u * u;
        u1 += fabs(u);
        const __half hi = __double2half(u);
        const __half m2hi = __double2half(-2.0 * (double)__half2float(hi));
        *reinterpret_cast<__half*>(&sm.a[k >> 4][canon_off(r, k & 15)]) = m2hi;
        if (P == 3) {
          const __half lo = __double2half(u - (double)__half2float(hi));
          const __half m2lo = __double2half(-2.0 * (double)__half2float(lo));
          int k1 = d + k, k2 = 2 * d + k;
          *reinterpret_cast<__half*>(&sm.a[k1 >> 4][canon_off(r, k1 & 15)]) = m2hi;
          *reinterpret_cast<__half*>(&sm.a[k2 >> 4][canon_off(r, k2 & 15)]) = m2lo;
        }
      }
      const __half th = __double2half(t);
      for (int j = 0; j < 3; ++j) {
        int kk = P * d + j;
        *reinterpret_cast<__half*>(&sm.a[kk >> 4][canon_off(r, kk & 15)]) = th;
      }
      // ---- error bound (DESIGN.md "CVT filter margin") -----------------------------------
      const double c2 = sqrt(__longlong_as_double((long long)hdr->cmax2_sq_bits));
      const double c1 = __longlong_as_double((long long)hdr->cmax1_bits);
      const double nmax = c2 * c2;
      const double un = sqrt(u2);
      const double rel = (P == 3) ? 3.2 * 0x1p-22 : 1.001 * 0x1p-10;
      const double e_prod = 2.0 * (rel * un * c2 + 1.01 * 0x1p-25 * (u1 + c1) + d * 0x1p-48);
      const double e_norm = t * (0x1p-32 * nmax + 0x1p-24);
      const double mag = 2.05 * un * c2 + t * nmax * 1.001;
      const double e_acc = (16.0 * ksteps + 1.0) * 0x1p-22 * mag;
      eps = (float)(2.1 * (e_prod + e_norm + e_acc) + 0x1p-20 * mag);
      eps = __fadd_ru(eps, eps * 1e-3f);
    } else if (q < B && finite && blockIdx.y == 0) {
      const uint32_t slot = atomicAdd(ws.n_outliers, 1u);
      ws.outliers[slot] = (int32_t)q;
    }
    sm.eps[r] = eps;
    sm.rowbest[r] = 0xFFFFFFFFu;
    if (q < B && blockIdx.y == 0) ws.eps[q] = eps;
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ===== producer: one bulk copy per B tile =============================================
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        mbar_expect_tx(&sm.full[s], tile_bytes);
        bulk_g2s(&sm.b[s][0][0], image + (size_t)(tile_begin + i) * tile_bytes, tile_bytes, &sm.full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ======================================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc();
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % kStages;
        const uint32_t ph = (i / kStages) & 1;
        const int acc = i & 1;
        const uint32_t aph = (i >> 1) & 1;
        mbar_wait(&sm.tempty[acc], aph ^ 1);
        mbar_wait(&sm.full[s], ph);
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k)
          umma_f16(tmem + acc * kTileN, make_desc(smem_u32(&sm.a[k][0])), make_desc(smem_u32(&sm.b[s][k][0])), idesc,
                   k > 0 ? 1u : 0u);
        umma_commit(&sm.empty[s]);    // smem stage reusable once these MMAs retire
        umma_commit(&sm.tfull[acc]);  // accumulator ready
      }
    }
  } else {
    // ===== epilogue ==========================================================================
    constexpr int kColsPerWarp = kTileN / (kEpiWarps / 4);  // 64
    const int quarter = warp & 3;
    const int cq = (warp - kEpiWarp0) >> 2;
    const int ew = warp - kEpiWarp0;
    const int r = quarter * 32 + lane;
    const int64_t q = q0 + r;
    const float eps = sm.eps[r];
    const bool live = (q < B) && eps > 0.f;
    // A candidate is a GROUP of 8 consecutive centroids (q, first column): the rescoring kernel
    // evaluates all 8 exactly, so the hot loop never has to locate individual columns.
    float best = 3.0e38f;
    float thr = live ? 3.0e38f : -INFINITY;  // dead rows never trigger; masked columns are +inf
    const int32_t C_total = (int32_t)hdr->C;
    const uint32_t region = (blockIdx.x * gridDim.y + blockIdx.y) % kRegions;
    int2* rcand = ws.cand + (size_t)region * ws.region_cap;
    float* rscore = ws.cand_score + (size_t)region * ws.region_cap;
    int2* wb = sm.wbuf[ew];
    float* wbs = sm.wscore[ew];
    uint32_t* wcnt = &sm.wcnt[ew];
    if (lane == 0) *wcnt = 0u;
    __syncwarp();
    auto flush = [&]() {  // warp-uniform
      __syncwarp();
      const uint32_t n = min(*wcnt, (uint32_t)kWarpBuf);
      if (n > 0) {
        uint32_t gbase = 0;
        if (lane == 0) gbase = atomicAdd(&ws.region_count[region], n);
        gbase = __shfl_sync(0xffffffffu, gbase, 0);
        for (uint32_t e = lane; e < n; e += 32) {
          if (gbase + e < ws.region_cap) {
            rcand[gbase + e] = wb[e];
            rscore[gbase + e] = wbs[e];
          } else {
            atomicOr(flags, QD_FLAG_CVT_OVERFLOW);
          }
        }
      }
      __syncwarp();
      if (lane == 0) *wcnt = 0u;
      __syncwarp();
    };
    static_assert(kColsPerWarp == 64, "the epilogue below loads two 32-column chunks per tile");
    static_assert(kColsPerWarp == 64, "rescore_kernel assumes 64-column candidates (kCandCols)");
    const int32_t qi = (int32_t)q;
    auto push = [&](int32_t col, float score) {
      const uint32_t pos = atomicAdd(wcnt, 1u);  // shared-memory atomic
      if (pos < (uint32_t)kWarpBuf) {
        wb[pos] = make_int2(qi, col);
        wbs[pos] = score;
      } else {
        const uint32_t slot = atomicAdd(&ws.region_count[region], 1u);
        if (slot < ws.region_cap) {
          rcand[slot] = make_int2(qi, col);
          rscore[slot] = score;
        } else {
          atomicOr(flags, QD_FLAG_CVT_OVERFLOW);
        }
      }
    };
    // minimum of one 32-column chunk with the 3-input min tree (16 FMNMX3-class ops)
    auto chunk_min = [&](const uint32_t (&raw)[32]) {
      float t[10];
#pragma unroll
      for (int g = 0; g < 10; ++g)
        t[g] = min3(__uint_as_float(raw[3 * g]), __uint_as_float(raw[3 * g + 1]), __uint_as_float(raw[3 * g + 2]));
      const float a = min3(t[0], t[1], t[2]), b = min3(t[3], t[4], t[5]), c = min3(t[6], t[7], t[8]);
      return min3(min3(a, b, c), t[9], fminf(__uint_as_float(raw[30]), __uint_as_float(raw[31])));
    };
    const int last_full = (C_total / kTileN);  // tiles with index < last_full have no padding columns
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      const uint32_t aph = (i >> 1) & 1;
      mbar_wait(&sm.tfull[acc], aph);
      tc_fence_after();
      const int32_t cbase = (tile_begin + i) * kTileN + cq * kColsPerWarp;
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kTileN + cq * kColsPerWarp);
      uint32_t ra[32], rb[32];
      tmem_ld32_nowait(taddr, ra);
      tmem_ld32_nowait(taddr + 32, rb);
      // pick up the row's running minimum from the other three warps while the loads fly
      const float shared_best = unford(sm.rowbest[r]);
      if (shared_best < best) {
        best = shared_best;
        thr = live ? best + eps : thr;
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.tempty[acc]);  // values are in registers: release the buffer early
      if (DBG) {
        if (ws.dbg_scores && q < B) {
          const int64_t ld = (int64_t)ntiles * kTileN;
          for (int j = 0; j < 32; ++j) {
            ws.dbg_scores[q * ld + cbase + j] = __uint_as_float(ra[j]);
            ws.dbg_scores[q * ld + cbase + 32 + j] = __uint_as_float(rb[j]);
          }
          ws.dbg_eps[q] = eps;
        }
      }
      if (tile_begin + i >= last_full) {  // tail tile: mask the padding columns (warp-uniform, once per CTA)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (cbase + j >= C_total) ra[j] = 0x7f800000u;
          if (cbase + 32 + j >= C_total) rb[j] = 0x7f800000u;
        }
      }
      // A candidate is this warp's whole 64-column slice of the tile (q, first column, its minimum):
      // the rescoring kernel filters stale records and evaluates the 64 centroids exactly, so the
      // hot loop never has to locate individual columns.
      const float m = fminf(chunk_min(ra), chunk_min(rb));
      const bool trig = m <= thr;
      if (__any_sync(0xffffffffu, trig)) {  // rare once the running minimum has settled
        if (trig) {
          best = fminf(best, m);
          thr = best + eps;
          push(cbase, m);
          atomicMin(&sm.rowbest[r], ford(best));
        }
        __syncwarp();
        if (*wcnt > (uint32_t)(kWarpBuf / 2)) flush();
      }
    }
    flush();
    if (live && best < 3.0e38f) atomicMin(&ws.best_approx[q], ford(best));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// exact rescoring
// ------------------------------------------------------------------------------------------
__global__ void search_init_kernel(int64_t B, SearchScratch ws, int32_t* __restrict__ idx) {
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    ws.bestd[i] = 0x7FF0000000000000ull;  // +inf
    ws.best_approx[i] = 0xFFFFFFFFu;
    idx[i] = INT32_MAX;
  }
  if (i0 < kRegions) ws.region_count[i0] = 0u;
  if (i0 == 0) *ws.n_outliers = 0u;
}

constexpr int kCandCols = 64;  // columns per candidate (= kTileN / (kEpiWarps / 4))

// Exact rescoring, pass 0. Each lane filters one candidate (stale running-minimum records: slice
// minimum above the query's final minimum + margin). Surviving slices are evaluated four at a time,
// one per 8-lane group, 8 columns per lane, in fp64 with NumPy's summation order; the group's
// (minimum distance, lowest column attaining it) goes back into the candidate record and the distance
// into bestd[q] with atomicMin. Pass 1 is then a per-candidate compare: no distance is computed twice.
template <typename T, int D>
__global__ void __launch_bounds__(256) rescore_kernel(const T* __restrict__ meas, const T* __restrict__ cent, int64_t C,
                                                     int d_rt, SearchScratch ws) {
  const int d = D > 0 ? D : d_rt;
  constexpr int XN = D > 0 ? D : QD_MAX_MEASURE_DIM;
  const int lane = threadIdx.x & 31;
  const int grp = lane >> 3, gl = lane & 7;
  const int warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (uint32_t region = blockIdx.x; region < kRegions; region += gridDim.x) {
    const uint32_t n = min(ws.region_count[region], ws.region_cap);
    int2* rc = ws.cand + (size_t)region * ws.region_cap;
    const float* rs = ws.cand_score + (size_t)region * ws.region_cap;
    unsigned long long* rb = ws.cand_bits + (size_t)region * ws.region_cap;
    for (uint32_t e0 = (blockIdx.y * nwarp + warp) * 32; e0 < n; e0 += gridDim.y * nwarp * 32) {
      const uint32_t e = e0 + lane;
      int2 qc = make_int2(0, 0);
      bool keep = false;
      if (e < n) {
        qc = rc[e];
        keep = rs[e] <= unford(ws.best_approx[qc.x]) + ws.eps[qc.x];
        if (!keep) rc[e].y = INT32_MAX;  // stale: pass 1 skips it
      }
      unsigned m = __ballot_sync(0xffffffffu, keep);
      while (m) {
        // this group's survivor: the grp-th set bit of m (if any)
        unsigned mm = m;
        for (int k = 0; k < grp; ++k) mm &= mm - 1;
        const int src = mm ? __ffs(mm) - 1 : -1;
        // drop the (up to) four survivors taken in this round
        for (int k = 0; k < 4 && m; ++k) m &= m - 1;
        const int q = __shfl_sync(0xffffffffu, qc.x, src < 0 ? 0 : src);
        const int c0 = __shfl_sync(0xffffffffu, qc.y, src < 0 ? 0 : src);
        unsigned long long mbits = 0xFFFFFFFFFFFFFFFFull;
        int32_t mcol = INT32_MAX;
        if (src >= 0) {
          double x[XN];
#pragma unroll
          for (int k = 0; k < XN; ++k)
            if (k < d) x[k] = (double)meas[(int64_t)q * d + k];
#pragma unroll 2
          for (int h = 0; h < kCandCols / 8; ++h) {
            const int32_t col = c0 + h * 8 + gl;
            if (col < C) {
              double c[XN];
#pragma unroll
              for (int k = 0; k < XN; ++k)
                if (k < d) c[k] = (double)cent[(int64_t)col * d + k];
              const unsigned long long bits = (unsigned long long)__double_as_longlong(sqdist_np<D>(x, c, d));
              if (bits < mbits) {  // ascending columns per lane: strict < keeps the lowest column
                mbits = bits;
                mcol = col;
              }
            }
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {  // (distance, column) lexicographic minimum over the 8-lane group
          const unsigned long long tb = __shfl_xor_sync(0xffffffffu, mbits, o);
          const int32_t tc = __shfl_xor_sync(0xffffffffu, mcol, o);
          if (tb < mbits || (tb == mbits && tc < mcol)) {
            mbits = tb;
            mcol = tc;
          }
        }
        if (src >= 0 && gl == 0) {
          rc[e0 + src].y = mcol;
          rb[e0 + src] = mbits;
          if (mcol != INT32_MAX) atomicMin(&ws.bestd[q], mbits);
        }
      }
    }
  }
}

// pass 1: idx[q] = lowest column among the candidates that attain bestd[q] (np.argmin's first minimum)
__global__ void __launch_bounds__(256) rescore_pick_kernel(SearchScratch ws, int32_t base, int32_t* __restrict__ idx) {
  for (uint32_t region = blockIdx.x; region < kRegions; region += gridDim.x) {
    const uint32_t n = min(ws.region_count[region], ws.region_cap);
    const int2* rc = ws.cand + (size_t)region * ws.region_cap;
    const unsigned long long* rb = ws.cand_bits + (size_t)region * ws.region_cap;
    for (uint32_t e = blockIdx.y * blockDim.x + threadIdx.x; e < n; e += gridDim.y * blockDim.x) {
      const int2 qc = rc[e];
      if (qc.y != INT32_MAX && rb[e] == ws.bestd[qc.x]) atomicMin(&idx[qc.x], qc.y + base);
    }
  }
}

template <typename T>
static int launch_rescore(const T* meas, const T* cent, int64_t C, int d, int32_t base, SearchScratch ws, int32_t* idx,
                          cudaStream_t s) {
  // blockIdx.x = region, blockIdx.y splits a region's candidates so that a warp sees at most a few
  // of them (the per-candidate work is a chain of dependent gathers)
  const dim3 grid(kRegions, 8);
#define QD_RS(DD) rescore_kernel<T, DD><<<grid, 256, 0, s>>>(meas, cent, C, d, ws)
  switch (d) {
    case 1: QD_RS(1); break;
    case 2: QD_RS(2); break;
    case 3: QD_RS(3); break;
    case 4: QD_RS(4); break;
    case 5: QD_RS(5); break;
    case 6: QD_RS(6); break;
    case 8: QD_RS(8); break;
    default: QD_RS(0);
  }
#undef QD_RS
  int rc = check_launch("rescore_kernel");
  if (rc) return rc;
  rescore_pick_kernel<<<dim3(kRegions, 2), 256, 0, s>>>(ws, base, idx);
  return check_launch("rescore_pick_kernel");
}

// queries that could not be represented in the fp16 image: one CTA scans all centroids
template <typename T>
__global__ void __launch_bounds__(256) outlier_scan_kernel(const T* __restrict__ meas, const T* __restrict__ cent,
                                                           int64_t C, int d, int32_t base, SearchScratch ws,
                                                           int32_t* __restrict__ idx) {
  __shared__ double sh_d[256];
  __shared__ int32_t sh_i[256];
  const uint32_t n = *ws.n_outliers;
  for (uint32_t o = blockIdx.x; o < n; o += gridDim.x) {
    const int64_t q = ws.outliers[o];
    double x[QD_MAX_MEASURE_DIM], c[QD_MAX_MEASURE_DIM];
    for (int k = 0; k < d; ++k) x[k] = (double)meas[q * d + k];
    double best = INFINITY;
    int32_t bi = INT32_MAX;
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
      for (int k = 0; k < d; ++k) c[k] = (double)cent[j * d + k];
      const double dd = sqdist_np<0>(x, c, d);
      if (dd < best) {
        best = dd;
        bi = (int32_t)j;
      }
    }
    sh_d[threadIdx.x] = best;
    sh_i[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        const double od = sh_d[threadIdx.x + s];
        const int32_t oi = sh_i[threadIdx.x + s];
        if (od < sh_d[threadIdx.x] || (od == sh_d[threadIdx.x] && oi < sh_i[threadIdx.x])) {
          sh_d[threadIdx.x] = od;
          sh_i[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      ws.bestd[q] = (unsigned long long)__double_as_longlong(sh_d[0]);
      idx[q] = sh_i[0] == INT32_MAX ? base : sh_i[0] + base;
    }
    __syncthreads();
  }
}

__global__ void search_finish_kernel(int64_t B, SearchScratch ws, int32_t* __restrict__ idx, int32_t base,
                                     double* __restrict__ dist_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    if (idx[i] == INT32_MAX) idx[i] = base;  // only reachable for non-finite queries (flag already set)
    if (dist_out) dist_out[i] = __longlong_as_double((long long)ws.bestd[i]);
  }
}

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }

static uint32_t region_cap_for(int64_t B) {
  int64_t cap = B / 16 + 4096;  // ~64 candidates per query overall, never fewer than 4096 per region
  if (cap > (1 << 22)) cap = 1 << 22;
  return (uint32_t)cap;
}

static size_t search_scratch_bytes(int64_t B) {
  return a256(kRegions * 4) + 256 + a256((size_t)B * 8) + 3 * a256((size_t)B * 4) +
         a256((size_t)kRegions * region_cap_for(B) * sizeof(int2)) +
         a256((size_t)kRegions * region_cap_for(B) * sizeof(float)) +
         a256((size_t)kRegions * region_cap_for(B) * sizeof(unsigned long long));
}

static SearchScratch carve_search(void* base, int64_t B) {
  SearchScratch ws;
  char* p = (char*)base;
  ws.region_count = (uint32_t*)p;
  p += a256(kRegions * 4);
  ws.n_outliers = (uint32_t*)p;
  p += 256;
  ws.bestd = (unsigned long long*)p;
  p += a256((size_t)B * 8);
  ws.outliers = (int32_t*)p;
  p += a256((size_t)B * 4);
  ws.best_approx = (uint32_t*)p;
  p += a256((size_t)B * 4);
  ws.eps = (float*)p;
  p += a256((size_t)B * 4);
  ws.region_cap = region_cap_for(B);
  ws.cand = (int2*)p;
  p += a256((size_t)kRegions * ws.region_cap * sizeof(int2));
  ws.cand_score = (float*)p;
  p += a256((size_t)kRegions * ws.region_cap * sizeof(float));
  ws.cand_bits = (unsigned long long*)p;
  ws.dbg_scores = nullptr;
  ws.dbg_eps = nullptr;
  return ws;
}

static float* g_dbg_scores = nullptr;
static float* g_dbg_eps = nullptr;

template <typename T>
int tc_impl(const T* meas, int64_t B, int d, const T* cent, int64_t C, const unsigned char* prepared, int32_t base,
            int32_t* idx, double* dist, uint32_t* flags, void* scratch, size_t scratch_bytes_given, cudaStream_t s) {
  QD_REQUIRE(scratch && scratch_bytes_given >= search_scratch_bytes(B), "CVT scratch too small (qd_cvt_scratch_bytes)");
  SearchScratch ws = carve_search(scratch, B);
  ws.dbg_scores = g_dbg_scores;
  ws.dbg_eps = g_dbg_eps;
  int P, ksteps;
  pick_split(d, P, ksteps);
  const int ntiles = (int)((C + kTileN - 1) / kTileN);
  const int qtiles = (int)((B + kTileM - 1) / kTileM);
  // split the centroid range when there are too few query tiles to fill the GPU
  int csplit = 1;
  if (qtiles < 2 * kNumSMs) {
    csplit = (2 * kNumSMs + qtiles - 1) / qtiles;
    if (csplit > ntiles) csplit = ntiles;
  }
  const int tiles_per_cta = (ntiles + csplit - 1) / csplit;
  csplit = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
  static bool attr_set[2] = {false, false};
  const size_t smem = sizeof(SmemLayout) + 1024;
  if (!attr_set[sizeof(T) == 8]) {
    QD_CHECK_CUDA(cudaFuncSetAttribute(cvt_tc_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QD_CHECK_CUDA(cudaFuncSetAttribute(cvt_tc_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[sizeof(T) == 8] = true;
  }
  int rc;
  search_init_kernel<<<grid_for(B > kRegions ? B : kRegions, 256), 256, 0, s>>>(B, ws, idx);
  if ((rc = check_launch("search_init_kernel"))) return rc;
  dim3 grid(qtiles, csplit);
  if (ws.dbg_scores)
    cvt_tc_kernel<T, true><<<grid, kThreads, smem, s>>>(meas, B, d, prepared, tiles_per_cta, ws, flags);
  else
    cvt_tc_kernel<T, false><<<grid, kThreads, smem, s>>>(meas, B, d, prepared, tiles_per_cta, ws, flags);
  if ((rc = check_launch("cvt_tc_kernel"))) return rc;
  if ((rc = launch_rescore<T>(meas, cent, C, d, base, ws, idx, s))) return rc;
  outlier_scan_kernel<T><<<kNumSMs, 256, 0, s>>>(meas, cent, C, d, base, ws, idx);
  if ((rc = check_launch("outlier_scan_kernel"))) return rc;
  search_finish_kernel<<<grid_for(B, 256), 256, 0, s>>>(B, ws, idx, base, dist);
  return check_launch("search_finish_kernel");
}

int cvt_tc_index_of(const void* measures, int dtype, int64_t B, int d, const void* centroids, int64_t C,
                    const void* prepared, int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags,
                    void* scratch, size_t scratch_bytes, cudaStream_t s) {
  if (dtype == QD_F64)
    return tc_impl<double>((const double*)measures, B, d, (const double*)centroids, C, (const unsigned char*)prepared,
                           index_base, idx_out, dist_out, flags, scratch, scratch_bytes, s);
  return tc_impl<float>((const float*)measures, B, d, (const float*)centroids, C, (const unsigned char*)prepared,
                        index_base, idx_out, dist_out, flags, scratch, scratch_bytes, s);
}

}  // namespace qd

extern "C" void qd_cvt_debug_buffers(float* scores, float* eps) {
  qd::g_dbg_scores = scores;
  qd::g_dbg_eps = eps;
}

extern "C" size_t qd_cvt_prepared_bytes(int64_t C, int d) {
  int P, ksteps;
  qd::pick_split(d, P, ksteps);
  const int64_t ntiles = (C + qd::kTileN - 1) / qd::kTileN;
  return qd::kHeaderBytes + (size_t)ntiles * ksteps * qd::kTileN * 32;
}

extern "C" size_t qd_cvt_scratch_bytes(int64_t B, int64_t C, int d) {
  (void)C;
  (void)d;
  return qd::search_scratch_bytes(B);
}

extern "C" int qd_cvt_prepare(const void* centroids, int dtype, int64_t C, int d, void* prepared, void* stream) {
  using namespace qd;
  QD_REQUIRE(centroids && prepared, "null pointer");
  QD_REQUIRE(d >= 1 && d <= QD_MAX_MEASURE_DIM, "measure_dim out of range [1, 32]");
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  QD_REQUIRE(C >= 1 && C < ((int64_t)1 << 31) - kTileN, "centroid count out of range");
  cudaStream_t s = (cudaStream_t)stream;
  PreparedHeader h;
  memset(&h, 0, sizeof(h));
  pick_split(d, h.P, h.ksteps);
  h.ntiles = (int32_t)((C + kTileN - 1) / kTileN);
  h.d = d;
  h.C = C;
  h.scale = 1.0;
  QD_CHECK_CUDA(cudaMemcpyAsync(prepared, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  PreparedHeader* dh = (PreparedHeader*)prepared;
  unsigned char* image = (unsigned char*)prepared + kHeaderBytes;
  const int64_t rows = (int64_t)h.ntiles * kTileN;
  int rc;
  if (dtype == QD_F64) {
    prep_absmax_kernel<double><<<grid_for(C * d, 256), 256, 0, s>>>((const double*)centroids, C * d, dh);
    if ((rc = check_launch("prep_absmax_kernel"))) return rc;
    prep_scale_kernel<<<1, 1, 0, s>>>(dh);
    if ((rc = check_launch("prep_scale_kernel"))) return rc;
    prep_image_kernel<double><<<grid_for(rows, 256), 256, 0, s>>>((const double*)centroids, C, d, dh, image);
  } else {
    prep_absmax_kernel<float><<<grid_for(C * d, 256), 256, 0, s>>>((const float*)centroids, C * d, dh);
    if ((rc = check_launch("prep_absmax_kernel"))) return rc;
    prep_scale_kernel<<<1, 1, 0, s>>>(dh);
    if ((rc = check_launch("prep_scale_kernel"))) return rc;
    prep_image_kernel<float><<<grid_for(rows, 256), 256, 0, s>>>((const float*)centroids, C, d, dh, image);
  }
  return check_launch("prep_image_kernel");
}
