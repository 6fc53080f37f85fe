// A13 / A14 at configs[3] scale (solution_dim = 10 000, batch 4096): the streaming side of the evolution
// strategies (reference ribs/emitters/opt/_sep_cma_es.py:140-191, _lm_ma_es.py:126-194).
//
//   * unit normals: Philox4x32-10 bits through a 4096-layer ziggurat (the method NumPy's
//     Generator.standard_normal uses, numpy/random/src/distributions: one table look-up, one fused multiply-add
//     and one compare for 99.88 % of the samples; wedges and the tail by rejection in fp64, parked per warp and
//     resolved 32 at a time). 51-bit uniforms, N(0,1) law, at a fraction of the instruction count of a
//     double-precision Box-Muller, which made the fill kernel compute-bound (213 us for 41 M normals against
//     50 us of HBM time; this generator: 82 us, bound by the IMAD.WIDE rate of Philox's multiplies).
//   * sep-CMA ask: generator fused with the transform -- X is the only HBM traffic (lambda * n * 8 bytes).
//   * LM-MA ask: the reference applies `itrs` rank-one transforms d <- (1-c_j) d + c_j m_j (m_j . d) one
//     after the other to every sample row (:136-144). With d_j = Q_j (z + sum_{i<j} t_i m_i),
//     Q_j = prod_{l<j} (1-c_l), the coefficients obey
//         t_j = c_j / (1-c_j) * (m_j . z + sum_{i<j} t_i (m_j . m_i)),
//     so the whole ask is P = Z M^T and G = M M^T (tall-skinny products, Z read once), a forward
//     substitution per row, and X = mean + sigma Q_m (Z + T M) (Z read once, X written once). Same
//     arithmetic up to rounding order (1e-15 relative; tests hold it to the oracle at 1e-9); the
//     row-by-row version re-read every m_j for every row out of L2 and took 2.1 ms at 20 vectors.
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "qd_common.cuh"

namespace qd {

// ---------------------------------------------------------------------------------------------
// ziggurat tables (Marsaglia & Tsang 2000): kZigLayers layers of equal area V under exp(-x^2/2). Layer 0 is
// the base strip [0, R] x [0, f(R)] plus the tail beyond R; layer i >= 1 is the rectangle
// [0, x_i] x [f(x_i), f(x_{i-1})], x_0 = 0 < x_1 < ... < x_{N-1} = R. 4096 layers (R = 4.3859) instead of the
// usual 256: a candidate misses the accept-at-once region of its layer with probability 0.12 % instead of
// 1.2 %, and the candidates that do miss are not resolved where they occur (one lane of 32 running the wedge
// code) but parked in a per-warp queue and resolved 32 at a time when the warp has finished its share.
// ---------------------------------------------------------------------------------------------
constexpr int kZigLayers = 4096;
constexpr int kZigThreads = 256;
constexpr int kZigQueue = 64;  // parked candidates per warp; a warp that collects more resolves the excess in place

struct ZigTables {
  double2 wb[kZigLayers];  // (width of layer i, accept-at-once bound x_{i-1}); layer 0: (V / f(R), R)
  double f[kZigLayers];    // density at the layer's right edge; f[0] = 1 (x = 0)
  double r, inv_r;
};

__device__ ZigTables g_zig;

static int zig_tables_ready() {
  static std::mutex mu;
  static bool done[64] = {};
  int dev = 0;
  QD_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && done[dev]) return QD_OK;
  auto dens = [](double x) { return std::exp(-0.5 * x * x); };
  auto area = [&](double R) { return R * dens(R) + std::sqrt(M_PI / 2.0) * std::erfc(R / std::sqrt(2.0)); };
  static double x[kZigLayers];
  // edges for a trial R; returns how far the top layer is from closing at the mode (V / x_1 + f(x_1) = f(0) = 1)
  auto closure = [&](double R) {
    const double V = area(R);
    x[kZigLayers - 1] = R;
    for (int i = kZigLayers - 2; i >= 1; --i) {
      const double a = V / x[i + 1] + dens(x[i + 1]);
      if (a >= 1.0) return 1.0;  // the layers pass the mode before the last one: R too small
      x[i] = std::sqrt(-2.0 * std::log(a));
    }
    return V / x[1] + dens(x[1]) - 1.0;
  };
  double lo = 2.0, hi = 7.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    (closure(mid) > 0.0 ? lo : hi) = mid;
  }
  const double R = 0.5 * (lo + hi);
  QD_REQUIRE(std::fabs(closure(R)) < 1e-9, "ziggurat tables do not close");
  x[0] = 0.0;
  static ZigTables t;
  const double q = area(R) / dens(R);  // virtual width of the base layer (rectangle + tail)
  t.wb[0] = make_double2(q, R);
  t.f[0] = 1.0;
  for (int i = 1; i < kZigLayers; ++i) {
    t.wb[i] = make_double2(x[i], x[i - 1]);
    t.f[i] = dens(x[i]);
  }
  t.r = R;
  t.inv_r = 1.0 / R;
  QD_CHECK_CUDA(cudaMemcpyToSymbol(g_zig, &t, sizeof(t)));
  if (dev < 64) done[dev] = true;
  return QD_OK;
}

// 32 x 32 -> (hi, lo) as one IMAD.WIDE; spelled in PTX so that the front end cannot promote the round's
// xors and key additions to 64-bit arithmetic (which costs two extra adds per multiply after ptxas)
__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%1, %0}, p;\n\t}" : "=r"(hi), "=r"(lo) : "r"(a), "r"(b));
}

__device__ __forceinline__ void philox_round4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  uint32_t hi0, lo0, hi1, lo1;
  mul_wide(0xD2511F53u, c[0], hi0, lo0);
  mul_wide(0xCD9E8D57u, c[2], hi1, lo1);
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}

// the ten round keys of a seed: the same for every counter. The generator kernels take them as a launch
// parameter, so the rounds read them as constant-bank operands: no instructions and no registers
struct PhiloxKeys {
  uint32_t a[10], b[10];
  __host__ __device__ __forceinline__ PhiloxKeys(uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      a[r] = k0 + (uint32_t)r * 0x9E3779B9u;
      b[r] = k1 + (uint32_t)r * 0xBB67AE85u;
    }
  }
};

__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], const PhiloxKeys& k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round4(c, k.a[r], k.b[r]);
}

__device__ __forceinline__ double open01(uint32_t lo, uint32_t hi) {  // (0, 1) from 64 bits
  return ((double)(long long)((((unsigned long long)hi << 32) | lo) >> 11) + 0.5) * 0x1p-53;
}

// A candidate is two Philox words: hi = 12 bits of layer index | the upper 20 mantissa bits, lo = the lower 32
// mantissa bits, taken as they come (no shifts); lo's bit 0 is also the sign, so the uniform has 51 free bits.
__device__ __forceinline__ uint32_t zig_layer(uint32_t hi) { return hi >> 20; }
__device__ __forceinline__ double zig_mant(uint32_t hi, uint32_t lo) {  // [1, 2)
  return __hiloint2double((int)((hi & 0x000FFFFFu) | 0x3FF00000u), (int)lo);
}
__device__ __forceinline__ double zig_signed(uint32_t lo, double x) {
  return __hiloint2double(__double2hiint(x) ^ (int)(lo << 31), __double2loint(x));
}

// A candidate that missed the accept-at-once region of its layer: the wedge test, or the tail beyond R for the
// base layer; a rejected candidate is replaced by fresh ones until one is accepted. Pure function of its
// arguments -- its random words come from Philox sub-streams of the same counter, keyed by the candidate's
// slot -- so the value does not depend on when or by which lane it is resolved.
__device__ __noinline__ double zig_fix(uint32_t lo, uint32_t hi, uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                       uint32_t slot) {
  const PhiloxKeys keys(k0, k1);
  for (uint32_t attempt = 0;; ++attempt) {
    uint32_t c[4] = {c0, c1, slot + 4u * attempt, 0x534C4F57u};  // uniforms of this attempt + the next candidate
    philox4x32_10(c, keys);
    const uint32_t idx = zig_layer(hi);
    const double2 wb = g_zig.wb[idx];
    const double x = fma(zig_mant(hi, lo), wb.x, -wb.x);
    if (x < wb.y) return zig_signed(lo, x);  // (only replacement candidates can land here)
    if (idx == 0) {
      const double R = g_zig.r, inv_r = g_zig.inv_r;
      for (uint32_t t = 0;; ++t) {
        uint32_t d[4] = {c0, c1, slot + 4u * attempt, 0x5441494Cu + t};
        philox4x32_10(d, keys);
        const double xx = -log(open01(d[0], d[1])) * inv_r;
        const double yy = -log(open01(d[2], d[3]));
        if (yy + yy > xx * xx) return zig_signed(lo, R + xx);
      }
    }
    const double fi = g_zig.f[idx], fa = g_zig.f[idx - 1];
    if (fi + open01(c[0], c[1]) * (fa - fi) < exp(-0.5 * x * x)) return zig_signed(lo, x);
    lo = c[2];
    hi = c[3];
  }
}

struct ZigMiss {  // a parked candidate: tag = (counter - launch offset) * 4 + slot
  uint32_t lo, hi;
  uint64_t tag;
};

// shared memory of a generator CTA: the (width, bound) table, then one queue and one fill count per warp
constexpr size_t kZigSmem =
    kZigLayers * sizeof(double2) + (kZigThreads / 32) * (kZigQueue * sizeof(ZigMiss) + sizeof(int));

struct ZigShared {
  const double2* wb;
  ZigMiss* queue;  // this warp's
  int* fill;       // this warp's
};

__device__ __forceinline__ ZigShared zig_shared_setup(unsigned char* smem) {
  double2* wb = reinterpret_cast<double2*>(smem);
  ZigMiss* q = reinterpret_cast<ZigMiss*>(wb + kZigLayers);
  int* fill = reinterpret_cast<int*>(q + (kZigThreads / 32) * kZigQueue);
  for (int i = threadIdx.x; i < kZigLayers; i += blockDim.x) wb[i] = g_zig.wb[i];
  if (threadIdx.x < kZigThreads / 32) fill[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  return {wb, q + warp * kZigQueue, fill + warp};
}

// Four unit normals from one counter: two Philox calls give four candidates, all four take the accept-at-once
// path without a branch; the rare candidate that missed leaves a placeholder in z and is parked.
__device__ __forceinline__ void zig_quad(const PhiloxKeys& keys, uint32_t k0, uint32_t k1, uint64_t ctr, uint64_t offset,
                                         const ZigShared& sh, double (&z)[4]) {
  const uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32);
  uint32_t a[4] = {c0, c1, 0u, 0x5A494731u}, b[4] = {c0, c1, 1u, 0x5A494731u};
  philox4x32_10(a, keys);
  philox4x32_10(b, keys);
  const uint32_t lo[4] = {a[0], a[2], b[0], b[2]}, hi[4] = {a[1], a[3], b[1], b[3]};
  unsigned miss = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 wb =
        *reinterpret_cast<const double2*>(reinterpret_cast<const unsigned char*>(sh.wb) + ((hi[i] >> 16) & 0xFFF0u));
    const double x = fma(zig_mant(hi[i], lo[i]), wb.x, -wb.x);  // (m - 1) * width, rounded once
    if (!(x < wb.y)) miss |= 1u << i;
    z[i] = zig_signed(lo[i], x);
  }
  if (miss) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (miss >> i & 1u) {
        const int slot = atomicAdd(sh.fill, 1);
        if (slot < kZigQueue)
          sh.queue[slot] = {lo[i], hi[i], (ctr - offset) * 4u + (uint64_t)i};
        else
          z[i] = zig_fix(lo[i], hi[i], k0, k1, c0, c1, (uint32_t)i);
      }
  }
}

// Resolves this warp's parked candidates, one per lane, and hands each value to patch(tag / 4, slot, z). Every
// lane of the warp must call it, after its last zig_quad; the placeholders the lanes stored are ordered before
// the patches by the warp barrier.
template <class Patch>
__device__ __forceinline__ void zig_drain(const ZigShared& sh, uint32_t k0, uint32_t k1, uint64_t offset, Patch patch) {
  __syncwarp();
  const int total = min(*sh.fill, kZigQueue);
  for (int e = threadIdx.x & 31; e < total; e += 32) {
    const ZigMiss m = sh.queue[e];
    const uint64_t t = m.tag >> 2, ctr = offset + t;
    const uint32_t slot = (uint32_t)(m.tag & 3u);
    patch(t, (int)slot, zig_fix(m.lo, m.hi, k0, k1, (uint32_t)ctr, (uint32_t)(ctr >> 32), slot));
  }
}

__device__ __forceinline__ void store4_cs(double* p, const double (&v)[4]) {  // 32-byte aligned, streaming
  __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
  __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(v[2], v[3]));
}

// out[4q .. 4q+3] come from counter offset + q
__global__ void __launch_bounds__(kZigThreads) fill_normal_zig_kernel(double* __restrict__ out, int64_t count,
                                                                      const __grid_constant__ PhiloxKeys keys,
                                                                      uint64_t offset) {
  extern __shared__ __align__(16) unsigned char zig_smem[];
  const ZigShared sh = zig_shared_setup(zig_smem);
  const uint32_t k0 = keys.a[0], k1 = keys.b[0];
  const int64_t quads = (count + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < quads; q += (int64_t)gridDim.x * blockDim.x) {
    double z[4];
    zig_quad(keys, k0, k1, offset + (uint64_t)q, offset, sh, z);
    if (4 * q + 3 < count) {
      store4_cs(out + 4 * q, z);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (4 * q + i < count) out[4 * q + i] = z[i];
    }
  }
  zig_drain(sh, k0, k1, offset, [&](uint64_t t, int slot, double zf) {
    const int64_t at = (int64_t)(4 * t) + slot;
    if (at < count) out[at] = zf;
  });
}

// ---------------------------------------------------------------------------------------------
// A13 ask with the generator fused in: X[row] = sqrt(C) * (sigma z) + mean   (_sep_cma_es.py:160-179)
// A thread owns four columns and walks down the rows: sqrt(C) and mean stay in registers.
// Sample k, quad jq draws from counter offset + k * ceil(n / 4) + jq.
// ---------------------------------------------------------------------------------------------
template <bool VEC, bool IDS>  // IDS: rows are redirected through row_ids
__global__ void __launch_bounds__(kZigThreads) sep_ask_rng_kernel(int64_t rows, int64_t n,
                                                                  const int32_t* __restrict__ row_ids,
                                                                  const double* __restrict__ mean,
                                                                  const double* __restrict__ C,
                                                                  const double* __restrict__ sigma_p,
                                                                  const __grid_constant__ PhiloxKeys keys,
                                                                  uint64_t offset, double* __restrict__ X,
                                                                  double* __restrict__ z_out, int mirror,
                                                                  const uint64_t dt /* gridDim.y * ceil(n / 4) */,
                                                                  const int64_t step /* gridDim.y * n */) {
  extern __shared__ __align__(16) unsigned char zig_smem[];
  const ZigShared sh = zig_shared_setup(zig_smem);
  const uint32_t k0 = keys.a[0], k1 = keys.b[0];
  const int64_t qpr = (n + 3) / 4;
  const int64_t jq = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t j0 = jq * 4;
  const double sigma = *sigma_p;
  double sd[4], mu[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool ok = j0 + i < n;
    sd[i] = ok ? sqrt(C[j0 + i]) : 0.0;
    mu[i] = ok ? mean[j0 + i] : 0.0;
  }
  if (jq < qpr) {
    // sample k of this thread: counter ctr, element `at` of a dense (rows, n) array; both advance by launch constants
    uint64_t ctr = offset + (uint64_t)blockIdx.y * (uint64_t)qpr + (uint64_t)jq;
    int64_t at = (int64_t)blockIdx.y * n + j0;
    const int nrows = (int)rows, stride = (int)gridDim.y;
    for (int k = blockIdx.y; k < nrows; k += stride, ctr += dt, at += step) {
      double z[4], x[4];
      zig_quad(keys, k0, k1, ctr, offset, sh, z);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        x[i] = __dadd_rn(__dmul_rn(sd[i], __dmul_rn(sigma, z[i])), mu[i]);  // rng.normal(0, sigma) = 0 + sigma z
      double* xr = X + (IDS ? (int64_t)row_ids[k] * n + j0 : at);
      if (VEC) {  // n % 4 == 0: every quad is whole and 32-byte aligned
        store4_cs(xr, x);
        if (z_out) store4_cs(z_out + at, z);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j0 + i < n) {
            xr[i] = x[i];
            if (z_out) z_out[at + i] = z[i];
          }
      }
      if (mirror) {  // mirrored sampling (_openai_es.py:163-172): sample k + rows is drawn from -z
        double y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = __dadd_rn(__dmul_rn(sd[i], __dmul_rn(sigma, -z[i])), mu[i]);
        double* xm = X + rows * n + at;
        if (VEC) {
          store4_cs(xm, y);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (j0 + i < n) xm[i] = y[i];
        }
      }
    }
  }
  zig_drain(sh, k0, k1, offset, [&](uint64_t t, int slot, double zf) {  // (sample, column) of a parked candidate
    const int64_t k = (int64_t)(t / (uint64_t)qpr), j = 4 * (int64_t)(t % (uint64_t)qpr) + slot;
    if (j >= n) return;
    const double s = sqrt(C[j]), m = mean[j];
    const int64_t row = IDS ? (int64_t)row_ids[k] : k;
    X[row * n + j] = __dadd_rn(__dmul_rn(s, __dmul_rn(sigma, zf)), m);
    if (z_out) z_out[k * n + j] = zf;
    if (mirror) X[(k + rows) * n + j] = __dadd_rn(__dmul_rn(s, __dmul_rn(sigma, -zf)), m);
  });
}

// ---------------------------------------------------------------------------------------------
// OpenAI-ES tell (ribs/emitters/opt/_openai_es.py:180-208 + AdamOpt.step, _adam_opt.py): the gradient estimate
// is a weighted sum of the noise rows (qd_weighted_rows); this kernel scales it, runs the Adam step on theta and
// leaves the block sums of |theta - theta_prev|^2 and |theta|^2 for the update ratio (check_stop).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) openai_tell_kernel(int64_t n, const double* __restrict__ grad_sum, double denom,
                                                          double* __restrict__ theta, double* __restrict__ m,
                                                          double* __restrict__ v, double a, double beta1, double beta2,
                                                          double eps, double l2, double* __restrict__ blocksum) {
  __shared__ double sh[2][8];
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double d2 = 0.0, t2 = 0.0;
  if (j < n) {
    const double th = theta[j];
    double g = -__ddiv_rn(grad_sum[j], denom);          // gradient /= half * sigma0; inverted: we maximise
    g = __dadd_rn(g, __dmul_rn(l2, th));                // L2 regularisation
    const double mn = __dadd_rn(__dmul_rn(beta1, m[j]), __dmul_rn(1.0 - beta1, g));
    const double vn = __dadd_rn(__dmul_rn(beta2, v[j]), __dmul_rn(1.0 - beta2, __dmul_rn(g, g)));
    const double step = __ddiv_rn(__dmul_rn(-a, mn), __dadd_rn(sqrt(vn), eps));
    const double tn = __dadd_rn(th, step);
    m[j] = mn;
    v[j] = vn;
    theta[j] = tn;
    const double df = tn - th;
    d2 = df * df;
    t2 = tn * tn;
  }
  for (int o = 16; o > 0; o >>= 1) {
    d2 += __shfl_down_sync(0xffffffffu, d2, o);
    t2 += __shfl_down_sync(0xffffffffu, t2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = d2;
    sh[1][threadIdx.x >> 5] = t2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a2 = 0.0, b2 = 0.0;
    for (int w = 0; w < 8; ++w) {
      a2 += sh[0][w];
      b2 += sh[1][w];
    }
    blocksum[2 * blockIdx.x] = a2;
    blocksum[2 * blockIdx.x + 1] = b2;
  }
}

// status[4] = |theta - theta_prev| / |theta| of the last step (fixed-order sum of the block sums)
__global__ void openai_ratio_kernel(const double* __restrict__ blocksum, int nblocks, double* __restrict__ status) {
  double a2 = 0.0, b2 = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    a2 += blocksum[2 * b];
    b2 += blocksum[2 * b + 1];
  }
  status[4] = sqrt(a2) / sqrt(b2);
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool pred) {  // !pred: zero fill
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {  // !pred: zero fill
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tall-skinny product: part[split][j][k] = sum over the split's columns c of A[k][c] * M[j][c]
// (A: rows x n, M: m x n, both row-major with leading dimension n) on the fp64 tensor pipe
// (mma.sync m8n8k4: 63.9 FMA/clk/SM measured against 56.9 for DFMA, tools/fp64_peak.cu, and a seventh of
// the shared-memory wavefronts the broadcast-FMA form needed, which is what bounded that one). A CTA owns
// 32 rows x 24 vectors; chunks of 64 columns arrive through a three-stage cp.async ring; each of the eight
// warps multiplies eight columns of every chunk (two k-steps of 4 x 3 MMAs) and the warps' accumulators
// are summed in warp order at the end (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int kTdRows = 32, kTdCols = 64, kTdVecs = 24, kTdStages = 3;
constexpr int kTdLd = kTdCols + 4;  // 68: the 8 x 4 fragment reads of a half-warp fall on distinct banks
constexpr int kTdStageDoubles = (kTdRows + kTdVecs) * kTdLd;
constexpr size_t kTdSmem = kTdStages * (size_t)kTdStageDoubles * sizeof(double);

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <bool WIDE>  // n even and 16-byte aligned operands: the ring is fed 16 bytes per copy instead of 8
__global__ void __launch_bounds__(256, 2) tall_dots_kernel(const double* __restrict__ A, int64_t rows, int64_t n,
                                                           const double* __restrict__ M, int m, int64_t cols_per_split,
                                                           double* __restrict__ part) {
  extern __shared__ __align__(16) double td_smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kTdRows;
  const int split = blockIdx.y;
  const int j0 = blockIdx.z * kTdVecs;
  const int64_t cbeg = split * cols_per_split;
  const int64_t cend = cbeg + cols_per_split < n ? cbeg + cols_per_split : n;
  const int nch = (int)((cend - cbeg + kTdCols - 1) / kTdCols);
  // a thread copies the same (row, column) slots of every chunk: 8 of A, 6 of M. Slot i: element i * 256 + t of
  // the 64-wide tile, i.e. row i * 4 + t / 64, column t % 64 (WIDE: 4 of A and 3 of M, two columns each: row
  // i * 8 + t / 32, columns 2 (t % 32) and the next; split edges and n are even, so a pair is in or out as a whole);
  // pointers and row predicates are loop invariants
  constexpr int kW = WIDE ? 2 : 1, kRowStep = 4 * kW;
  const int lc = WIDE ? 2 * (t & 31) : (t & 63), lr = WIDE ? t >> 5 : t >> 6;
  const double* a_src[kTdRows / kRowStep];
  const double* m_src[kTdVecs / kRowStep];
#pragma unroll
  for (int i = 0; i < kTdRows / kRowStep; ++i) {
    const int64_t row = row0 + i * kRowStep + lr;
    a_src[i] = row < rows ? A + row * n + cbeg + lc : nullptr;
  }
#pragma unroll
  for (int i = 0; i < kTdVecs / kRowStep; ++i) {
    const int j = j0 + i * kRowStep + lr;
    m_src[i] = j < m ? M + (int64_t)j * n + cbeg + lc : nullptr;
  }
  auto issue = [&](int ch) {
    double* As = td_smem + (ch % kTdStages) * kTdStageDoubles + lr * kTdLd + lc;
    double* Ms = As + kTdRows * kTdLd;
    const int64_t off = (int64_t)ch * kTdCols;
    const bool okc = cbeg + off + lc < cend;  // false only in the ragged last chunk
#pragma unroll
    for (int i = 0; i < kTdRows / kRowStep; ++i) {
      const bool ok = okc && a_src[i] != nullptr;
      if (WIDE)
        cp_async16(As + i * kRowStep * kTdLd, ok ? a_src[i] + off : A, ok);
      else
        cp_async8(As + i * kRowStep * kTdLd, ok ? a_src[i] + off : A, ok);
    }
#pragma unroll
    for (int i = 0; i < kTdVecs / kRowStep; ++i) {
      const bool ok = okc && m_src[i] != nullptr;
      if (WIDE)
        cp_async16(Ms + i * kRowStep * kTdLd, ok ? m_src[i] + off : M, ok);
      else
        cp_async8(Ms + i * kRowStep * kTdLd, ok ? m_src[i] + off : M, ok);
    }
  };
  double acc[kTdRows / 8][kTdVecs / 8][2];
#pragma unroll
  for (int rb = 0; rb < kTdRows / 8; ++rb)
#pragma unroll
    for (int nb = 0; nb < kTdVecs / 8; ++nb) acc[rb][nb][0] = acc[rb][nb][1] = 0.0;
#pragma unroll
  for (int st = 0; st < kTdStages - 1; ++st) {
    if (st < nch) issue(st);
    cp_async_commit();
  }
  const int fg = lane >> 2, ft = lane & 3;  // fragment coordinates: A[row fg][col ft], B[k ft][n fg]
  for (int ch = 0; ch < nch; ++ch) {
    cp_async_wait<kTdStages - 2>();
    __syncthreads();  // chunk ch has landed for every thread; the stage chunk ch-1 used is free again
    if (ch + kTdStages - 1 < nch) issue(ch + kTdStages - 1);
    cp_async_commit();
    const double* As = td_smem + (ch % kTdStages) * kTdStageDoubles;
    const double* Ms = As + kTdRows * kTdLd;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = warp * 8 + ks * 4 + ft;
      double a[kTdRows / 8], b[kTdVecs / 8];
#pragma unroll
      for (int rb = 0; rb < kTdRows / 8; ++rb) a[rb] = As[(rb * 8 + fg) * kTdLd + c];
#pragma unroll
      for (int nb = 0; nb < kTdVecs / 8; ++nb) b[nb] = Ms[(nb * 8 + fg) * kTdLd + c];
#pragma unroll
      for (int rb = 0; rb < kTdRows / 8; ++rb)
#pragma unroll
        for (int nb = 0; nb < kTdVecs / 8; ++nb) dmma884(acc[rb][nb], a[rb], b[nb]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  double* buf = td_smem;  // [vector j][row]: kTdVecs x 32 sums; C fragment: row fg, vectors 2 ft, 2 ft + 1
  for (int w = 0; w < 8; ++w) {
    if (warp == w) {
#pragma unroll
      for (int rb = 0; rb < kTdRows / 8; ++rb)
#pragma unroll
        for (int nb = 0; nb < kTdVecs / 8; ++nb)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int o = (nb * 8 + 2 * ft + e) * 32 + rb * 8 + fg;
            buf[o] = (w ? buf[o] : 0.0) + acc[rb][nb][e];
          }
    }
    __syncthreads();
  }
  for (int idx = t; idx < kTdVecs * 32; idx += 256) {
    const int j = idx >> 5, r = idx & 31;
    if (j0 + j < m && row0 + r < rows) part[((int64_t)split * m + j0 + j) * rows + row0 + r] = buf[idx];
  }
}

// out[i] = sum over splits of part[split][i]: one warp per element, fixed reduction tree (deterministic)
__global__ void __launch_bounds__(256) sum_splits_kernel(const double* __restrict__ part, int nsplit, int64_t count,
                                                         double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = w0; i < count; i += nw) {
    double s = 0.0;
    for (int k = lane; k < nsplit; k += 32) s += part[(int64_t)k * count + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[i] = s;
  }
}

// the same for many elements: one thread per element, the splits' loads issued together
__global__ void __launch_bounds__(256) sum_splits_wide_kernel(const double* __restrict__ part, int nsplit, int64_t count,
                                                              double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
#pragma unroll 8
    for (int k = 0; k < nsplit; ++k) s += part[(int64_t)k * count + i];
    out[i] = s;
  }
}

// forward substitution, one thread per sample row: t_j = c_j / (1 - c_j) * (p_j + sum_{i<j} t_i G[j][i]).
// Tt is vector-major (m x rows), so the rows of a warp read and write consecutive addresses. SMEM: the
// row's p / t values, the ratios and G sit in shared memory (m <= kCoefSmemMax), so the recurrence never
// waits on global memory; larger m re-reads Tt and G through L1.
constexpr int kCoefThreads = 128, kCoefSmemMax = 96;

template <bool SMEM>
__global__ void __launch_bounds__(kCoefThreads) lmma_coef_kernel(const double* __restrict__ P /*[m][rows]*/,
                                                                 const double* __restrict__ G,
                                                                 const double* __restrict__ cd, int m, int64_t rows,
                                                                 double* Tt, double* __restrict__ qm_out) {
  extern __shared__ double coef_smem[];
  double* ts = coef_smem;                              // [m][kCoefThreads]
  double* rho = ts + (SMEM ? m * kCoefThreads : 0);    // [m]  c_j / (1 - c_j)
  double* omc = rho + (SMEM ? m : 0);                  // [m]  1 - c_j
  double* gs = omc + (SMEM ? m : 0);                   // [m][m]
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (SMEM) {
    for (int j = threadIdx.x; j < m; j += kCoefThreads) {
      const double c = cd[j];
      omc[j] = 1.0 - c;
      rho[j] = c / (1.0 - c);
    }
    for (int e = threadIdx.x; e < m * m; e += kCoefThreads) gs[e] = G[e];
    if (k < rows) {
#pragma unroll 8
      for (int j = 0; j < m; ++j) ts[j * kCoefThreads + threadIdx.x] = P[(int64_t)j * rows + k];
    }
    __syncthreads();
    if (k == 0) {
      double q = 1.0;
      for (int j = 0; j < m; ++j) q *= omc[j];
      *qm_out = q;
    }
    if (k >= rows) return;
    for (int j = 0; j < m; ++j) {
      double a = ts[j * kCoefThreads + threadIdx.x], b = 0.0;
      const double* g = gs + j * m;
      int i = 0;
      for (; i + 1 < j; i += 2) {
        a = fma(ts[i * kCoefThreads + threadIdx.x], g[i], a);
        b = fma(ts[(i + 1) * kCoefThreads + threadIdx.x], g[i + 1], b);
      }
      if (i < j) a = fma(ts[i * kCoefThreads + threadIdx.x], g[i], a);
      const double tj = rho[j] * (a + b);
      ts[j * kCoefThreads + threadIdx.x] = tj;
      Tt[(int64_t)j * rows + k] = tj;
    }
    return;
  }
  if (k == 0) {
    double q = 1.0;
    for (int j = 0; j < m; ++j) q *= 1.0 - cd[j];
    *qm_out = q;
  }
  if (k >= rows) return;
  for (int j = 0; j < m; ++j) {
    double a = P[(int64_t)j * rows + k], b = 0.0;
    const double* __restrict__ g = G + (int64_t)j * m;
    int i = 0;
    for (; i + 1 < j; i += 2) {
      a = fma(Tt[(int64_t)i * rows + k], __ldg(g + i), a);
      b = fma(Tt[(int64_t)(i + 1) * rows + k], __ldg(g + i + 1), b);
    }
    if (i < j) a = fma(Tt[(int64_t)i * rows + k], __ldg(g + i), a);
    const double c = cd[j];
    Tt[(int64_t)j * rows + k] = c / (1.0 - c) * (a + b);
  }
}

// X[row][c] = mean[c] + sigma * Q_m * (z[k][c] + sum_j T[k][j] M[j][c]) on the fp64 tensor pipe: a CTA (four
// warps, four resident per SM so that their load / MMA / store phases interleave) owns 16 sample rows x 256
// columns, a warp 16 x 64 of them (2 x 8 accumulator tiles); k runs over the direction
// vectors, four per MMA. Every m_j value of the CTA's columns is used by exactly one lane, so B fragments are
// plain global loads (L2-resident M); the z tile streams in through cp.async while the MMAs run.
constexpr int kLoRows = 16, kLoThreads = 128, kLoCols = 2 * kLoThreads;
constexpr int kLoLd = kLoCols + 8;  // 264: conflict-free 16-byte fragment reads

template <bool VEC2>  // n even: rows stay 16-byte aligned, X leaves as double2
__global__ void __launch_bounds__(kLoThreads, 4) lmma_out_kernel(const double* __restrict__ z, int64_t rows, int64_t n,
                                                          const int32_t* __restrict__ row_ids,
                                                          const double* __restrict__ mean,
                                                          const double* __restrict__ status, const double* __restrict__ M,
                                                          int m, const double* __restrict__ Tt,
                                                          const double* __restrict__ qm_p, double* __restrict__ X) {
  extern __shared__ __align__(16) double zs[];  // [kLoRows][kLoLd]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t col0 = blockIdx.x * (int64_t)kLoCols;
  const int64_t row0 = (int64_t)blockIdx.y * kLoRows;
#pragma unroll
  for (int r = 0; r < kLoRows; ++r) {
    const bool okr = row0 + r < rows;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t c = col0 + h * kLoThreads + t;
      const bool ok = okr && c < n;
      cp_async8(&zs[r * kLoLd + h * kLoThreads + t], ok ? z + (row0 + r) * n + c : z, ok);
    }
  }
  cp_async_commit();
  const int fg = lane >> 2, ft = lane & 3;
  double acc[2][8][2];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
  const int64_t cw = col0 + warp * 64;  // first column of the warp
  // fragments of k-step j4: A[row fg][k ft] = T, B[k ft][n fg] = M; the next step's loads are issued before
  // this step's MMAs (the loop bound is a run-time value, so the compiler does not pipeline it itself)
  auto load_frags = [&](int j4, double (&a)[2], double (&b)[8]) {
    const int j = j4 + ft;
    const bool okj = j < m;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const int64_t row = row0 + mb * 8 + fg;
      a[mb] = (okj && row < rows) ? __ldg(Tt + (int64_t)j * rows + row) : 0.0;
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int64_t c = cw + nb * 8 + fg;
      b[nb] = (okj && c < n) ? __ldg(M + (int64_t)j * n + c) : 0.0;
    }
  };
  double a0[2], b0[8];
  if (m > 0) load_frags(0, a0, b0);
  for (int j4 = 0; j4 < m; j4 += 4) {
    double a1[2], b1[8];
    load_frags(j4 + 4, a1, b1);  // past the end: all predicates false, zeros
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) dmma884(acc[mb][nb], a0[mb], b0[nb]);
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) a0[mb] = a1[mb];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) b0[nb] = b1[nb];
  }
  cp_async_wait<0>();
  __syncthreads();
  const double sigma = status[0];
  const double qm = m > 0 ? *qm_p : 1.0;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb) {
    const int r = mb * 8 + fg;
    if (row0 + r >= rows) continue;
    const int64_t row = row_ids ? row_ids[row0 + r] : row0 + r;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const int cl = warp * 64 + nb * 8 + 2 * ft;  // C fragment: row fg, columns 2 ft, 2 ft + 1
      const int64_t c = col0 + cl;
      if (c >= n) continue;
      const double2 zz = *reinterpret_cast<const double2*>(&zs[r * kLoLd + cl]);
      const double d0 = m > 0 ? __dmul_rn(qm, __dadd_rn(zz.x, acc[mb][nb][0])) : zz.x;
      const double d1 = m > 0 ? __dmul_rn(qm, __dadd_rn(zz.y, acc[mb][nb][1])) : zz.y;
      const double x0 = __dadd_rn(mean[c], __dmul_rn(sigma, d0));  // mean + sigma * d  (:145)
      if (VEC2) {
        const double x1 = __dadd_rn(mean[c + 1], __dmul_rn(sigma, d1));
        __stcs(reinterpret_cast<double2*>(X + row * n + c), make_double2(x0, x1));
      } else {
        X[row * n + c] = x0;
        if (c + 1 < n) X[row * n + c + 1] = __dadd_rn(mean[c + 1], __dmul_rn(sigma, d1));
      }
    }
  }
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been opted in on
template <class K>
static int opt_in_smem(K kernel, size_t bytes, bool (&done)[64]) {
  int dev = 0;
  QD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return QD_OK;
  QD_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (dev < 64) done[dev] = true;
  return QD_OK;
}

// A generator kernel runs as one wave of resident CTAs (each stages the 64 KB layer table once and then strides
// over its share): the CTA count a device holds at once, queried on first use
template <class K>
static int zig_resident_ctas(K kernel, int (&ctas)[64], bool (&attr)[64], int* out) {
  int rc = opt_in_smem(kernel, kZigSmem, attr);
  if (rc) return rc;
  int dev = 0;
  QD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || ctas[dev] == 0) {
    int per_sm = 0, sms = 0;
    QD_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kZigThreads, kZigSmem));
    QD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    QD_REQUIRE(per_sm >= 1 && sms >= 1, "generator kernel does not fit an SM");
    *out = per_sm * sms;
    if (dev < 64) ctas[dev] = *out;
    return QD_OK;
  }
  *out = ctas[dev];
  return QD_OK;
}

static int launch_fill_zig(double* out, int64_t count, uint64_t seed, uint64_t offset, cudaStream_t s) {
  static int ctas[64] = {};
  static bool attr[64] = {};
  int resident = 0;
  int rc = zig_resident_ctas(fill_normal_zig_kernel, ctas, attr, &resident);
  if (rc) return rc;
  const int64_t want = ((count + 3) / 4 + kZigThreads - 1) / kZigThreads;
  fill_normal_zig_kernel<<<(unsigned)(want < resident ? want : resident), kZigThreads, kZigSmem, s>>>(
      out, count, PhiloxKeys((uint32_t)seed, (uint32_t)(seed >> 32)), offset);
  return check_launch("fill_normal_zig_kernel");
}

// launch geometry of the two tall-skinny products, shared by the workspace query and the launcher
struct TdPlan {
  int row_tiles, groups, nsplit;
  int64_t cols_per_split;
};

static TdPlan td_plan(int64_t rows, int64_t n, int64_t m) {
  TdPlan p;
  p.row_tiles = (int)((rows + kTdRows - 1) / kTdRows);
  p.groups = (int)((m + kTdVecs - 1) / kTdVecs);
  const int64_t chunks = (n + kTdCols - 1) / kTdCols;
  int64_t want = (4 * (int64_t)kNumSMs + (int64_t)p.row_tiles * p.groups - 1) / ((int64_t)p.row_tiles * p.groups);
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  if (want > chunks) want = chunks;
  const int64_t cps = (chunks + want - 1) / want;  // chunks per split
  p.nsplit = (int)((chunks + cps - 1) / cps);
  p.cols_per_split = cps * kTdCols;
  return p;
}

// Sample rows are independent given G = M M^T, so a large ask runs as row blocks that alternate between two
// side streams: the ALU-bound generator, the tensor-bound product and the HBM-bound output pass of
// neighbouring blocks overlap. Measured at n = 10 000, lambda = 4096 (327 MB of z): blocks of 160 MB (two
// blocks) 0.39-0.45 ms per ask, one block 0.40-0.46 ms, 40 MB blocks (L2-sized, eight blocks) 0.49-0.56 ms --
// the extra launches cost more than reading z back from L2 instead of HBM saves.
constexpr int64_t kLmmaBlockBytes = 160ll << 20;
constexpr int kLmmaSlots = 2;

static std::atomic<int64_t> g_lmma_block_bytes{0};  // 0: not set yet

static int64_t lmma_block_bytes() {  // QD_LMMA_BLOCK_MB / qd_lmma_set_block_bytes override the default
  int64_t v = g_lmma_block_bytes.load(std::memory_order_relaxed);
  if (v == 0) {
    const char* e = getenv("QD_LMMA_BLOCK_MB");
    const long mb = e ? atol(e) : 0;
    v = mb > 0 ? (int64_t)mb << 20 : kLmmaBlockBytes;
    g_lmma_block_bytes.store(v, std::memory_order_relaxed);
  }
  return v;
}

static int64_t lmma_block_rows(int64_t rows, int64_t n) {
  int64_t rb = lmma_block_bytes() / (n * (int64_t)sizeof(double)) / 32 * 32;
  if (rb < 32) rb = 32;
  return rb < rows ? rb : rows;
}

struct LmmaWs {  // offsets in doubles; the block part repeats once per slot
  size_t gpart, g, block0, ppart, p, tt, qm, block_stride, total;
};

static LmmaWs lmma_ws(int64_t rows, int64_t n, int64_t m) {
  LmmaWs w;
  const int64_t rb = lmma_block_rows(rows, n);
  const TdPlan pg = td_plan(m, n, m);
  const int64_t last = rows - (rows - 1) / rb * rb;  // rows of the final (possibly partial) block
  const size_t pp_full = (size_t)td_plan(rb, n, m).nsplit * rb, pp_last = (size_t)td_plan(last, n, m).nsplit * last;
  size_t off = 0;
  w.gpart = off;
  off += (size_t)pg.nsplit * m * m;
  w.g = off;
  off += (size_t)m * m;
  w.block0 = off;
  size_t b = 0;
  w.ppart = b;
  b += (pp_full > pp_last ? pp_full : pp_last) * m;
  w.p = b;
  b += (size_t)m * rb;
  w.tt = b;
  b += (size_t)m * rb;
  w.qm = b;
  b += 8;
  w.block_stride = b;
  w.total = off + kLmmaSlots * b;
  return w;
}

struct SideStreams {
  cudaStream_t s[kLmmaSlots];
  cudaEvent_t fork, gram, join[kLmmaSlots];
  bool ready = false;
};

static int side_streams(SideStreams** out) {
  static SideStreams pool[64];
  int dev = 0;
  QD_CHECK_CUDA(cudaGetDevice(&dev));
  QD_REQUIRE(dev < 64, "device ordinal out of range");
  SideStreams& p = pool[dev];
  if (!p.ready) {
    for (int i = 0; i < kLmmaSlots; ++i) {
      QD_CHECK_CUDA(cudaStreamCreateWithFlags(&p.s[i], cudaStreamNonBlocking));
      QD_CHECK_CUDA(cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming));
    }
    QD_CHECK_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    QD_CHECK_CUDA(cudaEventCreateWithFlags(&p.gram, cudaEventDisableTiming));
    p.ready = true;
  }
  *out = &p;
  return QD_OK;
}

static int launch_tall_dots(const double* A, int64_t rows, int64_t n, const double* M, int64_t m, double* part,
                            cudaStream_t s) {
  const TdPlan p = td_plan(rows, n, m);
  QD_REQUIRE(p.groups <= 65535 && p.nsplit <= 65535, "too many direction vectors");
  static bool attr[64] = {}, attr_wide[64] = {};
  const bool wide = n % 2 == 0 && (((uintptr_t)A | (uintptr_t)M) & 15u) == 0;
  int rc0 = wide ? opt_in_smem(tall_dots_kernel<true>, kTdSmem, attr_wide) : opt_in_smem(tall_dots_kernel<false>, kTdSmem, attr);
  if (rc0) return rc0;
  const dim3 grid(p.row_tiles, p.nsplit, p.groups);
  if (wide)
    tall_dots_kernel<true><<<grid, 256, kTdSmem, s>>>(A, rows, n, M, (int)m, p.cols_per_split, part);
  else
    tall_dots_kernel<false><<<grid, 256, kTdSmem, s>>>(A, rows, n, M, (int)m, p.cols_per_split, part);
  return check_launch("tall_dots_kernel");
}

}  // namespace qd

extern "C" int qd_fill_normal_zig(double* out, int64_t count, uint64_t seed, uint64_t offset, void* stream) {
  using namespace qd;
  if (count == 0) return QD_OK;
  QD_REQUIRE(out && count > 0, "bad arguments");
  int rc = zig_tables_ready();
  if (rc) return rc;
  return launch_fill_zig(out, count, seed, offset, (cudaStream_t)stream);
}

static int sep_ask_rng_impl(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean, const double* C,
                            const double* status, uint64_t seed, uint64_t offset, double* X, double* z_out,
                            int mirror, void* stream);

extern "C" int qd_sep_ask_rng(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean, const double* C,
                              const double* status, uint64_t seed, uint64_t offset, double* X, double* z_out,
                              void* stream) {
  return sep_ask_rng_impl(rows, n, row_ids, mean, C, status, seed, offset, X, z_out, 0, stream);
}

extern "C" int qd_mirror_ask_rng(int64_t half_rows, int64_t n, const double* mean, const double* scale,
                                 const double* status, uint64_t seed, uint64_t offset, double* X, double* z_out,
                                 void* stream) {
  return sep_ask_rng_impl(half_rows, n, nullptr, mean, scale, status, seed, offset, X, z_out, 1, stream);
}

extern "C" int qd_openai_tell(int64_t n, const double* grad_sum, double denom, double* theta, double* m, double* v,
                              double a, double beta1, double beta2, double eps, double l2, double* status,
                              double* workspace, void* stream) {
  using namespace qd;
  QD_REQUIRE(n >= 1 && grad_sum && theta && m && v && status && workspace && denom != 0.0, "bad arguments");
  const int nb = (int)((n + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  openai_tell_kernel<<<nb, 256, 0, s>>>(n, grad_sum, denom, theta, m, v, a, beta1, beta2, eps, l2, workspace);
  int rc = check_launch("openai_tell_kernel");
  if (rc) return rc;
  openai_ratio_kernel<<<1, 1, 0, s>>>(workspace, nb, status);
  return check_launch("openai_ratio_kernel");
}

static int sep_ask_rng_impl(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean, const double* C,
                            const double* status, uint64_t seed, uint64_t offset, double* X, double* z_out,
                            int mirror, void* stream) {
  using namespace qd;
  if (rows == 0) return QD_OK;
  QD_REQUIRE(mean && C && status && X && rows > 0 && n > 0, "bad arguments");
  int rc = zig_tables_ready();
  if (rc) return rc;
  const int64_t qpr = (n + 3) / 4;
  const int64_t gx = (qpr + kZigThreads - 1) / kZigThreads;
  // one wave of resident CTAs: a thread walks rows / gy samples with its columns' sqrt(C) and mean in registers
  static int ctas[4][64] = {};
  static bool attr[4][64] = {};
  const int variant = (n % 4 == 0 ? 2 : 0) + (row_ids ? 1 : 0);
  using Kernel = decltype(&sep_ask_rng_kernel<true, true>);
  const Kernel kernels[4] = {sep_ask_rng_kernel<false, false>, sep_ask_rng_kernel<false, true>,
                             sep_ask_rng_kernel<true, false>, sep_ask_rng_kernel<true, true>};
  int resident = 0;
  if ((rc = zig_resident_ctas(kernels[variant], ctas[variant], attr[variant], &resident))) return rc;
  QD_REQUIRE(gx <= 0x7fffffff && rows <= 0x7fffffff, "sizes out of range");
  int64_t gy = resident / gx;
  if (gy < 1) gy = 1;
  if (gy > rows) gy = rows;
  if (gy > 65535) gy = 65535;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  kernels[variant]<<<grid, kZigThreads, kZigSmem, (cudaStream_t)stream>>>(
      rows, n, row_ids, mean, C, status, PhiloxKeys((uint32_t)seed, (uint32_t)(seed >> 32)), offset, X, z_out, mirror,
      (uint64_t)gy * (uint64_t)qpr, gy * n);
  return check_launch("sep_ask_rng_kernel");
}

extern "C" int64_t qd_lmma_set_block_bytes(int64_t bytes) {
  const int64_t prev = qd::lmma_block_bytes();
  qd::g_lmma_block_bytes.store(bytes > 0 ? bytes : qd::kLmmaBlockBytes, std::memory_order_relaxed);
  return prev;
}

extern "C" size_t qd_lmma_ask_workspace_doubles(int64_t rows, int64_t n, int64_t itrs) {
  if (rows <= 0 || n <= 0 || itrs <= 0) return 8;
  return qd::lmma_ws(rows, n, itrs).total;
}

namespace qd {

// P, forward substitution and the output pass for one block of sample rows (G is ready in ws_g)
static int lmma_block(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                      const double* status, const double* M, const double* cd, int m, double* X, const double* ws_g,
                      double* wb, const LmmaWs& w, cudaStream_t s) {
  const double* Tt = nullptr;
  const double* qm = nullptr;
  int rc;
  if (m > 0) {
    if ((rc = launch_tall_dots(z, rows, n, M, m, wb + w.ppart, s))) return rc;  // P = Z M^T
    const int nsp = td_plan(rows, n, m).nsplit;
    sum_splits_wide_kernel<<<grid_for((int64_t)m * rows, 256), 256, 0, s>>>(wb + w.ppart, nsp, (int64_t)m * rows,
                                                                            wb + w.p);
    if ((rc = check_launch("sum_splits_wide_kernel"))) return rc;
    static bool attr[64] = {};
    if ((rc = opt_in_smem(lmma_coef_kernel<true>, 200 * 1024, attr))) return rc;
    const unsigned cgrid = (unsigned)((rows + kCoefThreads - 1) / kCoefThreads);
    const size_t tsm = ((size_t)m * kCoefThreads + 2 * m + (size_t)m * m) * sizeof(double);
    if (m <= kCoefSmemMax)
      lmma_coef_kernel<true><<<cgrid, kCoefThreads, tsm, s>>>(wb + w.p, ws_g, cd, m, rows, wb + w.tt, wb + w.qm);
    else
      lmma_coef_kernel<false><<<cgrid, kCoefThreads, 0, s>>>(wb + w.p, ws_g, cd, m, rows, wb + w.tt, wb + w.qm);
    if ((rc = check_launch("lmma_coef_kernel"))) return rc;
    Tt = wb + w.tt;
    qm = wb + w.qm;
  }
  const dim3 grid((unsigned)((n + kLoCols - 1) / kLoCols), (unsigned)((rows + kLoRows - 1) / kLoRows));
  QD_REQUIRE(grid.y <= 65535, "too many sample rows");
  constexpr size_t zsm = (size_t)kLoRows * kLoLd * sizeof(double);
  static bool attr_even[64] = {}, attr_odd[64] = {};
  if ((rc = opt_in_smem(lmma_out_kernel<true>, zsm, attr_even)) || (rc = opt_in_smem(lmma_out_kernel<false>, zsm, attr_odd)))
    return rc;
  if (n % 2 == 0)
    lmma_out_kernel<true><<<grid, kLoThreads, zsm, s>>>(z, rows, n, row_ids, mean, status, M, m, Tt, qm, X);
  else
    lmma_out_kernel<false><<<grid, kLoThreads, zsm, s>>>(z, rows, n, row_ids, mean, status, M, m, Tt, qm, X);
  return check_launch("lmma_out_kernel");
}

// generate != 0: z (rows x n) is filled first, block by block, with the stream qd_fill_normal_zig(z, rows * n, seed,
// offset) produces
static int lmma_ask_impl(double* z, int generate, uint64_t seed, uint64_t offset, int64_t rows, int64_t n,
                         const int32_t* row_ids, const double* mean, const double* status, const double* M,
                         const double* cd, int64_t itrs, double* X, double* workspace, cudaStream_t s) {
  if (rows == 0) return QD_OK;
  QD_REQUIRE(z && mean && status && X && rows > 0 && n > 0 && itrs >= 0, "bad arguments");
  QD_REQUIRE(itrs == 0 || (M && cd && workspace), "null pointer");
  QD_REQUIRE(itrs < ((int64_t)1 << 20) && rows < ((int64_t)1 << 31), "sizes out of range");
  int rc;
  if (generate && (rc = zig_tables_ready())) return rc;
  const int m = (int)itrs;
  LmmaWs w = {};
  double* ws = workspace;
  if (m > 0) w = lmma_ws(rows, n, m);
  const int64_t rb = lmma_block_rows(rows, n);
  const int64_t nblocks = (rows + rb - 1) / rb;
  SideStreams* side = nullptr;
  if (nblocks > 1) {  // the side streams start (with their generators) before G, which only the substitution needs
    if ((rc = side_streams(&side))) return rc;
    QD_CHECK_CUDA(cudaEventRecord(side->fork, s));
    for (int i = 0; i < kLmmaSlots; ++i) QD_CHECK_CUDA(cudaStreamWaitEvent(side->s[i], side->fork, 0));
  }
  if (m > 0) {
    if ((rc = launch_tall_dots(M, m, n, M, m, ws + w.gpart, s))) return rc;  // G = M M^T
    sum_splits_kernel<<<grid_for((int64_t)m * m * 32, 256), 256, 0, s>>>(ws + w.gpart, td_plan(m, n, m).nsplit,
                                                                   (int64_t)m * m, ws + w.g);
    if ((rc = check_launch("sum_splits_kernel"))) return rc;
    if (side) QD_CHECK_CUDA(cudaEventRecord(side->gram, s));
  }
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t r0 = b * rb, nr = rows - r0 < rb ? rows - r0 : rb;
    const int slot = (int)(b % kLmmaSlots);
    cudaStream_t bs = side ? side->s[slot] : s;
    double* zb = z + r0 * n;
    if (generate) {  // r0 * n is a multiple of 4 (r0 of 32): the block's quads are the flat fill's quads
      if ((rc = launch_fill_zig(zb, nr * n, seed, offset + (uint64_t)(r0 * n / 4), bs))) return rc;
    }
    if (side && m > 0 && b < kLmmaSlots) QD_CHECK_CUDA(cudaStreamWaitEvent(bs, side->gram, 0));
    rc = lmma_block(zb, nr, n, row_ids ? row_ids + r0 : nullptr, mean, status, M, cd, m, row_ids ? X : X + r0 * n,
                    m > 0 ? ws + w.g : nullptr, m > 0 ? ws + w.block0 + slot * w.block_stride : nullptr, w, bs);
    if (rc) return rc;
  }
  if (side) {
    for (int i = 0; i < kLmmaSlots; ++i) {
      QD_CHECK_CUDA(cudaEventRecord(side->join[i], side->s[i]));
      QD_CHECK_CUDA(cudaStreamWaitEvent(s, side->join[i], 0));
    }
  }
  return QD_OK;
}

}  // namespace qd

extern "C" int qd_lmma_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                           const double* status, const double* M, const double* cd, int64_t itrs, double* X,
                           double* workspace, void* stream) {
  return qd::lmma_ask_impl(const_cast<double*>(z), 0, 0, 0, rows, n, row_ids, mean, status, M, cd, itrs, X, workspace,
                           (cudaStream_t)stream);
}

extern "C" int qd_lmma_ask_rng(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                               const double* status, const double* M, const double* cd, int64_t itrs, uint64_t seed,
                               uint64_t offset, double* z_out, double* X, double* workspace, void* stream) {
  return qd::lmma_ask_impl(z_out, 1, seed, offset, rows, n, row_ids, mean, status, M, cd, itrs, X, workspace,
                           (cudaStream_t)stream);
}
