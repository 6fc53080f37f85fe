// A4-A8: the body of GridArchive.add / CVTArchive.add after index_of
// (reference ribs/archives/_grid_archive.py:623-712 == _cvt_archive.py:818-907, formerly
// batch_entries_with_threshold; store side ribs/archives/_array_store.py:535-608).
//
// Sort-free, deterministic design. Per-cell scratch in HBM (clean between calls):
//   best[c]  u64  order key of the highest insertable objective          (atomicMax)
//   win[c]   i32  lowest batch row holding that objective                (atomicMin)
//   cnt[c]   u32  number of insertable rows (CMA-MAE)                    (atomicAdd)
//   bins[c]  3xf64 exactly-summable split of the insertable objectives   (atomicAdd, exact)
// Passes over the batch (HBM-bound, B rows):
//   classify  (grid: fused with index_of) status / value vs the PRE-batch store; best, cnt, batch
//             |objective| max; an equal-objective collision in a cell raises the tie flag
//   resolve   only when needed: first-wins tie break (tie flag) and CMA-MAE objective sums
//   select    decides the winner of every touched cell and appends (row, cell) to a compact list
//   update    one thread per winner: threshold, occupancy, objective-sum delta, batch best
//   copy      flat HBM stream moving the winners' row fields into the store
//   bestcell  over the CELLS, only when the batch beat the archive's previous best objective:
//             lowest cell holding the new global best
//   append    new cells appended to occupied_list in ascending cell order (bitmap scan)
// A non-finite objective or measure (flag set by this or the index kernel) turns commit into
// a scratch clean-up, so the store is untouched when the host raises ValueError
// (ribs/_utils.py:18-31).
#include "grid_common.cuh"

namespace qd {

struct InsertGlobals {
  unsigned long long gbest_key;
  unsigned long long absmax_bits;
  int32_t best_cell;
  uint32_t n_insertable;
  uint32_t n_new;
  uint32_t flags;
  uint32_t tie_flag;  // two insertable rows of one cell share the cell's best objective
  uint32_t n_winners;
  double dsum[3];
  double store_absmax;  // persistent: max |objective| ever offered to the store
};

struct Scratch {
  unsigned long long* best;
  double* bins;  // 3 * capacity
  int32_t* win;
  uint32_t* cnt;
  uint32_t* newmask;
  uint32_t* blockcnt;
  int2* wlist;  // (batch row, cell) of this call's winners, at most one per cell
  InsertGlobals* g;
};

constexpr int kAppendThreads = 1024;
constexpr int kWordsPerThread = 4;
constexpr int kWordsPerBlock = kAppendThreads * kWordsPerThread;

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Scratch carve(void* base, int64_t N) {
  Scratch s;
  char* p = (char*)base;
  s.best = (unsigned long long*)p;
  p += align_up((size_t)N * 8, 256);
  s.bins = (double*)p;
  p += align_up((size_t)N * 24, 256);
  s.win = (int32_t*)p;
  p += align_up((size_t)N * 4, 256);
  s.cnt = (uint32_t*)p;
  p += align_up((size_t)N * 4, 256);
  int64_t words = (N + 31) / 32;
  s.newmask = (uint32_t*)p;
  p += align_up((size_t)words * 4, 256);
  int64_t nblk = (words + kWordsPerBlock - 1) / kWordsPerBlock;
  s.blockcnt = (uint32_t*)p;
  p += align_up((size_t)nblk * 4, 256);
  s.wlist = (int2*)p;
  p += align_up((size_t)N * 8, 256);
  s.g = (InsertGlobals*)p;
  return s;
}

static size_t scratch_bytes(int64_t N) {
  int64_t words = (N + 31) / 32;
  int64_t nblk = (words + kWordsPerBlock - 1) / kWordsPerBlock;
  return align_up((size_t)N * 8, 256) + align_up((size_t)N * 24, 256) + 2 * align_up((size_t)N * 4, 256) +
         align_up((size_t)words * 4, 256) + align_up((size_t)nblk * 4, 256) + align_up((size_t)N * 8, 256) +
         align_up(sizeof(InsertGlobals), 256);
}

__global__ void scratch_init_kernel(Scratch s, int64_t N) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = i; c < N; c += stride) {
    s.best[c] = 0ull;
    s.bins[3 * c] = s.bins[3 * c + 1] = s.bins[3 * c + 2] = 0.0;
    s.win[c] = INT32_MAX;
    s.cnt[c] = 0u;
  }
  for (int64_t w = i; w < (N + 31) / 32; w += stride) s.newmask[w] = 0u;
  if (i == 0) {
    InsertGlobals g;
    g.gbest_key = 0ull;
    g.absmax_bits = 0ull;
    g.best_cell = INT32_MAX;
    g.n_insertable = g.n_new = g.flags = g.tie_flag = g.n_winners = 0u;
    g.dsum[0] = g.dsum[1] = g.dsum[2] = 0.0;
    g.store_absmax = 0.0;
    *s.g = g;
  }
}

// ---------------------------------------------------------------------------------------
template <typename T, typename MT, int D, bool FUSED>
__global__ void __launch_bounds__(256) classify_kernel(int64_t B, int32_t* __restrict__ idx,
                                                       const MT* __restrict__ meas,
                                                       const __grid_constant__ GridParams gp,
                                                       const T* __restrict__ obj,
                                                       const uint8_t* __restrict__ occupied,
                                                       const T* __restrict__ thr, T tmin, bool cma_me, bool mae,
                                                       int32_t* __restrict__ status, T* __restrict__ value,
                                                       Scratch s) {
  uint32_t ncan = 0;
  double amax = 0.0;
  bool bad = false, bad_meas = false, tie = false;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    int32_t c;
    if constexpr (FUSED) {
      c = grid_cell_of<MT, D>(meas, i, gp, bad_meas);  // A2 fused into the first pass
      idx[i] = c;
    } else {
      c = idx[i];
    }
    const T o = obj[i];
    const bool fin = sizeof(T) == 8 ? finite_d((double)o) : finite_f((float)o);
    bad |= !fin;
    const bool occ = occupied[c] != 0;
    const T t = occ ? thr[c] : tmin;  // _grid_archive.py:628-630
    const bool can = fin && (o > t);  // :636
    status[i] = can ? (occ ? 1 : 2) : 0;
    const T tval = (can && !occ) ? (cma_me ? (T)0 : tmin) : t;  // :646-648
    value[i] = o - tval;                                        // :649
    if (can) {
      const unsigned long long key = ord_key((double)o);
      const unsigned long long old = atomicMax(&s.best[c], key);
      tie |= (old == key);
      if (mae) atomicAdd(&s.cnt[c], 1u);
      ++ncan;
    }
    if (fin) amax = fmax(amax, fabs((double)o));
  }
  if (tie) s.g->tie_flag = 1u;
  if (bad_meas) atomicOr(&s.g->flags, QD_FLAG_NONFINITE_MEASURES);
  // block reduction
  __shared__ uint32_t sh_n[8];
  __shared__ double sh_a[8];
  __shared__ int sh_bad;
  if (threadIdx.x == 0) sh_bad = 0;
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    ncan += __shfl_down_sync(0xffffffffu, ncan, o2);
    amax = fmax(amax, __shfl_down_sync(0xffffffffu, amax, o2));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    sh_n[threadIdx.x >> 5] = ncan;
    sh_a[threadIdx.x >> 5] = amax;
  }
  if (bad) sh_bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t n = 0;
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      n += sh_n[w];
      a = fmax(a, sh_a[w]);
    }
    if (n) atomicAdd(&s.g->n_insertable, n);
    if (a > 0.0) atomicMax(&s.g->absmax_bits, (unsigned long long)__double_as_longlong(a));
    if (sh_bad) atomicOr(&s.g->flags, QD_FLAG_NONFINITE_OBJECTIVE);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) resolve_kernel(int64_t B, const int32_t* __restrict__ idx,
                                                      const T* __restrict__ obj,
                                                      const int32_t* __restrict__ status, bool mae, Scratch s) {
  if (s.g->flags) return;
  if (s.g->n_insertable == 0) return;
  if (!mae && !s.g->tie_flag) return;  // unique cell maxima: commit identifies winners by their key
  BinConst bc;
  if (mae) bc = make_bins(__longlong_as_double((long long)s.g->absmax_bits), B);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    if (status[i] == 0) continue;
    const int32_t c = idx[i];
    const double o = (double)obj[i];
    if (ord_key(o) == s.best[c]) atomicMin(&s.win[c], (int32_t)i);  // first in batch wins ties (:693-696)
    if (mae) {
      double h0, h1, h2;
      bin_split(o, bc, h0, h1, h2);
      double* b = &s.bins[3 * (int64_t)c];
      if (h0 != 0.0) atomicAdd(b, h0);
      if (h1 != 0.0) atomicAdd(b + 1, h1);
      if (h2 != 0.0) atomicAdd(b + 2, h2);
    }
  }
}

struct FieldCopy {
  int n;
  const char* src[QD_MAX_FIELDS];
  char* dst[QD_MAX_FIELDS];
  int64_t row_bytes[QD_MAX_FIELDS];
  int vec[QD_MAX_FIELDS];  // 16, 8, 4 or 1
};

__device__ __forceinline__ void warp_copy_row(const char* __restrict__ src, char* __restrict__ dst, int64_t bytes,
                                              int vec, int lane) {
  if (vec == 16) {
    for (int64_t o = lane * 16; o < bytes; o += 32 * 16)
      *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(src + o);
  } else if (vec == 8) {
    for (int64_t o = lane * 8; o < bytes; o += 32 * 8)
      *reinterpret_cast<uint2*>(dst + o) = *reinterpret_cast<const uint2*>(src + o);
  } else if (vec == 4) {
    for (int64_t o = lane * 4; o < bytes; o += 32 * 4)
      *reinterpret_cast<uint32_t*>(dst + o) = *reinterpret_cast<const uint32_t*>(src + o);
  } else {
    for (int64_t o = lane; o < bytes; o += 32) dst[o] = src[o];
  }
}

// select: one pass over the rows that only decides who won and appends (row, cell) to the winner list
template <typename T>
__global__ void __launch_bounds__(256) select_kernel(int64_t B, const int32_t* __restrict__ idx,
                                                     const T* __restrict__ obj, const int32_t* __restrict__ status,
                                                     bool mae, Scratch s) {
  const int lane = threadIdx.x & 31;
  if (s.g->n_insertable == 0) return;
  if (s.g->flags) {  // validation failed: leave the store alone, clean the scratch
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
      if (status[i] == 0) continue;
      const int64_t c = idx[i];
      s.best[c] = 0ull;
      s.cnt[c] = 0u;
    }
    return;
  }
  const bool use_win = mae || s.g->tie_flag;  // resolve_kernel ran: win[] holds the first row of each cell maximum
  // block-wide append: one global atomic per 256 rows (a same-address atomic per warp would serialise)
  __shared__ uint32_t s_cnt[8];
  __shared__ uint32_t s_base;
  const int wi = threadIdx.x >> 5;
  for (int64_t b0 = (int64_t)blockIdx.x * 256; b0 < B; b0 += (int64_t)gridDim.x * 256) {
    const int64_t i = b0 + threadIdx.x;
    int32_t c = 0;
    bool iswin = false;
    if (i < B && status[i] != 0) {
      c = idx[i];
      iswin = use_win ? (s.win[c] == (int32_t)i) : (ord_key((double)obj[i]) == s.best[c]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, iswin);
    if (lane == 0) s_cnt[wi] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t n = s_cnt[k];
        s_cnt[k] = tot;
        tot += n;
      }
      s_base = tot ? atomicAdd(&s.g->n_winners, tot) : 0u;
    }
    __syncthreads();
    if (iswin) s.wlist[s_base + s_cnt[wi] + __popc(m & ((1u << lane) - 1u))] = make_int2((int32_t)i, c);
    __syncthreads();
  }
}

// (1 - lr)^k for integral k >= 1 by binary powering (deterministic; k = 1 returns the base exactly)
__device__ __forceinline__ double powi(double b, uint32_t k) {
  double r = 1.0;
  while (true) {
    if (k & 1u) r = __dmul_rn(r, b);
    k >>= 1;
    if (!k) break;
    b = __dmul_rn(b, b);
  }
  return r;
}

// update: one thread per winner -- threshold, objective, occupancy, objective-sum delta, batch best
template <typename T>
__global__ void __launch_bounds__(256) update_kernel(int64_t B, const T* __restrict__ obj,
                                                     uint8_t* __restrict__ occupied, T* __restrict__ store_obj,
                                                     T* __restrict__ store_thr, T tmin, T lr, bool mae, Scratch s) {
  if (s.g->flags) return;
  const int64_t W = s.g->n_winners;
  if (W == 0) return;
  const int lane = threadIdx.x & 31;
  const bool use_win = mae || s.g->tie_flag;
  BinConst bc = make_bins(__longlong_as_double((long long)s.g->absmax_bits) + s.g->store_absmax, B);
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  uint32_t nnew = 0;
  unsigned long long kbest = 0ull;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < W; t += (int64_t)gridDim.x * blockDim.x) {
    const int2 rc = s.wlist[t];
    const int32_t c = rc.y;
    const T o = obj[rc.x];
    const bool occ = occupied[c] != 0;
    const T old_obj = occ ? store_obj[c] : (T)0;
    T newthr;
    if (!mae) {
      newthr = o;  // :663-665
    } else {       // _compute_thresholds, :473-516 (fp64 arithmetic after NumPy promotion)
      const uint32_t kc = s.cnt[c];
      const double k = (double)kc;
      const double* b = &s.bins[3 * (int64_t)c];
      const double S = (double)(T)__dadd_rn(__dadd_rn(b[0], b[1]), b[2]);
      const double tcell = (double)(occ ? store_thr[c] : tmin);
      const double ratio = powi((double)((T)1 - lr), kc);
      newthr = (T)__dadd_rn(__dmul_rn(ratio, tcell), __dmul_rn(__ddiv_rn(S, k), __dsub_rn(1.0, ratio)));
    }
    store_thr[c] = newthr;
    store_obj[c] = o;
    if (!occ) {
      occupied[c] = 1;
      atomicOr(&s.newmask[c >> 5], 1u << (c & 31));
      ++nnew;
    }
    double h0, h1, h2;
    bin_split((double)(T)(o - old_obj), bc, h0, h1, h2);  // :707-710
    d0 += h0;
    d1 += h1;
    d2 += h2;
    const unsigned long long k2 = ord_key((double)o);
    kbest = k2 > kbest ? k2 : kbest;
    // scratch back to its clean state
    s.best[c] = 0ull;
    if (use_win) s.win[c] = INT32_MAX;
    if (mae) {
      s.cnt[c] = 0u;
      s.bins[3 * (int64_t)c] = s.bins[3 * (int64_t)c + 1] = s.bins[3 * (int64_t)c + 2] = 0.0;
    }
  }
  // warp -> block -> global accumulation (bins add exactly, so the grouping does not matter)
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    d0 += __shfl_down_sync(0xffffffffu, d0, o2);
    d1 += __shfl_down_sync(0xffffffffu, d1, o2);
    d2 += __shfl_down_sync(0xffffffffu, d2, o2);
    nnew += __shfl_down_sync(0xffffffffu, nnew, o2);
    unsigned long long ok = __shfl_down_sync(0xffffffffu, kbest, o2);
    kbest = ok > kbest ? ok : kbest;
  }
  __shared__ double sh_d[3][8];
  __shared__ uint32_t sh_n[8];
  __shared__ unsigned long long sh_k[8];
  const int wi = threadIdx.x >> 5;
  if (lane == 0) {
    sh_d[0][wi] = d0;
    sh_d[1][wi] = d1;
    sh_d[2][wi] = d2;
    sh_n[wi] = nnew;
    sh_k[wi] = kbest;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      d0 += sh_d[0][k];
      d1 += sh_d[1][k];
      d2 += sh_d[2][k];
      nnew += sh_n[k];
      kbest = sh_k[k] > kbest ? sh_k[k] : kbest;
    }
    if (d0 != 0.0) atomicAdd(&s.g->dsum[0], d0);
    if (d1 != 0.0) atomicAdd(&s.g->dsum[1], d1);
    if (d2 != 0.0) atomicAdd(&s.g->dsum[2], d2);
    if (nnew) atomicAdd(&s.g->n_new, nnew);
    if (kbest) atomicMax(&s.g->gbest_key, kbest);
  }
}

// Row scatter of the winners (_array_store.py:605-608): every (winner, 16-byte chunk) pair is one
// work item of a flat grid-stride loop, four loads in flight per thread, so the kernel is a plain
// HBM stream: 2 * row_bytes per winner.
__global__ void __launch_bounds__(256) copy_rows_kernel(const __grid_constant__ FieldCopy fc, Scratch s) {
  if (s.g->flags) return;
  const int64_t W = s.g->n_winners;
  if (W == 0) return;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int f = 0; f < fc.n; ++f) {
    const int vec = fc.vec[f];
    const int64_t rb = fc.row_bytes[f];
    const char* __restrict__ src = fc.src[f];
    char* __restrict__ dst = fc.dst[f];
    const int64_t nch = rb / vec;
    const int64_t total = W * nch;
    if (vec == 16) {
      for (int64_t base = tid; base < total; base += 4 * nthreads) {
        uint4 v[4];
        int64_t doff[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t item = base + u * nthreads;
          doff[u] = -1;
          if (item < total) {
            const int64_t wn = item / nch, ch = item - wn * nch;
            const int2 rc = s.wlist[wn];
            v[u] = *reinterpret_cast<const uint4*>(src + (int64_t)rc.x * rb + ch * 16);
            doff[u] = (int64_t)rc.y * rb + ch * 16;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (doff[u] >= 0) *reinterpret_cast<uint4*>(dst + doff[u]) = v[u];
      }
    } else {
      for (int64_t item = tid; item < total; item += nthreads) {
        const int64_t wn = item / nch, ch = item - wn * nch;
        const int2 rc = s.wlist[wn];
        const char* sp = src + (int64_t)rc.x * rb + ch * vec;
        char* dp = dst + (int64_t)rc.y * rb + ch * vec;
        if (vec == 8)
          *reinterpret_cast<uint2*>(dp) = *reinterpret_cast<const uint2*>(sp);
        else if (vec == 4)
          *reinterpret_cast<uint32_t*>(dp) = *reinterpret_cast<const uint32_t*>(sp);
        else
          *dp = *sp;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bestcell_kernel(int64_t N, const uint8_t* __restrict__ occupied,
                                                       const T* __restrict__ store_obj, double prev_obj_max,
                                                       Scratch s) {
  if (s.g->flags || s.g->n_insertable == 0) return;
  const unsigned long long gk = s.g->gbest_key;
  if (gk == 0ull) return;
  const double gbest = ord_unkey(gk);
  if (prev_obj_max == prev_obj_max && !(gbest > prev_obj_max)) return;  // not NaN and no new record
  int32_t m = INT32_MAX;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x)
    if (occupied[c] && (double)store_obj[c] == gbest) m = min(m, (int32_t)c);
  m = __reduce_min_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m != INT32_MAX) atomicMin(&s.g->best_cell, m);
}

// ---- occupied_list append: ascending cell order (_array_store.py:586-600) --------------
__global__ void __launch_bounds__(kAppendThreads) append_count_kernel(int64_t words, Scratch s) {
  const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
  uint32_t n = 0;
#pragma unroll
  for (int k = 0; k < kWordsPerThread; ++k)
    if (w0 + k < words) n += __popc(s.newmask[w0 + k]);
  n = __reduce_add_sync(0xffffffffu, n);
  __shared__ uint32_t sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = sh[threadIdx.x];
    v = __reduce_add_sync(0xffffffffu, v);
    if (threadIdx.x == 0) s.blockcnt[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kAppendThreads) append_emit_kernel(int64_t words, int32_t* __restrict__ occupied_list,
                                                                     int64_t n_before, int64_t B, Scratch s,
                                                                     qd_add_result* result) {
  __shared__ uint32_t sh[32];
  __shared__ uint32_t sh_base;
  // exclusive offset of this block = sum of earlier block counts
  uint32_t part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += kAppendThreads) part += s.blockcnt[b];
  part = __reduce_add_sync(0xffffffffu, part);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = __reduce_add_sync(0xffffffffu, sh[threadIdx.x]);
    if (threadIdx.x == 0) sh_base = v;
  }
  __syncthreads();
  const uint32_t block_base = sh_base;
  __syncthreads();
  // in-block exclusive scan of per-thread popcounts
  const int64_t w0 = (int64_t)blockIdx.x * kWordsPerBlock + (int64_t)threadIdx.x * kWordsPerThread;
  uint32_t wv[kWordsPerThread];
  uint32_t n = 0;
#pragma unroll
  for (int k = 0; k < kWordsPerThread; ++k) {
    wv[k] = (w0 + k < words) ? s.newmask[w0 + k] : 0u;
    n += __popc(wv[k]);
  }
  uint32_t incl = n;
  const int lane = threadIdx.x & 31;
  for (int o2 = 1; o2 < 32; o2 <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o2);
    if (lane >= o2) incl += t;
  }
  if (lane == 31) sh[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v = sh[threadIdx.x];
    uint32_t iv = v;
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, iv, o2);
      if (lane >= o2) iv += t;
    }
    sh[threadIdx.x] = iv - v;  // exclusive warp offsets
  }
  __syncthreads();
  int64_t pos = n_before + block_base + sh[threadIdx.x >> 5] + (incl - n);
#pragma unroll
  for (int k = 0; k < kWordsPerThread; ++k) {
    uint32_t m = wv[k];
    if (m) {
      s.newmask[w0 + k] = 0u;
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        occupied_list[pos++] = (int32_t)((w0 + k) * 32 + bit);
      }
    }
  }
  // block 0 publishes the result and re-arms the globals for the next call
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    InsertGlobals g = *s.g;
    qd_add_result r;
    r.n_insertable = (int32_t)g.n_insertable;
    r.n_new = (int32_t)g.n_new;
    r.best_cell = g.best_cell == INT32_MAX ? -1 : g.best_cell;
    r.flags = g.flags;
    r.best_objective = g.gbest_key ? ord_unkey(g.gbest_key) : 0.0;
    r.delta_sum[0] = g.dsum[0];
    r.delta_sum[1] = g.dsum[1];
    r.delta_sum[2] = g.dsum[2];
    r.batch_absmax = __longlong_as_double((long long)g.absmax_bits);
    *result = r;
    if (!g.flags) g.store_absmax = fmax(g.store_absmax, r.batch_absmax);
    g.gbest_key = 0ull;
    g.absmax_bits = 0ull;
    g.best_cell = INT32_MAX;
    g.n_insertable = g.n_new = g.flags = g.tie_flag = g.n_winners = 0u;
    g.dsum[0] = g.dsum[1] = g.dsum[2] = 0.0;
    *s.g = g;
  }
}

static int pick_vec(const void* a, const void* b, int64_t row_bytes) {
  uintptr_t x = (uintptr_t)a | (uintptr_t)b | (uintptr_t)row_bytes;
  if ((x & 15) == 0) return 16;
  if ((x & 7) == 0) return 8;
  if ((x & 3) == 0) return 4;
  return 1;
}

struct GridSpec {  // non-null measures => fused grid index
  const void* measures = nullptr;
  int meas_dtype = QD_F64;
  GridParams gp;
};

template <typename T, typename MT, int D, bool FUSED>
static void launch_classify(int grid, cudaStream_t stream, int64_t B, int32_t* idx, const GridSpec& gs, const T* obj,
                            const qd_store* st, T tmin, bool cma_me, bool mae, int32_t* status, T* value, Scratch s) {
  classify_kernel<T, MT, D, FUSED><<<grid, 256, 0, stream>>>(B, idx, (const MT*)gs.measures, gs.gp, obj, st->occupied,
                                                             (const T*)st->threshold, tmin, cma_me, mae, status,
                                                             value, s);
}

template <typename T>
int insert_impl(const qd_store* st, int64_t B, int32_t* idx, const GridSpec& gs, const void* objective,
                const void* const* field_src, double lr, double tmin, int64_t n_before, double prev_obj_max,
                int32_t* status, void* value, qd_add_result* result, cudaStream_t stream) {
  const int64_t N = st->capacity;
  Scratch s = carve(st->scratch, N);
  const bool cma_me = (tmin == -INFINITY);
  const bool mae = !cma_me;
  FieldCopy fc;
  fc.n = st->n_fields;
  for (int f = 0; f < fc.n; ++f) {
    fc.src[f] = (const char*)field_src[f];
    fc.dst[f] = (char*)st->field_dst[f];
    fc.row_bytes[f] = st->field_row_bytes[f];
    fc.vec[f] = pick_vec(fc.src[f], fc.dst[f], fc.row_bytes[f]);
  }
  const T* obj = (const T*)objective;
  const int grid = grid_for(B, 256, 8);
  if (gs.measures == nullptr) {
    launch_classify<T, double, 0, false>(grid, stream, B, idx, gs, obj, st, (T)tmin, cma_me, mae, status, (T*)value, s);
  } else if (gs.meas_dtype == QD_F64) {
    if (gs.gp.d == 2)
      launch_classify<T, double, 2, true>(grid, stream, B, idx, gs, obj, st, (T)tmin, cma_me, mae, status, (T*)value, s);
    else
      launch_classify<T, double, 0, true>(grid, stream, B, idx, gs, obj, st, (T)tmin, cma_me, mae, status, (T*)value, s);
  } else {
    if (gs.gp.d == 2)
      launch_classify<T, float, 2, true>(grid, stream, B, idx, gs, obj, st, (T)tmin, cma_me, mae, status, (T*)value, s);
    else
      launch_classify<T, float, 0, true>(grid, stream, B, idx, gs, obj, st, (T)tmin, cma_me, mae, status, (T*)value, s);
  }
  int rc = check_launch("classify_kernel");
  if (rc) return rc;
  resolve_kernel<T><<<grid, 256, 0, stream>>>(B, idx, obj, status, mae, s);
  if ((rc = check_launch("resolve_kernel"))) return rc;
  select_kernel<T><<<grid, 256, 0, stream>>>(B, idx, obj, status, mae, s);
  if ((rc = check_launch("select_kernel"))) return rc;
  update_kernel<T><<<grid_for(B < N ? B : N, 256, 8), 256, 0, stream>>>(B, obj, st->occupied, (T*)st->objective,
                                                                       (T*)st->threshold, (T)tmin, (T)lr, mae, s);
  if ((rc = check_launch("update_kernel"))) return rc;
  if (fc.n > 0) {
    copy_rows_kernel<<<kNumSMs * 8, 256, 0, stream>>>(fc, s);
    if ((rc = check_launch("copy_rows_kernel"))) return rc;
  }
  bestcell_kernel<T><<<grid_for(N, 256, 4), 256, 0, stream>>>(N, st->occupied, (const T*)st->objective, prev_obj_max, s);
  if ((rc = check_launch("bestcell_kernel"))) return rc;
  const int64_t words = (N + 31) / 32;
  const int nblk = (int)((words + kWordsPerBlock - 1) / kWordsPerBlock);
  if (nblk > 1) {
    append_count_kernel<<<nblk, kAppendThreads, 0, stream>>>(words, s);
    if ((rc = check_launch("append_count_kernel"))) return rc;
  }
  append_emit_kernel<<<nblk, kAppendThreads, 0, stream>>>(words, st->occupied_list, n_before, B, s, result);
  return check_launch("append_emit_kernel");
}

static int check_insert_args(const qd_store* store, int64_t B, const void* objective, const int32_t* status_out,
                             const void* value_out, const qd_add_result* result) {
  QD_REQUIRE(store && result, "null store/result");
  QD_REQUIRE(store->capacity >= 1 && store->capacity < ((int64_t)1 << 31), "capacity must fit int32");
  QD_REQUIRE(store->n_fields >= 0 && store->n_fields <= QD_MAX_FIELDS, "too many fields");
  QD_REQUIRE(B >= 0 && B < ((int64_t)1 << 31), "batch size must fit int32");
  QD_REQUIRE(store->obj_dtype == QD_F32 || store->obj_dtype == QD_F64, "objective dtype must be f32/f64");
  QD_REQUIRE(store->scratch && store->occupied && store->occupied_list && store->objective && store->threshold,
             "null store pointer");
  if (B > 0) QD_REQUIRE(objective && status_out && value_out, "null batch pointer");
  return QD_OK;
}

}  // namespace qd

extern "C" size_t qd_insert_scratch_bytes(int64_t capacity) { return qd::scratch_bytes(capacity); }

extern "C" uint32_t* qd_insert_flags_ptr(void* scratch, int64_t capacity) {
  return &qd::carve(scratch, capacity).g->flags;
}

extern "C" int qd_insert_scratch_init(void* scratch, int64_t capacity, void* stream) {
  using namespace qd;
  QD_REQUIRE(scratch && capacity >= 1, "bad scratch arguments");
  Scratch s = carve(scratch, capacity);
  scratch_init_kernel<<<grid_for(capacity, 256), 256, 0, (cudaStream_t)stream>>>(s, capacity);
  return check_launch("scratch_init_kernel");
}

extern "C" int qd_insert(const qd_store* store, int64_t B, const int32_t* idx, const void* objective,
                         const void* const* field_src, double learning_rate, double threshold_min,
                         int64_t n_occupied_before, double prev_obj_max, int32_t* status_out, void* value_out,
                         qd_add_result* result, void* stream) {
  using namespace qd;
  int rc = check_insert_args(store, B, objective, status_out, value_out, result);
  if (rc) return rc;
  if (B > 0) QD_REQUIRE(idx, "null idx");
  GridSpec gs;
  cudaStream_t s = (cudaStream_t)stream;
  if (store->obj_dtype == QD_F64)
    return insert_impl<double>(store, B, const_cast<int32_t*>(idx), gs, objective, field_src, learning_rate,
                               threshold_min, n_occupied_before, prev_obj_max, status_out, value_out, result, s);
  return insert_impl<float>(store, B, const_cast<int32_t*>(idx), gs, objective, field_src, learning_rate,
                            threshold_min, n_occupied_before, prev_obj_max, status_out, value_out, result, s);
}

extern "C" int qd_grid_insert(const qd_store* store, int64_t B, const void* measures, int meas_dtype, int d,
                              const int32_t* dims, const double* lower, const double* interval, double epsilon,
                              const void* objective, const void* const* field_src, double learning_rate,
                              double threshold_min, int64_t n_occupied_before, double prev_obj_max, int32_t* idx_out,
                              int32_t* status_out, void* value_out, qd_add_result* result, void* stream) {
  using namespace qd;
  int rc = check_insert_args(store, B, objective, status_out, value_out, result);
  if (rc) return rc;
  QD_REQUIRE(meas_dtype == QD_F32 || meas_dtype == QD_F64, "measures dtype must be f32/f64");
  if (B > 0) QD_REQUIRE(measures && idx_out, "null measures / idx_out");
  GridSpec gs;
  if ((rc = fill_grid_params(gs.gp, d, dims, lower, interval, epsilon))) return rc;
  gs.measures = measures;
  gs.meas_dtype = meas_dtype;
  cudaStream_t s = (cudaStream_t)stream;
  if (store->obj_dtype == QD_F64)
    return insert_impl<double>(store, B, idx_out, gs, objective, field_src, learning_rate, threshold_min,
                               n_occupied_before, prev_obj_max, status_out, value_out, result, s);
  return insert_impl<float>(store, B, idx_out, gs, objective, field_src, learning_rate, threshold_min,
                            n_occupied_before, prev_obj_max, status_out, value_out, result, s);
}
