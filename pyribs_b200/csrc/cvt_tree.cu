// A3, measure_dim > 3: exact nearest centroid through a bounding-box tree (method 3).
//
// Reference: ribs/archives/_cvt_archive.py:579-609. The reference's default search is
// scipy.spatial.KDTree.query(k=1) over a tree it builds in __init__ (:410-429) -- a spatial
// index that looks at ~10^3 of 10^6 centroids per query, not a B x C contraction. This file is
// its GPU form for measure dimensions where the bucket grid of cvt_bucket.cu (<= 3-D) stops
// working and an exhaustive contraction (cvt_tc.cu) does 10^2..10^3 times the necessary work.
//
// Structure (qd_cvt_tree_prepare, built once where the reference builds its KD-tree): the
// centroids are put in KD order by recursive median splits along the widest axis; every 32
// consecutive points form a LEAF (one point per lane), every 32 consecutive leaves a level-1
// node, every 32 of those a level-2 node, ... (splits fall on multiples of the largest power of
// 32 below the leaf count, so an aligned group of 32^k leaves is a subtree and node j of level l
// is simply leaf >> 5l -- no pointers). A node stores the axis-aligned boxes of its 32 children
// as fp32 rounded outwards, struct-of-arrays, so a warp reads them as coalesced 128-byte rows.
//
// Search: ONE WARP PER QUERY. At a node lane j computes a rigorous fp32 lower bound (directed
// rounding) of the squared distance from the query to child j's box; children are visited best
// first (one redux.min over keys = bound bits | lane) and the walk of a node stops at the first
// child whose bound exceeds the best squared distance found so far. At a leaf lane j evaluates
// the exact fp64 squared distance to point j in NumPy's summation order (cvt_common.cuh); the
// (distance, centroid index) minimum is lexicographic, so the result is bit-identical to the
// fp64 scan and to np.argmin over the unsorted centroid array (_cvt_archive.py:632) -- lowest
// centroid index on exact ties. A subtree is skipped only if its bound is STRICTLY above the best
// distance inflated by 2^-21, far more than the rounding of either side (see prune note below).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "cvt_common.cuh"

namespace qd {

constexpr int kLeaf = 32;         // points per leaf (= lanes)
constexpr int kFan = 32;          // children per node (= lanes)
constexpr int kTreeMaxLev = 7;    // 32^6 leaves x 32 points > 2^31
constexpr size_t kTreeHeaderBytes = 256;
constexpr int kTreeWarps = 8;     // warps (= queries in flight) per CTA

struct TreeHeader {
  int64_t C;
  int32_t d, nleaves, nlev, pad;
  int32_t n[kTreeMaxLev + 1];         // boxes per level; n[0] = leaves, n[nlev-1] = 1
  uint64_t off_box[kTreeMaxLev + 1];  // byte offset of the level-l boxes, grouped by parent (l < nlev-1)
  uint64_t off_pts, off_ids, total;
  uint64_t off_pf;   // the leaf points again, rounded to fp32 (same layout as off_pts): the leaf pre-filter
  uint64_t off_rad;  // d floats: max |p_k - fp32(p_k)| over all points, rounded up
};
static_assert(sizeof(TreeHeader) <= kTreeHeaderBytes, "header too large");

static TreeHeader tree_layout(int64_t C, int d) {
  TreeHeader h;
  memset(&h, 0, sizeof(h));
  h.C = C;
  h.d = d;
  h.nleaves = (int32_t)((C + kLeaf - 1) / kLeaf);
  int l = 0;
  h.n[0] = h.nleaves;
  while (h.n[l] > 1) {
    h.n[l + 1] = (h.n[l] + kFan - 1) / kFan;
    ++l;
  }
  h.nlev = l + 1;
  uint64_t off = kTreeHeaderBytes;
  h.off_pts = off;
  off += (uint64_t)h.nleaves * d * kLeaf * sizeof(double);
  h.off_ids = off;
  off += (uint64_t)h.nleaves * kLeaf * sizeof(int32_t);
  const int dp = (d + 1) / 2;  // fp32 data is stored as PAIRS of coordinates (packed f32x2 arithmetic in the walk)
  for (int k = 0; k + 1 < h.nlev; ++k) {
    h.off_box[k] = off;
    off += (uint64_t)h.n[k + 1] * 2 * dp * kFan * sizeof(float2);
  }
  off = (off + 255) / 256 * 256;
  h.off_pf = off;
  off += (uint64_t)h.nleaves * dp * kLeaf * sizeof(float2);
  off = (off + 255) / 256 * 256;
  h.off_rad = off;
  off += (uint64_t)QD_MAX_MEASURE_DIM * sizeof(float);
  h.total = (off + 255) / 256 * 256;
  return h;
}

// ---- host build ------------------------------------------------------------------------------
static float f32_down(double v) {
  float f = (float)v;
  if ((double)f > v) f = nextafterf(f, -INFINITY);
  return f;
}
static float f32_up(double v) {
  float f = (float)v;
  if ((double)f < v) f = nextafterf(f, INFINITY);
  return f;
}

// KD order of `n` points (ids in idx[0..n)) that will fill `nleaves` leaves: split along the widest axis at
// a multiple of the largest power of kFan below nleaves, recurse.
static void kd_split(const double* pts, int d, int32_t* idx, int64_t n, int64_t nleaves, int depth) {
  if (nleaves <= 1 || n <= 1) return;
  int64_t u = 1;
  while (u * kFan < nleaves) u *= kFan;
  const int64_t nu = (nleaves + u - 1) / u;
  const int64_t left_leaves = (nu / 2) * u;
  const int64_t k = left_leaves * kLeaf;
  if (k <= 0 || k >= n) return;
  // widest axis of the node (large nodes: of an evenly spaced sample, the choice only shapes the boxes)
  const int64_t step = n > 8192 ? n / 4096 : 1;
  int dim = 0;
  double widest = -1.0;
  for (int a = 0; a < d; ++a) {
    double lo = INFINITY, hi = -INFINITY;
    for (int64_t i = 0; i < n; i += step) {
      const double v = pts[(int64_t)idx[i] * d + a];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    if (hi - lo > widest) {
      widest = hi - lo;
      dim = a;
    }
  }
  {  // select on a contiguous (coordinate, id) array: nth_element through the index would miss the cache on every compare
    std::vector<std::pair<double, int32_t>> keyed((size_t)n);
    for (int64_t i = 0; i < n; ++i) keyed[i] = {pts[(int64_t)idx[i] * d + dim], idx[i]};
    std::nth_element(keyed.begin(), keyed.begin() + k, keyed.end());
    for (int64_t i = 0; i < n; ++i) idx[i] = keyed[i].second;
  }
  if (depth < 4 && n > (1 << 16)) {
    std::thread t(kd_split, pts, d, idx, k, left_leaves, depth + 1);
    kd_split(pts, d, idx + k, n - k, nleaves - left_leaves, depth + 1);
    t.join();
  } else {
    kd_split(pts, d, idx, k, left_leaves, depth + 1);
    kd_split(pts, d, idx + k, n - k, nleaves - left_leaves, depth + 1);
  }
}

template <typename T>
static void tree_build_host(const T* cent, int64_t C, int d, unsigned char* blob) {
  const TreeHeader h = tree_layout(C, d);
  memcpy(blob, &h, sizeof(h));
  std::vector<double> pts((size_t)C * d);
  for (int64_t i = 0; i < C * d; ++i) pts[i] = (double)cent[i];
  std::vector<int32_t> order((size_t)C);
  for (int64_t i = 0; i < C; ++i) order[i] = (int32_t)i;
  auto T0 = std::chrono::steady_clock::now();
  kd_split(pts.data(), d, order.data(), C, h.nleaves, 0);
  auto T1 = std::chrono::steady_clock::now();
  if (getenv("QD_TREE_TIMING")) fprintf(stderr, "kd_split %.3f s\n", std::chrono::duration<double>(T1 - T0).count());
  double* lp = reinterpret_cast<double*>(blob + h.off_pts);
  int32_t* li = reinterpret_cast<int32_t*>(blob + h.off_ids);
  const int dp = (d + 1) / 2;
  // fp32 copy of the points: pair j of leaf L and slot s at ((L * dp + j) * kLeaf + s), components (2j, 2j + 1);
  // the missing partner of an odd last coordinate is 0 (the query's interval for it is [0, 0]: no contribution)
  float* lf = reinterpret_cast<float*>(blob + h.off_pf);
  auto pf_at = [&](int64_t leaf, int a, int j) -> float& { return lf[(((leaf * dp + (a >> 1)) * kLeaf + j) << 1) + (a & 1)]; };
  float* rad = reinterpret_cast<float*>(blob + h.off_rad);
  std::vector<double> radd((size_t)d, 0.0);
  // leaves + their boxes (level 0)
  std::vector<float> lo((size_t)h.n[0] * d), hi((size_t)h.n[0] * d);
  for (int64_t leaf = 0; leaf < h.nleaves; ++leaf) {
    for (int a = 0; a < d; ++a) {
      lo[leaf * d + a] = INFINITY;
      hi[leaf * d + a] = -INFINITY;
    }
    for (int j = 0; j < kLeaf; ++j) {
      const int64_t p = leaf * kLeaf + j;
      if (p < C) {
        const int32_t id = order[p];
        li[leaf * kLeaf + j] = id;
        for (int a = 0; a < d; ++a) {
          const double v = pts[(int64_t)id * d + a];
          lp[(leaf * d + a) * kLeaf + j] = v;
          const float vf = (float)v;
          pf_at(leaf, a, j) = vf;
          if (std::isfinite(v) && std::isfinite(vf)) radd[a] = std::max(radd[a], std::fabs(v - (double)vf));
          lo[leaf * d + a] = std::min(lo[leaf * d + a], f32_down(v));
          hi[leaf * d + a] = std::max(hi[leaf * d + a], f32_up(v));
        }
      } else {  // padding of the last leaf: infinitely far away, never the minimum
        li[leaf * kLeaf + j] = INT32_MAX;
        for (int a = 0; a < d; ++a) {
          lp[(leaf * d + a) * kLeaf + j] = INFINITY;
          pf_at(leaf, a, j) = INFINITY;
        }
        if (d & 1) pf_at(leaf, d, j) = INFINITY;
      }
    }
  }
  for (int a = 0; a < d; ++a) rad[a] = f32_up(radd[a]);
  for (int l = 0; l + 1 < h.nlev; ++l) {
    float* bx = reinterpret_cast<float*>(blob + h.off_box[l]);
    const int64_t groups = h.n[l + 1];
    std::vector<float> plo((size_t)groups * d, INFINITY), phi((size_t)groups * d, -INFINITY);
    for (int64_t g = 0; g < groups; ++g) {
      // group g: lo pairs [dp][kFan] then hi pairs [dp][kFan], each entry a float2 of coordinates (2j, 2j + 1)
      float* gb = bx + (size_t)g * 2 * dp * kFan * 2;
      for (int s = 0; s < kFan; ++s) {
        const int64_t c = g * kFan + s;
        const bool real = c < h.n[l];
        for (int a = 0; a < 2 * dp; ++a) {
          float l0 = real ? 0.f : INFINITY, h0 = real ? 0.f : -INFINITY;  // (odd d: the pad coordinate)
          if (a < d) {
            l0 = real ? lo[c * d + a] : INFINITY;
            h0 = real ? hi[c * d + a] : -INFINITY;
            plo[g * d + a] = std::min(plo[g * d + a], l0);
            phi[g * d + a] = std::max(phi[g * d + a], h0);
          }
          gb[(((a >> 1) * kFan + s) << 1) + (a & 1)] = l0;
          gb[((((dp + (a >> 1)) * kFan) + s) << 1) + (a & 1)] = h0;
        }
      }
    }
    lo.swap(plo);
    hi.swap(phi);
  }
}

// ---- search ------------------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FADD2 / FFMA2 with directed rounding): the walk issues on 76 % of its cycles, and the
// box / point bounds are two subtractions and one fused multiply-add per coordinate -- two coordinates per instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2_rd(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2_rd(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// acc += max(a - xhi, xlo - b, 0)^2, both halves, everything rounded down (a = b = the point for a leaf slot,
// a = box lo / b = box hi for a child box)
__device__ __forceinline__ f32x2 gap2_acc(f32x2 a, f32x2 b, f32x2 xlo, f32x2 xhi, f32x2 acc) {
  float t0, t1, u0, u1;
  unpack2(sub2_rd(a, xhi), t0, t1);
  unpack2(sub2_rd(xlo, b), u0, u1);
  const f32x2 g = pack2(fmaxf(fmaxf(t0, u0), 0.f), fmaxf(fmaxf(t1, u1), 0.f));
  return fma2_rd(g, g, acc);
}
__device__ __forceinline__ float sum2_rd(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return __fadd_rd(lo, hi);
}

struct TreeView {
  const double* pts;
  const f32x2* pf;  // [leaf][pair][slot]
  const float* rad;
  const int32_t* ids;
  const f32x2* box[kTreeMaxLev];  // [group][lo | hi][pair][child]
  int32_t nlev, d;
};

// Prune note. lbk <= (fp32, every operation rounded down, query coordinates widened to an enclosing fp32
// interval) <= the true squared distance from the query to anything inside the box. fl(dist^2) of cvt_common.cuh
// is a sum of non-negative terms, each off by at most three roundings: fl(dist^2) >= true * (1 - 2^-48). bestf >=
// best * (1 + 2^-22). Hence lbk > bestf implies fl(dist^2(p)) > best for every p in the box: p is neither the
// minimum nor an exact tie.
//
// HOME = true is the ordering pre-pass: a greedy descent to the leaf with the smallest bound at every level; the
// leaf number is the query's position along the KD order of the centroids. Large batches are counting-sorted by
// it (tree_scan_kernel / tree_scatter_kernel) and every CTA then walks a contiguous run of the sorted queries with
// all of its warps, so that the nodes and leaves the queries of one SM touch overlap and come out of L1 instead of
// L2 (the walk is bound by L2 round trips: ~125 KB per query unsorted).
template <typename T, int D, bool HOME, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 32 ? 1 : ((D >= 1 && D <= 8) ? 4 : 2))
    cvt_tree_kernel(const T* __restrict__ meas, int64_t B, TreeView tv, int32_t base, const int32_t* __restrict__ order,
                    int64_t chunk, int32_t* __restrict__ idx, double* __restrict__ dist_out, uint32_t* flags,
                    uint32_t* __restrict__ home_key, uint32_t* __restrict__ home_count) {
  constexpr int XN = D > 0 ? D : QD_MAX_MEASURE_DIM;
  const int d = D > 0 ? D : tv.d;
  extern __shared__ uint32_t skey_raw[];  // [warps][kTreeMaxLev][32] per-lane slots: child keys of the nodes on the path
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ const f32x2* sbox[kTreeMaxLev];  // (indexing the kernel parameter dynamically would spill it to local memory)
  __shared__ int s_next;  // sorted mode: the CTA's warps draw the queries of its run one by one (their costs differ)
  if (threadIdx.x < kTreeMaxLev) sbox[threadIdx.x] = tv.box[threadIdx.x];
  if (threadIdx.x == 0) s_next = nw;
  __syncthreads();
  uint32_t(*key)[32] = reinterpret_cast<uint32_t(*)[32]>(skey_raw + (size_t)warp * kTreeMaxLev * 32);
  // the query's fp64 coordinates live in shared memory (behind the key slots): only the exact evaluation of a leaf
  // reads them (broadcast), the walk itself runs on the fp32 interval below
  double* sx = reinterpret_cast<double*>(skey_raw + (size_t)nw * kTreeMaxLev * 32) + (size_t)warp * XN;
  bool bad = false;
  const bool dynamic = order != nullptr;  // then gridDim.x * chunk >= B: one run per CTA
  for (int64_t c0 = (int64_t)blockIdx.x * chunk; c0 < B; c0 += (int64_t)gridDim.x * chunk) {
    const int64_t c1 = min(B, c0 + chunk);
    for (int64_t pos = c0 + warp; pos < c1;) {
      const int64_t q = order ? (int64_t)__ldg(order + pos) : pos;
      // next query of this warp: drawn now so that the shared-memory atomic is off the critical path
      int64_t next_pos = pos + nw;
      if (dynamic) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&s_next, 1);
        next_pos = c0 + __shfl_sync(0xffffffffu, t, 0);
      }
      pos = next_pos;
      // fp32 interval around the query, widened by the rounding radius of the fp32 copies of the leaf points: the
      // boxes AND the fp32 points are then compared against an interval that certainly contains (x - p + pf)
      constexpr int XP = (XN + 1) / 2;
      const int dp = (d + 1) >> 1;
      f32x2 xlo[XP], xhi[XP];  // pairs of coordinates; the partner of an odd last one is the interval [0, 0]
      bool qbad = false;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < XP; ++j)
        if (j < dp) {
          float lo2[2] = {0.f, 0.f}, hi2[2] = {0.f, 0.f};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 2 * j + h;
            if (k < d) {
              const double xk = (double)meas[q * d + k];
              if (lane == 0) sx[k] = xk;
              qbad |= !finite_d(xk);
              const float r = __ldg(tv.rad + k);
              lo2[h] = __fsub_rd(__double2float_rd(xk), r);
              hi2[h] = __fadd_ru(__double2float_ru(xk), r);
            }
          }
          xlo[j] = pack2(lo2[0], lo2[1]);
          xhi[j] = pack2(hi2[0], hi2[1]);
        }
      __syncwarp();
      bad |= qbad;
      double best = INFINITY;
      int32_t bi = INT32_MAX;
      float bestf = INFINITY;

      auto eval_leaf = [&](int leaf) {
        const double* p = tv.pts + ((int64_t)leaf * d) * kLeaf + lane;
        double c[XN], x[XN];
#pragma unroll
        for (int k = 0; k < XN; ++k)
          if (k < d) c[k] = __ldg(p + k * kLeaf);
#pragma unroll
        for (int k = 0; k < XN; ++k)
          if (k < d) x[k] = sx[k];
        const double dd = sqdist_np<D>(x, c, d);
        const bool cand = dd <= best;
        if (__any_sync(0xffffffffu, cand)) {
          int32_t id = cand ? __ldg(tv.ids + (int64_t)leaf * kLeaf + lane) : INT32_MAX;
          unsigned long long bits = cand ? (unsigned long long)__double_as_longlong(dd) : ~0ull;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bits, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, id, o);
            if (ob < bits || (ob == bits && oi < id)) {
              bits = ob;
              id = oi;
            }
          }
          const double nd = __longlong_as_double((long long)bits);
          if (nd < best || (nd == best && id < bi)) {
            best = nd;
            bi = id;
            bestf = __fmul_ru(__double2float_ru(best), 1.0000005f);
          }
        }
      };
      // Leaf pre-filter. Lane j holds a rigorous fp32 lower bound of the squared distance to point j (the point's
      // fp32 copy against the widened interval, every operation rounded down -- the argument of the prune note with
      // a degenerate box); the 2 KB of exact fp64 coordinates are read only if some point of the leaf can still be
      // the minimum or tie it. After the first few leaves of a walk that is rare: the walk reads half the bytes,
      // and two leaves per step halve its chain of dependent round trips.
      auto leaf_bound = [&](const f32x2* __restrict__ pp) -> float {
        f32x2 acc = 0ull;
#pragma unroll
        for (int j = 0; j < XP; ++j)
          if (j < dp) {
            const f32x2 c = __ldg(pp + j * kLeaf);
            acc = gap2_acc(c, c, xlo[j], xhi[j], acc);
          }
        return sum2_rd(acc);
      };
      // keys of the 32 children (boxes of level `lvl`, group `parent`): lower-bound bits | lane
      auto node_key = [&](int lvl, int parent) -> uint32_t {
        const f32x2* p = sbox[lvl] + (int64_t)parent * (2 * dp * kFan) + lane;
        f32x2 acc = 0ull;
#pragma unroll
        for (int j = 0; j < XP; ++j)
          if (j < dp) acc = gap2_acc(__ldg(p + j * kFan), __ldg(p + (dp + j) * kFan), xlo[j], xhi[j], acc);
        return (__float_as_uint(sum2_rd(acc)) & 0x7fffffe0u) | (uint32_t)lane;
      };

      if constexpr (HOME) {
        int cur = 0;
        if (!qbad)
          for (int level = tv.nlev - 1; level >= 1; --level)
            cur = cur * kFan + (int)(__reduce_min_sync(0xffffffffu, node_key(level - 1, cur)) & 31u);
        if (lane == 0) {
          home_key[q] = (uint32_t)cur;
          atomicAdd(&home_count[cur + 1], 1u);
        }
        continue;
      }
      if (!qbad) {
        if (tv.nlev == 1) {
          eval_leaf(0);
        } else {
          int level = tv.nlev - 1;  // level of the current node; its children are the boxes of level - 1
          int cur = 0;              // index of the current node within its level
          key[level][lane] = node_key(level - 1, 0);
          uint32_t k = __reduce_min_sync(0xffffffffu, key[level][lane]);
          while (true) {
            if (k >= 0x7f800000u || __uint_as_float(k & 0xffffffe0u) > bestf) {  // nothing left below this node
              if (++level >= tv.nlev) break;
              cur >>= 5;
              k = __reduce_min_sync(0xffffffffu, key[level][lane]);
              continue;
            }
            const int child = (int)(k & 31u);
            if (lane == child) key[level][lane] = 0xffffffffu;  // visited
            const int cnode = cur * kFan + child;
            if (level == 1) {
              // second leaf of this step: the next child in bound order, if it passes now
              const uint32_t k2 = __reduce_min_sync(0xffffffffu, key[1][lane]);
              const bool two = k2 < 0x7f800000u && !(__uint_as_float(k2 & 0xffffffe0u) > bestf);
              const int cnode2 = cur * kFan + (int)(k2 & 31u);
              if (two && lane == (int)(k2 & 31u)) key[1][lane] = 0xffffffffu;
              const float b1 = leaf_bound(tv.pf + ((int64_t)cnode * dp) * kLeaf + lane);
              float b2 = INFINITY;
              if (two) b2 = leaf_bound(tv.pf + ((int64_t)cnode2 * dp) * kLeaf + lane);
              if (__any_sync(0xffffffffu, !(b1 > bestf))) eval_leaf(cnode);
              if (two && __any_sync(0xffffffffu, !(b2 > bestf))) eval_leaf(cnode2);
              k = __reduce_min_sync(0xffffffffu, key[1][lane]);
            } else {
              --level;
              cur = cnode;
              key[level][lane] = node_key(level - 1, cur);
              k = __reduce_min_sync(0xffffffffu, key[level][lane]);
            }
          }
        }
      }
      if (lane == 0) {
        idx[q] = qbad ? base : ((bi == INT32_MAX ? 0 : bi) + base);
        if (dist_out) dist_out[q] = qbad ? INFINITY : best;
      }
    }
  }
  if (!HOME && __syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, QD_FLAG_NONFINITE_MEASURES);
}

// inclusive scan of the per-leaf query counts in place (count[0] = 0, count[l + 1] = queries of leaf l): one block,
// every thread owns a contiguous segment, one block-wide scan of the 1024 segment sums
__global__ void __launch_bounds__(1024) tree_scan_kernel(uint32_t* __restrict__ count, int64_t n) {
  __shared__ uint32_t warp_sum[32];
  const int64_t per = (n + 1023) / 1024;
  const int64_t i0 = 1 + (int64_t)threadIdx.x * per, i1 = min(n + 1, i0 + per);
  uint32_t tot = 0;
  for (int64_t i = i0; i < i1; ++i) tot += count[i];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sum[wi] = incl;
  __syncthreads();
  if (wi == 0) {
    const uint32_t v = warp_sum[lane];
    uint32_t iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, iv, o);
      if (lane >= o) iv += t;
    }
    warp_sum[lane] = iv - v;
  }
  __syncthreads();
  uint32_t run = warp_sum[wi] + incl - tot;
  for (int64_t i = i0; i < i1; ++i) {
    run += count[i];
    count[i] = run;
  }
}

__global__ void tree_scatter_kernel(const uint32_t* __restrict__ home_key, uint32_t* __restrict__ cursor, int64_t B,
                                    int32_t* __restrict__ order) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < B; q += (int64_t)gridDim.x * blockDim.x)
    order[atomicAdd(&cursor[home_key[q]], 1u)] = (int32_t)q;  // cursor[l] starts at the scanned offset of leaf l
}

constexpr int64_t kTreeSortMin = 16384;  // smaller batches are walked in input order

static size_t a256(size_t x) { return (x + 255) / 256 * 256; }
static size_t tree_scratch_bytes(int64_t B, int64_t C) {
  if (B < kTreeSortMin) return 0;
  const int64_t nleaves = (C + kLeaf - 1) / kLeaf;
  return 2 * a256((size_t)B * 4) + a256((size_t)(nleaves + 2) * 4);
}

static size_t tree_smem(int warps, int D) {  // per warp: key slots of the path + the query's fp64 coordinates
  return (size_t)warps * kTreeMaxLev * 32 * 4 + (size_t)warps * (D > 0 ? D : QD_MAX_MEASURE_DIM) * 8;
}

template <typename T, int D>
static int launch_tree_d(const T* meas, int64_t B, const TreeView& v, const TreeHeader& h, int32_t base, int32_t* idx,
                         double* dist, uint32_t* flags, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  const bool sorted = B >= kTreeSortMin && h.nlev > 1 && scratch && scratch_bytes >= tree_scratch_bytes(B, h.C);
  int rc;
  if (!sorted) {
    // input order: 8 warps per CTA, one query per warp and pass
    const int64_t blocks = (B + kTreeWarps - 1) / kTreeWarps;
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)kNumSMs * 8);
    cvt_tree_kernel<T, D, false, kTreeWarps><<<grid, kTreeWarps * 32, tree_smem(kTreeWarps, D), s>>>(
        meas, B, v, base, nullptr, kTreeWarps, idx, dist, flags, nullptr, nullptr);
    return check_launch("cvt_tree_kernel");
  }
  char* p = (char*)scratch;
  uint32_t* home_key = (uint32_t*)p;
  p += a256((size_t)B * 4);
  int32_t* order = (int32_t*)p;
  p += a256((size_t)B * 4);
  uint32_t* count = (uint32_t*)p;
  QD_CHECK_CUDA(cudaMemsetAsync(count, 0, (size_t)(h.nleaves + 2) * 4, s));
  {
    const int64_t blocks = (B + kTreeWarps - 1) / kTreeWarps;
    const int grid = (int)std::min<int64_t>(blocks, (int64_t)kNumSMs * 8);
    cvt_tree_kernel<T, D, true, kTreeWarps><<<grid, kTreeWarps * 32, tree_smem(kTreeWarps, D), s>>>(
        meas, B, v, base, nullptr, kTreeWarps, idx, dist, flags, home_key, count);
    if ((rc = check_launch("cvt_tree_kernel(home)"))) return rc;
  }
  // (a variant with the counts staged in 200 KB of shared memory measured slower: 45 vs 26 us under ncu -- the
  // carve-out switch between its neighbours costs more than the strided walks it saves)
  tree_scan_kernel<<<1, 1024, 0, s>>>(count, h.nleaves);
  if ((rc = check_launch("tree_scan_kernel"))) return rc;
  tree_scatter_kernel<<<grid_for(B, 256), 256, 0, s>>>(home_key, count, B, order);
  if ((rc = check_launch("tree_scatter_kernel"))) return rc;
  // every CTA walks a contiguous run of the sorted queries with all of its warps: one CTA of 32 warps per SM
  // where the per-thread state fits 64 registers (measure_dim <= 8), two of 8 warps otherwise
  constexpr int kBig = (D >= 1 && D <= 8) ? 32 : kTreeWarps;
  // Runs of >= ~16 queries per warp (a CTA ends when its slowest warp does: short runs are mostly tail -- 125k
  // queries in runs of 64 ran at half the per-query rate of 1M), at most 16 runs per SM.
  int64_t runs_per_sm = B / ((int64_t)kNumSMs * kBig * 16);
  runs_per_sm = runs_per_sm < 1 ? 1 : (runs_per_sm > 16 ? 16 : runs_per_sm);
  int64_t chunk = (B + (int64_t)kNumSMs * runs_per_sm - 1) / ((int64_t)kNumSMs * runs_per_sm);
  chunk = (chunk + kBig - 1) / kBig * kBig;
  const int grid = (int)((B + chunk - 1) / chunk);
  cvt_tree_kernel<T, D, false, kBig><<<grid, kBig * 32, tree_smem(kBig, D), s>>>(
      meas, B, v, base, order, chunk, idx, dist, flags, nullptr, nullptr);
  return check_launch("cvt_tree_kernel(sorted)");
}

template <typename T>
static int launch_tree(const T* meas, int64_t B, int d, const unsigned char* tree, const TreeHeader& h, int32_t base,
                       int32_t* idx, double* dist, uint32_t* flags, void* scratch, size_t scratch_bytes,
                       cudaStream_t s) {
  TreeView v;
  v.pts = reinterpret_cast<const double*>(tree + h.off_pts);
  v.pf = reinterpret_cast<const f32x2*>(tree + h.off_pf);
  v.rad = reinterpret_cast<const float*>(tree + h.off_rad);
  v.ids = reinterpret_cast<const int32_t*>(tree + h.off_ids);
  for (int l = 0; l < kTreeMaxLev; ++l)
    v.box[l] = l + 1 < h.nlev ? reinterpret_cast<const f32x2*>(tree + h.off_box[l]) : nullptr;
  v.nlev = h.nlev;
  v.d = d;
#define QD_TREE(DD) return launch_tree_d<T, DD>(meas, B, v, h, base, idx, dist, flags, scratch, scratch_bytes, s)
  switch (d) {
    case 2: QD_TREE(2);
    case 3: QD_TREE(3);
    case 4: QD_TREE(4);
    case 5: QD_TREE(5);
    case 6: QD_TREE(6);
    case 7: QD_TREE(7);
    case 8: QD_TREE(8);
    case 10: QD_TREE(10);
    case 12: QD_TREE(12);
    case 16: QD_TREE(16);
    default: QD_TREE(0);
  }
#undef QD_TREE
}

int cvt_tree_index_of(const void* measures, int dtype, int64_t B, int d, int64_t C, const void* tree,
                      int32_t index_base, int32_t* idx_out, double* dist_out, uint32_t* flags, void* scratch,
                      size_t scratch_bytes, cudaStream_t s) {
  const TreeHeader h = tree_layout(C, d);  // the layout is a pure function of (C, d): no device read needed
  if (dtype == QD_F64)
    return launch_tree<double>((const double*)measures, B, d, (const unsigned char*)tree, h, index_base, idx_out,
                               dist_out, flags, scratch, scratch_bytes, s);
  return launch_tree<float>((const float*)measures, B, d, (const unsigned char*)tree, h, index_base, idx_out, dist_out,
                            flags, scratch, scratch_bytes, s);
}

}  // namespace qd

extern "C" size_t qd_cvt_tree_scratch_bytes(int64_t B, int64_t C) { return qd::tree_scratch_bytes(B, C); }

extern "C" size_t qd_cvt_tree_bytes(int64_t C, int d) {
  if (C < 1 || d < 1 || d > QD_MAX_MEASURE_DIM) return 0;
  return (size_t)qd::tree_layout(C, d).total;
}

extern "C" int qd_cvt_tree_build_host(const void* centroids, int dtype, int64_t C, int d, void* tree) {
  using namespace qd;
  QD_REQUIRE(d >= 1 && d <= QD_MAX_MEASURE_DIM, "measure_dim out of range [1, 32]");
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  QD_REQUIRE(C >= 1 && C < ((int64_t)1 << 31) - kLeaf, "centroid count out of range");
  QD_REQUIRE(centroids && tree, "null pointer");
  memset(tree, 0, (size_t)tree_layout(C, d).total);
  if (dtype == QD_F64)
    tree_build_host<double>((const double*)centroids, C, d, (unsigned char*)tree);
  else
    tree_build_host<float>((const float*)centroids, C, d, (unsigned char*)tree);
  return QD_OK;
}

extern "C" int qd_cvt_tree_prepare(const void* centroids, int dtype, int64_t C, int d, void* tree, void* stream) {
  using namespace qd;
  QD_REQUIRE(d >= 1 && d <= QD_MAX_MEASURE_DIM, "measure_dim out of range [1, 32]");
  QD_REQUIRE(dtype == QD_F32 || dtype == QD_F64, "dtype must be QD_F32 or QD_F64");
  QD_REQUIRE(C >= 1 && C < ((int64_t)1 << 31) - kLeaf, "centroid count out of range");
  QD_REQUIRE(centroids && tree, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t esz = dtype == QD_F64 ? 8 : 4;
  std::vector<unsigned char> host_c((size_t)C * d * esz);
  QD_CHECK_CUDA(cudaMemcpyAsync(host_c.data(), centroids, host_c.size(), cudaMemcpyDeviceToHost, s));
  QD_CHECK_CUDA(cudaStreamSynchronize(s));
  const size_t total = (size_t)tree_layout(C, d).total;
  std::vector<unsigned char> blob(total);
  int rc = qd_cvt_tree_build_host(host_c.data(), dtype, C, d, blob.data());
  if (rc) return rc;
  QD_CHECK_CUDA(cudaMemcpyAsync(tree, blob.data(), total, cudaMemcpyHostToDevice, s));
  QD_CHECK_CUDA(cudaStreamSynchronize(s));  // `blob` is pageable and dies with this frame
  return QD_OK;
}
