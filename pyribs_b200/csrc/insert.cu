// A4-A8: the body of GridArchive.add / CVTArchive.add after index_of
// (reference ribs/archives/_grid_archive.py:623-712 == _cvt_archive.py:818-907, formerly
// batch_entries_with_threshold; store side ribs/archives/_array_store.py:535-608; stats
// _grid_archive.py:325-360).
//
// ONE persistent cooperative kernel per add(): every CTA is resident (grid <= SMs x occupancy),
// the phases below are separated by grid-wide barriers instead of kernel boundaries, and all
// cross-call state (element count, objective sum, best objective, best-elite snapshot) lives in
// device memory, so add() needs no host round trip between calls.
//
// Sort-free and deterministic. Per-cell scratch: one 32-byte record (one L2 sector), clean
// between calls:
//   best  u64  order key of the highest insertable objective            (atomicMax)
//   pick  u64  (key with its low `rb` bits replaced by the inverted batch row) (atomicMax): the first
//              row among those with the highest TRUNCATED key. It is the cell's winner -- first in batch
//              among the rows holding the maximum, _grid_archive.py:693-696 -- whenever its own key equals
//              `best`; otherwise two rows of the cell differ only below bit rb (relative 2^-(52-rb)) and the
//              cell is settled by the resolve pass (mark) below
//   w0,w1 i64  two-level fixed-point sum of the insertable objectives   (atomicAdd, CMA-MAE);
//              the low bits of w1 carry the row count
// The threshold array doubles as the occupancy test of the row passes: unoccupied cells hold NaN
// (the reference never reads the threshold of an unoccupied cell, _grid_archive.py:628-630), so
// a row costs ONE scattered load + its atomics. Scattered accesses (~1.3-2 clk/lane/SM), not HBM
// bytes, bound the row pass; everything else streams. Round 1 found the winner with a second row
// pass (gather `best`, atomicMin(row): 20 us of 171 at 2M rows); `pick` costs one more RED in the
// first pass and makes that pass an exception path.
//
// Phases (B rows, W winners, N cells):
//   0 absmax    CMA-MAE only: max |objective| fixes the fixed-point scale of the sums
//   1 classify  (grid: fused with index_of) status / value vs the PRE-batch store; best; pick; sums
//   2 mark      list mode, and cell mode only for cells update found ambiguous: rows holding their
//               cell's maximum compete with their row number, lowest wins
//   3 select    list mode only (cells >> rows): winners -> compact (row, cell) list
//   4 update    cell mode: one thread per cell, coalesced, emits the winner list in runs sorted by cell;
//               list mode: one thread per winner. Threshold, objective, occupancy, sum delta
//   6 best/cnt  lowest winning cell holding the new best objective; new-cell bitmap counts
//   5 copy      TMA bulk copies of the winners' row fields into the store (HBM stream)
//   7 publish   occupied_list append in ascending cell order (_array_store.py:586-600), stats,
//               best-elite snapshot, qd_add_result; re-arms the per-call globals
// A non-finite objective or measure turns phases 3.. into a scratch clean-up, so the store is
// untouched when the host raises ValueError (ribs/_utils.py:18-31).
#include <cooperative_groups.h>

#include "grid_common.cuh"

namespace cg = cooperative_groups;

namespace qd {

constexpr int kThreads = 512;
constexpr int kMaxCtas = kNumSMs * 4;  // upper bound on the persistent grid

struct __align__(32) Cell {
  unsigned long long best;
  unsigned long long pick;
  long long w0;
  long long w1;
};

// `pick` arithmetic. rb = bits of a batch row number; keys of finite objectives never reach the all-ones
// truncation (that is a NaN pattern), which therefore marks a cell whose winner the mark pass decides.
struct PickConst {
  unsigned long long rowmask;  // 2^rb - 1
  unsigned long long amb;      // ~rowmask: base of the mark-pass entries
};
__host__ __device__ inline PickConst make_pick(int64_t B) {
  int rb = 1;
  while (((int64_t)1 << rb) < B) ++rb;
  PickConst p;
  p.rowmask = (1ull << rb) - 1ull;
  p.amb = ~p.rowmask;
  return p;
}

struct InsertState {
  // ---- per call (re-armed by the publish phase) ----
  unsigned long long gbest_key;
  unsigned long long absmax_bits;
  int32_t best_cell;
  uint32_t n_insertable;
  uint32_t n_new;
  uint32_t flags;
  uint32_t pad0;
  uint32_t n_winners;
  uint32_t n_ambiguous;  // cells whose pick was not their maximum (per call, re-armed by publish)
  uint32_t pad1;
  double dsum[3];
  // ---- persistent ----
  double store_absmax;            // max |objective| ever offered to the store
  long long n_occupied;           // len(store)
  unsigned long long obj_max_key; // order key of stats.obj_max, 0 = no elite yet
  double objective_sum;           // running sum of elite objectives, rounded as the objective dtype
  unsigned long long n_commits;   // calls that inserted something (ArrayStore.updates[ADD])
  unsigned long long best_serial; // bumps whenever the best-elite snapshot is replaced
  unsigned long long n_failed;    // calls dropped by validation
  uint32_t last_failed_flags;     // flags of the most recent such call
  uint32_t pad;
  unsigned long long t_phase[10]; // %globaltimer at the phase boundaries of the last call (diagnostics)
  qd_add_result* result_mirror;   // page-locked host copy of every call's qd_add_result (qd_insert_set_result_mirror) or NULL
  // ---- per call, re-armed by publish ----
  unsigned long long best_pack;   // (cell << 32 | batch row) of the lowest winning cell holding the new maximum
  // ---- pipelined commit (QD_INSERT_ROWS / QD_INSERT_COPY): winners of the row phases, one count per list buffer
  uint32_t copy_w[2];

};

struct Scratch {
  Cell* cell;
  uint32_t* cnt;  // per-cell row count, only when it does not fit the low bits of w1 (B >= 2^24)
  uint32_t* newmask;
  uint32_t* ctacnt;  // kMaxCtas
  int2* wlist;       // (batch row, cell) of this call's winners, at most one per cell
  // row copy: next undistributed piece of every bulk field, per list buffer [2][QD_MAX_FIELDS] (re-armed by the
  // publish phase of the launch that fills that buffer)
  unsigned long long* copy_next;
  InsertState* g;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Scratch carve(void* base, int64_t N) {
  Scratch s;
  char* p = (char*)base;
  s.cell = (Cell*)p;
  p += align_up((size_t)N * sizeof(Cell), 256);
  s.cnt = (uint32_t*)p;
  p += align_up((size_t)N * 4, 256);
  const int64_t words = (N + 31) / 32;
  s.newmask = (uint32_t*)p;
  p += align_up((size_t)words * 4, 256);
  s.ctacnt = (uint32_t*)p;
  p += align_up((size_t)kMaxCtas * 4, 256);
  s.wlist = (int2*)p;  // two buffers of N entries (a pipelined row copy reads one while the next call fills the other)
  p += 2 * align_up((size_t)N * 8, 256);
  s.copy_next = (unsigned long long*)p;
  p += align_up((size_t)2 * QD_MAX_FIELDS * 8, 256);
  s.g = (InsertState*)p;
  return s;
}

static size_t scratch_bytes(int64_t N) {
  const int64_t words = (N + 31) / 32;
  return align_up((size_t)N * sizeof(Cell), 256) + align_up((size_t)N * 4, 256) + align_up((size_t)words * 4, 256) +
         align_up((size_t)kMaxCtas * 4, 256) + 2 * align_up((size_t)N * 8, 256) +
         align_up((size_t)2 * QD_MAX_FIELDS * 8, 256) +
         align_up(sizeof(InsertState), 256);
}

__global__ void scratch_init_kernel(Scratch s, int64_t N) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = i; c < N; c += stride) {
    Cell z;
    z.best = z.pick = 0ull;
    z.w0 = z.w1 = 0ll;
    s.cell[c] = z;
    s.cnt[c] = 0u;
  }
  for (int64_t w = i; w < (N + 31) / 32; w += stride) s.newmask[w] = 0u;
  if (i == 0) {
    InsertState g;
    memset(&g, 0, sizeof(g));
    g.best_cell = INT32_MAX;
    g.best_pack = ~0ull;
    *s.g = g;
    for (int k = 0; k < 2 * QD_MAX_FIELDS; ++k) s.copy_next[k] = 0ull;
  }
}

// ---- fixed-point accumulation of the insertable objectives of a cell (CMA-MAE) -----------
// x * 2^s0 is rounded to an integer q0 (|q0| <= 2^(62-b), 2^b > B rows: the int64 sum cannot
// overflow); the exact remainder, scaled again, gives q1. Integer adds are associative, so the
// cell sums do not depend on the order the atomics land in. The mean of a cell is exact to
// 2^-(lvl0+lvl1+1) of the batch maximum (2^-54 at 16M rows, better below) -- tighter than the
// reference's sequential fp64 sum.
struct FxConst {
  double scale0, scale1;  // x -> level-0 integer; level-0 remainder -> level-1 integer
  double inv0, inv1;      // back to doubles
  int b;                  // count bits at the bottom of w1 (0: separate counter)
};

__host__ __device__ inline FxConst make_fx(double absmax, int64_t B) {
  int e;
  if (!(absmax > 0.0)) absmax = 1e-280;
  frexp(absmax, &e);  // absmax < 2^e
  if (e < -900) e = -900;
  int b = 1;
  while (((int64_t)1 << b) <= B) ++b;
  FxConst k;
  const int lvl0 = 62 - b;
  const bool packed = b <= 24;
  const int lvl1 = packed ? 63 - 2 * b : 62 - b;
  k.b = packed ? b : 0;
  k.scale0 = ldexp(1.0, lvl0 - e);
  k.scale1 = ldexp(1.0, lvl1);
  k.inv0 = ldexp(1.0, e - lvl0);
  k.inv1 = ldexp(1.0, e - lvl0 - lvl1);
  return k;
}

__device__ __forceinline__ void fx_add(Cell* c, uint32_t* cnt, double x, const FxConst& k) {
  const double xs = __dmul_rn(x, k.scale0);
  const long long q0 = __double2ll_rn(xs);
  const double r = __dsub_rn(xs, (double)q0);  // exact
  const long long q1 = __double2ll_rn(__dmul_rn(r, k.scale1));
  if (q0) atomicAdd((unsigned long long*)&c->w0, (unsigned long long)q0);
  if (k.b) {
    atomicAdd((unsigned long long*)&c->w1, ((unsigned long long)q1 << k.b) + 1ull);
  } else {
    if (q1) atomicAdd((unsigned long long*)&c->w1, (unsigned long long)q1);
    atomicAdd(cnt, 1u);
  }
}

__device__ __forceinline__ void fx_read(const Cell& c, const uint32_t* cnt, const FxConst& k, double& sum,
                                        uint32_t& count) {
  long long s1 = c.w1;
  if (k.b) {
    count = (uint32_t)((unsigned long long)s1 & ((1ull << k.b) - 1ull));
    s1 = (s1 - (long long)count) >> k.b;
  } else {
    count = *cnt;
  }
  sum = __dadd_rn(__dmul_rn((double)c.w0, k.inv0), __dmul_rn((double)s1, k.inv1));
}

struct FieldCopy {
  int n;
  const char* src[QD_MAX_FIELDS];
  char* dst[QD_MAX_FIELDS];
  int64_t row_bytes[QD_MAX_FIELDS];
  int64_t snap_off[QD_MAX_FIELDS];  // offset of the field inside the best-elite snapshot
  int vec[QD_MAX_FIELDS];           // 16, 8, 4 or 1
  int host_src[QD_MAX_FIELDS];      // source rows live in page-locked HOST memory (read over PCIe)
  // data-parallel add(): batch row g lives on rank g / rows_per_rank; peer_tab[rank * n + f] is that rank's
  // slice of field f, mapped into this GPU's address space (NVLink peer memory). NULL: src[] holds all rows.
  const char* const* peer_tab;
  int64_t rows_per_rank;
};

__device__ __forceinline__ const char* src_row(const FieldCopy& fc, int f, int row, int64_t rb) {
  if (fc.peer_tab == nullptr) return fc.src[f] + (int64_t)row * rb;
  const int r = (int)(row / fc.rows_per_rank);
  const int64_t l = row - (int64_t)r * fc.rows_per_rank;
  return fc.peer_tab[r * fc.n + f] + l * rb;
}

template <typename T>
struct InsertArgs {
  int64_t B, N;
  int32_t* idx;
  const void* measures;  // fused grid index only
  const T* obj;
  uint8_t* occupied;
  int32_t* occupied_list;
  T* store_obj;
  T* store_thr;
  T tmin, lr;
  int mae;
  int32_t* status;
  T* value;
  char* snapshot;  // best-elite snapshot buffer or NULL
  qd_add_result* result;
  Scratch s;
  // owner-computes insertion (qd_insert_owned_cells): this store holds the cells [cell_lo, cell_hi) of a sharded
  // archive. Rows of other cells are skipped, cell numbers are local from then on, and status / value of a row go to
  // the rank its batch slice came from (out_tab[2 * r] = status, [2 * r + 1] = value of rank r, peer memory).
  int32_t cell_lo, cell_hi;  // cell_hi == 0: the whole archive (idx is local already)
  void* const* out_tab;      // [dev] or NULL: status / value stay in the arrays above
  int64_t out_rows_per_rank;
};

template <typename T>
__device__ __forceinline__ void store_out(const InsertArgs<T>& a, int64_t i, int32_t st, T v) {
  if (a.out_tab == nullptr) {
    a.status[i] = st;
    a.value[i] = v;
  } else {
    const uint32_t r = (uint32_t)i / (uint32_t)a.out_rows_per_rank;
    const uint32_t l = (uint32_t)i - r * (uint32_t)a.out_rows_per_rank;
    reinterpret_cast<int32_t*>(a.out_tab[2 * r])[l] = st;
    reinterpret_cast<T*>(a.out_tab[2 * r + 1])[l] = v;
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

template <typename T>
__device__ __forceinline__ bool is_finite_t(T o) {
  return sizeof(T) == 8 ? finite_d((double)o) : finite_f((float)o);
}

// Row passes keep kRows independent rows in flight per thread (rows tid, tid+S, tid+2S, ... of one
// sweep, S = threads in the grid): with 1024 threads per SM a single outstanding load per thread
// cannot cover the HBM / L2 latency.
constexpr int kRows = 4;

// ---- phase 0 ----------------------------------------------------------------------------------
// block-wide max of non-negative doubles through `sh` (kThreads / 32 entries); result valid in thread 0
__device__ __forceinline__ double block_max(double v, double* sh) {
  for (int o2 = 16; o2 > 0; o2 >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o2));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.0;
    for (int o2 = 16; o2 > 0; o2 >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o2));
  }
  return v;
}

template <typename T>
__device__ void phase_absmax(const InsertArgs<T>& a, double* sh_d) {
  double amax = 0.0;
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < a.B; base += 8 * S) {
    T o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t i = base + u * S;
      o[u] = i < a.B ? a.obj[i] : (T)0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (is_finite_t(o[u])) amax = fmax(amax, fabs((double)o[u]));
  }
  amax = block_max(amax, sh_d);
  if (threadIdx.x == 0 && amax > 0.0)
    atomicMax(&a.s.g->absmax_bits, (unsigned long long)__double_as_longlong(amax));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- phase 1 ----------------------------------------------------------------------------------
template <typename T, typename MT, int D, bool FUSED>
__device__ void phase_classify(const InsertArgs<T>& a, const GridParams& gp, bool cell_mode, double* sh_d,
                               uint32_t* sh_u) {
  const bool mae = a.mae;
  const PickConst pc = make_pick(a.B);
  const bool cma_me = !mae;
  FxConst fx;
  if (mae) fx = make_fx(__longlong_as_double((long long)a.s.g->absmax_bits), a.B);
  uint32_t ncan = 0;
  double amax = 0.0;
  bool bad = false, bad_meas = false;
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < a.B; base += kRows * S) {
    int32_t c[kRows];
    T o[kRows], tc[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const int64_t i = base + u * S;
      c[u] = -1;
      if (i < a.B) {
        if constexpr (FUSED)
          c[u] = grid_cell_of<MT, D>((const MT*)a.measures, i, gp, bad_meas);  // A2 fused into the first pass
        else
          c[u] = a.idx[i];
        o[u] = a.obj[i];
        if (a.cell_hi > 0) {
          // every rank looks at every objective (one verdict about the batch on all ranks), but only at the rows of
          // its own cells from here on; the later passes see the local cell number or -1
          bad |= !is_finite_t(o[u]);
          c[u] = (c[u] >= a.cell_lo && c[u] < a.cell_hi) ? c[u] - a.cell_lo : -1;
          a.idx[i] = c[u];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u)
      if (c[u] >= 0) tc[u] = a.store_thr[c[u]];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      if (c[u] < 0) continue;
      const int64_t i = base + u * S;
      if constexpr (FUSED) a.idx[i] = c[u];
      const bool fin = is_finite_t(o[u]);
      bad |= !fin;
      const bool occ = (tc[u] == tc[u]);   // unoccupied cells hold NaN
      const T t = occ ? tc[u] : a.tmin;    // _grid_archive.py:628-630
      const bool can = fin && (o[u] > t);  // :636
      const T tval = (can && !occ) ? (cma_me ? (T)0 : a.tmin) : t;  // :646-648
      store_out<T>(a, i, can ? (occ ? 1 : 2) : 0, o[u] - tval);     // :649
      if (can) {
        Cell* cell = &a.s.cell[c[u]];
        const unsigned long long key = ord_key((double)o[u]);
        atomicMax(&cell->best, key);  // results unused: fire-and-forget REDs
        if (cell_mode) atomicMax(&cell->pick, (key & pc.amb) | (pc.rowmask - (unsigned long long)i));
        if (mae) fx_add(cell, &a.s.cnt[c[u]], (double)o[u], fx);
        ++ncan;
      }
      if (!mae && fin) amax = fmax(amax, fabs((double)o[u]));
    }
  }
  // one set of global atomics per CTA (per-warp same-address atomics serialise in L2)
  const int any_bad_meas = __syncthreads_or(bad_meas), any_bad = __syncthreads_or(bad);
  amax = block_max(amax, sh_d);
  ncan = __reduce_add_sync(0xffffffffu, ncan);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh_u[threadIdx.x >> 5] = ncan;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t n = 0;
    for (int k = 0; k < kThreads / 32; ++k) n += sh_u[k];
    if (n) atomicAdd(&a.s.g->n_insertable, n);
    if (amax > 0.0) atomicMax(&a.s.g->absmax_bits, (unsigned long long)__double_as_longlong(amax));
    if (any_bad_meas) atomicOr(&a.s.g->flags, QD_FLAG_NONFINITE_MEASURES);
    if (any_bad) atomicOr(&a.s.g->flags, QD_FLAG_NONFINITE_OBJECTIVE);
  }
  __syncthreads();
}

// ---- phase 2 ----------------------------------------------------------------------------------
// Every row whose objective equals its cell's maximum competes for the cell with its row number: the lowest
// row wins (first in batch wins ties, _grid_archive.py:693-696). A row holding the maximum of its cell is
// insertable by construction (insertability depends on (objective, cell) only), and cells without
// insertable rows keep best == 0, which is no finite objective's key -- so `status` is not read.
// List mode runs it for every cell; cell mode (only_flagged) only for the cells whose `pick` the update
// phase found not to hold the maximum and re-armed to pc.amb.
template <typename T>
__device__ void phase_mark_winner(const InsertArgs<T>& a, bool only_flagged) {
  const PickConst pc = make_pick(a.B);
  const int64_t S = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; base < a.B; base += kRows * S) {
    int32_t c[kRows];
    T o[kRows];
    ulonglong2 rec[kRows];
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const int64_t i = base + u * S;
      c[u] = -1;
      if (i < a.B) {
        c[u] = a.idx[i];
        o[u] = a.obj[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u)
      if (c[u] >= 0) rec[u] = *reinterpret_cast<const ulonglong2*>(&a.s.cell[c[u]]);  // (best, pick)
#pragma unroll
    for (int u = 0; u < kRows; ++u)
      if (c[u] >= 0 && ord_key((double)o[u]) == rec[u].x && (!only_flagged || rec[u].y >= pc.amb))
        atomicMax(&a.s.cell[c[u]].pick, pc.amb | (pc.rowmask - (unsigned long long)(base + u * S)));
  }
}

// ---- phase 3 ----------------------------------------------------------------------------------
template <typename T>
__device__ void phase_cleanup(const InsertArgs<T>& a) {  // validation failed: leave the store alone
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.B; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = a.idx[i];
    if (c < 0) continue;  // (a row of another rank's cells)
    Cell* cell = &a.s.cell[c];
    if (cell->best == 0ull) continue;  // untouched (no insertable row), or reset by another row of the cell already
    cell->best = cell->pick = 0ull;
    cell->w0 = cell->w1 = 0ll;
    a.s.cnt[c] = 0u;
  }
}

// list mode (cells >> rows): winners append (row, cell) to the compact list, one global atomic per
// kThreads rows
template <typename T>
__device__ void phase_select(const InsertArgs<T>& a, uint32_t* s_cnt, uint32_t* s_base) {
  const PickConst pc = make_pick(a.B);
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  constexpr int NW = kThreads / 32;
  for (int64_t b0 = (int64_t)blockIdx.x * kThreads; b0 < a.B; b0 += (int64_t)gridDim.x * kThreads) {
    const int64_t i = b0 + threadIdx.x;
    int32_t c = 0;
    bool iswin = false;
    if (i < a.B) {  // `pick` stays 0 for cells without an insertable row: no need to read status
      c = a.idx[i];
      iswin = c >= 0 && a.s.cell[c].pick == (pc.amb | (pc.rowmask - (unsigned long long)i));
    }
    const unsigned m = __ballot_sync(0xffffffffu, iswin);
    if (lane == 0) s_cnt[wi] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const uint32_t n = s_cnt[k];
        s_cnt[k] = tot;
        tot += n;
      }
      *s_base = tot ? atomicAdd(&a.s.g->n_winners, tot) : 0u;
    }
    __syncthreads();
    if (iswin) a.s.wlist[*s_base + s_cnt[wi] + __popc(m & ((1u << lane) - 1u))] = make_int2((int32_t)i, c);
    __syncthreads();
  }
}

// ---- phase 4 ----------------------------------------------------------------------------------
// What a winner does to its cell (threshold, objective, occupancy, objective-sum delta); returns through
// the accumulators. `cell` is the record before cleaning.
template <typename T>
struct UpdateAcc {
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  uint32_t nnew = 0;
  unsigned long long kbest = 0ull;
};

// Returns true when the cell was unoccupied (ATOMIC_MASK = false: the caller owns the cell's bitmap word and
// sets the bit itself).
template <typename T, bool ATOMIC_MASK>
__device__ __forceinline__ bool update_cell(const InsertArgs<T>& a, int32_t c, T o, const Cell& cell, bool mae,
                                            const FxConst& fx, const BinConst& bc, UpdateAcc<T>& acc, T tc,
                                            T stored_obj) {
  const bool occ = (tc == tc);
  const T old_obj = occ ? stored_obj : (T)0;
  T newthr;
  if (!mae) {
    newthr = o;  // :663-665
  } else {       // _compute_thresholds, :473-516 (fp64 arithmetic after NumPy promotion)
    double S;
    uint32_t kc;
    fx_read(cell, &a.s.cnt[c], fx, S, kc);
    S = (double)(T)S;
    const double k = (double)kc;
    const double tcell = (double)(occ ? tc : a.tmin);
    const double ratio = powi((double)((T)1 - a.lr), kc);
    newthr = (T)__dadd_rn(__dmul_rn(ratio, tcell), __dmul_rn(__ddiv_rn(S, k), __dsub_rn(1.0, ratio)));
  }
  a.store_thr[c] = newthr;
  a.store_obj[c] = o;
  if (!occ) {
    a.occupied[c] = 1;
    if (ATOMIC_MASK) atomicOr(&a.s.newmask[c >> 5], 1u << (c & 31));
    ++acc.nnew;
  }
  double h0, h1, h2;
  bin_split((double)(T)(o - old_obj), bc, h0, h1, h2);  // :707-710
  acc.d0 += h0;
  acc.d1 += h1;
  acc.d2 += h2;
  const unsigned long long k2 = ord_key((double)o);
  acc.kbest = k2 > acc.kbest ? k2 : acc.kbest;
  Cell z;  // scratch back to its clean state
  z.best = z.pick = 0ull;
  z.w0 = z.w1 = 0ll;
  a.s.cell[c] = z;
  if (mae && !fx.b) a.s.cnt[c] = 0u;
  return !occ;
}

// warp -> block -> global accumulation (bins add exactly, so the grouping does not matter)
template <typename T>
__device__ void update_reduce(const InsertArgs<T>& a, UpdateAcc<T> acc, double (*sh_d)[kThreads / 32], uint32_t* sh_n,
                              unsigned long long* sh_k) {
  const int lane = threadIdx.x & 31;
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    acc.d0 += __shfl_down_sync(0xffffffffu, acc.d0, o2);
    acc.d1 += __shfl_down_sync(0xffffffffu, acc.d1, o2);
    acc.d2 += __shfl_down_sync(0xffffffffu, acc.d2, o2);
    acc.nnew += __shfl_down_sync(0xffffffffu, acc.nnew, o2);
    const unsigned long long ok = __shfl_down_sync(0xffffffffu, acc.kbest, o2);
    acc.kbest = ok > acc.kbest ? ok : acc.kbest;
  }
  const int wi = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
    sh_d[0][wi] = acc.d0;
    sh_d[1][wi] = acc.d1;
    sh_d[2][wi] = acc.d2;
    sh_n[wi] = acc.nnew;
    sh_k[wi] = acc.kbest;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kThreads / 32; ++k) {
      acc.d0 += sh_d[0][k];
      acc.d1 += sh_d[1][k];
      acc.d2 += sh_d[2][k];
      acc.nnew += sh_n[k];
      acc.kbest = sh_k[k] > acc.kbest ? sh_k[k] : acc.kbest;
    }
    if (acc.d0 != 0.0) atomicAdd(&a.s.g->dsum[0], acc.d0);
    if (acc.d1 != 0.0) atomicAdd(&a.s.g->dsum[1], acc.d1);
    if (acc.d2 != 0.0) atomicAdd(&a.s.g->dsum[2], acc.d2);
    if (acc.nnew) atomicAdd(&a.s.g->n_new, acc.nnew);
    if (acc.kbest) atomicMax(&a.s.g->gbest_key, acc.kbest);
  }
  __syncthreads();
}

// list mode: one thread per winner
template <typename T>
__device__ void phase_update_list(const InsertArgs<T>& a, double (*sh_d)[kThreads / 32], uint32_t* sh_n,
                                  unsigned long long* sh_k) {
  const InsertState* g = a.s.g;
  const int64_t W = g->n_winners;
  const bool mae = a.mae;
  const double batch_absmax = __longlong_as_double((long long)g->absmax_bits);
  const BinConst bc = make_bins(batch_absmax + g->store_absmax, a.B);
  FxConst fx;
  if (mae) fx = make_fx(batch_absmax, a.B);
  UpdateAcc<T> acc;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < W; t += (int64_t)gridDim.x * blockDim.x) {
    const int2 rc = a.s.wlist[t];
    const Cell cell = a.s.cell[rc.y];
    update_cell<T, true>(a, rc.y, a.obj[rc.x], cell, mae, fx, bc, acc, a.store_thr[rc.y], a.store_obj[rc.y]);
  }
  update_reduce<T>(a, acc, sh_d, sh_n, sh_k);
}

// Short rows (every field below the bulk-copy threshold, e.g. configs[1]: 80-byte solutions, 16-byte measures):
// the thread that updates a cell also moves its winner's rows, so small archives skip the copy phase and one
// grid barrier altogether (at 100k rows the insertion is a chain of barrier and L2 latencies, not bandwidth).
constexpr int kInlineRowMax = 256;  // == kBulkMinRow: rows the bulk path would not take anyway

__device__ __forceinline__ bool inline_copy_ok(const FieldCopy& fc) {
  if (fc.peer_tab != nullptr || fc.n == 0) return false;
  int64_t total = 0;
  for (int f = 0; f < fc.n; ++f) {
    if (fc.host_src[f] || fc.row_bytes[f] >= kInlineRowMax) return false;
    total += fc.row_bytes[f];
  }
  return total <= 512;
}

__device__ __forceinline__ void copy_row_inline(const FieldCopy& fc, int row, int cell) {
  for (int f = 0; f < fc.n; ++f) {
    const int64_t rb = fc.row_bytes[f];
    const char* __restrict__ sp = fc.src[f] + (int64_t)row * rb;
    char* __restrict__ dp = fc.dst[f] + (int64_t)cell * rb;
    const int vec = fc.vec[f];
    if (vec == 16) {
      for (int64_t o = 0; o < rb; o += 16) *reinterpret_cast<uint4*>(dp + o) = *reinterpret_cast<const uint4*>(sp + o);
    } else if (vec == 8) {
      for (int64_t o = 0; o < rb; o += 8) *reinterpret_cast<uint2*>(dp + o) = *reinterpret_cast<const uint2*>(sp + o);
    } else if (vec == 4) {
      for (int64_t o = 0; o < rb; o += 4) *reinterpret_cast<uint32_t*>(dp + o) = *reinterpret_cast<const uint32_t*>(sp + o);
    } else {
      for (int64_t o = 0; o < rb; ++o) dp[o] = sp[o];
    }
  }
}

// cell mode: one thread per CELL, everything coalesced. The winner's objective is decoded from the cell's
// key, its row comes from `pick`; the (row, cell) pairs go to the list in cell order within a CTA trip (one
// reservation per trip), so the row copy writes the store in runs of ascending addresses.
// second_pass = false: every cell the batch touched. A cell whose pick does not hold the maximum (two rows of the
// cell with keys equal above bit rb, see Cell) is left for the mark pass: pick := pc.amb, counted in n_ambiguous.
// second_pass = true: exactly those cells, after the mark pass.
template <typename T>
__device__ void phase_update_cells(const InsertArgs<T>& a, const FieldCopy& fc, bool inline_copy, bool second_pass,
                                   double (*sh_d)[kThreads / 32], uint32_t* sh_n, unsigned long long* sh_k,
                                   uint32_t* s_cnt, uint32_t* s_base) {
  const InsertState* g = a.s.g;
  const bool mae = a.mae;
  const double batch_absmax = __longlong_as_double((long long)g->absmax_bits);
  const BinConst bc = make_bins(batch_absmax + g->store_absmax, a.B);
  const PickConst pc = make_pick(a.B);
  // keys of float objectives have 29 zero bits at the bottom: the truncation drops nothing
  const bool exact_pick = sizeof(T) == 4 && pc.rowmask < (1ull << 29);
  FxConst fx;
  if (mae) fx = make_fx(batch_absmax, a.B);
  UpdateAcc<T> acc;
  uint32_t namb = 0;
  const int lane = threadIdx.x & 31;
  // U tiles per trip, every load issued before the first use: the phase is 1-2 trips per thread, i.e. a chain of
  // L2 round trips (record -> pick's objective -> list slot), not bandwidth. (U = 2 spilled and measured slower.)
  constexpr int U = 1;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t c0 = (int64_t)blockIdx.x * kThreads; c0 < a.N; c0 += U * stride) {
    Cell cell[U];
    T tc[U], so[U], po[U];
    int32_t row[U];
    bool has[U], fresh[U];
    unsigned m[U], mnew[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {  // a warp covers 32 aligned cells = one word of the new-cell bitmap
      const int64_t c = c0 + u * stride + threadIdx.x;
      cell[u].best = cell[u].pick = 0ull;
      tc[u] = so[u] = (T)0;
      if (c < a.N) {
        cell[u] = a.s.cell[c];
        tc[u] = a.store_thr[c];
        so[u] = a.store_obj[c];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      has[u] = cell[u].best != 0ull;
      if (second_pass) has[u] = has[u] && cell[u].pick >= pc.amb;
      row[u] = (int32_t)(pc.rowmask - (cell[u].pick & pc.rowmask));
      po[u] = (has[u] && !second_pass && !exact_pick) ? a.obj[row[u]] : (T)0;
    }
    uint32_t nwin = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = c0 + u * stride + threadIdx.x;
      if (has[u] && !second_pass && !exact_pick && ord_key((double)po[u]) != cell[u].best) {
        a.s.cell[c].pick = pc.amb;  // nobody else writes the record in this phase
        ++namb;
        has[u] = false;
      }
      if (has[u] && inline_copy) copy_row_inline(fc, row[u], (int)c);  // _array_store.py:605-608
      fresh[u] = false;
      if (has[u])
        fresh[u] = update_cell<T, false>(a, (int32_t)c, (T)ord_unkey(cell[u].best), cell[u], mae, fx, bc, acc, tc[u], so[u]);
      mnew[u] = __ballot_sync(0xffffffffu, fresh[u]);
      m[u] = __ballot_sync(0xffffffffu, has[u]);
      nwin += __popc(m[u]);
    }
    // one list reservation per CTA and trip (a reservation per warp is 8 000 atomics on one address per add, which
    // the L2 serialises: measured +8 us)
    if (lane == 0) s_cnt[threadIdx.x >> 5] = nwin;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int k = 0; k < kThreads / 32; ++k) {
        const uint32_t n = s_cnt[k];
        s_cnt[k] = tot;
        tot += n;
      }
      *s_base = tot ? atomicAdd(&a.s.g->n_winners, tot) : 0u;
    }
    __syncthreads();
    uint32_t base = *s_base + s_cnt[threadIdx.x >> 5];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t c = c0 + u * stride + threadIdx.x;
      if (lane == 0 && mnew[u]) {
        if (second_pass) atomicOr(&a.s.newmask[c >> 5], mnew[u]);  // (the first pass may have set other bits)
        else a.s.newmask[c >> 5] = mnew[u];
      }
      if (has[u]) a.s.wlist[base + __popc(m[u] & ((1u << lane) - 1u))] = make_int2(row[u], (int32_t)c);
      base += __popc(m[u]);
    }
    __syncthreads();
  }
  namb = __reduce_add_sync(0xffffffffu, namb);
  if (lane == 0 && namb) atomicAdd(&a.s.g->n_ambiguous, namb);
  update_reduce<T>(a, acc, sh_d, sh_n, sh_k);
}

// ---- phase 5: row scatter of the winners (_array_store.py:605-608) --------------------------------
// TMA bulk copies: a lane moves one piece of a winner row (a whole row up to 2 KB) global -> shared
// (cp.async.bulk, completion on the warp's mbarrier) and shared -> global (bulk store group). No
// registers hold data in flight, so each SM keeps kCopyWarps x kSlotBytes per CTA outstanding --
// about three times what register staging at this occupancy could -- and the phase is a plain HBM
// stream of 2 * row_bytes per winner. Fields whose rows are not 16-byte aligned take the scalar path.
constexpr int kCopyWarps = 4;
constexpr int kSlotBytes = 24 * 1024;
constexpr int kPieceMax = 2048;
constexpr int kCopySmem = kCopyWarps * kSlotBytes;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

constexpr int kBulkMinRow = 256;  // rows shorter than this are cheaper through registers

__device__ __forceinline__ bool is_bulk_field(const FieldCopy& fc, int f) {
  return fc.vec[f] == 16 && fc.row_bytes[f] >= kBulkMinRow && !fc.host_src[f] && fc.peer_tab == nullptr;
}

// small / unaligned rows: a group of 2^k threads (the smallest power of two covering the row's chunks, at
// most a warp) moves one row, four rows in flight per thread; `tid` of `nthreads` is the caller's team
__device__ void copy_rows_small(const FieldCopy& fc, int f, const Scratch& s, int64_t W, int64_t tid,
                                int64_t nthreads) {
  const int vec = fc.vec[f];
  const int64_t rb = fc.row_bytes[f];
  char* __restrict__ dst = fc.dst[f];
  const int64_t nch = rb / vec;
  int lg = 0;
  while ((1 << lg) < nch && lg < 5) ++lg;
  const int gsz = 1 << lg;
  const int sub = (int)(tid & (gsz - 1));
  const int64_t grp = tid >> lg, ngrp = nthreads >> lg;
  if (ngrp == 0) return;
  for (int64_t w0 = grp; w0 < W; w0 += 4 * ngrp) {
    int2 rc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t w = w0 + u * ngrp;
      rc[u] = w < W ? s.wlist[w] : make_int2(-1, -1);
    }
    const char* sp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sp[u] = rc[u].x >= 0 ? src_row(fc, f, rc[u].x, rb) : nullptr;
    for (int64_t ch = sub; ch < nch; ch += gsz) {
      if (vec == 16) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0) v[u] = __ldcs(reinterpret_cast<const uint4*>(sp[u]) + ch);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0) __stcs(reinterpret_cast<uint4*>(dst + (int64_t)rc[u].y * rb) + ch, v[u]);
      } else if (vec == 8) {
        uint2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0) v[u] = *(reinterpret_cast<const uint2*>(sp[u]) + ch);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0) *(reinterpret_cast<uint2*>(dst + (int64_t)rc[u].y * rb) + ch) = v[u];
      } else if (vec == 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0)
            *(reinterpret_cast<uint32_t*>(dst + (int64_t)rc[u].y * rb) + ch) =
                *(reinterpret_cast<const uint32_t*>(sp[u]) + ch);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (rc[u].x >= 0) dst[(int64_t)rc[u].y * rb + ch] = sp[u][ch];
      }
    }
  }
}

// Shape of the bulk path: kBulkWarps x kBulkStages = kCopyWarps slots of 24 KB per CTA. Measured at 2M rows / 249k
// winners of 800 B (tools/grid_insert_probe.py): 4 warps x 1 stage on both CTAs of an SM 76 us, 2 warps x 2 stages
// 76 us, 2 warps x 2 stages on ONE elected CTA per SM 97 us -- although a contiguous 2 x 200 MB copy is fastest in
// that last shape (tools/overlap_probe.cu: 5.9 vs 5.5 TB/s): scattered 800-byte rows want more requests in flight.
constexpr int kBulkWarps = 4;
constexpr int kBulkStages = 1;
static_assert(kBulkWarps * kBulkStages == kCopyWarps, "one mbarrier and one 24 KB slot per (warp, stage)");

__device__ void phase_copy(const FieldCopy& fc, const Scratch& s, int64_t W, int buf, unsigned char* slots,
                           uint64_t* mbar, uint32_t& parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  bool any_bulk = false;
  for (int f = 0; f < fc.n; ++f) any_bulk |= is_bulk_field(fc, f);
  if (!any_bulk || warp >= kBulkWarps) {
    // the register path: the warps the bulk copy leaves idle
    const int lean = any_bulk ? kThreads - 32 * kBulkWarps : kThreads;
    const int local = threadIdx.x - (kThreads - lean);
    if (local >= 0)
      for (int f = 0; f < fc.n; ++f)
        if (!is_bulk_field(fc, f))
          copy_rows_small(fc, f, s, W, (int64_t)blockIdx.x * lean + local, (int64_t)gridDim.x * lean);
    return;
  }
  for (int f = 0; f < fc.n; ++f) {
    if (!is_bulk_field(fc, f)) continue;
    const int64_t rb = fc.row_bytes[f];
    const char* __restrict__ src = fc.src[f];
    char* __restrict__ dst = fc.dst[f];
    const int piece = (int)(rb < kPieceMax ? rb : kPieceMax);
    const int64_t ppr = (rb + piece - 1) / piece;  // pieces per row
    const int L = kSlotBytes / piece < 32 ? kSlotBytes / piece : 32;
    const int64_t total = W * ppr;
    // Rounds are drawn from a counter, not dealt out: SMs do not copy at the same speed (with a static split the
    // phase ended 13 us after its first CTA was done).
    unsigned long long* next = &s.copy_next[buf * QD_MAX_FIELDS + f];
    int64_t it[kBulkStages];
    char* dp[kBulkStages];
    uint32_t len[kBulkStages];
    auto issue = [&](int st) {
      unsigned long long drawn = 0;
      if (lane == 0) drawn = atomicAdd(next, (unsigned long long)L);
      it[st] = (int64_t)__shfl_sync(0xffffffffu, drawn, 0);
      len[st] = 0;
      dp[st] = nullptr;
      if (it[st] >= total) return;
      const int64_t item = it[st] + lane;
      const char* sp = nullptr;
      if (lane < L && item < total) {
        const int64_t w = ppr == 1 ? item : item / ppr;
        const int64_t off = (item - w * ppr) * piece;
        len[st] = (uint32_t)(rb - off < piece ? rb - off : piece);
        const int2 rc = s.wlist[w];
        sp = src + (int64_t)rc.x * rb + off;
        dp[st] = dst + (int64_t)rc.y * rb + off;
      }
      uint64_t* bar = &mbar[warp * kBulkStages + st];
      const uint32_t round_bytes = __reduce_add_sync(0xffffffffu, len[st]);
      if (lane == 0) mbar_expect_tx(bar, round_bytes);
      __syncwarp();
      if (len[st]) bulk_g2s(slots + (warp * kBulkStages + st) * kSlotBytes + lane * piece, sp, len[st], bar);
    };
#pragma unroll
    for (int st = 0; st < kBulkStages; ++st) issue(st);
    for (int st = 0; it[st] < total; st = (st + 1) % kBulkStages) {
      mbar_wait(&mbar[warp * kBulkStages + st], (parity >> st) & 1u);
      parity ^= 1u << st;
      if (len[st]) bulk_s2g(dp[st], slots + (warp * kBulkStages + st) * kSlotBytes + lane * piece, len[st]);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the slot may be refilled
      __syncwarp();
      issue(st);
    }
  }
  // bulk stores are complete (not just read) and ordered before the generic-proxy accesses that follow
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
}


// ---- phase 6 ----------------------------------------------------------------------------------
// Words of the new-cell bitmap owned by one CTA: a contiguous chunk, so that the concatenation of
// the CTAs' outputs in CTA order is the ascending cell order np.nonzero produces.
__device__ __forceinline__ void word_chunk(int64_t words, int64_t& w0, int64_t& w1) {
  const int64_t per = (words + gridDim.x - 1) / gridDim.x;
  w0 = min((int64_t)blockIdx.x * per, words);
  w1 = min(w0 + per, words);
}

template <typename T>
__device__ void phase_best_and_count(const InsertArgs<T>& a, bool new_record, bool any_new, uint32_t* sh) {
  const InsertState* g = a.s.g;
  if (new_record) {  // lowest winning cell holding the batch maximum (_stats_update's argmax, :331-341)
    const int64_t W = g->n_winners;
    const unsigned long long gk = g->gbest_key;
    unsigned long long m = ~0ull;  // (cell << 32 | batch row): the minimum is the lowest cell, and names its winner's row
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < W; t += (int64_t)gridDim.x * blockDim.x) {
      const int2 rc = a.s.wlist[t];
      if (ord_key((double)a.obj[rc.x]) == gk) {
        const unsigned long long pk = ((unsigned long long)(uint32_t)rc.y << 32) | (uint32_t)rc.x;
        m = pk < m ? pk : m;
      }
    }
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
      const unsigned long long o = __shfl_down_sync(0xffffffffu, m, o2);
      m = o < m ? o : m;
    }
    if ((threadIdx.x & 31) == 0 && m != ~0ull) atomicMin(&a.s.g->best_pack, m);
  }
  if (any_new) {
    int64_t w0, w1;
    word_chunk((a.N + 31) / 32, w0, w1);
    uint32_t n = 0;
    for (int64_t w = w0 + threadIdx.x; w < w1; w += blockDim.x) n += __popc(a.s.newmask[w]);
    n = __reduce_add_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0u;
      v = __reduce_add_sync(0xffffffffu, v);
      if (threadIdx.x == 0) a.s.ctacnt[blockIdx.x] = v;
    }
    __syncthreads();
  }
}

// ---- phase 7 ----------------------------------------------------------------------------------
// from_source: the best elite's rows are taken from the batch (the row copy has not run yet: QD_INSERT_ROWS);
// save_w >= 0: the winner count is kept in copy_w[save_w] for the QD_INSERT_COPY launch that follows
template <typename T>
__device__ void phase_publish(const InsertArgs<T>& a, const FieldCopy& fc, bool committed, bool new_record,
                              bool any_new, int64_t n_before, uint32_t* sh, uint32_t* sh_base, bool from_source,
                              int save_w, int rearm_buf) {
  InsertState* gp = a.s.g;
  const int lane = threadIdx.x & 31;
  if (any_new) {
    // exclusive offset of this CTA = sum of the counts of the CTAs before it
    uint32_t part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += blockDim.x) part += a.s.ctacnt[b];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0u;
      v = __reduce_add_sync(0xffffffffu, v);
      if (threadIdx.x == 0) *sh_base = v;
    }
    __syncthreads();
    int64_t pos0 = n_before + *sh_base;
    __syncthreads();
    int64_t w0, w1;
    word_chunk((a.N + 31) / 32, w0, w1);
    for (int64_t wb = w0; wb < w1; wb += blockDim.x) {  // ordered tiles of one word per thread
      const int64_t w = wb + threadIdx.x;
      uint32_t m = w < w1 ? a.s.newmask[w] : 0u;
      const uint32_t n = __popc(m);
      uint32_t incl = n;
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o2);
        if (lane >= o2) incl += t;
      }
      if (lane == 31) sh[threadIdx.x >> 5] = incl;
      __syncthreads();
      if (threadIdx.x < 32) {
        const uint32_t v = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0u;
        uint32_t iv = v;
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, iv, o2);
          if (lane >= o2) iv += t;
        }
        sh[threadIdx.x] = iv - v;  // exclusive warp offsets
        if (threadIdx.x == 31) *sh_base = iv;
      }
      __syncthreads();
      int64_t pos = pos0 + sh[threadIdx.x >> 5] + (incl - n);
      if (m) {
        a.s.newmask[w] = 0u;
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          a.occupied_list[pos++] = (int32_t)(w * 32 + bit);
        }
      }
      pos0 += *sh_base;
      __syncthreads();
    }
  }
  if (blockIdx.x != 0) return;
  // CTA 0: best-elite snapshot, then stats + result + re-arm
  const unsigned long long best_pack = gp->best_pack;
  const int32_t best_cell = best_pack != ~0ull ? (int32_t)(best_pack >> 32) : INT32_MAX;
  const bool snap = committed && new_record && best_cell != INT32_MAX && a.snapshot != nullptr;
  if (snap) {
    for (int f = 0; f < fc.n; ++f) {
      const char* sp = from_source ? src_row(fc, f, (int)(uint32_t)best_pack, fc.row_bytes[f])
                                   : fc.dst[f] + (int64_t)best_cell * fc.row_bytes[f];
      char* dp = a.snapshot + fc.snap_off[f];
      for (int64_t o = threadIdx.x; o < fc.row_bytes[f]; o += blockDim.x) dp[o] = sp[o];
    }
  }
  __syncthreads();  // every warp has read best_pack before thread 0 re-arms it
  if (threadIdx.x < QD_MAX_FIELDS) a.s.copy_next[rearm_buf * QD_MAX_FIELDS + threadIdx.x] = 0ull;
  if (threadIdx.x == 0) {
    InsertState g = *gp;
    const double delta = (double)(T)__dadd_rn(__dadd_rn(g.dsum[0], g.dsum[1]), g.dsum[2]);
    if (committed) {
      g.n_occupied += g.n_new;
      g.objective_sum = (double)(T)(g.objective_sum + delta);  // _grid_archive.py:707-712
      g.n_commits += 1;
      g.store_absmax = fmax(g.store_absmax, __longlong_as_double((long long)g.absmax_bits));
      if (new_record) {
        g.obj_max_key = g.gbest_key;
        g.best_serial += 1;
        if (a.snapshot) {  // header: objective, threshold (as doubles), cell
          double* hd = (double*)a.snapshot;
          hd[0] = (double)a.store_obj[best_cell];
          hd[1] = (double)a.store_thr[best_cell];
          *(int32_t*)(hd + 2) = best_cell;
        }
      }
    }
    if (g.flags) {
      g.n_failed += 1;
      g.last_failed_flags = g.flags;
    }
    qd_add_result r;
    r.n_insertable = (int32_t)g.n_insertable;
    r.n_new = committed ? (int32_t)g.n_new : 0;
    r.best_cell = (committed && new_record && best_cell != INT32_MAX) ? best_cell : -1;
    r.flags = g.flags;
    r.best_objective = g.gbest_key ? ord_unkey(g.gbest_key) : 0.0;
    r.delta_sum[0] = g.dsum[0];
    r.delta_sum[1] = g.dsum[1];
    r.delta_sum[2] = g.dsum[2];
    r.batch_absmax = __longlong_as_double((long long)g.absmax_bits);
    r.n_occupied = g.n_occupied;
    r.objective_sum = g.objective_sum;
    r.obj_max = g.obj_max_key ? ord_unkey(g.obj_max_key) : 0.0;
    r.has_obj_max = g.obj_max_key ? 1 : 0;
    r.last_failed_flags = g.last_failed_flags;
    r.n_commits = g.n_commits;
    r.best_serial = g.best_serial;
    r.n_failed = g.n_failed;
    r.n_winners = committed ? g.n_winners : 0;
    *a.result = r;
    if (g.result_mirror) *g.result_mirror = r;  // the host reads it after its stream synchronize: no D2H call
    if (save_w >= 0) g.copy_w[save_w] = committed ? g.n_winners : 0u;
    g.gbest_key = 0ull;
    g.absmax_bits = 0ull;
    g.best_cell = INT32_MAX;
    g.best_pack = ~0ull;
    g.n_insertable = g.n_new = g.flags = g.n_winners = g.n_ambiguous = 0u;

    g.dsum[0] = g.dsum[1] = g.dsum[2] = 0.0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g.t_phase[8]));
    *gp = g;
  }
}

// ---- the kernel ------------------------------------------------------------------------------------
// Phases [phase_lo, phase_hi] run inside one launch; a launch boundary stands in for the barrier
// when add() is split in two (host batches: the row copy waits for the H2D of the solutions).
constexpr int64_t kSingleMaxRows = 2048;      // one CTA handles the whole batch up to here ...
constexpr int64_t kSingleMaxCells = 1 << 18;  // ... in archives whose cell passes one CTA walks in a few microseconds
constexpr int64_t kSingleMaxWork = 1 << 15;   // ... and at most 512 KB of winner rows to move (16-byte units)

// SINGLE: the whole add() in ONE CTA, launched plainly -- small batches (the reference's own benchmark protocol
// adds 100 rows at a time, benchmarks/cvt_add.py:79-86) are a chain of barrier latencies, and a block barrier
// plus fence costs a tenth of a cooperative grid barrier (and the launch itself is cheaper).
template <typename T, typename MT, int D, bool FUSED, bool SINGLE>
__global__ void __launch_bounds__(kThreads, 2) insert_kernel(const __grid_constant__ InsertArgs<T> a_in,
                                                           const __grid_constant__ GridParams gp,
                                                           const __grid_constant__ FieldCopy fc, int phase_lo,
                                                           int phase_hi, int mode, int buf) {
  // mode 0: phases [phase_lo, phase_hi] in order. mode 1 (QD_INSERT_ROWS): everything but the row copy -- phases
  // 0-4, then best / count and publish with the best elite taken from the batch; the winner list stays in list
  // buffer `buf` (a.s.wlist points at it). mode 2 (QD_INSERT_COPY): only the row copy of that list, a plain
  // launch on another stream that overlaps the row phases of the next add().
  auto gsync = [&]() {
    if constexpr (SINGLE) {
      __threadfence();  // (MEMBAR + L1 invalidate: REDs of the other threads land in L2)
      __syncthreads();
      __threadfence();
    } else {
      cg::this_grid().sync();
    }
  };
  extern __shared__ __align__(128) unsigned char copy_slots[];
  __shared__ uint64_t copy_bar[kCopyWarps];
  uint32_t copy_parity = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < kCopyWarps; ++k) mbar_init(&copy_bar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  __shared__ uint32_t sh_u[64];
  __shared__ uint32_t sh_base;
  __shared__ double sh_d[3][kThreads / 32];
  __shared__ unsigned long long sh_k[kThreads / 32];
  InsertArgs<T> a_staged;
  const InsertArgs<T>* ap = &a_in;
  if constexpr (SINGLE) {
    // the objectives are read by up to four phases; a small batch handed over in page-locked host memory would pay
    // a PCIe round trip in each: one coalesced read into shared memory instead
    __shared__ T obj_stage[kSingleMaxRows];
    for (int64_t i = threadIdx.x; i < a_in.B; i += kThreads) obj_stage[i] = a_in.obj[i];
    a_staged = a_in;
    a_staged.obj = obj_stage;
    ap = &a_staged;
    __syncthreads();
  }
  const InsertArgs<T>& a = *ap;
  const InsertState* g = a.s.g;
  if (mode == 2) {
    const int64_t W = g->copy_w[buf];
    if (W > 0 && fc.n > 0) phase_copy(fc, a.s, W, buf, copy_slots, copy_bar, copy_parity);
    return;
  }
  const bool mae = a.mae;
  const int64_t n_before = g->n_occupied;  // CTA 0 advances it in the publish phase
  auto stamp = [&](int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.s.g->t_phase[k] = t;
    }
  };
  if (phase_lo <= 0) stamp(0);
  if (phase_lo <= 0 && mae) {
    phase_absmax<T>(a, &sh_d[0][0]);
    gsync();
  }
  if (phase_lo <= 0) stamp(1);
  // cell mode: the update runs over the cells (coalesced) and emits a cell-sorted winner list; list mode
  // (archives much larger than the batch) never touches cells the batch did not
  const bool cell_mode = a.N <= 16 * a.B;
  if (phase_lo <= 1 && phase_hi >= 1) {
    phase_classify<T, MT, D, FUSED>(a, gp, cell_mode, &sh_d[0][0], sh_u);
    gsync();
  }
  if (phase_lo <= 1) stamp(2);
  // the globals below were last written before a barrier / launch boundary: uniform over the grid
  const bool failed = g->flags != 0u;
  const bool any = g->n_insertable != 0u;
  const bool committed = any && !failed;
  if (phase_lo <= 2 && phase_hi >= 2 && committed && !cell_mode) {
    phase_mark_winner<T>(a, false);
    gsync();
  }
  if (phase_lo <= 2) stamp(3);
  if (phase_lo <= 3 && phase_hi >= 3 && any && (failed || !cell_mode)) {
    if (failed)
      phase_cleanup<T>(a);
    else
      phase_select<T>(a, sh_u, &sh_base);
    gsync();
  }
  if (phase_lo <= 3) stamp(4);
  // one launch, cell mode, short rows: the update phase copies the winners' rows itself
  const bool inline_copy = mode == 0 && phase_lo <= 4 && phase_hi >= 5 && cell_mode && inline_copy_ok(fc);
  if (phase_lo <= 4 && phase_hi >= 4 && committed) {
    if (cell_mode) {
      phase_update_cells<T>(a, fc, inline_copy, false, sh_d, sh_u + 32, sh_k, sh_u, &sh_base);
      stamp(9);  // (diagnostics: CTA 0 is through its cells; the rest of the phase is the barrier)
      gsync();
      if (g->n_ambiguous != 0u) {  // near-ties below bit rb inside a cell: the exact row pass, for those cells only
        phase_mark_winner<T>(a, true);
        gsync();
        phase_update_cells<T>(a, fc, inline_copy, true, sh_d, sh_u + 32, sh_k, sh_u, &sh_base);
        gsync();
      }
    } else {
      phase_update_list<T>(a, sh_d, sh_u + 32, sh_k);
      gsync();
    }
  }
  if (phase_lo <= 4) stamp(5);
  if (phase_hi < 5 && mode == 0) return;
  // winner list, best key and new-cell bitmap are complete (barrier above / launch boundary)
  const unsigned long long gk = g->gbest_key, pk = g->obj_max_key;
  const bool new_record = committed && gk != 0ull && (pk == 0ull || gk > pk);
  const bool any_new = committed && g->n_new != 0u;
  if (new_record || any_new) phase_best_and_count<T>(a, new_record, any_new, sh_u);
  if (mode == 0 && committed && fc.n > 0 && !inline_copy)
    phase_copy(fc, a.s, g->n_winners, buf, copy_slots, copy_bar, copy_parity);
  stamp(6);
  if (phase_hi < 6 && mode == 0) return;
  gsync();
  stamp(7);
  phase_publish<T>(a, fc, committed, new_record, any_new, n_before, sh_u, &sh_base, mode == 1, mode == 1 ? buf : -1,
                   buf);
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
static int launch_insert(InsertArgs<T> a, const GridParams& gp, const FieldCopy& fc, int64_t work, int part,
                         cudaStream_t stream) {
  static int max_ctas = 0, sm_count = kNumSMs;  // per instantiation
  if (max_ctas == 0) {

    int per_sm = 0, dev = 0, sms = kNumSMs;
    QD_CHECK_CUDA(cudaFuncSetAttribute(insert_kernel<T, MT, D, FUSED, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kCopySmem));
    QD_CHECK_CUDA(cudaFuncSetAttribute(insert_kernel<T, MT, D, FUSED, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kCopySmem));
    QD_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, insert_kernel<T, MT, D, FUSED, false>,
                                                                kThreads, kCopySmem));
    QD_CHECK_CUDA(cudaGetDevice(&dev));
    QD_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (per_sm < 1) {
      set_error("insert_kernel cannot be made resident");
      return QD_ERR_CUDA;
    }
    int64_t m = (int64_t)per_sm * sms;
    max_ctas = (int)(m < kMaxCtas ? m : kMaxCtas);
    sm_count = sms;
  }
  const int kind = part & 0xff, buf = (part >> 8) & 1;
  a.s.wlist += (int64_t)buf * (int64_t)(align_up((size_t)a.N * 8, 256) / 8);
  int64_t want = (work + kThreads - 1) / kThreads;
  if (want < 8) want = 8;
  // the two halves of a pipelined add() share the SMs: one CTA per SM each
  const int cap = (kind == QD_INSERT_ROWS || kind == QD_INSERT_COPY) ? (sm_count < max_ctas ? sm_count : max_ctas)
                                                                     : max_ctas;
  const int grid = (int)(want < cap ? want : cap);
  int lo = kind == QD_INSERT_COMMIT ? 5 : 0, hi = (kind == QD_INSERT_CLASSIFY || kind == QD_INSERT_ROWS) ? 4 : 7;
  int mode = kind == QD_INSERT_ROWS ? 1 : kind == QD_INSERT_COPY ? 2 : 0;
  int bufi = buf;
  if (mode == 2) {  // no grid barrier inside: a plain launch, free to run next to the following add()'s row phases
    insert_kernel<T, MT, D, FUSED, false><<<grid, kThreads, kCopySmem, stream>>>(a, gp, fc, 5, 5, mode, bufi);
    return check_launch("insert_kernel(copy)");
  }
  if (mode == 0 && a.B <= kSingleMaxRows && a.N <= kSingleMaxCells && work <= kSingleMaxWork && fc.peer_tab == nullptr) {
    insert_kernel<T, MT, D, FUSED, true><<<1, kThreads, kCopySmem, stream>>>(a, gp, fc, lo, hi, mode, bufi);
    return check_launch("insert_kernel(single)");
  }
  void* params[] = {(void*)&a, (void*)&gp, (void*)&fc, (void*)&lo, (void*)&hi, (void*)&mode, (void*)&bufi};
  QD_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)insert_kernel<T, MT, D, FUSED, false>, dim3(grid),
                                            dim3(kThreads), params, mode == 1 ? 0 : kCopySmem, stream));
  return check_launch("insert_kernel");
}

template <typename T>
int insert_impl(const qd_store* st, int64_t B, int32_t* idx, const GridSpec& gs, const void* objective,
                const void* const* field_src, double lr, double tmin, int32_t* status, void* value,
                qd_add_result* result, int part, cudaStream_t stream, const void* const* peer_tab = nullptr,
                int64_t rows_per_rank = 0, int32_t cell_lo = 0, int32_t cell_hi = 0, void* const* out_tab = nullptr) {
  const int64_t N = st->capacity;
  InsertArgs<T> a;
  a.B = B;
  a.N = N;
  a.idx = idx;
  a.measures = gs.measures;
  a.obj = (const T*)objective;
  a.occupied = st->occupied;
  a.occupied_list = st->occupied_list;
  a.store_obj = (T*)st->objective;
  a.store_thr = (T*)st->threshold;
  a.tmin = (T)tmin;
  a.lr = (T)lr;
  a.mae = (tmin == -INFINITY) ? 0 : 1;
  a.status = status;
  a.value = (T*)value;
  a.snapshot = (char*)st->best_snapshot;
  a.result = result;
  a.cell_lo = cell_lo;
  a.cell_hi = cell_hi;
  a.out_tab = out_tab;
  a.out_rows_per_rank = rows_per_rank;
  a.s = carve(st->scratch, N);
  FieldCopy fc;
  fc.n = st->n_fields;
  fc.peer_tab = reinterpret_cast<const char* const*>(peer_tab);
  fc.rows_per_rank = rows_per_rank;
  int64_t max_row = 0;
  int64_t off = 32;  // snapshot header: objective, threshold, cell
  for (int f = 0; f < fc.n; ++f) {
    fc.src[f] = peer_tab ? nullptr : (const char*)field_src[f];
    fc.dst[f] = (char*)st->field_dst[f];
    fc.row_bytes[f] = st->field_row_bytes[f];
    fc.host_src[f] = 0;
    // A source in page-locked host memory (zero-copy: only the winners' rows cross PCIe, read by the copy
    // phase itself) is addressed through its device alias.
    cudaPointerAttributes pa;
    if (fc.src[f] && cudaPointerGetAttributes(&pa, fc.src[f]) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
      QD_REQUIRE(pa.devicePointer != nullptr, "host field source is not mapped into the device address space");
      fc.src[f] = (const char*)pa.devicePointer;
      fc.host_src[f] = 1;
    } else {
      cudaGetLastError();  // unregistered host pointers make the query fail: not ours to handle here
    }
    fc.vec[f] = pick_vec(fc.src[f], fc.dst[f], fc.row_bytes[f]);
    fc.snap_off[f] = off;
    off += (int64_t)align_up((size_t)fc.row_bytes[f], 16);
    max_row += fc.row_bytes[f];
  }
  // enough CTAs for the row passes and for the copy of up to min(B, N) winner rows
  const int64_t wmax = B < N ? B : N;
  int64_t work = B;
  if (wmax * (max_row / 16) > work) work = wmax * (max_row / 16);
  if (gs.measures == nullptr) return launch_insert<T, double, 0, false>(a, gs.gp, fc, work, part, stream);
  if (gs.meas_dtype == QD_F64) {
    if (gs.gp.d == 2) return launch_insert<T, double, 2, true>(a, gs.gp, fc, work, part, stream);
    return launch_insert<T, double, 0, true>(a, gs.gp, fc, work, part, stream);
  }
  if (gs.gp.d == 2) return launch_insert<T, float, 2, true>(a, gs.gp, fc, work, part, stream);
  return launch_insert<T, float, 0, true>(a, gs.gp, fc, work, part, stream);
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

// ---- store state helpers -------------------------------------------------------------------------
template <typename T>
__global__ void store_clear_kernel(int64_t N, uint8_t* occupied, T* thr, Scratch s) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t c = i; c < N; c += (int64_t)gridDim.x * blockDim.x) {
    occupied[c] = 0;
    thr[c] = (T)NAN;
  }
  if (i == 0) {
    InsertState* g = s.g;
    g->n_occupied = 0;
    g->obj_max_key = 0ull;
    g->objective_sum = 0.0;
    g->store_absmax = 0.0;
  }
}

template <typename T>
__global__ void store_mask_thresholds_kernel(int64_t N, const uint8_t* occupied, T* thr) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x)
    if (!occupied[c]) thr[c] = (T)NAN;
}

__global__ void store_set_state_kernel(Scratch s, long long n_occupied, double objective_sum, int has_max,
                                       double obj_max, double absmax) {
  InsertState* g = s.g;
  g->n_occupied = n_occupied;
  g->objective_sum = objective_sum;
  g->obj_max_key = has_max ? ord_key(obj_max) : 0ull;
  g->store_absmax = absmax;
}

__global__ void store_read_state_kernel(Scratch s, qd_add_result* out) {
  InsertState* g = s.g;
  qd_add_result r;
  memset(&r, 0, sizeof(r));
  r.best_cell = -1;
  r.n_occupied = g->n_occupied;
  r.objective_sum = g->objective_sum;
  r.obj_max = g->obj_max_key ? ord_unkey(g->obj_max_key) : 0.0;
  r.has_obj_max = g->obj_max_key ? 1 : 0;
  r.last_failed_flags = g->last_failed_flags;
  r.n_commits = g->n_commits;
  r.best_serial = g->best_serial;
  r.n_failed = g->n_failed;
  *out = r;
}

}  // namespace qd

extern "C" size_t qd_insert_scratch_bytes(int64_t capacity) { return qd::scratch_bytes(capacity); }

extern "C" uint32_t* qd_insert_flags_ptr(void* scratch, int64_t capacity) {
  return &qd::carve(scratch, capacity).g->flags;
}

extern "C" int qd_insert_phase_times(const void* scratch, int64_t capacity, uint64_t* out_ns, void* stream) {
  using namespace qd;
  QD_REQUIRE(scratch && out_ns && capacity >= 1, "bad arguments");
  Scratch s = carve(const_cast<void*>(scratch), capacity);
  QD_CHECK_CUDA(cudaMemcpyAsync(out_ns, s.g->t_phase, 10 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                (cudaStream_t)stream));
  QD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return QD_OK;
}

extern "C" int qd_insert_set_result_mirror(void* scratch, int64_t capacity, qd_add_result* host_mirror, void* stream) {
  using namespace qd;
  QD_REQUIRE(scratch && capacity >= 1, "bad scratch arguments");
  Scratch s = carve(scratch, capacity);
  if (host_mirror) {
    cudaPointerAttributes pa;
    QD_CHECK_CUDA(cudaPointerGetAttributes(&pa, host_mirror));
    QD_REQUIRE(pa.type == cudaMemoryTypeHost && pa.devicePointer != nullptr, "mirror must be page-locked host memory");
    host_mirror = (qd_add_result*)pa.devicePointer;
  }
  QD_CHECK_CUDA(cudaMemcpyAsync(&s.g->result_mirror, &host_mirror, sizeof(host_mirror), cudaMemcpyHostToDevice,
                                (cudaStream_t)stream));
  QD_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // `host_mirror` is a stack variable
  return QD_OK;
}

extern "C" int qd_insert_scratch_init(void* scratch, int64_t capacity, void* stream) {
  using namespace qd;
  QD_REQUIRE(scratch && capacity >= 1, "bad scratch arguments");
  Scratch s = carve(scratch, capacity);
  scratch_init_kernel<<<grid_for(capacity, 256), 256, 0, (cudaStream_t)stream>>>(s, capacity);
  return check_launch("scratch_init_kernel");
}

extern "C" size_t qd_best_snapshot_bytes(const qd_store* store) {
  if (!store) return 0;
  size_t off = 32;
  for (int f = 0; f < store->n_fields; ++f) off += qd::align_up((size_t)store->field_row_bytes[f], 16);
  return off;
}

extern "C" int qd_store_clear(const qd_store* store, void* stream) {
  using namespace qd;
  QD_REQUIRE(store && store->scratch && store->occupied && store->threshold, "null store pointer");
  Scratch s = carve(store->scratch, store->capacity);
  const int grid = grid_for(store->capacity, 256);
  if (store->obj_dtype == QD_F64)
    store_clear_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(store->capacity, store->occupied,
                                                                       (double*)store->threshold, s);
  else
    store_clear_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(store->capacity, store->occupied,
                                                                      (float*)store->threshold, s);
  return check_launch("store_clear_kernel");
}

extern "C" int qd_store_set_state(const qd_store* store, int64_t n_occupied, double objective_sum, int has_obj_max,
                                  double obj_max, double objective_absmax, void* stream) {
  using namespace qd;
  QD_REQUIRE(store && store->scratch && store->occupied && store->threshold, "null store pointer");
  Scratch s = carve(store->scratch, store->capacity);
  const int grid = grid_for(store->capacity, 256);
  if (store->obj_dtype == QD_F64)
    store_mask_thresholds_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(store->capacity, store->occupied,
                                                                                 (double*)store->threshold);
  else
    store_mask_thresholds_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(store->capacity, store->occupied,
                                                                                (float*)store->threshold);
  int rc = check_launch("store_mask_thresholds_kernel");
  if (rc) return rc;
  store_set_state_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(s, (long long)n_occupied, objective_sum, has_obj_max,
                                                          obj_max, objective_absmax);
  return check_launch("store_set_state_kernel");
}

extern "C" int qd_store_read_state(const qd_store* store, qd_add_result* out, void* stream) {
  using namespace qd;
  QD_REQUIRE(store && store->scratch && out, "null pointer");
  Scratch s = carve(store->scratch, store->capacity);
  store_read_state_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(s, out);
  return check_launch("store_read_state_kernel");
}

extern "C" int qd_insert(const qd_store* store, int64_t B, const int32_t* idx, const void* objective,
                         const void* const* field_src, double learning_rate, double threshold_min,
                         int32_t* status_out, void* value_out, qd_add_result* result, int part, void* stream) {
  return qd_insert_sharded_rows(store, B, idx, objective, field_src, nullptr, 0, learning_rate, threshold_min,
                                status_out, value_out, result, part, stream);
}

extern "C" int qd_insert_sharded_rows(const qd_store* store, int64_t B, const int32_t* idx, const void* objective,
                                      const void* const* field_src, const void* const* peer_field_src,
                                      int64_t rows_per_rank, double learning_rate, double threshold_min,
                                      int32_t* status_out, void* value_out, qd_add_result* result, int part,
                                      void* stream) {
  using namespace qd;
  int rc = check_insert_args(store, B, objective, status_out, value_out, result);
  if (rc) return rc;
  QD_REQUIRE((part & 0xff) >= QD_INSERT_WHOLE && (part & 0xff) <= QD_INSERT_COPY && (part & ~0x1ff) == 0, "bad part");
  if (B > 0) QD_REQUIRE(idx, "null idx");
  if (peer_field_src) QD_REQUIRE(rows_per_rank >= 1, "rows_per_rank must be positive");
  GridSpec gs;
  gs.gp.d = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (store->obj_dtype == QD_F64)
    return insert_impl<double>(store, B, const_cast<int32_t*>(idx), gs, objective, field_src, learning_rate,
                               threshold_min, status_out, value_out, result, part, s, peer_field_src, rows_per_rank);
  return insert_impl<float>(store, B, const_cast<int32_t*>(idx), gs, objective, field_src, learning_rate,
                            threshold_min, status_out, value_out, result, part, s, peer_field_src, rows_per_rank);
}

extern "C" int qd_insert_owned_cells(const qd_store* store, int64_t B, int32_t* idx, const void* objective,
                                     const void* const* peer_field_src, int64_t rows_per_rank, int32_t cell_lo,
                                     int32_t cell_hi, void* const* out_tab, double learning_rate,
                                     double threshold_min, qd_add_result* result, void* stream) {
  using namespace qd;
  QD_REQUIRE(store && result, "null store/result");
  QD_REQUIRE(idx && objective && peer_field_src && out_tab, "null pointer");
  QD_REQUIRE(rows_per_rank >= 1 && B >= 1 && B < ((int64_t)1 << 31), "bad batch shape");
  QD_REQUIRE(cell_lo >= 0 && cell_hi > cell_lo && (int64_t)cell_hi - cell_lo == store->capacity,
             "the store must hold exactly the cells [cell_lo, cell_hi)");
  // (status / value go through out_tab: check_insert_args only needs non-null placeholders)
  int rc = check_insert_args(store, B, objective, reinterpret_cast<const int32_t*>(out_tab), out_tab, result);
  if (rc) return rc;
  GridSpec gs;
  gs.gp.d = 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (store->obj_dtype == QD_F64)
    return insert_impl<double>(store, B, idx, gs, objective, nullptr, learning_rate, threshold_min, nullptr, nullptr,
                               result, QD_INSERT_WHOLE, s, peer_field_src, rows_per_rank, cell_lo, cell_hi, out_tab);
  return insert_impl<float>(store, B, idx, gs, objective, nullptr, learning_rate, threshold_min, nullptr, nullptr,
                            result, QD_INSERT_WHOLE, s, peer_field_src, rows_per_rank, cell_lo, cell_hi, out_tab);
}

extern "C" int qd_grid_insert(const qd_store* store, int64_t B, const void* measures, int meas_dtype, int d,
                              const int32_t* dims, const double* lower, const double* interval, double epsilon,
                              const void* objective, const void* const* field_src, double learning_rate,
                              double threshold_min, int32_t* idx_out, int32_t* status_out, void* value_out,
                              qd_add_result* result, int part, void* stream) {
  using namespace qd;
  int rc = check_insert_args(store, B, objective, status_out, value_out, result);
  if (rc) return rc;
  QD_REQUIRE(meas_dtype == QD_F32 || meas_dtype == QD_F64, "measures dtype must be f32/f64");
  QD_REQUIRE((part & 0xff) >= QD_INSERT_WHOLE && (part & 0xff) <= QD_INSERT_COPY && (part & ~0x1ff) == 0, "bad part");
  if (B > 0) QD_REQUIRE(measures && idx_out, "null measures / idx_out");
  GridSpec gs;
  if ((rc = fill_grid_params(gs.gp, d, dims, lower, interval, epsilon))) return rc;
  gs.measures = measures;
  gs.meas_dtype = meas_dtype;
  cudaStream_t s = (cudaStream_t)stream;
  if (store->obj_dtype == QD_F64)
    return insert_impl<double>(store, B, idx_out, gs, objective, field_src, learning_rate, threshold_min, status_out,
                               value_out, result, part, s);
  return insert_impl<float>(store, B, idx_out, gs, objective, field_src, learning_rate, threshold_min, status_out,
                            value_out, result, part, s);
}
