/*
 * qd_b200.h -- C ABI of the B200-native quality-diversity hot path.
 *
 * The reference (icaros-usc/pyribs) is pure Python and has no FFI for this path; the
 * boundary below is what a reference-side binding (ctypes) binds in
 * place of the NumPy/SciPy bodies cited on each entry point (paths relative to the
 * reference checkout).
 *
 * Conventions
 *   - every pointer marked [dev] is a CUDA device pointer on the current device, every
 *     pointer marked [host] is ordinary host memory read during the call;
 *   - no entry point allocates device memory: callers pass workspaces whose sizes the
 *     *_bytes() functions report;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises unless stated;
 *   - return value: QD_OK (0) or a negative QD_ERR_* code; qd_last_error() gives the
 *     text for the calling thread. No C++ exception crosses the ABI.
 */
#ifndef QD_B200_H_
#define QD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QD_ABI_VERSION 6

#define QD_OK 0
#define QD_ERR_INVALID_ARG (-1)
#define QD_ERR_CUDA (-2)
#define QD_ERR_UNSUPPORTED (-3)

/* element types of archive fields */
#define QD_F32 0
#define QD_F64 1

#define QD_MAX_MEASURE_DIM 32
#define QD_MAX_FIELDS 16

/* bits of the device-side `flags` word written by the index / insert kernels */
#define QD_FLAG_NONFINITE_MEASURES 1u
#define QD_FLAG_NONFINITE_OBJECTIVE 2u
#define QD_FLAG_CVT_OVERFLOW 4u /* tensor-core search ran out of candidate space: rerun with method 1 */

int qd_abi_version(void);
const char* qd_last_error(void);
/* Number of kernels this library has launched in the calling process (all threads). */
int64_t qd_launch_count(void);
/* cudaStreamSynchronize(stream): the one blocking call of add() (the host needs qd_add_result). */
int qd_stream_synchronize(void* stream);
/* cudaMemcpyAsync / cudaEventRecord / cudaStreamWaitEvent on raw handles (kind: 1 H2D, 2 D2H, 3 D2D, else
 * default), so that a binding can order its copies around the kernels with one foreign call each. */
int qd_memcpy_async(void* dst, const void* src, size_t nbytes, int kind, void* stream);
int qd_event_record(void* event /*cudaEvent_t*/, void* stream);
int qd_stream_wait_event(void* stream, void* event /*cudaEvent_t*/);

/* ------------------------------------------------------------------------------------
 * A2  GridArchive.index_of      ribs/archives/_grid_archive.py:362-410, 432-450
 *   idx[i] = ravel(clip(trunc_int32((dims*(m-lb)+eps)/interval), 0, dims-1))
 *   `lower`, `interval`, `epsilon` hold the archive-dtype values widened to double.
 *   flags |= QD_FLAG_NONFINITE_MEASURES when any measure is inf/NaN
 *   (ribs/_utils.py:18-31 check_finite).
 * ---------------------------------------------------------------------------------- */
int qd_grid_index_of(const void* measures /*[dev] B x d*/, int dtype, int64_t B, int d,
                     const int32_t* dims /*[host] d*/, const double* lower /*[host] d*/,
                     const double* interval /*[host] d*/, double epsilon,
                     int32_t* idx_out /*[dev] B*/, uint32_t* flags /*[dev] 1*/, void* stream);

/* ------------------------------------------------------------------------------------
 * A3  CVTArchive.index_of       ribs/archives/_cvt_archive.py:579-644
 *   Nearest centroid under squared Euclidean distance, evaluated in fp64 with NumPy's
 *   summation order; exact ties resolve to the lowest centroid index (np.argmin,
 *   _cvt_archive.py:632).
 *
 *   qd_cvt_prepare builds the tensor-core operand image of a centroid shard once
 *   (CVTArchive.__init__, _cvt_archive.py:410-429 builds the KD-tree at this point).
 *   qd_cvt_index_of searches one shard: cell ids are `index_base + local row`;
 *   `dist_out` (optional) receives the exact fp64 squared distance of the winner so
 *   shards on different GPUs can be merged with a min-reduction.
 *   method: 0 = tensor-core filter (tcgen05) + fp64 rescoring, 1 = plain fp64 scan,
 *           2 = bucket grid (measure_dim <= 3; the GPU analogue of the reference's default
 *               KD-tree, _cvt_archive.py:605-609): `prepared` is then the structure built by
 *               qd_cvt_bucket_prepare.
 *           3 = bounding-box tree (any measure_dim; the same analogue for measure_dim > 3,
 *               one warp per query, exact fp64 leaves): `prepared` is the structure built by
 *               qd_cvt_tree_prepare. All four return identical indices.
 * ---------------------------------------------------------------------------------- */
size_t qd_cvt_prepared_bytes(int64_t C, int d);
int qd_cvt_prepare(const void* centroids /*[dev] C x d*/, int dtype, int64_t C, int d,
                   void* prepared /*[dev]*/, void* stream);
/* Bucket grid over a centroid shard (built where the reference builds its KD-tree,
 * _cvt_archive.py:410-429); 0 bytes = unsupported measure_dim. */
size_t qd_cvt_bucket_bytes(int64_t C, int d);
int qd_cvt_bucket_prepare(const void* centroids /*[dev] C x d*/, int dtype, int64_t C, int d,
                          void* buckets /*[dev] qd_cvt_bucket_bytes*/, void* stream);
/* Bounding-box tree over a centroid shard (KD order, 32 points per leaf, 32 children per node), built on the
 * host where the reference builds its KD-tree (_cvt_archive.py:410-429) and copied to `tree`. The _host variant
 * takes and fills HOST memory (used by qd_cvt_tree_prepare and by the CPU tests of the structure). */
size_t qd_cvt_tree_bytes(int64_t C, int d);
/* Scratch of a method-3 search (`scratch` of qd_cvt_index_of): large batches are counting-sorted by their home
 * leaf first so that neighbouring warps walk neighbouring parts of the tree. 0 = none needed; without it the
 * search runs in input order (same results). */
size_t qd_cvt_tree_scratch_bytes(int64_t B, int64_t C);
int qd_cvt_tree_prepare(const void* centroids /*[dev] C x d*/, int dtype, int64_t C, int d,
                        void* tree /*[dev] qd_cvt_tree_bytes*/, void* stream);
int qd_cvt_tree_build_host(const void* centroids /*[host] C x d*/, int dtype, int64_t C, int d,
                           void* tree /*[host] qd_cvt_tree_bytes*/);
size_t qd_cvt_scratch_bytes(int64_t B, int64_t C, int d);
int qd_cvt_index_of(const void* measures /*[dev] B x d*/, int dtype, int64_t B, int d,
                    const void* centroids /*[dev] C x d*/, int64_t C,
                    const void* prepared /*[dev] or NULL for method 1*/, int32_t index_base,
                    int method, int32_t* idx_out /*[dev] B*/, double* dist_out /*[dev] B or NULL*/,
                    uint32_t* flags /*[dev] 1*/, void* scratch /*[dev]*/, size_t scratch_bytes,
                    void* stream);
/* Diagnostics: when non-NULL, the next method-0 searches also dump the raw tensor-core scores
 * ([dev] B x 256*ceil(C/256) floats) and per-query margins ([dev] B floats). Pass NULLs to
 * switch off. Not for production use. */
void qd_cvt_debug_buffers(float* scores /*[dev]*/, float* eps /*[dev]*/);
/* Merge step of the sharded search: keeps idx where dist == global min, else INT32_MAX
 * (followed by an integer min-reduction across ranks). */
int qd_cvt_mask_losers(const double* local_dist /*[dev] B*/, const double* global_min /*[dev] B*/,
                       int32_t* idx_inout /*[dev] B*/, int64_t B, void* stream);

/* ------------------------------------------------------------------------------------
 * A4-A8  body of GridArchive.add / CVTArchive.add (the former
 * batch_entries_with_threshold)     ribs/archives/_grid_archive.py:606-714
 *                                   ribs/archives/_array_store.py:535-608
 * ---------------------------------------------------------------------------------- */
typedef struct qd_add_result {
  int32_t n_insertable; /* rows with status != 0 */
  int32_t n_new;        /* newly occupied cells */
  int32_t best_cell;    /* cell of the highest-objective winner (lowest cell on ties); -1 when the
                           batch did not beat the archive's best objective */
  uint32_t flags;       /* QD_FLAG_* of THIS call */
  double best_objective;
  double delta_sum[3]; /* exact 3-bin split of sum(new objective - old objective) */
  double batch_absmax; /* max |objective| of the batch */
  /* store state after the call (kept on the device between calls, so add() needs no host
   * round trip): what _stats_update (_grid_archive.py:325-360) consumes */
  int64_t n_occupied;     /* len(store) */
  double objective_sum;   /* running sum of the elites' objectives, rounded in the objective dtype */
  double obj_max;         /* valid when has_obj_max */
  int32_t has_obj_max;
  uint32_t last_failed_flags; /* `flags` of the most recent call that failed validation */
  uint64_t n_commits;     /* calls that inserted at least one row (ArrayStore.updates[ADD]) */
  uint64_t best_serial;   /* bumps whenever the best-elite snapshot was replaced */
  uint64_t n_failed;      /* calls dropped because of a non-finite objective / measure: a host that
                             did not wait for every call raises ValueError when this advances */
  uint64_t n_winners;     /* rows of THIS call written into the store (= cells that received an elite) */
} qd_add_result;

typedef struct qd_store {
  int64_t capacity;
  int obj_dtype;              /* dtype of objective / threshold */
  uint8_t* occupied;          /* [dev] capacity */
  int32_t* occupied_list;     /* [dev] capacity, insertion order */
  void* objective;            /* [dev] capacity */
  void* threshold;            /* [dev] capacity; NaN marks an unoccupied cell (qd_store_clear) */
  int n_fields;               /* row fields copied verbatim: solution, measures, extras */
  void* field_dst[QD_MAX_FIELDS];        /* [dev] capacity x row_bytes */
  int64_t field_row_bytes[QD_MAX_FIELDS];
  void* scratch;              /* [dev] qd_insert_scratch_bytes(capacity), zeroed once by
                                 qd_insert_scratch_init; every call leaves it clean */
  void* best_snapshot;        /* [dev] qd_best_snapshot_bytes(store) or NULL: copy of the best
                                 elite's row taken the moment it became the best
                                 (_grid_archive.py:331-341). Layout: double objective, double
                                 threshold, int32 cell, pad to 32 bytes, then every row field at
                                 16-byte aligned offsets in field order */
} qd_store;

size_t qd_insert_scratch_bytes(int64_t capacity);
int qd_insert_scratch_init(void* scratch /*[dev]*/, int64_t capacity, void* stream);
/* Every insertion on this store also writes its qd_add_result to `host_mirror` (page-locked host memory, or NULL
 * to stop): the host reads it after its stream synchronize without a device->host copy call. The reference
 * returns add_info synchronously (_grid_archive.py:606-714); this is the cheapest way to learn what it needs. */
int qd_insert_set_result_mirror(void* scratch /*[dev]*/, int64_t capacity, qd_add_result* host_mirror, void* stream);
/* Device address of the validation flag word inside `scratch`. Pass it as `flags` to the
 * index kernels of the same add() so that a non-finite measure aborts the commit on the
 * device (the store stays untouched, the host raises ValueError after reading
 * qd_add_result.flags). qd_insert clears the word at the end of every call. */
uint32_t* qd_insert_flags_ptr(void* scratch /*[dev]*/, int64_t capacity);
/* *out = OR of in[0..n): merges the validation flags gathered from all ranks of a data-parallel add. */
int qd_flags_or(const uint32_t* in /*[dev] n*/, int n, uint32_t* out /*[dev] 1*/, void* stream);

/* Diagnostics: %globaltimer (ns) at the phase boundaries of the last insertion on this store
 * (absmax, classify, mark [list mode], select [list mode], update, best/count + copy, its barrier, publish start,
 * publish end; slot 9: CTA 0 through its cells in the update phase).
 * Blocks until the stream is idle. */
int qd_insert_phase_times(const void* scratch /*[dev]*/, int64_t capacity, uint64_t* out_ns /*[host] 10*/,
                          void* stream);
size_t qd_best_snapshot_bytes(const qd_store* store);
/* ArrayStore.clear (_array_store.py:610-616) + _stats_reset: occupancy 0, thresholds NaN, element
 * count / objective sum / best objective reset. Also the initialiser of a fresh store. */
int qd_store_clear(const qd_store* store, void* stream);
/* Restores the device-side store state after the fields were loaded from a checkpoint
 * (re-marks the thresholds of unoccupied cells). */
int qd_store_set_state(const qd_store* store, int64_t n_occupied, double objective_sum, int has_obj_max,
                       double obj_max, double objective_absmax, void* stream);
/* Writes the store-state half of qd_add_result (per-call fields zero) to out [dev]. */
int qd_store_read_state(const qd_store* store, qd_add_result* out /*[dev]*/, void* stream);

/* Exchange step of the data-parallel add(): rank `rank` writes its slice of every batch
 * field (src[f], slice_bytes[f] bytes) into the global field array of EVERY peer, at
 * peer_bases[p] + dst_off[f] + rank * slice_bytes[f] -- P2P stores over NVLink into symmetric memory
 * (peer_bases[rank] is the local buffer). The caller follows it with a cross-rank barrier before any
 * rank reads its global arrays. */
int qd_dp_push(int n_fields, const void* const* src /*[host] dev ptrs*/, const int64_t* slice_bytes /*[host]*/,
               const int64_t* dst_off /*[host]*/, void* const* peer_bases /*[host] n_peers dev ptrs*/, int n_peers,
               int rank, void* stream);

/* One batched insertion = ONE persistent cooperative kernel. `idx` comes from qd_grid_index_of /
 * qd_cvt_index_of. New cells are appended to occupied_list in ascending cell order after the
 * store's current element count (device state).
 * learning_rate / threshold_min as validated by ribs/archives/_utils.py:76-93
 * (threshold_min == -inf selects the CMA-ME / MAP-Elites rule).
 * part: QD_INSERT_WHOLE runs everything in one launch. A caller whose row fields (field_src:
 * solutions, measures, extras) are still in flight on another stream issues QD_INSERT_CLASSIFY
 * (needs idx / objective only; status_out / value_out and the thresholds are final afterwards),
 * makes `stream` wait for the rows, then issues QD_INSERT_COMMIT (row copy, occupied_list,
 * stats, result) with the same arguments.
 * result [dev]: one qd_add_result, valid once the stream has run the call. */
#define QD_INSERT_WHOLE 0
#define QD_INSERT_CLASSIFY 1
#define QD_INSERT_COMMIT 2
/* Pipelined commit: QD_INSERT_ROWS does everything except the row copy (status / value, thresholds, objectives,
 * occupancy, occupied_list, stats, result, best-elite snapshot taken from the batch) and leaves the winner list
 * in one of two buffers; QD_INSERT_COPY, issued afterwards with the same arguments on ANOTHER stream (ordered
 * after the ROWS launch by an event), copies the winners' rows into the store. It has no grid barrier and shares
 * the SMs with the ROWS launch of the next add(): the HBM-bound copy hides behind the scattered-access-bound
 * row passes. The caller alternates the buffers (OR in QD_INSERT_BUF1 for every other add), keeps the batch
 * alive until the copy has run, issues copies in order on one stream, and joins that stream before anything reads
 * the store's row fields or before a QD_INSERT_WHOLE / _COMMIT launch. */
#define QD_INSERT_ROWS 3
#define QD_INSERT_COPY 4
#define QD_INSERT_BUF1 0x100
int qd_insert(const qd_store* store, int64_t B, const int32_t* idx /*[dev] B*/,
              const void* objective /*[dev] B*/, const void* const* field_src /*[host] n_fields dev ptrs*/,
              double learning_rate, double threshold_min,
              int32_t* status_out /*[dev] B*/, void* value_out /*[dev] B*/,
              qd_add_result* result /*[dev]*/, int part, void* stream);
/* Data-parallel variant: the batch is the concatenation of every rank's rows_per_rank rows;
 * idx / objective are global arrays (exchanged beforehand, 12 B per row), but the row fields stay where they
 * were produced: peer_field_src [dev] is an (n_ranks x n_fields) table of device pointers -- rank r's slice of
 * field f, mapped into this GPU's address space (NVLink peer memory) -- and the copy phase of the insert kernel
 * gathers just the WINNERS' rows from their owners. peer_field_src == NULL: plain qd_insert (field_src used). */
int qd_insert_sharded_rows(const qd_store* store, int64_t B, const int32_t* idx, const void* objective,
                           const void* const* field_src, const void* const* peer_field_src /*[dev]*/,
                           int64_t rows_per_rank, double learning_rate, double threshold_min,
                           int32_t* status_out, void* value_out, qd_add_result* result, int part, void* stream);
/* Owner-computes insertion (scatter target ribs/archives/_array_store.py:535-608): `store`
 * holds only the cells [cell_lo, cell_hi) of an archive whose cells are sharded over the ranks by contiguous range.
 * Every rank passes the GLOBAL batch -- idx / objective of all B rows (exchanged beforehand, 12 B per row), the
 * row fields through the peer table as in qd_insert_sharded_rows -- and this rank classifies, resolves and commits
 * only the rows whose cell it owns: per-rank work is B / n_ranks rows and its own winners' row copies. `idx` is
 * rewritten in place to local cell numbers (-1: not this rank's). Status / value of row i go to the rank its batch
 * slice came from: out_tab [dev] = { status_0, value_0, status_1, value_1, ... }, rank r's arrays of rows_per_rank
 * entries in peer memory; a cross-rank barrier after this call completes them. Validation verdicts are global: every
 * rank checks every objective. */
int qd_insert_owned_cells(const qd_store* store, int64_t B, int32_t* idx /*[dev] B, rewritten*/,
                          const void* objective /*[dev] B*/, const void* const* peer_field_src /*[dev]*/,
                          int64_t rows_per_rank, int32_t cell_lo, int32_t cell_hi, void* const* out_tab /*[dev]*/,
                          double learning_rate, double threshold_min, qd_add_result* result, void* stream);
/* GridArchive.add with A2 fused into the first pass (measures are read once, idx_out is
 * written for the later passes and for the caller). Arguments as qd_grid_index_of + qd_insert. */
int qd_grid_insert(const qd_store* store, int64_t B, const void* measures /*[dev] B x d*/, int meas_dtype,
                   int d, const int32_t* dims /*[host]*/, const double* lower /*[host]*/,
                   const double* interval /*[host]*/, double epsilon, const void* objective,
                   const void* const* field_src, double learning_rate, double threshold_min,
                   int32_t* idx_out /*[dev] B*/, int32_t* status_out, void* value_out,
                   qd_add_result* result, int part, void* stream);

/* A6  ArrayStore.retrieve      ribs/archives/_array_store.py:330-485
 * Gathers rows `idx` of every listed field plus the occupied flag. */
int qd_store_gather(int64_t M, const int32_t* idx /*[dev] M*/, int n_fields,
                    const void* const* field_src /*[host] dev ptrs, capacity rows*/,
                    void* const* field_out /*[host] dev ptrs, M rows*/,
                    const int64_t* field_row_bytes /*[host]*/,
                    const uint8_t* occupied /*[dev] or NULL*/, uint8_t* occupied_out /*[dev] or NULL*/,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * A12-A14  evolution strategies   ribs/emitters/opt/_cma_es.py, _sep_cma_es.py,
 *                                 _lm_ma_es.py
 * All strategy state is fp64 in device memory. `status` is a device block of 8 doubles:
 *   [0] sigma  [1] ||ps||^2 of the last tell  [2] max / [3] min of the (diagonal) covariance
 * so that ask -> tell -> ask never synchronises with the host. `z` holds unit normals
 * (rows x n), filled by qd_fill_normal (Philox4x32-10 + Box-Muller) or injected by the
 * caller for parity runs. `row_ids` (optional, [dev] rows int32) redirects sample k to
 * X[row_ids[k]] -- the bounds-resampling loop of the reference (_cma_es.py:208-225)
 * re-asks only the rows that fell outside the bounds.
 * ---------------------------------------------------------------------------------- */
int qd_fill_normal(double* out /*[dev] count*/, int64_t count, uint64_t seed, uint64_t offset, void* stream);
/* Same contract with a cheaper generator for the bandwidth-bound strategies (sep-CMA-ES, LM-MA-ES at
 * solution_dim ~ 10^4): Philox4x32-10 bits through a 256-layer ziggurat (the method of NumPy's
 * Generator.standard_normal); out[4q..4q+3] are drawn from counter offset + q. */
int qd_fill_normal_zig(double* out /*[dev] count*/, int64_t count, uint64_t seed, uint64_t offset, void* stream);
/* oob[k] = any coordinate of X[row_ids[k]] outside [lower, upper]   (_cma_es.py:190-194) */
int qd_es_out_of_bounds(const double* X, int64_t rows, int64_t n, const int32_t* row_ids,
                        const double* lower /*[dev] n*/, const double* upper /*[dev] n*/,
                        int32_t* oob /*[dev] rows*/, void* stream);
/* Weighted parent reduction shared by the three tells:
 *   out_mean[j] = sum_k w[k] X[rank[k]][j]                       (_cma_es.py:297)
 *   out_sq[j]   = sum_k w[k] (X[rank[k]][j] - old_mean[j])^2     (_sep_cma_es.py:303)
 * out_mean / out_sq / old_mean may be NULL. */
size_t qd_wsum_partial_bytes(int64_t mu, int64_t n);
int qd_weighted_rows(const double* X /*[dev] lam x n*/, int64_t n, const int64_t* rank /*[dev] mu*/,
                     const double* w /*[dev] mu*/, int64_t mu, const double* old_mean,
                     double* out_mean, double* out_sq, void* partial /*[dev]*/, void* stream);
/* scalar strategy parameters, computed on the host exactly as _calc_strat_params does */
typedef struct qd_cma_scalars {
  double cc, cs, c1, cmu, mueff, damps, wsum;
  double hsig_exp; /* 2 * current_eval / batch_size, after the increment (_cma_es.py:305) */
} qd_cma_scalars;
size_t qd_es_workspace_doubles(int64_t n);
/* sep-CMA-ES   (_sep_cma_es.py:140-191 ask, :257-311 tell) */
int qd_sep_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
               const double* C /*[dev] n diagonal*/, const double* status, double* X, void* stream);
/* ask with the generator fused in (no z round trip through HBM): sample k, columns 4q..4q+3 draw from
 * counter offset + k * ceil(n / 4) + q of qd_fill_normal_zig's stream family; z_out (optional, rows x n)
 * receives the unit normals. */
int qd_sep_ask_rng(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean, const double* C,
                   const double* status, uint64_t seed, uint64_t offset, double* X, double* z_out /*or NULL*/,
                   void* stream);
/* OpenAI-ES   (ribs/emitters/opt/_openai_es.py:127-208, AdamOpt ribs/emitters/opt/_adam_opt.py)
 * mirrored ask: X[k] = mean + scale * (sigma z_k), X[k + half_rows] = mean + scale * (sigma * -z_k), z_out = z
 * (half_rows x n), sigma = status[0]; same counter layout as qd_sep_ask_rng.
 * tell: theta <- Adam step on g = -(grad_sum / denom) + l2 * theta, a = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
 * computed by the caller; status[4] receives |theta - theta_prev| / |theta|. workspace: qd_es_workspace_doubles(n). */
int qd_mirror_ask_rng(int64_t half_rows, int64_t n, const double* mean, const double* scale /*[dev] n, sqrt'ed*/,
                      const double* status, uint64_t seed, uint64_t offset, double* X /*[dev] 2 half_rows x n*/,
                      double* z_out /*[dev] half_rows x n or NULL*/, void* stream);
int qd_openai_tell(int64_t n, const double* grad_sum /*[dev] n*/, double denom, double* theta, double* m, double* v,
                   double a, double beta1, double beta2, double eps, double l2, double* status, double* workspace,
                   void* stream);
int qd_sep_tell(int64_t n, const double* new_mean, const double* rank_mu /*[dev] n*/, double* mean, double* pc,
                double* ps, double* C, double* status, const qd_cma_scalars* sc /*[host]*/,
                double* workspace /*[dev] qd_es_workspace_doubles(n)*/, void* stream);
/* CMA-ES       (_cma_es.py:51-82 eigen refresh, :181-237 ask, :274-328 tell) */
int qd_sym_max(int64_t n, double* A /*[dev] n x n*/, void* stream);          /* A = max(A, A^T) */
/* Batched symmetric eigendecomposition, n <= 150 (parallel cyclic Jacobi, one CTA per matrix):
 * eig ascending, eigenvectors in the COLUMNS of V (the LAPACK convention np.linalg.eigh follows,
 * _cma_es.py:70). sweeps (optional, [dev] batch) receives the number of sweeps used. */
int qd_sym_eigh_batched(int64_t n, int64_t batch, const double* A /*[dev] batch x n x n*/,
                        double* eig /*[dev] batch x n*/, double* V /*[dev] batch x n x n*/,
                        int* sweeps /*[dev] batch or NULL*/, void* stream);
/* ------------------------------------------------------------------------------------
 * Variation operators of the MAP-Elites emitters (next rows of the scope table)
 *   GaussianEmitter.ask   ribs/emitters/_gaussian_emitter.py:139-166
 *   IsoLineEmitter.ask    ribs/emitters/_iso_line_emitter.py:156-201
 * X[k] = clip(P[a_k] + sigma * z[k] (+ line_sigma * line_z[k] * (P[b_k] - P[a_k])), lower, upper), P = the
 * archive's solution field; parent lists NULL: every parent is x0 (empty archive). iso_sigma holds one value
 * or n values (sigma_is_vector) in the solution dtype. line_z == NULL selects the Gaussian operator.
 * ---------------------------------------------------------------------------------- */
int qd_emit_variation(const void* store_solution /*[dev] capacity x n*/, int dtype, int64_t n, int64_t B,
                      const int32_t* parent_a /*[dev] B or NULL*/, const int32_t* parent_b /*[dev] B or NULL*/,
                      const void* x0 /*[dev] n or NULL*/, const double* z /*[dev] B x n unit normals*/,
                      const double* line_z /*[dev] B or NULL*/, const void* iso_sigma /*[dev] 1 or n*/,
                      int sigma_is_vector, double line_sigma, const void* lower /*[dev] n*/,
                      const void* upper /*[dev] n*/, void* X_out /*[dev] B x n*/, void* stream);

/* ------------------------------------------------------------------------------------
 * k_means_centroids   ribs/archives/_cvt_archive.py:36-134 (sklearn.cluster.k_means, Lloyd)
 * One update step: samples X (n x d) with their nearest centroid idx (from qd_cvt_index_of) -> new centroids
 * (k x d, dtype of X), the per-centroid sums (k x d doubles; the host relocates empty clusters with them),
 * member counts and the squared shift of every centroid. Order-independent exact accumulation: reproducible.
 * bins: qd_kmeans_workspace_doubles(k, d) doubles; absmax: a bound on |X|.
 * ---------------------------------------------------------------------------------- */
size_t qd_kmeans_workspace_doubles(int64_t k, int64_t d);
int qd_kmeans_update(const void* X /*[dev] n x d*/, int dtype, int64_t n, int64_t d, const int32_t* idx /*[dev] n*/,
                     int64_t k, double absmax, double* bins /*[dev]*/, uint32_t* count /*[dev] k*/,
                     const void* old_centroids /*[dev] k x d*/, void* new_centroids /*[dev] k x d*/,
                     double* sums /*[dev] k x d*/, double* shift /*[dev] k*/, void* stream);

/* CMA-ES pool: E strategies of identical (n, lambda) whose state is stacked in device memory; every
 * kernel is batched over the emitters (the many-emitter regime: ribs/schedulers/_scheduler.py:180-213,
 * 361-412 loops ask/tell over the emitters in Python). Arithmetic identical to the single-strategy entry
 * points below, so a pooled strategy and a stand-alone one fed the same normals agree bit for bit.
 * The covariance is updated in place. `ids` selects the emitters an operation applies to. */
typedef struct qd_cma_pool {
  int64_t E, n, lam;
  double* mean;     /* [dev] E x n */
  double* pc;       /* [dev] E x n */
  double* ps;       /* [dev] E x n */
  double* C;        /* [dev] E x n x n */
  double* B;        /* [dev] E x n x n eigenvectors (columns) */
  double* eig;      /* [dev] E x n */
  double* T;        /* [dev] E x n x n  B sqrt(eig) */
  double* invsqrt;  /* [dev] E x n x n */
  double* status;   /* [dev] E x 8 */
  double* new_mean; /* [dev] E x n scratch */
  double* ws;       /* [dev] E x ws_stride scratch, ws_stride >= n + 8 */
  double* X;        /* [dev] E x lam x n samples of the last ask */
  double* z;        /* [dev] E x lam x n unit normals of the last ask */
  int64_t ws_stride;
} qd_cma_pool;
/* lazy eigensystem refresh (_cma_es.py:51-82) of the listed emitters, n <= 150 */
int qd_cma_pool_refresh(const qd_cma_pool* pool /*[host]*/, const int32_t* ids /*[dev] n_ids*/, int n_ids,
                        void* stream);
/* ask of every emitter; emitter e draws from Philox stream (seeds[e], offsets[e]) -- the stream a
 * stand-alone strategy with that seed would use. seeds == NULL: pool->z was filled by the caller. */
int qd_cma_pool_ask(const qd_cma_pool* pool /*[host]*/, const uint64_t* seeds /*[dev] E*/,
                    const uint64_t* offsets /*[dev] E*/, void* stream);
/* tell of the listed emitters: rank / w are E x lam (first mu[e] entries of row e used), sc E entries */
int qd_cma_pool_tell(const qd_cma_pool* pool /*[host]*/, const int32_t* ids /*[dev] n_ids*/, int n_ids,
                     const int64_t* rank /*[dev]*/, const double* w /*[dev]*/, const int64_t* mu /*[dev] E*/,
                     const qd_cma_scalars* sc /*[dev] E*/, void* stream);
int qd_cma_transform(int64_t n, const double* Bmat, const double* eigvals, double* T, void* stream); /* T = B*sqrt(eig) */
int qd_cma_invsqrt(int64_t n, const double* Bmat, const double* eigvals, double* invsqrt, void* stream);
int qd_cma_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
               const double* T /*[dev] n x n*/, const double* status, double* X, void* stream);
/* Cnext receives the updated covariance (C is read-only); mean/pc/ps/status update in place. */
int qd_cma_tell(int64_t n, const double* X, const int64_t* rank, const double* w, int64_t mu,
                const double* new_mean, const double* invsqrt, double* mean, double* pc, double* ps,
                const double* Cmat, double* Cnext, double* status, const qd_cma_scalars* sc /*[host]*/,
                double* workspace, void* stream);
/* LM-MA-ES     (_lm_ma_es.py:126-194 ask, :209-239 tell) */
/* `workspace`: qd_lmma_ask_workspace_doubles(rows, n, itrs) doubles of device scratch. The `itrs` rank-one
 * transforms of :136-144 are evaluated as Z M^T, M M^T (tall-skinny products), a forward substitution per
 * sample and one pass X = mean + sigma Q (Z + T M): same arithmetic up to rounding order. */
size_t qd_lmma_ask_workspace_doubles(int64_t rows, int64_t n, int64_t itrs);
/* A large ask runs as row blocks of at most this many bytes of z, alternating between two internal streams
 * (default 160 MB; <= 0 restores it). Returns the previous value; call before sizing the workspace. */
int64_t qd_lmma_set_block_bytes(int64_t bytes);
int qd_lmma_ask(const double* z, int64_t rows, int64_t n, const int32_t* row_ids, const double* mean,
                const double* status, const double* M /*[dev] n_vectors x n*/, const double* cd /*[dev]*/,
                int64_t itrs, double* X, double* workspace, void* stream);
/* the same with the unit normals generated on the way: z_out (rows x n, needed by tell) receives exactly what
 * qd_fill_normal_zig(z_out, rows * n, seed, offset) produces */
int qd_lmma_ask_rng(int64_t rows, int64_t n, const int32_t* row_ids, const double* mean, const double* status,
                    const double* M, const double* cd, int64_t itrs, uint64_t seed, uint64_t offset,
                    double* z_out, double* X, double* workspace, void* stream);
int qd_lmma_tell(int64_t n, int64_t n_vectors, const double* new_mean, const double* z_mean, double* mean,
                 double* ps, double* M, const double* cc /*[dev] n_vectors*/, double csigma, double mueff,
                 double* status, double* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QD_B200_H_ */
