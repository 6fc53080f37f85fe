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

#define QD_ABI_VERSION 1

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
 *   method: 0 = tensor-core filter (tcgen05) + fp64 rescoring, 1 = plain fp64 scan.
 * ---------------------------------------------------------------------------------- */
size_t qd_cvt_prepared_bytes(int64_t C, int d);
int qd_cvt_prepare(const void* centroids /*[dev] C x d*/, int dtype, int64_t C, int d,
                   void* prepared /*[dev]*/, void* stream);
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
  int32_t best_cell;    /* cell of the highest-objective winner (lowest cell on ties) */
  uint32_t flags;       /* QD_FLAG_* */
  double best_objective;
  double delta_sum[3]; /* exact 3-bin split of sum(new objective - old objective) */
  double batch_absmax; /* max |objective| of the batch */
} qd_add_result;

typedef struct qd_store {
  int64_t capacity;
  int obj_dtype;              /* dtype of objective / threshold */
  uint8_t* occupied;          /* [dev] capacity */
  int32_t* occupied_list;     /* [dev] capacity, insertion order */
  void* objective;            /* [dev] capacity */
  void* threshold;            /* [dev] capacity */
  int n_fields;               /* row fields copied verbatim: solution, measures, extras */
  void* field_dst[QD_MAX_FIELDS];        /* [dev] capacity x row_bytes */
  int64_t field_row_bytes[QD_MAX_FIELDS];
  void* scratch;              /* [dev] qd_insert_scratch_bytes(capacity), zeroed once by
                                 qd_insert_scratch_init; every call leaves it clean */
} qd_store;

size_t qd_insert_scratch_bytes(int64_t capacity);
int qd_insert_scratch_init(void* scratch /*[dev]*/, int64_t capacity, void* stream);
/* Device address of the validation flag word inside `scratch`. Pass it as `flags` to the
 * index kernels of the same add() so that a non-finite measure aborts the commit on the
 * device (the store stays untouched, the host raises ValueError after reading
 * qd_add_result.flags). qd_insert clears the word at the end of every call. */
uint32_t* qd_insert_flags_ptr(void* scratch /*[dev]*/, int64_t capacity);

/* One batched insertion. `idx` comes from qd_grid_index_of / qd_cvt_index_of.
 * n_occupied_before is the store's element count before the call; new cells are appended
 * to occupied_list[n_occupied_before ...] in ascending cell order.
 * learning_rate / threshold_min as validated by ribs/archives/_utils.py:76-93
 * (threshold_min == -inf selects the CMA-ME / MAP-Elites rule).
 * result [dev]: one qd_add_result, read by the host after the stream is synchronised. */
int qd_insert(const qd_store* store, int64_t B, const int32_t* idx /*[dev] B*/,
              const void* objective /*[dev] B*/, const void* const* field_src /*[host] n_fields dev ptrs*/,
              double learning_rate, double threshold_min, int64_t n_occupied_before,
              int32_t* status_out /*[dev] B*/, void* value_out /*[dev] B*/,
              qd_add_result* result /*[dev]*/, void* stream);

/* A6  ArrayStore.retrieve      ribs/archives/_array_store.py:330-485
 * Gathers rows `idx` of every listed field plus the occupied flag. */
int qd_store_gather(int64_t M, const int32_t* idx /*[dev] M*/, int n_fields,
                    const void* const* field_src /*[host] dev ptrs, capacity rows*/,
                    void* const* field_out /*[host] dev ptrs, M rows*/,
                    const int64_t* field_row_bytes /*[host]*/,
                    const uint8_t* occupied /*[dev] or NULL*/, uint8_t* occupied_out /*[dev] or NULL*/,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QD_B200_H_ */
