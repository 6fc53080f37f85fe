"""DeviceArchive: everything GridArchive and CVTArchive share.

Implements the `ArchiveBase` API of the reference (ribs/archives/_archive_base.py:17-397)
on top of DeviceStore and the qd_insert kernel sequence. Subclasses provide
`_index_of_device(measures, flags_ptr)`.

Semantics of `add` follow ribs/archives/_grid_archive.py:518-714 (identical to
_cvt_archive.py:713-909); see SURVEY.md appendix A for the restatement.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import sys

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import (
    as_dim_tuple,
    check_batch_shape,
    check_finite,
    check_shape,
    is_device_tensor,
    torch_dtype,
    validate_batch,
    validate_single,
)
from pyribs_b200.archives import _shard_util
from pyribs_b200.archives._archive_stats import ArchiveStats
from pyribs_b200.archives._device_store import DeviceStore, current_stream_ptr

_RESERVED = {"solution", "objective", "measures", "threshold", "index"}
_NULL_CTX = contextlib.nullcontext()


def parse_all_dtypes(dtype, solution_dtype, objective_dtype, measures_dtype):
    """ribs/archives/_utils.py:28-73."""
    if isinstance(dtype, dict):
        raise ValueError(
            "Passing a dict as `dtype` is deprecated in pyribs 0.9.0. Please use "
            "`solution_dtype`, `objective_dtype`, and/or `measures_dtype` instead."
        )
    individual = solution_dtype is not None or objective_dtype is not None or measures_dtype is not None
    if dtype is not None and individual:
        raise ValueError(
            "dtype cannot be used at the same time as solution_dtype, objective_dtype, or measures_dtype. "
            "Either pass dtype on its own, or pass solution_dtype, objective_dtype, and/or measures_dtype."
        )
    default = np.float64
    if individual:
        return tuple(np.dtype(default if d is None else d) for d in (solution_dtype, objective_dtype, measures_dtype))
    d = np.dtype(default if dtype is None else dtype)
    return d, d, d


def validate_cma_mae_settings(learning_rate, threshold_min, dtype):
    """ribs/archives/_utils.py:76-93."""
    if threshold_min != -np.inf and learning_rate is None:
        raise ValueError(
            "threshold_min was set without setting learning_rate. Please note that threshold_min is only "
            "used in CMA-MAE; it is not intended to be used for only filtering archive solutions. To run "
            "CMA-MAE, please also set learning_rate."
        )
    if learning_rate is None:
        learning_rate = 1.0
    if threshold_min == -np.inf and learning_rate != 1.0:
        raise ValueError("threshold_min can only be -np.inf if learning_rate is 1.0")
    return np.asarray(learning_rate, dtype=dtype), np.asarray(threshold_min, dtype=dtype)


def fill_sentinel_values(occupied, data):
    """ribs/archives/_utils.py:96-112."""
    unoccupied = ~occupied
    for name, arr in data.items():
        if name == "index":
            fill = -1
        elif np.issubdtype(arr.dtype, np.integer):
            fill = 0
        else:
            fill = np.nan
        arr[unoccupied] = fill


# reference count of an array held only by a list slot, as `sys.getrefcount(entry[k])` sees it (calibrated
# once, so the test below does not depend on how this CPython version counts call temporaries)
_PROBE = [np.empty(1)]
_IDLE_REFS = sys.getrefcount(_PROBE[0])


# host batches up to these sizes are staged in one page-locked block that the kernels read in place (_add_small);
# below csrc/insert.cu's kSingleMaxRows / kSingleMaxWork one CTA then inserts the whole batch
_SMALL_ROWS = 16384
_SMALL_BYTES = 4 << 20
_CHUNK_MIN_ROWS = 1 << 18  # host batches from here on overlap the transfer of their measures with the search


def _pool_entry_free(entry):
    """True when nothing but the pool refers to the master arrays entry[2], entry[3]. Every view handed to
    a caller (and every view of such a view) has the master as its `.base`, so a larger count means a result
    is still alive somewhere."""
    return sys.getrefcount(entry[2]) <= _IDLE_REFS and sys.getrefcount(entry[3]) <= _IDLE_REFS


class DeviceArchive:
    def __init__(self, *, solution_dim, measure_dim, cells, learning_rate, threshold_min, qd_score_offset, seed,
                 solution_dtype, objective_dtype, measures_dtype, dtype, extra_fields, device=None,
                 data_parallel=False, process_group=None, shard_store=False):
        self._rng = np.random.default_rng(seed)
        # data-parallel insertion: every rank passes its own slice of the global batch to add();
        # indices are computed locally, then (index, objective, row fields) are all-gathered over
        # NCCL and every rank commits the same global batch to its replica (bit-identical replicas:
        # the insert kernels are deterministic).
        self._dp_world, self._dp_rank, self._dp_group = 1, 0, process_group
        if data_parallel:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("data_parallel=True needs an initialised torch.distributed process group")
            self._dp_world = dist.get_world_size(process_group)
            self._dp_rank = dist.get_rank(process_group)
        # owner-computes store: the cells are sharded over the ranks by contiguous range, every
        # rank classifies / commits only the rows of its own cells and holds only their elites (see _add_owned,
        # csrc/insert.cu qd_insert_owned_cells). Reads of the whole archive are collectives.
        if shard_store and not data_parallel:
            raise ValueError("shard_store=True needs data_parallel=True (the cells are sharded over the ranks)")
        self._shard_store = bool(shard_store) and self._dp_world > 1
        self._cells_global = int(cells)
        self._cell_lo, self._cell_hi = 0, int(cells)
        if self._shard_store:
            self._cell_lo, self._cell_hi = _shard_util.shard_cell_range(cells, self._dp_world, self._dp_rank)
            if self._cell_hi <= self._cell_lo:
                raise ValueError("shard_store=True needs at least one cell per rank")
            cells = self._cell_hi - self._cell_lo
        self._global_cache = None  # (adds seen, gathered per-rank state) of the last collective stats read
        self._adds_enqueued = 0
        self._solution_dim = as_dim_tuple(solution_dim) if not np.isscalar(solution_dim) else int(solution_dim)
        self._objective_dim = ()
        self._measure_dim = int(measure_dim)
        extra_fields = extra_fields or {}
        if _RESERVED & extra_fields.keys():
            raise ValueError(f"The following names are not allowed in extra_fields: {_RESERVED}")
        sdt, odt, mdt = parse_all_dtypes(dtype, solution_dtype, objective_dtype, measures_dtype)
        for what, dt in (("objective", odt), ("measures", mdt)):
            if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
                raise TypeError(f"{what} dtype must be float32 or float64 on the GPU path, got {dt}")
        for name, (_, dt) in extra_fields.items():
            torch_dtype(dt)  # raises TypeError for object dtype etc.
        self._store = DeviceStore(
            {
                "solution": (self._solution_dim, sdt),
                "objective": ((), odt),
                "measures": (self._measure_dim, mdt),
                "threshold": ((), odt),  # same dtype as objective: they share arithmetic
                **extra_fields,
            },
            capacity=cells,
            device=device,
        )
        self._device = self._store.device
        self._learning_rate, self._threshold_min = validate_cma_mae_settings(learning_rate, threshold_min, odt)
        self._qd_score_offset = np.asarray(qd_score_offset, dtype=odt)
        self._lr_f = float(self._learning_rate)
        self._tmin_f = float(self._threshold_min)
        self._best_elite = None
        self._best_serial = 0
        self._st = (0, 0.0, None, 0)
        self._stats_dirty = False
        self._objective_sum = None
        self._stats = None
        self._stats_reset()
        self._pinned = {}
        self._outbuf = {}
        self._copy_stream = None
        self._copy_event = None
        self._fused_index_insert = False  # GridArchive: index_of fused into the first insert pass
        self._sync_device_adds = False    # True: add() of CUDA tensors also waits and validates like NumPy input
        self._inflight = None
        self._needs_flag_check = False
        self._dp_symm = None  # {batch signature: symmetric-memory exchange state}; False = unavailable
        self._zero_copy_rows = True   # read winners' rows straight from page-locked host arrays when that saves PCIe traffic
        self._dev_fast = True         # CUDA-tensor batches in the store's dtypes take the lean path (_add_device)
        self._dev_idx = None          # its reusable cell-index buffer
        self._small_adds = True       # host batches of <= _SMALL_ROWS rows take the two-launch path (_add_small)
        self._small = {}              # its per-batch-size plans
        self._last_zero_copy = False
        self._zero_copy_bytes = 0
        self._zc = None
        self._shared_idx = None  # (device idx tensor, B) handed over by an archive of the same geometry (result_archive)
        self._last_idx = None    # (device idx tensor, B) of the last completed, waited-for add()
        self._extra_names = {n for n in self._store.field_list if n not in ("solution", "objective", "measures", "threshold")}
        self._pin_cache = {}
        self._row_bytes_total = sum(self._store._row_bytes(n) for n in self._store.row_fields)
        self._force_scan = False
        self._overflow_fallbacks = 0
        # pipelined commit (see `pipeline_commit`)
        self._pipeline_commit = False
        self._pipe = None          # (side stream, rows-done event, [copy-done event of list buffer 0, 1])
        self._pipe_calls = 0
        self._pipe_pending = None  # copy-done event nobody has waited for yet
        self._attach_store(self._store)

    def _attach_store(self, store):
        """Adopts `store` (construction, retessellate, unpickling): the store's hooks point back at this archive
        and every piece of host state that described the previous store starts over."""
        self._store = store
        self._tdt = {n: store._fields[n].dtype for n in ("solution", "objective", "measures")}
        self._np_dt = tuple(store._np_dtypes[n] for n in ("solution", "objective", "measures"))
        store._copy_join = self._join_copy
        store._owner_sync = self._sync
        self._st = (0, 0.0, None, 0)
        self._best_serial = 0
        self._best_elite = None
        self._last_idx = None
        self._shared_idx = None
        self._pipe_pending = None
        self._inflight = None

    # state that cannot (streams, events, ctypes arrays, page-locked pools, process groups) or need not (device
    # search structures, scratch) travel through pickle; rebuilt by __setstate__ / on first use
    _TRANSIENT = {"_pinned": dict, "_outbuf": dict, "_copy_stream": None, "_copy_event": None, "_inflight": None,
                  "_dp_symm": None, "_zc": None, "_shared_idx": None, "_last_idx": None, "_pin_cache": dict, "_small": dict, "_dev_idx": None,
                  "_symm_last": None, "_global_cache": None, "_chunk_events": None,
                  "_pipe": None, "_pipe_pending": None, "_dp_group": None}

    def __getstate__(self):
        """Checkpoint (the reference's archives pickle as plain NumPy holders; SURVEY.md lists checkpoint / resume):
        device tensors travel as NumPy through DeviceStore.__getstate__, the running stats and the best elite as the
        host values `stats` / `best_elite` report."""
        self._sync()
        self._join_copy()
        if self._stats_dirty:
            self._build_stats()
        d = {k: v for k, v in self.__dict__.items() if k not in self._TRANSIENT and k not in self._transient_extra()}
        d["_best_elite"] = self.best_elite
        return d

    def _transient_extra(self):
        return ()

    def __setstate__(self, d):
        self.__dict__.update(d)
        for k, v in self._TRANSIENT.items():
            setattr(self, k, v() if callable(v) else v)
        self._device = self._store.device
        best, st = self._best_elite, self._st
        self._attach_store(self._store)
        self._rebuild_device_side()
        # re-seed the device-side running state with the archive's own values (the best objective of a CMA-MAE
        # archive can exceed every stored row's) and adopt what the device reports back
        store = self._store
        n = len(store)
        absmax = float(store._fields["objective"][store._occupied.bool()].abs().max().item()) if n else 0.0
        store._owner_sync = None
        store._push_state(st[1], st[2], absmax)
        store._owner_sync = self._sync
        self._apply_state(store._last)
        self._best_elite = best
        self._best_serial = self._st[3]

    def _rebuild_device_side(self):
        """Subclasses: ABI constants / search structures dropped by __getstate__."""

    # ---- properties ---------------------------------------------------------------------
    @property
    def solution_dim(self):
        return self._solution_dim

    @property
    def objective_dim(self):
        return self._objective_dim

    @property
    def measure_dim(self):
        return self._measure_dim

    @property
    def field_list(self):
        return self._store.field_list_with_index

    @property
    def dtypes(self):
        return self._store.dtypes_with_index

    @property
    def stats(self):
        if self._shard_store:  # collective: the shards' running sums in rank (= cell) order
            n, osum, omax, _ = _shard_util.merge_shard_stats(self._global_state(), self.dtypes["objective"])
            self._build_stats((n, osum, omax, 0))
            return self._stats
        self._sync()
        if self._stats_dirty:
            self._build_stats()
        return self._stats

    @property
    def empty(self):
        return len(self) == 0

    @property
    def best_elite(self):
        if self._shard_store:
            # collective: the shard holding the highest objective hands out its record (lowest rank = lowest cells on
            # equal maxima; a tie between maxima reached in different add() calls on different ranks is not ordered
            # by time as the reference would)
            import torch.distributed as dist

            owner = _shard_util.merge_shard_stats(self._global_state(), self.dtypes["objective"])[3]
            if owner < 0:
                return None
            box = [None]
            if self._dp_rank == owner:
                if self._st[3] != self._best_serial:
                    self._best_serial = self._st[3]
                    self._best_elite = self._store.read_snapshot()
                box[0] = dict(self._best_elite, index=np.int32(self._best_elite["index"] + self._cell_lo))
            group = self._dp_group
            dist.broadcast_object_list(box, src=owner if group is None else dist.get_global_rank(group, owner), group=group)
            return box[0]
        self._sync()
        if self._st[3] != self._best_serial:  # the device replaced its best-elite snapshot since the last read
            self._best_serial = self._st[3]
            self._best_elite = self._store.read_snapshot()
        return self._best_elite

    @property
    def cells(self):
        return self._cells_global if self._shard_store else self._store.capacity

    @property
    def learning_rate(self):
        return self._learning_rate

    @property
    def threshold_min(self):
        return self._threshold_min

    @property
    def qd_score_offset(self):
        return self._qd_score_offset

    @property
    def device(self):
        return self._device

    def __len__(self):
        if self._shard_store:
            return int(self._global_state()[:, 0].sum())
        self._sync()
        return len(self._store)

    def __iter__(self):
        if self._shard_store:  # collective snapshot, then a plain host iteration
            d = self.data()
            return iter([{k: v[i] for k, v in d.items()} for i in range(len(d["index"]))])
        self._sync()
        return iter(self._store)

    # ---- device-resident state -> host mirror ------------------------------------------------
    _ERR_OBJ = "All elements of objective must be finite (infinity and NaN values are not supported)."
    _ERR_MEAS = "All elements of measures must be finite (infinity and NaN values are not supported)."

    # ---- pipelined commit -------------------------------------------------------------------------------
    @property
    def pipeline_commit(self):
        """Opt-in software pipelining of back-to-back `add()` calls on CUDA tensors (single process). When set,
        `add()` runs every phase except the copy of the winners' rows on the caller's stream (QD_INSERT_ROWS:
        status / value, thresholds, objectives, occupancy, stats, best elite) and the row copy on a side stream
        (QD_INSERT_COPY), where it overlaps the row phases of the next `add()`: the copy is HBM-bound, the row
        passes are bound by scattered accesses, and together they fill the machine. Results are identical to the
        default mode. Contract: like any asynchronous copy, the batch tensors must not be overwritten until the
        copy has run -- until the next read of the archive (`data`, `retrieve*`, `sample_elites`, iteration,
        `clear`, pickling; they wait for it) or an explicit `archive.join()`; freeing them is safe (their memory
        is tied to the side stream)."""
        return self._pipeline_commit

    @pipeline_commit.setter
    def pipeline_commit(self, on):
        if not on:
            self._join_copy()
        self._pipeline_commit = bool(on)

    def join(self):
        """Makes the current stream wait for a row copy still running on the side stream."""
        self._join_copy()

    def _join_copy(self):
        ev = self._pipe_pending
        if ev is not None:
            self._pipe_pending = None
            torch.cuda.current_stream(self._device).wait_event(ev)

    def _pipe_state(self):
        if self._pipe is None:
            self._pipe = (torch.cuda.Stream(device=self._device), torch.cuda.Event(),
                          [torch.cuda.Event(), torch.cuda.Event()])
        return self._pipe

    def _raise_for_flags(self, flags):
        if flags & _lib.FLAG_NONFINITE_OBJECTIVE:
            raise ValueError(self._ERR_OBJ)
        if flags & _lib.FLAG_NONFINITE_MEASURES:
            raise ValueError(self._ERR_MEAS)

    def _sync(self):
        """Insertions given CUDA tensors are only enqueued (the element count, objective sum, best
        objective and best-elite snapshot live on the device, csrc/insert.cu); anything that needs
        them on the host lands here. A batch dropped by the device-side finite check surfaces as
        the reference's ValueError at this point."""
        store = self._store
        store.join_row_copy()
        if not store._pending:
            return
        flags = store._read_back()
        self._apply_state(store._last)
        self._raise_for_flags(flags)

    def _apply_state(self, res):
        """Adopts the device-side running state; the ArchiveStats record / best elite are built on access."""
        self._st = (int(res.n_occupied), float(res.objective_sum), float(res.obj_max) if res.has_obj_max else None,
                    int(res.best_serial))
        self._stats_dirty = True

    def _build_stats(self, st=None):
        """_stats_update (_grid_archive.py:325-360) from the device-side running state."""
        self._stats_dirty = False
        n, osum, omax, _ = self._st if st is None else st
        if n == 0 or omax is None:
            return
        odt = self.dtypes["objective"]
        self._objective_sum = np.asarray(osum, dtype=odt)
        qd = self._objective_sum - np.asarray(n, dtype=odt) * self._qd_score_offset
        self._stats = ArchiveStats(
            num_elites=n,
            coverage=np.asarray(n / self.cells, dtype=odt),
            qd_score=qd,
            norm_qd_score=np.asarray(qd / self.cells, dtype=odt),
            obj_max=np.dtype(odt).type(omax),
            obj_mean=np.asarray(self._objective_sum / n, dtype=odt),
        )

    # ---- stats (A8, _grid_archive.py:311-360) ---------------------------------------------
    def _stats_reset(self):
        odt = self.dtypes["objective"]
        self._best_elite = None
        self._best_serial = self._st[3]
        self._stats_dirty = False
        self._objective_sum = np.asarray(0.0, dtype=odt)
        self._stats = ArchiveStats(
            num_elites=0,
            coverage=np.asarray(0.0, dtype=odt),
            qd_score=np.asarray(0.0, dtype=odt),
            norm_qd_score=np.asarray(0.0, dtype=odt),
            obj_max=None,
            obj_mean=None,
        )

    # ---- host <-> device plumbing ----------------------------------------------------------
    def _to_device(self, arr, np_dtype, name="_"):
        """Host array -> device tensor through a reusable pinned staging buffer (one per
        field, so that in-flight async copies of one add() never share a buffer)."""
        if isinstance(arr, torch.Tensor):
            want = torch_dtype(np_dtype)
            if arr.device == self._device and arr.dtype == want and arr.is_contiguous():
                return arr
            return arr.to(device=self._device, dtype=want).contiguous()
        arr = np.ascontiguousarray(arr, dtype=np_dtype)
        if arr.size == 0:
            return torch.empty(arr.shape, dtype=torch_dtype(np_dtype), device=self._device)
        src = torch.from_numpy(arr)
        if arr.nbytes < (1 << 16):
            return src.to(self._device)
        if src.is_pinned():  # caller already staged the batch in page-locked memory
            return src.to(self._device, non_blocking=True)
        key = (name, np.dtype(np_dtype))
        buf = self._pinned.get(key)
        if buf is None or buf.numel() < arr.size:
            buf = torch.empty(max(arr.size, 1), dtype=torch_dtype(np_dtype)).pin_memory()
            self._pinned[key] = buf
        stage = buf[: arr.size].view(arr.shape)
        stage.copy_(src)
        return stage.to(self._device, non_blocking=True)

    def _index_of_device(self, measures: torch.Tensor, flags_ptr: int, out=None) -> torch.Tensor:
        raise NotImplementedError

    def _search_chunked(self, measures, B):
        """index_of over a large host array, overlapped with its own host->device transfer: piece k is copied on the
        side stream (through a page-locked staging piece when the array is pageable) and searched on the caller's
        stream as soon as it has landed. Page-locked arrays travel as a small first piece (the search can start
        after an eighth of the transfer) and the rest; pageable ones as four equal pieces, so that the host's
        staging memcpy of piece k+1 runs next to the search of piece k. Returns (device measures, device idx)."""
        mdt = self.dtypes["measures"]
        arr = np.ascontiguousarray(measures, dtype=mdt)
        src = torch.from_numpy(arr)
        pinned = src.is_pinned()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self._device)
            self._copy_event = torch.cuda.Event()
        meas = torch.empty(arr.shape, dtype=torch_dtype(mdt), device=self._device)
        idx = torch.empty(B, dtype=torch.int32, device=self._device)
        bounds = [0, B // 8, B] if pinned else [min(B, k * -(-B // 4)) for k in range(5)]
        pieces = len(bounds) - 1
        cur = torch.cuda.current_stream()
        events = self.__dict__.get("_chunk_events")
        if not events or len(events) < pieces:
            events = self._chunk_events = [torch.cuda.Event() for _ in range(pieces)]
        stage = None
        if not pinned:
            key = ("measures", np.dtype(mdt))
            stage = self._pinned.get(key)
            if stage is None or stage.numel() < arr.size:
                stage = self._pinned[key] = torch.empty(max(arr.size, 1), dtype=torch_dtype(mdt)).pin_memory()
            stage = stage[: arr.size].view(arr.shape)
        # `meas` may be a block the allocator just took back from work still queued on `cur`, and the staging buffer's
        # previous use was consumed in `cur`'s order: the side stream starts behind both
        self._copy_stream.wait_stream(cur)
        for k in range(pieces):
            lo, hi = bounds[k], bounds[k + 1]
            if lo >= hi:
                continue
            piece = src[lo:hi]
            if stage is not None:
                stage[lo:hi].copy_(piece)  # host memcpy of this piece while the previous pieces travel / are searched
                piece = stage[lo:hi]
            with torch.cuda.stream(self._copy_stream):
                meas[lo:hi].copy_(piece, non_blocking=True)
                events[k].record(self._copy_stream)
            cur.wait_event(events[k])
            self._index_of_device(meas[lo:hi], self._store._flags_ptr, out=idx[lo:hi])
        meas.record_stream(self._copy_stream)
        return meas, idx

    def same_geometry(self, other):
        """True when `other` maps every measure vector to the same cell index as this archive (then a batch
        indexed for one of them needs no second index_of: Scheduler's result_archive, _scheduler.py:260-264)."""
        return False

    def _host_measures_ok(self):
        """True when the index kernel of the next add() may read the measures straight from page-locked host
        memory (one coalesced pass over them): the fused grid index and the CVT bucket search."""
        return self._fused_index_insert

    def _all_gather_many(self, tensors):
        """All-gathers several equally-sharded tensors over NCCL."""
        import torch.distributed as dist

        W = self._dp_world
        outs = [torch.empty((W * t.shape[0], *t.shape[1:]), dtype=t.dtype, device=t.device) for t in tensors]
        ins = [t.contiguous() for t in tensors]
        for o, t in zip(outs, ins):  # (torch's coalescing manager measured slower than separate launches)
            dist.all_gather_into_tensor(o, t, group=self._dp_group)
        return outs

    # ---- exchange step of the data-parallel add() over NVLink peer memory --------------------------
    def _symm_exchange(self, idx, dev, B_local):
        """Exchange step of the data-parallel add() over NVLink peer memory (torch symmetric memory = P2P-mapped
        cudaMalloc). Every rank pushes its slice of (flags, idx, objective) -- 12 B per row -- straight into the
        global arrays of every peer (csrc/abi.cu qd_dp_push) and parks its row fields (solutions, measures,
        extras) in its own symmetric buffer; one cross-rank barrier. The row fields are NOT exchanged: at most
        `cells` rows can win, and the copy phase of the insert kernel gathers exactly the winners' rows from their
        owners' buffers (qd_insert_sharded_rows). Two buffer sets alternate so that a rank running ahead never
        overwrites memory a slower peer's insert kernel is still reading. Returns (global idx, global objective,
        device pointer of the peer table, global batch size), or None when symmetric memory cannot be set up
        (the NCCL all-gathers below are used instead)."""
        if self._dp_symm is False:
            return None
        import torch.distributed as dist

        store = self._store
        rows = store.row_fields
        key = B_local  # field shapes / dtypes are fixed per archive
        st = self._dp_symm.get(key) if self._dp_symm else None
        R = self._dp_world
        if st is None:
            try:
                import torch.distributed._symmetric_memory as symm

                group = self._dp_group if self._dp_group is not None else dist.group.WORLD
                osz = dev["objective"].element_size()
                small = [4, 4 * B_local, osz * B_local]           # flags, idx, objective: replicated R times
                local = [store._row_bytes(n) * B_local for n in rows]  # row fields: this rank's slice only
                offs, total = [], 0
                for nb in small:
                    offs.append(total)
                    total += (R * nb + 255) // 256 * 256
                loffs = []
                for nb in local:
                    loffs.append(total)
                    total += (nb + 255) // 256 * 256
                # owner-computes insertion: status / value of this rank's rows, written by the ranks that own their cells
                ooffs = []
                for nb in (4 * B_local, osz * B_local):
                    ooffs.append(total)
                    total += (nb + 255) // 256 * 256
                bufs, handles = [], []
                for _ in range(2):
                    t = symm.empty(total, dtype=torch.uint8, device=self._device)
                    handles.append(symm.rendezvous(t, group))
                    bufs.append(t)
                peers, own, tabs, views, outs = [], [], [], [], []
                for G, h in zip(bufs, handles):
                    ptrs = [int(p) for p in h.buffer_ptrs]
                    otab = np.array([[ptrs[r] + ooffs[0], ptrs[r] + ooffs[1]] for r in range(R)], dtype=np.int64)
                    outs.append((torch.from_numpy(otab).to(self._device),
                                 G[ooffs[0]:ooffs[0] + 4 * B_local].view(torch.int32),
                                 G[ooffs[1]:ooffs[1] + osz * B_local].view(dev["objective"].dtype)))
                    peers.append((C.c_void_p * R)(*ptrs))
                    own.append((C.c_void_p * 1)(ptrs[self._dp_rank]))
                    tab = np.array([[ptrs[r] + lo for lo in loffs] for r in range(R)], dtype=np.int64)
                    tabs.append(torch.from_numpy(tab).to(self._device))
                    views.append((G[offs[1]:offs[1] + R * small[1]].view(torch.int32),
                                  G[offs[2]:offs[2] + R * small[2]].view(dev["objective"].dtype),
                                  G[offs[0]:offs[0] + 4 * R].data_ptr()))
                st = dict(handles=handles, peers=peers, own=own, tabs=tabs, views=views, bufs=bufs, turn=0, outs=outs,
                          c_small=(C.c_int64 * 3)(*small), c_offs=(C.c_int64 * 3)(*offs),
                          c_local=(C.c_int64 * max(len(local), 1))(*local),
                          c_loffs=(C.c_int64 * max(len(loffs), 1))(*loffs))
                if self._dp_symm is None:
                    self._dp_symm = {}
                self._dp_symm[key] = st
            except Exception:  # noqa: BLE001  (no P2P / IPC on this box)
                self._dp_symm = False
                return None
        turn = st["turn"]
        st["turn"] = 1 - turn
        L = _lib.lib()
        sp = current_stream_ptr()
        c_small = (C.c_void_p * 3)(store._flags_ptr, idx.data_ptr(), dev["objective"].data_ptr())
        _lib.check(L.qd_dp_push(3, c_small, st["c_small"], st["c_offs"], st["peers"][turn], R, self._dp_rank, sp))
        if rows:
            c_rows = (C.c_void_p * len(rows))(*[dev[n].data_ptr() for n in rows])
            _lib.check(L.qd_dp_push(len(rows), c_rows, st["c_local"], st["c_loffs"], st["own"][turn], 1, 0, sp))
        st["handles"][turn].barrier()
        g_idx, g_obj, flags_ptr = st["views"][turn]
        _lib.check(L.qd_flags_or(flags_ptr, R, store._flags_ptr, sp))
        self._symm_last = (st["handles"][turn], st["outs"][turn])  # (_add_owned: the reverse route of status / value)
        return g_idx, g_obj, st["tabs"][turn].data_ptr(), R * B_local

    def _gather_global_batch(self, idx, dev, B_local, skip=()):
        """All-gathers the per-rank slices (equal sizes on every rank) into the global batch and
        ORs the validation flags across ranks. `idx` may be None (already global)."""
        store = self._store
        off = store._flags_ptr - store._scratch.data_ptr()
        flags = store._scratch[off:off + 4].view(torch.int32)
        names = [n for n in dev if n not in skip]
        tensors = [flags] + ([idx] if idx is not None else []) + [dev[n] for n in names]
        outs = self._all_gather_many(tensors)
        _lib.check(_lib.lib().qd_flags_or(outs[0].data_ptr(), int(outs[0].numel()), flags.data_ptr(),
                                          current_stream_ptr()))
        k = 1
        g_idx = idx
        if idx is not None:
            g_idx = outs[k]
            k += 1
        g_dev = dict(dev)
        for n in names:
            g_dev[n] = outs[k]
            k += 1
        return g_idx, g_dev, self._dp_world * B_local

    def _out_buffers(self, B, dev_in):
        """status / value outputs: (device status, device value, pinned status, pinned value, pooled).

        Device-mode calls get fresh device tensors (the caller keeps them). Host-mode calls reuse
        device tensors per batch size and hand out page-locked host buffers from a small pool: the
        NumPy arrays add() returns are VIEWS of such a buffer (no extra 1.2 MB memcpy per 100k rows),
        and a buffer goes back into rotation only when CPython's reference count shows that the
        caller dropped every array derived from it. If the caller hoards results the pool runs dry
        and add() falls back to copying out of a private staging buffer."""
        odt = torch_dtype(self.dtypes["objective"])
        if dev_in:
            return (torch.empty(B, dtype=torch.int32, device=self._device),
                    torch.empty(B, dtype=odt, device=self._device), None, None, None)
        buf = self._outbuf.get(B)
        if buf is None:
            if len(self._outbuf) > 8:
                self._outbuf.clear()
            buf = {"dev": (torch.empty(B, dtype=torch.int32, device=self._device),
                           torch.empty(B, dtype=odt, device=self._device)), "pool": [], "staging": None}
            self._outbuf[B] = buf
        for entry in buf["pool"]:
            if _pool_entry_free(entry):
                return (*buf["dev"], entry[0], entry[1], entry)
        if len(buf["pool"]) < 6:
            st, va = torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=odt).pin_memory()
            entry = [st, va, st.numpy(), va.numpy()]  # the arrays handed out are views of entry[2:4]
            buf["pool"].append(entry)
            return (*buf["dev"], st, va, entry)
        if buf["staging"] is None:
            buf["staging"] = (torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, dtype=odt).pin_memory())
        return (*buf["dev"], *buf["staging"], None)

    # ---- index_of ---------------------------------------------------------------------------
    def index_of(self, measures):
        mdt = self.dtypes["measures"]
        dev_in = is_device_tensor(measures)
        if not dev_in:
            measures = np.asarray(measures, dtype=mdt)
        check_batch_shape(measures, "measures", self.measure_dim, "measure_dim")
        m = self._to_device(measures, mdt)
        flags = torch.zeros(1, dtype=torch.int32, device=self._device)
        with torch.cuda.device(self._device):
            idx = self._index_of_device(m, flags.data_ptr())
        fl = int(flags.item())
        if fl & _lib.FLAG_NONFINITE_MEASURES:
            raise ValueError("All elements of measures must be finite (infinity and NaN values are not supported).")
        if fl & _lib.FLAG_CVT_OVERFLOW:  # candidate list overflowed: redo with the exact scan
            self._overflow_fallbacks += 1
            self._force_scan = True
            try:
                return self.index_of(measures)
            finally:
                self._force_scan = False
        return idx if dev_in else idx.cpu().numpy()

    def index_of_single(self, measures):
        measures = np.asarray(measures, dtype=self.dtypes["measures"])
        check_shape(measures, "measures", self.measure_dim, "measure_dim")
        check_finite(measures, "measures")
        return self.index_of(measures[None])[0]

    # ---- add (A5) -----------------------------------------------------------------------------
    def add(self, solution, objective, measures, **fields):
        dev_in = is_device_tensor(solution)
        # a hand-over from an archive of the same geometry is good for this call only, also when the call raises
        preset, self._shared_idx = self._shared_idx, None
        if dev_in and self._dev_fast:
            out = self._add_device(solution, objective, measures, fields, preset)
            if out is not None:
                return out
        if (not fields and type(solution) is np.ndarray and type(objective) is np.ndarray
                and type(measures) is np.ndarray and self._small_adds and self._plain_batch_ok(solution, objective, measures)):
            # a small, well-formed NumPy batch (the common case of the reference's own loop and benchmark): the checks
            # validate_batch would make have all passed, straight to the staged path
            B = solution.shape[0]
            self._last_idx = None
            if preset is not None and preset[1] != B:
                preset = None
            return self._add_small({"solution": solution, "objective": objective, "measures": measures}, B, preset)
        data = validate_batch(self, {"solution": solution, "objective": objective, "measures": measures, **fields})
        if self._extra_names != set(fields):
            raise ValueError(
                f"`data` has keys {list(data)} but should have the same keys as this archive's store, i.e., "
                f"{[n for n in self._store.field_list if n != 'threshold']}. This error may occur if the archive "
                "has extra_fields but the fields were not passed to archive.add() or scheduler.tell()."
            )
        B = int(data["solution"].shape[0])
        odt = self.dtypes["objective"]
        self._last_idx = None
        if preset is not None and preset[1] != B:
            preset = None
        if B == 0:
            if dev_in:
                return {"status": torch.empty(0, dtype=torch.int32, device=self._device),
                        "value": torch.empty(0, dtype=torch_dtype(odt), device=self._device)}
            return {"status": np.empty(0, dtype=np.int32), "value": np.empty(0, dtype=odt)}
        store = self._store
        if self._shard_store:
            return self._add_owned(data, B, dev_in)
        if (not dev_in and B <= _SMALL_ROWS and B * self._row_bytes_total <= _SMALL_BYTES and self._dp_world == 1
                and self._small_adds and self._small_index_ok(B)
                and torch.cuda.current_device() == self._device.index):
            return self._add_small(data, B, preset)
        if (not dev_in and self._dp_world == 1 and self.cells * 2 <= B and self._zero_copy_rows
                and (preset is not None or self._host_measures_ok())
                and torch.cuda.current_device() == self._device.index):
            out = self._add_zero_copy(data, B, preset)
            if out is not None:
                return out
        # (entering a device context costs more than the launches below: only when it is not current already)
        with (_NULL_CTX if torch.cuda.current_device() == self._device.index else torch.cuda.device(self._device)):
            fused = self._dp_world == 1 and self._fused_index_insert and preset is None
            global_search = self._dp_world > 1 and getattr(self, "_world", 1) > 1  # centroid-sharded CVT
            if self._dp_world > 1:
                preset = None
            rows_event = None
            self._last_zero_copy = False
            # measures + objective first: the index kernels only need those, so the (much larger)
            # solution / extra-field rows travel host->device on a side stream while the search runs
            idx = None
            if (not dev_in and not fused and preset is None and self._dp_world == 1 and B >= _CHUNK_MIN_ROWS
                    and isinstance(data["measures"], np.ndarray)):
                # large host batch: the measures travel in pieces and the search of piece k runs while piece k+1 is
                # still on the PCIe bus (and, for pageable arrays, still being staged by the host)
                dev = {"objective": self._to_device(data["objective"], store.dtypes["objective"], "objective")}
                dev["measures"], idx = self._search_chunked(data["measures"], B)
            else:
                dev = {name: self._to_device(data[name], store.dtypes[name], name) for name in ("measures", "objective")}
            if global_search:
                # every rank searches ALL queries against its centroid shard: gather the measures first
                dev["measures"] = self._all_gather_many([dev["measures"]])[0]
            if preset is not None:  # an archive of the same geometry has just indexed this very batch
                idx = preset[0]
            elif idx is None:
                idx = None if fused else self._index_of_device(dev["measures"], store._flags_ptr)
            rest = [n for n in data if n not in dev]
            if dev_in or all(isinstance(data[n], torch.Tensor) for n in rest):
                for n in rest:
                    dev[n] = self._to_device(data[n], store.dtypes[n], n)
            else:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self._device)
                    self._copy_event = torch.cuda.Event()
                with torch.cuda.stream(self._copy_stream):
                    for n in rest:
                        dev[n] = self._to_device(data[n], store.dtypes[n], n)
                    self._copy_event.record(self._copy_stream)
                if self._dp_world > 1:
                    torch.cuda.current_stream().wait_event(self._copy_event)
                else:  # only the row-copy phase of the insertion waits for the solutions (qd_insert)
                    rows_event = self._copy_event
            B_local = B
            peer_tab = None
            if self._dp_world > 1:
                if global_search:  # idx and measures are already global
                    _, dev, B = self._gather_global_batch(None, dev, B_local, skip=("measures",))
                else:
                    dev = {n: dev[n].contiguous() for n in dev}
                    got = self._symm_exchange(idx, dev, B_local)
                    if got is not None:  # row fields stay on their ranks: the insert kernel gathers the winners' rows
                        idx, g_obj, peer_tab, B = got
                        dev = {"objective": g_obj}
                    else:
                        idx, dev, B = self._gather_global_batch(idx, dev, B_local)
            status, value, status_h, value_h, pooled = self._out_buffers(B, dev_in)
            host_out = False
            names = store.row_fields
            src = None if peer_tab else (C.c_void_p * max(len(names), 1))(*[dev[n].data_ptr() for n in names])

            def insert(part):
                if fused:
                    self._grid_insert(dev, B, src, status, value, part)
                else:
                    _lib.check(_lib.lib().qd_insert_sharded_rows(
                        C.byref(store.abi_store()), B, idx.data_ptr(), dev["objective"].data_ptr(), src, peer_tab,
                        B_local, self._lr_f, self._tmin_f, status.data_ptr(), value.data_ptr(),
                        store._result_dev.data_ptr(), part, current_stream_ptr()))

            pipelined = (self._pipeline_commit and rows_event is None and dev_in and self._dp_world == 1
                         and all(is_device_tensor(v) for v in data.values()))
            if pipelined:
                side, rows_done, copy_done = self._pipe_state()
                cur = torch.cuda.current_stream()
                buf = self._pipe_calls & 1
                bits = _lib.INSERT_BUF1 if buf else 0
                if self._pipe_calls >= 2:  # the copy that read this list buffer two calls ago
                    cur.wait_event(copy_done[buf])
                probe = self.__dict__.get("_pipe_probe")  # tools/pipe_probe.py: timed events around both launches
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if probe is not None else None
                if ev:
                    ev[0].record(cur)
                insert(_lib.INSERT_ROWS | bits)
                rows_done.record(cur)
                if ev:
                    ev[1].record(cur)
                side.wait_event(rows_done)
                with torch.cuda.stream(side):
                    if ev:
                        ev[2].record(side)
                    insert(_lib.INSERT_COPY | bits)
                    copy_done[buf].record(side)
                    if ev:
                        ev[3].record(side)
                        probe.append(ev)
                for t in dev.values():  # the allocator must not hand the batch out again before the copy has read it
                    t.record_stream(side)
                self._pipe_calls += 1
                self._pipe_pending = copy_done[buf]
            elif rows_event is None:
                self._join_copy()
                insert(_lib.INSERT_WHOLE)
            else:
                self._join_copy()
                # status / value are final after the first part: they travel back while the solutions
                # are still travelling in (full-duplex PCIe); only the row copy waits for them
                insert(_lib.INSERT_CLASSIFY)
                status_h.copy_(status, non_blocking=True)
                value_h.copy_(value, non_blocking=True)
                torch.cuda.current_stream().wait_event(rows_event)
                insert(_lib.INSERT_COMMIT)
            store._pending = True
            lo, hi = self._dp_rank * B_local, (self._dp_rank + 1) * B_local  # this rank's slice of the batch
            mixed = any(not is_device_tensor(v) for v in data.values())  # host arrays ride on reusable staging buffers
            if dev_in and not mixed and not self._sync_device_adds and not self._needs_flag_check:
                # CUDA tensors in, CUDA tensors out: nothing to wait for. Stats / len / the finite
                # check are read back lazily (_sync).
                self._inflight = dev  # keeps the batch alive until the stream has consumed it
                return {"status": status[lo:hi], "value": value[lo:hi]} if self._dp_world > 1 else \
                    {"status": status, "value": value}
            store._result_host.copy_(store._result_dev, non_blocking=True)
            if not dev_in and rows_event is None and not host_out:
                status_h.copy_(status, non_blocking=True)
                value_h.copy_(value, non_blocking=True)
            _lib.check(_lib.lib().qd_stream_synchronize(current_stream_ptr()))
        res = _lib.AddResult.from_buffer_copy(store._result_host.numpy().tobytes())
        flags = store.consume(res) | res.flags
        if flags & _lib.FLAG_CVT_OVERFLOW and not flags & (_lib.FLAG_NONFINITE_OBJECTIVE | _lib.FLAG_NONFINITE_MEASURES):
            # candidate list of the tensor-core search overflowed; nothing was committed: redo with the exact scan
            self._overflow_fallbacks += 1
            self._force_scan = True
            try:
                return self.add(data["solution"], data["objective"], data["measures"],
                                **{k: data[k] for k in fields})
            finally:
                self._force_scan = False
        self._raise_for_flags(flags)
        self._apply_state(res)
        if self._dp_world == 1:
            self._last_idx = (idx if idx is not None else self._idx_buf[:B], B)
        if dev_in:
            return {"status": status[lo:hi], "value": value[lo:hi]} if self._dp_world > 1 else \
                {"status": status, "value": value}
        if pooled is not None:  # views of a pooled page-locked buffer (see _out_buffers)
            return {"status": pooled[2][lo:hi], "value": pooled[3][lo:hi]}
        return {"status": status_h.numpy()[lo:hi].copy(), "value": value_h.numpy()[lo:hi].copy()}

    # ---- add() into a store sharded by cell range (owner computes) ------------------------------------------
    def _add_owned(self, data, B_local, dev_in):
        """SURVEY.md 8(e) row 2. Every rank passes its slice of the global batch; the slices' (cell, objective) are
        exchanged over NVLink peer memory exactly as for the replicated store (12 B per row), but every rank then
        classifies, resolves and commits only the rows whose cell it OWNS (qd_insert_owned_cells), pulls only its own
        winners' rows from their owners, and writes status / value of a row back into the symmetric buffer of the
        rank the row came from; a second cross-rank barrier completes that reverse route. Per-rank insert work and
        row traffic are 1 / n_ranks of the replicated store's."""
        store = self._store
        with (_NULL_CTX if torch.cuda.current_device() == self._device.index else torch.cuda.device(self._device)):
            self._join_copy()
            # measures + objective first: the search needs nothing else, so the (much larger) solution / extra-field
            # rows of a host batch travel on a side stream while it runs
            dev = {name: self._to_device(data[name], store.dtypes[name], name).contiguous()
                   for name in ("measures", "objective")}
            idx = self._index_of_device(dev["measures"], store._flags_ptr)
            rest = [n for n in data if n not in dev]
            if dev_in or all(isinstance(data[n], torch.Tensor) for n in rest):
                for n in rest:
                    dev[n] = self._to_device(data[n], store.dtypes[n], n).contiguous()
            else:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self._device)
                    self._copy_event = torch.cuda.Event()
                with torch.cuda.stream(self._copy_stream):
                    for n in rest:
                        dev[n] = self._to_device(data[n], store.dtypes[n], n).contiguous()
                    self._copy_event.record(self._copy_stream)
                torch.cuda.current_stream().wait_event(self._copy_event)
            got = self._symm_exchange(idx, dev, B_local)
            if got is None:
                raise RuntimeError("shard_store=True needs NVLink peer memory (torch symmetric memory) between the ranks")
            g_idx, g_obj, peer_tab, B = got
            handle, (out_tab, status_v, value_v) = self._symm_last
            _lib.check(_lib.lib().qd_insert_owned_cells(
                C.byref(store.abi_store()), B, g_idx.data_ptr(), g_obj.data_ptr(), peer_tab, B_local, self._cell_lo,
                self._cell_hi, out_tab.data_ptr(), self._lr_f, self._tmin_f, store._result_dev.data_ptr(),
                current_stream_ptr()))
            handle.barrier()  # every owner's status / value of my rows has landed
            status, value = status_v.clone(), value_v.clone()  # (the buffer set is written again two adds from now)
            store._pending = True
            self._adds_enqueued += 1
            self._inflight = dev
            if dev_in and not self._sync_device_adds:
                return {"status": status, "value": value}
            out = {"status": status.cpu().numpy(), "value": value.cpu().numpy()}
        self._sync()  # validation verdict (global: every rank saw every objective and the OR of the index flags)
        return out

    def _global_state(self):
        """COLLECTIVE (every rank, after the same add() calls): per-rank (elites, objective sum, best objective,
        has-best) of a sharded store, gathered once per add."""
        import torch.distributed as dist

        self._sync()
        if self._global_cache is not None and self._global_cache[0] == self._adds_enqueued:
            return self._global_cache[1]
        n, osum, omax, _ = self._st
        rows = _shard_util.all_gather_rows(
            (n, osum, omax if omax is not None else float("-inf"), 0.0 if omax is None else 1.0), self._dp_world,
            self._dp_group, self._device)
        self._global_cache = (self._adds_enqueued, rows)
        return rows

    def _no_sharded_read(self, what):
        raise NotImplementedError(
            f"{what} is not available on a shard_store=True archive (every rank holds only the elites of its own "
            "cells); use data() -- a collective -- or a replicated data_parallel archive")

    # ---- add() of CUDA tensors that need nothing done to them --------------------------------------------
    def _add_device(self, solution, objective, measures, fields, preset):
        """The asynchronous route with the host work cut to what it has to be: the batch is already on this
        archive's device in the store's dtypes, so it is checked (same errors as validate_batch would raise: any
        mismatch falls back to the general path, which raises them) and handed to the two launches. configs[1]'s
        device-resident step is bound by this function, not by its kernels (36 us)."""
        if (self._dp_world != 1 or self._pipeline_commit or self._sync_device_adds
                or self._extra_names != fields.keys() or torch.cuda.current_device() != self._device.index):
            return None
        store = self._store
        tdt = self._tdt
        B = solution.shape[0] if (solution.dim() == 2 and isinstance(self._solution_dim, int)) else -1
        if (B <= 0 or solution.shape[1] != self._solution_dim or solution.dtype != tdt["solution"]
                or not solution.is_contiguous() or solution.device != self._device
                or not is_device_tensor(objective) or objective.dim() != 1 or objective.shape[0] != B
                or objective.dtype != tdt["objective"] or not objective.is_contiguous()
                or not is_device_tensor(measures) or measures.dim() != 2 or measures.shape[0] != B
                or measures.shape[1] != self._measure_dim or measures.dtype != tdt["measures"]
                or not measures.is_contiguous()):
            return None
        for n, t in fields.items():
            want = store._fields[n]
            if (not is_device_tensor(t) or t.shape[0] != B or tuple(t.shape[1:]) != tuple(want.shape[1:])
                    or t.dtype != want.dtype or not t.is_contiguous()):
                return None
        if preset is not None and preset[1] != B:
            preset = None
        if not self._small_index_ok(B):  # (tensor-core search: its overflow flag needs the general path's re-run)
            return None
        self._join_copy()
        self._last_zero_copy = False
        out = torch.empty(B * 12 if tdt["objective"] == torch.float64 else B * 8, dtype=torch.uint8, device=self._device)
        vbytes = out.numel() - 4 * B
        value, status = out[:vbytes].view(tdt["objective"]), out[vbytes:].view(torch.int32)
        dev = {"solution": solution, "objective": objective, "measures": measures, **fields}
        src = (C.c_void_p * max(len(store.row_fields), 1))(*[dev[n].data_ptr() for n in store.row_fields])
        fused = self._fused_index_insert and preset is None
        if fused:
            self._grid_insert(dev, B, src, status, value, _lib.INSERT_WHOLE)
            idx = None
        else:
            if preset is not None:
                idx = preset[0]
            else:
                # one reusable index buffer per stream: the insert kernel of the previous add() still reads it, and
                # only work queued on the same stream is ordered behind that
                sp_now = current_stream_ptr()
                buf = self._dev_idx if self.__dict__.get("_dev_idx_stream") == sp_now else None
                if buf is None or buf.numel() < B:
                    buf = self._dev_idx = torch.empty(B, dtype=torch.int32, device=self._device)
                    self._dev_idx_stream = sp_now
                idx = self._index_of_device(measures, store._flags_ptr, out=buf[:B] if buf.numel() != B else buf)
            _lib.check(_lib.lib().qd_insert(C.byref(store.abi_store()), B, idx.data_ptr(), objective.data_ptr(), src,
                                            self._lr_f, self._tmin_f, status.data_ptr(), value.data_ptr(),
                                            store._result_dev.data_ptr(), _lib.INSERT_WHOLE, current_stream_ptr()))
        store._pending = True
        self._inflight = dev  # keeps the batch alive until the stream has consumed it
        self._last_idx = (idx if idx is not None else self._idx_buf[:B], B)
        return {"status": status, "value": value}

    # ---- add() of small host batches ---------------------------------------------------------------------
    def _plain_batch_ok(self, solution, objective, measures):
        """True for a NumPy batch that needs no conversion or further validation and is small enough for _add_small."""
        nd = self._np_dt
        B = solution.shape[0] if solution.ndim == 2 else 0
        return (0 < B <= _SMALL_ROWS and not self._extra_names and self._dp_world == 1
                and isinstance(self._solution_dim, int) and solution.shape[1] == self._solution_dim
                and solution.dtype == nd[0] and objective.ndim == 1 and objective.shape[0] == B and objective.dtype == nd[1]
                and measures.ndim == 2 and measures.shape[0] == B and measures.shape[1] == self._measure_dim
                and measures.dtype == nd[2] and B * self._row_bytes_total <= _SMALL_BYTES and self._small_index_ok(B)
                and torch.cuda.current_device() == self._device.index)

    def _small_index_ok(self, B):
        """True when the index kernel of a B-row add may read the measures straight from page-locked host memory."""
        return True

    def _small_plan(self, B):
        """Everything a B-row host add() needs, built once per batch size: a page-locked block the NumPy inputs
        are copied into and the kernels read (and write status / value) directly, and the ABI argument arrays."""
        plan = self._small.get(B)
        if plan is not None:
            return plan
        if len(self._small) > 16:
            self._small.clear()
        store = self._store
        names = ["measures", "objective"] + [n for n in store.row_fields if n != "measures"]
        odt = np.dtype(self.dtypes["objective"])
        spec = [(n, store._np_dtypes[n], tuple(store._fields[n].shape[1:])) for n in names]
        spec += [("status", np.dtype(np.int32), ()), ("value", odt, ())]
        offs, total = [], 0
        for _, dt, shp in spec:
            offs.append(total)
            total += (int(np.prod((B, *shp))) * dt.itemsize + 255) // 256 * 256
        block = torch.zeros(total, dtype=torch.uint8).pin_memory()
        host = block.numpy()
        views, ptrs = {}, {}
        for (n, dt, shp), off in zip(spec, offs):
            views[n] = host[off:off + int(np.prod((B, *shp))) * dt.itemsize].view(dt).reshape((B, *shp))
            ptrs[n] = block.data_ptr() + off
        rows = store.row_fields
        plan = self._small[B] = {
            "block": block, "views": views, "ptrs": ptrs, "inputs": names,
            "src": (C.c_void_p * max(len(rows), 1))(*[ptrs[n] for n in rows]),
            "meas": self._HostRows(ptrs["measures"], (B, self.measure_dim)),
            "obj": self._HostRows(ptrs["objective"], (B,)),
            "status": self._HostRows(ptrs["status"], (B,)), "value": self._HostRows(ptrs["value"], (B,)),
            "idx": torch.empty(B, dtype=torch.int32, device=self._device),
        }
        return plan

    def _add_small(self, data, B, preset=None):
        """Host batches of up to a few thousand rows (the reference's own benchmark adds 100 rows at a time,
        benchmarks/cvt_add.py:79-86): the cost of such an add() is launch and copy-call latency, not work. The
        inputs are copied into one page-locked block (host memcpy), the index kernel and the single-CTA insert
        kernel read them there and write status / value / qd_add_result back into page-locked memory themselves:
        two launches and one stream synchronize, no copy calls."""
        store = self._store
        plan = self._small_plan(B)
        views = plan["views"]
        for n in plan["inputs"]:
            np.copyto(views[n], data[n], casting="unsafe")
        self._join_copy()
        self._last_zero_copy = False
        L = _lib.lib()
        sp = current_stream_ptr()
        fused = self._fused_index_insert and preset is None
        if fused:
            self._grid_insert({"measures": plan["meas"], "objective": plan["obj"]}, B, plan["src"], plan["status"],
                              plan["value"], _lib.INSERT_WHOLE)
            idx = None
        else:
            idx = preset[0] if preset is not None else self._index_of_device(plan["meas"], store._flags_ptr,
                                                                            out=plan["idx"])
            _lib.check(L.qd_insert(C.byref(store.abi_store()), B, idx.data_ptr(), plan["ptrs"]["objective"], plan["src"],
                                   self._lr_f, self._tmin_f, plan["ptrs"]["status"], plan["ptrs"]["value"],
                                   store._result_dev.data_ptr(), _lib.INSERT_WHOLE, sp))
        store._pending = True
        _lib.check(L.qd_stream_synchronize(sp))
        res = store._result_live
        flags = store.consume(res) | res.flags
        if flags & _lib.FLAG_CVT_OVERFLOW and not flags & (_lib.FLAG_NONFINITE_OBJECTIVE | _lib.FLAG_NONFINITE_MEASURES):
            self._overflow_fallbacks += 1
            self._force_scan = True
            try:
                return self._add_small(data, B, preset)
            finally:
                self._force_scan = False
        self._raise_for_flags(flags)
        self._apply_state(res)
        self._last_idx = (idx if idx is not None else self._idx_buf[:B], B)
        return {"status": views["status"].copy(), "value": views["value"].copy()}

    # ---- add() of page-locked host batches: zero-copy plan ----------------------------------------------
    class _HostRows:
        """A page-locked host array seen by the kernels: just an address and a shape."""
        __slots__ = ("ptr", "shape")

        def __init__(self, ptr, shape):
            self.ptr, self.shape = ptr, shape

        def data_ptr(self):
            return self.ptr

    def _is_pinned(self, arr):
        key = (arr.__array_interface__["data"][0], arr.nbytes)
        hit = self._pin_cache.get(key)
        if hit is None:
            if len(self._pin_cache) > 512:
                self._pin_cache.clear()
            hit = self._pin_cache[key] = bool(torch.from_numpy(arr).is_pinned())
        return hit

    def _add_zero_copy(self, data, B, preset=None):
        """At most `cells` rows of a batch can win, so shipping every solution row to the GPU is wasted PCIe
        traffic. With page-locked inputs: the search / classify phases stream the measures straight from host
        memory (coalesced zero-copy reads, no DMA set-up latency) while the objective vector travels by DMA
        next to them; the copy phase of the insert kernel gathers just the winners' rows of the solution /
        extra fields from host memory; status / value are written by the kernel directly into the page-locked
        buffers the caller receives. Returns None when the batch is not eligible (pageable / strided arrays)."""
        store = self._store
        dts = store._np_dtypes
        for n, arr in data.items():
            if (not isinstance(arr, np.ndarray) or not arr.flags.c_contiguous or arr.dtype != dts[n]
                    or not self._is_pinned(arr)):
                return None
        self._join_copy()
        L = _lib.lib()
        zc = self._zc
        if zc is None:
            cs = torch.cuda.Stream(device=self._device)
            ev = torch.cuda.Event()
            ev.record(cs)
            zc = self._zc = {"stream": cs, "sp": cs.cuda_stream, "ev": ev, "ep": ev.cuda_event, "obj": {},
                             "res": _lib.AddResult.from_address(store._result_host.data_ptr()), "store": store}
        elif zc["store"] is not store:  # retessellate / resize replaced the store
            zc["res"] = _lib.AddResult.from_address(store._result_host.data_ptr())
            zc["store"] = store
        sp = current_stream_ptr()
        obj_h = data["objective"]
        obj_d = zc["obj"].get(B)
        if obj_d is None:
            if len(zc["obj"]) > 8:
                zc["obj"].clear()
            obj_d = zc["obj"][B] = torch.empty(B, dtype=torch_dtype(store.dtypes["objective"]), device=self._device)
        _lib.check(L.qd_memcpy_async(obj_d.data_ptr(), obj_h.__array_interface__["data"][0], obj_h.nbytes, 1, zc["sp"]))
        _lib.check(L.qd_event_record(zc["ep"], zc["sp"]))
        meas = data["measures"]
        host_meas = self._HostRows(meas.__array_interface__["data"][0], meas.shape)
        fused = self._fused_index_insert and preset is None
        if preset is not None:
            idx = preset[0]
        else:
            idx = None if fused else self._index_of_device(host_meas, store._flags_ptr)
        _lib.check(L.qd_stream_wait_event(sp, zc["ep"]))
        status, value, status_h, value_h, pooled = self._out_buffers(B, False)
        if pooled is not None:  # the kernel writes straight into the buffers the caller gets
            status, value = status_h, value_h
        dev = {"measures": host_meas, "objective": obj_d}
        names = store.row_fields
        src = (C.c_void_p * max(len(names), 1))(*[data[n].__array_interface__["data"][0] for n in names])
        if fused:
            self._grid_insert(dev, B, src, status, value, _lib.INSERT_WHOLE)
        else:
            _lib.check(L.qd_insert(C.byref(store.abi_store()), B, idx.data_ptr(), obj_d.data_ptr(), src, self._lr_f,
                                   self._tmin_f, status.data_ptr(), value.data_ptr(), store._result_dev.data_ptr(),
                                   _lib.INSERT_WHOLE, sp))
        store._pending = True
        if pooled is None:
            status_h.copy_(status, non_blocking=True)
            value_h.copy_(value, non_blocking=True)
        _lib.check(L.qd_memcpy_async(store._result_host.data_ptr(), store._result_dev.data_ptr(),
                                     C.sizeof(_lib.AddResult), 2, sp))
        _lib.check(L.qd_stream_synchronize(sp))
        res = zc["res"]  # a live view of the page-locked result block
        flags = store.consume(res) | res.flags
        self._raise_for_flags(flags)
        self._apply_state(res)
        self._last_zero_copy = True
        self._zero_copy_bytes += int(res.n_winners) * self._row_bytes_total
        self._last_idx = (idx if idx is not None else self._idx_buf[:B], B)
        if pooled is not None:
            return {"status": pooled[2][0:B], "value": pooled[3][0:B]}
        return {"status": status_h.numpy().copy(), "value": value_h.numpy().copy()}

    def add_single(self, solution, objective, measures, **fields):
        """Scalar rule of _grid_archive.py:716-848, run as a batch of one (identical statuses;
        the threshold t(1-lr)+f*lr equals the batch rule with k=1 up to rounding)."""
        data = validate_single(self, {"solution": solution, "objective": objective, "measures": measures, **fields})
        info = self.add(
            np.asarray(data["solution"])[None], np.asarray(data["objective"])[None], np.asarray(data["measures"])[None],
            **{k: np.asarray(v)[None] for k, v in fields.items()},
        )
        return {"status": info["status"][0], "value": info["value"][0]}

    def clear(self):
        self._sync()
        self._store.clear()
        self._stats_reset()
        self._global_cache = None
        self._adds_enqueued += 1

    # ---- read side (A10, _grid_archive.py:850-924) ---------------------------------------------
    def retrieve(self, measures):
        measures = np.asarray(measures, dtype=self.dtypes["measures"])
        check_batch_shape(measures, "measures", self.measure_dim, "measure_dim")
        idx = self.index_of(measures)
        if self._shard_store:
            return self._retrieve_cells_sharded(np.asarray(idx))
        occupied, data = self._store.retrieve(idx)
        fill_sentinel_values(occupied, data)
        return occupied, data

    def _retrieve_cells_sharded(self, cells):
        """COLLECTIVE retrieve by archive-wide cell number: every rank looks up the cells it owns, the answers are
        exchanged (host objects: this is a read API for a handful of rows, not a data path) and stitched."""
        import torch.distributed as dist

        self._sync()
        cells = np.asarray(cells, dtype=np.int64)
        mine = (cells >= self._cell_lo) & (cells < self._cell_hi)
        occ, data = self._store.retrieve(np.where(mine, cells - self._cell_lo, 0).astype(np.int32))
        parts = [None] * self._dp_world
        dist.all_gather_object(parts, (mine, np.asarray(occ) & mine, data), group=self._dp_group)
        occupied, out = _shard_util.stitch_shard_retrieve(cells, parts)
        fill_sentinel_values(occupied, out)
        return occupied, out

    def _sharded_occupied_list(self):
        """COLLECTIVE: the archive-wide occupied list, shard after shard."""
        import torch.distributed as dist

        self._sync()
        local = (self._store._sync_occupied_list()[: len(self._store)].astype(np.int64) + self._cell_lo).astype(np.int32)
        parts = [None] * self._dp_world
        dist.all_gather_object(parts, local, group=self._dp_group)
        return np.concatenate(parts)

    def retrieve_single(self, measures):
        measures = np.asarray(measures, dtype=self.dtypes["measures"])
        check_shape(measures, "measures", self.measure_dim, "measure_dim")
        check_finite(measures, "measures")
        occupied, data = self.retrieve(measures[None])
        return occupied[0], {f: a[0] for f, a in data.items()}

    def data(self, fields=None, return_type="dict"):
        self._sync()  # adopts the state of insertions still in flight (and raises for a batch the device dropped)
        if self._shard_store:
            # collective: the shards' elites, rank after rank (= ascending cell ranges, insertion order within a
            # shard), with archive-wide cell indices
            import torch.distributed as dist

            if return_type != "dict":
                self._no_sharded_read(f"data(return_type={return_type!r})")
            single = isinstance(fields, str)
            want = None if fields is None else ([fields] if single else list(fields))
            local = self._store.data(None if want is None else sorted(set(want) | {"index"}), "dict")
            local["index"] = (local["index"] + self._cell_lo).astype(np.int32)
            parts = [None] * self._dp_world
            dist.all_gather_object(parts, local, group=self._dp_group)
            out = _shard_util.stitch_shard_data(parts, want)
            return out[fields] if single else out
        return self._store.data(fields, return_type)

    def sample_elite_cells(self, n, replace=True):
        """The cell indices `sample_elites(n)` would retrieve (same draw from the archive's generator,
        _grid_archive.py:912-924), without fetching the rows: device-side consumers gather them in place."""
        if self.empty:
            raise IndexError("No elements in archive.")
        if self._shard_store:  # collective; the ranks' generators were seeded alike and draw alike
            glist = self._sharded_occupied_list()
            if not replace and n > len(glist):
                raise ValueError(
                    "Cannot take a larger sample than the number of elites in the archive when 'replace=False'")
            return np.ascontiguousarray(glist[self._rng.choice(len(glist), size=n, replace=replace)])
        if not replace and n > len(self._store):
            raise ValueError(
                "Cannot take a larger sample than the number of elites in the archive when 'replace=False'")
        random_indices = self._rng.choice(len(self._store), size=n, replace=replace)
        return np.ascontiguousarray(self._store._sync_occupied_list()[random_indices])

    def sample_elites(self, n, replace=True):
        if self._shard_store:
            return self._retrieve_cells_sharded(self.sample_elite_cells(n, replace))[1]
        if self.empty:
            raise IndexError("No elements in archive.")
        self._sync()
        if not replace and n > len(self._store):
            raise ValueError(
                "Cannot take a larger sample than the number of elites in the archive when 'replace=False'")
        random_indices = self._rng.choice(len(self._store), size=n, replace=replace)
        selected = self._store._sync_occupied_list()[random_indices]
        return self._store.retrieve(selected)[1]
