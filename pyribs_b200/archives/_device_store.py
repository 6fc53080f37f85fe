"""DeviceStore: the HBM-resident struct-of-arrays behind every archive.

Drop-in for the reference's `ArrayStore` read API (ribs/archives/_array_store.py:71-651):
`retrieve`, `data`, `occupied`, `occupied_list`, `field_desc`, `dtypes`, iteration with
the same "modified during iteration" RuntimeError, `clear`, `resize`. The write side is not
`ArrayStore.add`: insertion is the fused `qd_insert` kernel sequence driven by
`DeviceArchive.add`, which reproduces `_array_store.py:560-608` (ascending append of new
cells to `occupied_list`, last field scatter) on the device.

Layout in HBM (capacity = number of cells):
    fields[name]   (capacity, *shape) contiguous, row-major, field dtype
    occupied       (capacity,) uint8
    occupied_list  (capacity,) int32, first n_occupied entries valid, insertion order
    scratch        per-cell reduction scratch of qd_insert (see csrc/insert.cu)
"""
from __future__ import annotations

import ctypes as C
import itertools
import numbers

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import arr_readonly, torch_dtype


def current_stream_ptr() -> int:
    """Raw cudaStream_t of torch's current stream (the C accessor is ~10x cheaper than
    torch.cuda.current_stream())."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


class DeviceStoreIterator:
    """Iterates entries in insertion order (ArrayStoreIterator, _array_store.py:33-68)."""

    def __init__(self, store: "DeviceStore"):
        self.store = store
        self.iter_idx = 0
        store.sync()
        self.state = list(store._updates)
        self._snapshot = None

    def __iter__(self):
        return self

    def __next__(self):
        self.store.sync()
        if self.state != self.store._updates:
            raise RuntimeError("ArrayStore was modified with add() or clear() during iteration.")
        if self.iter_idx >= len(self.store):
            raise StopIteration
        if self._snapshot is None:  # one D2H of all occupied rows instead of one per entry
            self._snapshot = self.store.data()
        d = {"index": self._snapshot["index"][self.iter_idx]}
        for name in self.store._fields:
            d[name] = self._snapshot[name][self.iter_idx]
        self.iter_idx += 1
        return d


class DeviceStore:
    def __init__(self, field_desc, capacity, device=None):
        if not torch.cuda.is_available():
            raise _lib.QdError("pyribs_b200 needs a CUDA device (there is no CPU fallback)")
        _lib.lib()  # fail loudly if the native library is missing
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        capacity = int(capacity)
        self._capacity = capacity
        self._n_occupied = 0
        self._updates = [0, 0]
        self._fields = {}
        self._np_dtypes = {}
        for name, (field_shape, dtype) in field_desc.items():
            if name == "index":
                raise ValueError(f"`{name}` is a reserved field name.")
            if not name.isidentifier():
                raise ValueError(f"Field names must be valid identifiers: `{name}`")
            if isinstance(field_shape, numbers.Integral):
                field_shape = (field_shape,)
            np_dt = np.dtype(dtype)
            self._np_dtypes[name] = np_dt
            self._fields[name] = torch.empty((capacity, *field_shape), dtype=torch_dtype(np_dt), device=self._device)
        with torch.cuda.device(self._device):
            self._occupied = torch.zeros(capacity, dtype=torch.uint8, device=self._device)
            self._occupied_list = torch.empty(capacity, dtype=torch.int32, device=self._device)
            nbytes = _lib.lib().qd_insert_scratch_bytes(capacity)
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            _lib.check(_lib.lib().qd_insert_scratch_init(self._scratch.data_ptr(), capacity, current_stream_ptr()))
            self._flags_ptr = _lib.lib().qd_insert_flags_ptr(self._scratch.data_ptr(), capacity)
            self._result_dev = torch.zeros(C.sizeof(_lib.AddResult), dtype=torch.uint8, device=self._device)
            self._result_host = torch.zeros(C.sizeof(_lib.AddResult), dtype=torch.uint8).pin_memory()
            # every insertion also writes its result block into this page-locked buffer (read after a stream
            # synchronize, no device->host copy call); `_result_live` is a live view of it
            _lib.check(_lib.lib().qd_insert_set_result_mirror(self._scratch.data_ptr(), capacity,
                                                              self._result_host.data_ptr(), current_stream_ptr()))
            self._result_live = _lib.AddResult.from_address(self._result_host.data_ptr())
            self._abi = None
            self._snapshot = None
            nsnap = int(_lib.lib().qd_best_snapshot_bytes(C.byref(self.abi_store())))
            self._snapshot = torch.zeros(nsnap, dtype=torch.uint8, device=self._device)
            self._abi = None
            # occupancy 0, thresholds NaN (= unoccupied), device-side element count / stats reset
            _lib.check(_lib.lib().qd_store_clear(C.byref(self.abi_store()), current_stream_ptr()))
        self._occupied_list_host = np.empty(capacity, dtype=np.int32)
        self._occupied_list_synced = 0
        # Device-resident store state (csrc/insert.cu InsertState) mirrored lazily on the host:
        # `_pending` is set by every enqueued insertion whose qd_add_result was not read back yet.
        self._pending = False
        self._commits_base = 0
        self._failed_seen = 0
        self._last = None  # last consumed qd_add_result

    # ---- ABI view ---------------------------------------------------------------------
    @property
    def row_fields(self):
        """Fields copied row-wise by the commit kernel (all but objective / threshold)."""
        return [n for n in self._fields if n not in ("objective", "threshold")]

    def abi_store(self) -> _lib.Store:
        if self._abi is None:
            st = _lib.Store()
            st.capacity = self._capacity
            st.obj_dtype = _lib.QD_F64 if self._np_dtypes["objective"] == np.float64 else _lib.QD_F32
            st.occupied = self._occupied.data_ptr()
            st.occupied_list = self._occupied_list.data_ptr()
            st.objective = self._fields["objective"].data_ptr()
            st.threshold = self._fields["threshold"].data_ptr()
            names = self.row_fields
            if len(names) > _lib.QD_MAX_FIELDS:
                raise ValueError(f"at most {_lib.QD_MAX_FIELDS} row fields are supported")
            st.n_fields = len(names)
            for i, n in enumerate(names):
                t = self._fields[n]
                st.field_dst[i] = t.data_ptr()
                st.field_row_bytes[i] = t[0].numel() * t.element_size() if t.dim() > 1 else t.element_size()
            st.scratch = self._scratch.data_ptr()
            st.best_snapshot = self._snapshot.data_ptr() if self._snapshot is not None else None
            self._abi = st
        return self._abi

    # ---- host mirror of the device-side state -------------------------------------------------
    def consume(self, res):
        """Adopts a qd_add_result read back from the device. Returns the flags of a call that was
        dropped by validation and not reported yet (0 otherwise)."""
        self._pending = False
        self._last = res
        self._n_occupied = int(res.n_occupied)
        self._updates[0] = self._commits_base + int(res.n_commits)
        if res.n_failed > self._failed_seen:
            self._failed_seen = int(res.n_failed)
            return int(res.last_failed_flags) or int(res.flags)
        return 0

    def join_row_copy(self):
        """A pipelined add() (DeviceArchive.pipeline_commit) copies the winners' rows on a side stream; anything that
        reads or rewrites the row fields first makes the current stream wait for it."""
        hook = self.__dict__.get("_copy_join")
        if hook is not None:
            hook()

    def sync(self):
        """Reads the device-side store state back if insertions are still in flight. Returns the
        unreported validation flags (see consume). A store that belongs to an archive hands the read-back to
        the archive (`_owner_sync` = DeviceArchive._sync), which also adopts the running stats and raises the
        reference's ValueError for a batch the device-side finite check dropped -- whichever of `len(store)`,
        `store.data()`, iteration ... notices the pending insertion first."""
        self.join_row_copy()
        if not self._pending:
            return 0
        owner = self.__dict__.get("_owner_sync")
        if owner is not None and not self.__dict__.get("_in_owner_sync", False):
            self._in_owner_sync = True
            try:
                owner()
            finally:
                self._in_owner_sync = False
            return 0
        return self._read_back()

    def _read_back(self):
        with torch.cuda.device(self._device):
            self._result_host.copy_(self._result_dev, non_blocking=True)
            _lib.check(_lib.lib().qd_stream_synchronize(current_stream_ptr()))
        return self.consume(_lib.AddResult.from_buffer_copy(self._result_host.numpy().tobytes()))

    def read_snapshot(self):
        """Best-elite snapshot taken on the device (layout: include/qd_b200.h qd_store.best_snapshot)."""
        raw = self._snapshot.cpu().numpy()
        head = raw[:24].view(np.float64)
        out = {"objective": np.asarray(head[0], dtype=self._np_dtypes["objective"])[()],
               "threshold": np.asarray(head[1], dtype=self._np_dtypes["threshold"])[()],
               "index": np.int32(raw[16:20].view(np.int32)[0])}
        off = 32
        for n in self.row_fields:
            t = self._fields[n]
            rb = self._row_bytes(n)
            out[n] = raw[off:off + rb].view(self._np_dtypes[n]).reshape(tuple(t.shape[1:])).copy()
            if out[n].ndim == 0:
                out[n] = out[n][()]
            off += (rb + 15) // 16 * 16
        return {k: out[k] for k in [*self._fields, "index"]}

    # ---- properties (ArrayStore API) ----------------------------------------------------
    def __len__(self):
        if self._pending:
            self.sync()
        return self._n_occupied

    def __iter__(self):
        return DeviceStoreIterator(self)

    @property
    def device(self):
        return self._device

    @property
    def capacity(self):
        return self._capacity

    @property
    def occupied(self):
        return arr_readonly(self._occupied.cpu().numpy().astype(bool))

    @property
    def updates(self):
        self.sync()
        return list(self._updates)

    @property
    def occupied_device(self):
        return self._occupied

    @property
    def occupied_list(self):
        return arr_readonly(self._sync_occupied_list()[: self._n_occupied].copy())

    def _sync_occupied_list(self):
        n = len(self)
        if self._occupied_list_synced < n:
            s = self._occupied_list_synced
            self._occupied_list_host[s:n] = self._occupied_list[s:n].cpu().numpy()
            self._occupied_list_synced = n
        return self._occupied_list_host

    @property
    def field_desc(self):
        return {name: (tuple(t.shape[1:]), self._np_dtypes[name]) for name, t in self._fields.items()}

    @property
    def dtypes(self):
        return dict(self._np_dtypes)

    @property
    def dtypes_with_index(self):
        d = self.__dict__.get("_dtypes_wi")
        if d is None:
            d = self.__dict__["_dtypes_wi"] = {**self._np_dtypes, "index": np.dtype(np.int32)}
        return d

    @property
    def field_list(self):
        return list(self._fields)

    @property
    def field_list_with_index(self):
        return [*self._fields, "index"]

    # ---- read side ----------------------------------------------------------------------
    def gather_device(self, indices: torch.Tensor, names):
        """Gather kernel (A6): returns (occupied uint8 tensor, {name: tensor}) on the device."""
        M = int(indices.numel())
        names = list(dict.fromkeys(names))
        out = {n: torch.empty((M, *self._fields[n].shape[1:]), dtype=self._fields[n].dtype, device=self._device)
               for n in names}
        occ = torch.empty(M, dtype=torch.uint8, device=self._device)
        nf = len(names)
        if nf > _lib.QD_MAX_FIELDS:
            raise ValueError("too many fields")
        src = (C.c_void_p * max(nf, 1))(*[self._fields[n].data_ptr() for n in names])
        dst = (C.c_void_p * max(nf, 1))(*[out[n].data_ptr() for n in names])
        rb = (C.c_int64 * max(nf, 1))(*[self._row_bytes(n) for n in names])
        self.join_row_copy()
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().qd_store_gather(M, indices.data_ptr(), nf, src, dst, rb, self._occupied.data_ptr(),
                                                  occ.data_ptr(), current_stream_ptr()))
        return occ, out

    def _row_bytes(self, name):
        t = self._fields[name]
        return (t[0].numel() if t.dim() > 1 else 1) * t.element_size()

    def retrieve(self, indices, fields=None, return_type="dict"):
        """ArrayStore.retrieve (_array_store.py:330-485) with NumPy outputs."""
        single = isinstance(fields, str)
        if isinstance(indices, torch.Tensor):
            idx_dev = indices.to(device=self._device, dtype=torch.int32)
            idx_np = None
        else:
            idx_np = np.asarray(indices, dtype=np.int32)
            idx_dev = torch.from_numpy(np.ascontiguousarray(idx_np)).to(self._device)
        if not single and return_type not in ("dict", "tuple", "pandas"):
            raise ValueError(f"Invalid return_type {return_type}.")
        if single:
            fields = [fields]
        elif fields is None:
            fields = list(itertools.chain(self._fields, ["index"]))
        else:
            fields = list(fields)
        for name in fields:
            if name != "index" and name not in self._fields:
                raise ValueError(f"`{name}` is not a field in this ArrayStore.")
        occ_d, got = self.gather_device(idx_dev, [n for n in fields if n != "index"])
        occupied = occ_d.cpu().numpy().astype(bool)
        host = {n: t.cpu().numpy() for n, t in got.items()}
        if idx_np is None:
            idx_np = idx_dev.cpu().numpy()
        arrays = [np.array(idx_np, copy=True) if n == "index" else host[n] for n in fields]
        if single:
            return occupied, arrays[0]
        if return_type == "dict":
            return occupied, dict(zip(fields, arrays))
        if return_type == "tuple":
            return occupied, tuple(arrays)
        import pandas as pd

        cols = {}
        for n, arr in zip(fields, arrays):
            if arr.ndim == 1:
                cols[n] = arr
            elif arr.ndim == 2:
                for i in range(arr.shape[1]):
                    cols[f"{n}_{i}"] = arr[:, i]
            else:
                raise ValueError(f"Field `{n}` has shape {arr.shape[1:]} -- cannot convert fields with shape >1D to Pandas")
        return occupied, pd.DataFrame(cols, copy=False)

    def data(self, fields=None, return_type="dict"):
        return self.retrieve(self._occupied_list[: len(self)], fields, return_type)[1]

    # ---- mutation -------------------------------------------------------------------------
    def clear(self):
        self.sync()
        self._updates[1] += 1
        self._n_occupied = 0
        self._occupied_list_synced = 0
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().qd_store_clear(C.byref(self.abi_store()), current_stream_ptr()))

    def _push_state(self, objective_sum, obj_max, absmax):
        """Re-seeds the device-side state (after resize / unpickling)."""
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().qd_store_set_state(
                C.byref(self.abi_store()), self._n_occupied, float(objective_sum), int(obj_max is not None),
                float(obj_max) if obj_max is not None else 0.0, float(absmax), current_stream_ptr()))
            _lib.check(_lib.lib().qd_store_read_state(C.byref(self.abi_store()), self._result_dev.data_ptr(),
                                                      current_stream_ptr()))
        self._pending = True
        self._commits_base = self._updates[0]
        self._failed_seen = 0
        self.sync()

    def resize(self, capacity):
        capacity = int(capacity)
        if capacity <= self._capacity:
            raise ValueError(f"New capacity ({capacity}) must be greater than current capacity ({self._capacity}.")
        self.sync()
        last = self._last
        old = self._capacity
        self._capacity = capacity
        occ = torch.zeros(capacity, dtype=torch.uint8, device=self._device)
        occ[:old] = self._occupied
        self._occupied = occ
        ol = torch.empty(capacity, dtype=torch.int32, device=self._device)
        ol[:old] = self._occupied_list
        self._occupied_list = ol
        host = np.empty(capacity, dtype=np.int32)
        host[:old] = self._occupied_list_host
        self._occupied_list_host = host
        for name, t in self._fields.items():
            nt = torch.empty((capacity, *t.shape[1:]), dtype=t.dtype, device=self._device)
            nt[:old] = t
            self._fields[name] = nt
        nbytes = _lib.lib().qd_insert_scratch_bytes(capacity)
        self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().qd_insert_scratch_init(self._scratch.data_ptr(), capacity, current_stream_ptr()))
        self._flags_ptr = _lib.lib().qd_insert_flags_ptr(self._scratch.data_ptr(), capacity)
        self._abi = None
        osum = last.objective_sum if last is not None else 0.0
        omax = last.obj_max if last is not None and last.has_obj_max else None
        absmax = float(self._fields["objective"][self._occupied.bool()].abs().max().item()) if self._n_occupied else 0.0
        self._push_state(osum, omax, absmax)

    # ---- pickling: device tensors travel as NumPy (checkpoint / resume) ---------------------
    def __getstate__(self):
        self.sync()
        st = {
            "field_desc": self.field_desc,
            "capacity": self._capacity,
            "n": self._n_occupied,
            "updates": list(self._updates),
            "occupied": self._occupied.cpu().numpy(),
            "occupied_list": self._occupied_list.cpu().numpy(),
            "fields": {n: t.cpu().numpy() for n, t in self._fields.items()},
        }
        return st

    def __setstate__(self, st):
        self.__init__(st["field_desc"], st["capacity"])
        self._n_occupied = st["n"]
        self._updates = st["updates"]
        self._occupied.copy_(torch.from_numpy(st["occupied"]))
        self._occupied_list.copy_(torch.from_numpy(st["occupied_list"]))
        for n, a in st["fields"].items():
            self._fields[n].copy_(torch.from_numpy(a))
        # objective sum / best objective: a bare store re-derives them from its rows; an owning archive re-seeds
        # them with its own running values afterwards (DeviceArchive.__setstate__)
        occ = st["occupied"].astype(bool)
        obj = st["fields"]["objective"][occ] if "objective" in st["fields"] else np.zeros(0)
        self._push_state(float(obj.sum()) if obj.size else 0.0, float(obj.max()) if obj.size else None,
                         float(np.abs(obj).max()) if obj.size else 0.0)
