"""GridArchive on the B200 (API of ribs/archives/_grid_archive.py:32-976)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import check_batch_shape, check_is_1d
from pyribs_b200.archives._archive_base import DeviceArchive
from pyribs_b200.archives._device_store import DeviceStore, current_stream_ptr


class GridArchive(DeviceArchive):
    """Uniform n-D grid over measure space; every cell keeps its best solution.

    Constructor arguments, properties and method semantics follow the reference
    (`_grid_archive.py:105-182`). All cell data lives in HBM; `add`, `index_of`,
    `retrieve`, `data` and `sample_elites` run hand-written sm_100a kernels through the C
    ABI in include/qd_b200.h. Host arrays in -> NumPy out; CUDA tensors in -> CUDA tensors
    out (`status` / `value` stay on the device).
    """

    def __init__(self, *, solution_dim, dims, ranges, learning_rate=None, threshold_min=-np.inf, epsilon=1e-6,
                 qd_score_offset=0.0, seed=None, solution_dtype=None, objective_dtype=None, measures_dtype=None,
                 dtype=None, extra_fields=None, device=None, data_parallel=False, process_group=None,
                 shard_store=False):
        self._dims = np.array(dims, dtype=np.int32)
        if len(self._dims) != len(ranges):
            raise ValueError(
                f"dims (length {len(self._dims)}) and ranges (length {len(ranges)}) must be the same length")
        if len(self._dims) > 32:
            raise ValueError("at most 32 measure dimensions are supported on the GPU path")
        DeviceArchive.__init__(
            self, solution_dim=solution_dim, measure_dim=len(self._dims), cells=int(np.prod(self._dims)),
            learning_rate=learning_rate, threshold_min=threshold_min, qd_score_offset=qd_score_offset, seed=seed,
            solution_dtype=solution_dtype, objective_dtype=objective_dtype, measures_dtype=measures_dtype,
            dtype=dtype, extra_fields=extra_fields, device=device, data_parallel=data_parallel, shard_store=shard_store,
            process_group=process_group)
        mdt = self.dtypes["measures"]
        lo, hi = zip(*ranges)
        self._lower_bounds = np.array(lo, dtype=mdt)
        self._upper_bounds = np.array(hi, dtype=mdt)
        self._interval_size = self._upper_bounds - self._lower_bounds
        self._boundaries = self._compute_boundaries(self._dims, self._lower_bounds, self._upper_bounds)
        self._epsilon = np.asarray(epsilon, dtype=mdt)
        self._refresh_abi_consts()
        self._fused_index_insert = True
        self._idx_buf = None

    def _refresh_abi_consts(self):
        d = len(self._dims)
        self._c_dims = (C.c_int32 * d)(*[int(v) for v in self._dims])
        self._c_lower = (C.c_double * d)(*[float(v) for v in self._lower_bounds])
        self._c_interval = (C.c_double * d)(*[float(v) for v in self._interval_size])
        self._meas_code = _lib.QD_F64 if self.dtypes["measures"] == np.float64 else _lib.QD_F32
        self._eps_f = float(self._epsilon)

    @staticmethod
    def _compute_boundaries(dims, lower_bounds, upper_bounds):
        return [np.linspace(lo, hi, int(n) + 1) for n, lo, hi in zip(dims, lower_bounds, upper_bounds)]

    # ---- properties -----------------------------------------------------------------------
    @property
    def dims(self):
        return self._dims

    @property
    def lower_bounds(self):
        return self._lower_bounds

    @property
    def upper_bounds(self):
        return self._upper_bounds

    @property
    def interval_size(self):
        return self._interval_size

    @property
    def boundaries(self):
        return self._boundaries

    @property
    def epsilon(self):
        return self._epsilon

    # ---- A2 ---------------------------------------------------------------------------------
    def _index_of_device(self, measures: torch.Tensor, flags_ptr: int, out=None) -> torch.Tensor:
        B = int(measures.shape[0])
        idx = out if out is not None else torch.empty(B, dtype=torch.int32, device=self._device)
        _lib.check(_lib.lib().qd_grid_index_of(
            measures.data_ptr(), self._meas_code, B, len(self._dims), self._c_dims, self._c_lower, self._c_interval,
            float(self._epsilon), idx.data_ptr(), flags_ptr, current_stream_ptr()))
        return idx

    def same_geometry(self, other):
        return (type(other) is type(self) and np.array_equal(self._dims, other._dims)
                and np.array_equal(self._lower_bounds, other._lower_bounds)
                and np.array_equal(self._upper_bounds, other._upper_bounds) and self._epsilon == other._epsilon
                and self.dtypes["measures"] == other.dtypes["measures"] and self._device == other._device)

    def _grid_insert(self, dev, B, src, status, value, part):
        """A2 + A4-A8 in one ABI call (qd_grid_insert): measures are read once."""
        store = self._store
        if self._idx_buf is None or self._idx_buf.numel() < B:
            self._idx_buf = torch.empty(B, dtype=torch.int32, device=self._device)
        _lib.check(_lib.lib().qd_grid_insert(
            C.byref(store.abi_store()), B, dev["measures"].data_ptr(), self._meas_code, len(self._dims), self._c_dims,
            self._c_lower, self._c_interval, self._eps_f, dev["objective"].data_ptr(), src,
            self._lr_f, self._tmin_f, self._idx_buf.data_ptr(),
            status.data_ptr(), value.data_ptr(), store._result_dev.data_ptr(), part, current_stream_ptr()))

    def grid_to_int_index(self, grid_indices):
        grid_indices = np.asarray(grid_indices)
        check_batch_shape(grid_indices, "grid_indices", self.measure_dim, "measure_dim")
        return np.ravel_multi_index(grid_indices.T, self._dims).astype(np.int32)

    def int_to_grid_index(self, int_indices):
        int_indices = np.asarray(int_indices)
        check_is_1d(int_indices, "int_indices")
        return np.asarray(np.unravel_index(int_indices, self._dims), dtype=np.int32).T

    def retessellate(self, new_dims):
        """_grid_archive.py:926-976: re-grid and re-insert the current elites."""
        if not np.isclose(self.learning_rate, 1):
            raise ValueError("Cannot retessellate an archive with learning rate not equal to 1.")
        if len(new_dims) != self.measure_dim:
            raise ValueError(
                f"The measure space dimensionality indicated in `new_dims` is {len(new_dims)}, but this archive "
                f"has a measure space dimensionality of {self.measure_dim}.")
        cur = self.data()
        del cur["index"]
        del cur["threshold"]
        best, (_, osum, omax, _) = self.best_elite, self._st
        self._dims = np.array(new_dims, dtype=np.int32)
        self._boundaries = self._compute_boundaries(self._dims, self._lower_bounds, self._upper_bounds)
        store = DeviceStore(self._store.field_desc, capacity=int(np.prod(self._dims)), device=self._device)
        self._attach_store(store)
        self._refresh_abi_consts()
        self._idx_buf = None
        # Like the reference, which swaps the store and re-adds WITHOUT resetting its statistics
        # (_grid_archive.py:970-976): the running objective sum and the best objective carry over into the new
        # store's device-side state, so qd_score counts the survivors on top of the old sum and best_elite stays
        # the old record (old cell index included) until a strictly better solution arrives.
        absmax = float(np.abs(cur["objective"]).max()) if len(cur["objective"]) else 0.0
        store._owner_sync = None
        store._push_state(osum, omax, absmax)
        store._owner_sync = self._sync
        self._apply_state(store._last)
        self._best_elite, self._best_serial = best, self._st[3]
        if self._stats_dirty:
            self._build_stats()
        self.add(**cur)

    def _transient_extra(self):
        return ("_c_dims", "_c_lower", "_c_interval", "_idx_buf")

    def _rebuild_device_side(self):
        self._refresh_abi_consts()
        self._idx_buf = None
