"""CategoricalArchive on the device.

Reference: ribs/archives/_categorical_archive.py:31-820 -- a grid whose axes are categorical: a host-side
category -> index map (`index_of`, :302-328) in front of the same `add` body as GridArchive (:400-596). Here it IS
a GridArchive over the ranges [0, n_i): a row's measures travel to the device as `code + 0.5`, so the fused
index + insert kernel (csrc/insert.cu) puts it in cell `code` with no arithmetic ambiguity, and everything the
reference's add does (statuses against the pre-batch archive, first-in-batch ties, CMA-MAE thresholds, stats, best
elite) is the kernel path the other archives use. The measures field on the device holds the codes; every read
(`data`, `retrieve*`, `sample_elites`, `best_elite`, iteration) decodes them back into the category objects.

Not supported (raises TypeError like every archive of this package): object-dtype *solutions* (e.g. strings) --
they have no device representation.
"""
from __future__ import annotations

import contextlib

import numpy as np

from pyribs_b200._utils import check_batch_shape, check_shape, validate_single
from pyribs_b200.archives._archive_base import fill_sentinel_values, parse_all_dtypes
from pyribs_b200.archives._grid_archive import GridArchive


class CategoricalArchive(GridArchive):
    def __init__(self, *, solution_dim, categories, learning_rate=None, threshold_min=-np.inf, qd_score_offset=0.0,
                 seed=None, solution_dtype=None, objective_dtype=None, measures_dtype=None, dtype=None,
                 extra_fields=None, device=None):
        self._public_dtypes = False
        self._categories = [list(axis) for axis in categories]
        self._category_to_idx = []
        for axis in self._categories:
            if len(set(axis)) != len(axis):
                raise ValueError("categories of a dimension must be unique")
            self._category_to_idx.append(dict(zip(axis, range(len(axis)))))
        # `dtype` and the individual dtypes exclude each other exactly as in the reference (:131-134); the measures
        # dtype is object for the caller whatever was passed, and a float of category codes on the device
        sdt, odt, _ = parse_all_dtypes(dtype, solution_dtype, objective_dtype, measures_dtype)
        dims = [len(axis) for axis in self._categories]
        GridArchive.__init__(
            self, solution_dim=solution_dim, dims=dims, ranges=[(0.0, float(n)) for n in dims],
            learning_rate=learning_rate, threshold_min=threshold_min, qd_score_offset=qd_score_offset, seed=seed,
            solution_dtype=sdt, objective_dtype=odt, measures_dtype=np.float64, extra_fields=extra_fields,
            device=device)
        self._public_dtypes = True

    # ---- properties --------------------------------------------------------------------------------
    @property
    def categories(self):
        return self._categories

    @property
    def dtypes(self):
        d = dict(GridArchive.dtypes.fget(self))
        if self._public_dtypes:
            d["measures"] = np.dtype(object)
        return d

    @contextlib.contextmanager
    def _coded(self):
        """Inside: the base class sees the measures field as what it is on the device (float codes)."""
        prev, self._public_dtypes = self._public_dtypes, False
        try:
            yield
        finally:
            self._public_dtypes = prev

    # ---- category <-> code ---------------------------------------------------------------------------
    def _grid_indices(self, measures):
        return [[self._category_to_idx[i][m] for i, m in enumerate(row)] for row in measures]  # :321-324

    def _encode(self, measures):
        """(batch, measure_dim) category objects -> code + 0.5 (the centre of the category's unit cell)."""
        return np.asarray(self._grid_indices(measures), dtype=np.float64).reshape(len(measures), self.measure_dim) + 0.5

    def _decode_measures(self, coded):
        coded = np.asarray(coded)
        out = np.empty(coded.shape, dtype=object)
        flat_in, flat_out = coded.reshape(-1, self.measure_dim), out.reshape(-1, self.measure_dim)
        for r in range(flat_in.shape[0]):
            for i in range(self.measure_dim):
                v = flat_in[r, i]
                flat_out[r, i] = self._categories[i][int(v)] if np.isfinite(v) else None  # unoccupied: sentinel None
        return out

    def _decode(self, data):
        if isinstance(data, dict) and "measures" in data and data["measures"] is not None:
            data = dict(data)
            data["measures"] = self._decode_measures(data["measures"])
        return data

    # ---- index_of (:302-354) ---------------------------------------------------------------------------
    def index_of(self, measures):
        measures = np.asarray(measures, dtype=object)
        check_batch_shape(measures, "measures", self.measure_dim, "measure_dim")
        return self.grid_to_int_index(self._grid_indices(measures)) if len(measures) else np.empty(0, dtype=np.int32)

    def index_of_single(self, measures):
        measures = np.asarray(measures, dtype=object)
        check_shape(measures, "measures", self.measure_dim, "measure_dim")
        return self.index_of(measures[None])[0]

    # ---- writes ------------------------------------------------------------------------------------
    def add(self, solution, objective, measures, **fields):
        measures = np.asarray(measures, dtype=object)
        check_batch_shape(measures, "measures", self.measure_dim, "measure_dim")
        coded = self._encode(measures)
        with self._coded():
            return GridArchive.add(self, solution, objective, coded, **fields)

    def add_single(self, solution, objective, measures, **fields):
        measures = np.asarray(measures, dtype=object)
        check_shape(measures, "measures", self.measure_dim, "measure_dim")
        coded = self._encode(measures[None])
        with self._coded():  # (the base add_single would route through this class's add and encode twice)
            data = validate_single(self, {"solution": solution, "objective": objective, "measures": coded[0], **fields})
            info = GridArchive.add(self, np.asarray(data["solution"])[None], np.asarray(data["objective"])[None], coded,
                                   **{k: np.asarray(v)[None] for k, v in fields.items()})
        return {"status": info["status"][0], "value": info["value"][0]}

    # ---- reads -------------------------------------------------------------------------------------
    def retrieve(self, measures):
        measures = np.asarray(measures, dtype=object)
        check_batch_shape(measures, "measures", self.measure_dim, "measure_dim")
        occupied, data = self._store.retrieve(self.index_of(measures))
        fill_sentinel_values(occupied, data)
        return occupied, self._decode(data)

    def retrieve_single(self, measures):
        measures = np.asarray(measures, dtype=object)
        check_shape(measures, "measures", self.measure_dim, "measure_dim")
        occupied, data = self.retrieve(measures[None])
        return occupied[0], {f: a[0] for f, a in data.items()}

    def data(self, fields=None, return_type="dict"):
        if return_type == "pandas":
            raise NotImplementedError("return_type='pandas' is not available for CategoricalArchive on the GPU path")
        with self._coded():
            out = GridArchive.data(self, fields, return_type)
        if isinstance(fields, str):
            return self._decode_measures(out) if fields == "measures" else out
        if return_type == "tuple":
            names = list(self.field_list) if fields is None else list(fields)
            return tuple(self._decode_measures(a) if n == "measures" else a for n, a in zip(names, out))
        return self._decode(out)

    def sample_elites(self, n, replace=True):
        with self._coded():
            return self._decode(GridArchive.sample_elites(self, n, replace))

    @property
    def best_elite(self):
        with self._coded():
            best = GridArchive.best_elite.fget(self)
        if best is None:
            return None
        best = dict(best)
        best["measures"] = self._decode_measures(np.asarray(best["measures"])[None])[0]
        return best

    def __iter__(self):
        with self._coded():
            rows = list(GridArchive.__iter__(self))
        for row in rows:
            row = dict(row)
            row["measures"] = self._decode_measures(np.asarray(row["measures"])[None])[0]
            yield row

    def retessellate(self, new_dims):
        raise NotImplementedError("CategoricalArchive has no retessellate()")
