"""CPU oracle for the pyribs hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy/SciPy restatement of the reference algorithm for batched archive insertion
(`GridArchive` / `CVTArchive` `index_of` + `add`) and the CMA-family ES ask/tell. Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import this module; `pyribs_b200/` never does (the product path fails loudly
when the CUDA library is missing).

Parity status: PINNED. `oracle/gen_golden.py` imports the unmodified reference from
`/root/reference` (through `oracle/shim/`) and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` replays every golden through this file and demands
bit-equality (indices, statuses, winners, occupied_list order) or <=1e-12 relative
(floating state). When `/root/reference` is present the same test module also checks
this file against the live reference on fresh random inputs.

Each function cites the reference lines it follows (paths relative to
`/root/reference`).
"""

from __future__ import annotations

import numpy as np

try:  # numba is a reference dependency (ribs/emitters/opt/_cma_es.py:8); keep optional.
    import numba as nb

    _HAVE_NUMBA = True
except Exception:  # pragma: no cover
    _HAVE_NUMBA = False

try:
    from scipy.spatial import KDTree
except Exception:  # pragma: no cover
    KDTree = None


# --------------------------------------------------------------------------------------
# Group reductions: numpy_groupies.aggregate_nb(len / sum / argmax)
# (call sites ribs/archives/_grid_archive.py:495,499,693). Third-party dependency
# `numpy_groupies>=0.9.16` (pyproject.toml:35, no lockfile) is absent from the image;
# its published numba-backend semantics are restated here: one forward pass in batch
# order, strict `<` for argmax (first occurrence wins), input-dtype accumulation.
# --------------------------------------------------------------------------------------
def _group_pass_py(cells, obj, size):
    cnt = np.zeros(size, dtype=np.int64)
    tot = np.zeros(size, dtype=obj.dtype)
    win = np.full(size, -1, dtype=np.int64)
    best = np.zeros(size, dtype=obj.dtype)
    for i in range(cells.shape[0]):
        c = cells[i]
        v = obj[i]
        if cnt[c] == 0:
            best[c] = v
            win[c] = i
        elif best[c] < v:
            best[c] = v
            win[c] = i
        cnt[c] += 1
        tot[c] += v
    return cnt, tot, win


_group_pass = nb.njit(cache=False)(_group_pass_py) if _HAVE_NUMBA else _group_pass_py


def group_reduce(cells: np.ndarray, obj: np.ndarray):
    """(count, sum, argmax-first) per cell; arrays of length max(cells)+1."""
    size = int(cells.max()) + 1
    return _group_pass(np.ascontiguousarray(cells), np.ascontiguousarray(obj), size)


# --------------------------------------------------------------------------------------
# A2: GridArchive.index_of   (ribs/archives/_grid_archive.py:362-410, 432-450)
# --------------------------------------------------------------------------------------
def grid_index_of(measures, dims, lower_bounds, interval_size, epsilon):
    """Affine map -> C truncation to int32 -> clip -> row-major ravel.

    `dims` is int32, so `dims * (m - lb)` promotes to float64 even for float32 archives
    while `(m - lb)` itself is evaluated in the measures dtype (_grid_archive.py:401-404).
    """
    measures = np.asarray(measures)
    dims = np.asarray(dims, dtype=np.int32)
    scaled = (dims * (measures - lower_bounds) + epsilon) / interval_size
    with np.errstate(invalid="ignore"):
        g = scaled.astype(np.int32)  # out-of-range doubles become INT_MIN on x86
    g = np.clip(g, 0, dims - 1)
    return np.ravel_multi_index(g.T, dims).astype(np.int32)


# --------------------------------------------------------------------------------------
# A3: CVTArchive.index_of   (ribs/archives/_cvt_archive.py:579-644)
# --------------------------------------------------------------------------------------
class CentroidIndex:
    """Nearest-centroid lookup. `method` mirrors `nearest_neighbors=`."""

    def __init__(self, centroids, method="scipy_kd_tree", workers=1, chunk_size=None):
        self.centroids = np.asarray(centroids)
        self.method = method
        self.workers = workers
        self.chunk_size = chunk_size
        if method == "scipy_kd_tree":  # _cvt_archive.py:411-417
            self.tree = KDTree(self.centroids)
        elif method != "brute_force":
            raise ValueError(method)

    def index_of(self, measures):
        measures = np.asarray(measures)
        if self.method == "scipy_kd_tree":  # _cvt_archive.py:605-609
            _, idx = self.tree.query(measures, workers=self.workers)
            return idx.astype(np.int32)
        # brute force, _cvt_archive.py:610-632: sum((x-c)^2) then first-min argmin.
        chunk = self.chunk_size or max(1, min(len(measures), (1 << 24) // max(1, len(self.centroids))))
        out = np.empty(len(measures), dtype=np.int32)
        for s in range(0, len(measures), chunk):
            diff = measures[s : s + chunk, None, :] - self.centroids[None, :, :]
            out[s : s + chunk] = np.argmin(np.sum(np.square(diff), axis=2), axis=1)
        return out


# --------------------------------------------------------------------------------------
# A6/A7: ArrayStore   (ribs/archives/_array_store.py:129-162, 330-485, 535-608)
# --------------------------------------------------------------------------------------
class OracleStore:
    def __init__(self, field_desc, capacity):
        self.capacity = int(capacity)
        self.occupied = np.zeros(self.capacity, dtype=bool)
        self.occupied_list = np.empty(self.capacity, dtype=np.int32)
        self.n_occupied = 0
        self.updates = [0, 0]
        self.fields = {}
        for name, (shape, dtype) in field_desc.items():
            shape = (shape,) if np.isscalar(shape) else tuple(shape)
            self.fields[name] = np.empty((self.capacity, *shape), dtype=dtype)

    def retrieve(self, indices):
        """Full-field gather, as the reference does (_array_store.py:414-460)."""
        indices = np.asarray(indices, dtype=np.int32)
        occ = self.occupied[indices]
        data = {name: arr[indices] for name, arr in self.fields.items()}
        data["index"] = indices.copy()
        return occ, data

    def add(self, indices, data):
        """_array_store.py:560-608: ascending append of new cells, then scatter."""
        self.updates[0] += 1
        if len(indices) == 0:
            return
        mask = np.zeros(self.capacity, dtype=bool)
        mask[indices] = True
        uniq = np.nonzero(mask)[0]
        new = uniq[~self.occupied[uniq]]
        self.occupied[new] = True
        self.occupied_list[self.n_occupied : self.n_occupied + len(new)] = new
        self.n_occupied += len(new)
        for name, arr in self.fields.items():
            arr[indices] = np.asarray(data[name], dtype=arr.dtype)

    def clear(self):
        self.updates[1] += 1
        self.n_occupied = 0
        self.occupied[:] = False


# --------------------------------------------------------------------------------------
# A4: CMA-MAE batch threshold rule   (ribs/archives/_grid_archive.py:473-516)
# --------------------------------------------------------------------------------------
def compute_thresholds(cells, objective, cur_threshold, learning_rate, dtype):
    if len(cells) == 0:
        return np.empty(0, dtype=dtype)
    cnt, tot, _ = group_reduce(cells, objective)
    k = cnt[cells]
    s = tot[cells]
    ratio = np.asarray(1.0 - learning_rate, dtype=dtype) ** k
    return ratio * cur_threshold + (s / k) * (1 - ratio)


# --------------------------------------------------------------------------------------
# A5 + A8: the body of `add`   (ribs/archives/_grid_archive.py:606-714)
# --------------------------------------------------------------------------------------
class OracleArchive:
    """Archive state + `add` / `add_single`; `index_fn` supplies A2 or A3."""

    def __init__(
        self,
        *,
        solution_dim,
        measure_dim,
        cells,
        index_fn,
        learning_rate=None,
        threshold_min=-np.inf,
        qd_score_offset=0.0,
        dtype=np.float64,
        extra_fields=None,
        seed=None,
    ):
        sdim = (solution_dim,) if np.isscalar(solution_dim) else tuple(solution_dim)
        self.dtype = np.dtype(dtype)
        desc = {
            "solution": (sdim, dtype),
            "objective": ((), dtype),
            "measures": ((measure_dim,), dtype),
            "threshold": ((), dtype),
            **(extra_fields or {}),
        }
        self.store = OracleStore(desc, cells)
        self.cells = int(cells)
        self.index_fn = index_fn
        # ribs/archives/_utils.py:76-93
        if threshold_min != -np.inf and learning_rate is None:
            raise ValueError("threshold_min was set without setting learning_rate.")
        if learning_rate is None:
            learning_rate = 1.0
        if threshold_min == -np.inf and learning_rate != 1.0:
            raise ValueError("threshold_min can only be -np.inf if learning_rate is 1.0")
        self.learning_rate = np.asarray(learning_rate, dtype=dtype)
        self.threshold_min = np.asarray(threshold_min, dtype=dtype)
        self.qd_score_offset = np.asarray(qd_score_offset, dtype=dtype)
        self.rng = np.random.default_rng(seed)
        self.objective_sum = np.asarray(0.0, dtype=dtype)
        self.obj_max = None
        self.best_elite = None

    def __len__(self):
        return self.store.n_occupied

    # _grid_archive.py:325-360
    def _stats_update(self, new_sum, best_index):
        _, row = self.store.retrieve([best_index])
        row = {k: v[0] for k, v in row.items()}
        if self.obj_max is None or row["objective"] > self.obj_max:
            self.best_elite = row
            self.obj_max = row["objective"]
        self.objective_sum = new_sum

    def stats(self):
        n = len(self)
        qd = self.objective_sum - np.asarray(n, dtype=self.dtype) * self.qd_score_offset
        return {
            "num_elites": n,
            "coverage": np.asarray(n / self.cells, dtype=self.dtype),
            "qd_score": qd,
            "norm_qd_score": np.asarray(qd / self.cells, dtype=self.dtype),
            "obj_max": self.obj_max,
            "obj_mean": None if n == 0 else np.asarray(self.objective_sum / n, dtype=self.dtype),
        }

    def add(self, solution, objective, measures, **extras):
        data = {
            "solution": np.asarray(solution, dtype=self.store.fields["solution"].dtype),
            "objective": np.asarray(objective, dtype=self.dtype),
            "measures": np.asarray(measures, dtype=self.store.fields["measures"].dtype),
            **{k: np.asarray(v) for k, v in extras.items()},
        }
        if not np.all(np.isfinite(data["objective"])):
            raise ValueError("All elements of objective must be finite")
        if not np.all(np.isfinite(data["measures"])):
            raise ValueError("All elements of measures must be finite")
        cells = self.index_fn(data["measures"])
        B = len(cells)
        # :628-630 gather + threshold_min for empty cells
        occ, cur = self.store.retrieve(cells)
        thr = cur["threshold"]
        thr[~occ] = self.threshold_min
        # :636-649 status / value
        can = data["objective"] > thr
        is_new = can & ~occ
        status = np.zeros(B, dtype=np.int32)
        status[is_new] = 2
        status[can & occ] = 1
        thr[is_new] = 0.0 if self.threshold_min == -np.inf else self.threshold_min
        value = data["objective"] - thr
        info = {"status": status, "value": value}
        if not np.any(can):  # :653-654
            return info
        # :658-660 compaction of every field
        cells_c = cells[can]
        data_c = {k: v[can] for k, v in data.items()}
        thr_c = thr[can]
        # :663-677
        if self.threshold_min == -np.inf:
            new_thr = data_c["objective"]
        else:
            new_thr = compute_thresholds(cells_c, data_c["objective"], thr_c, self.learning_rate, self.dtype)
        # :693-701 per-cell winner (first max), ascending-cell order
        _, _, win = group_reduce(cells_c, data_c["objective"])
        win = win[win != -1]
        cells_w = cells_c[win]
        data_w = {k: v[win] for k, v in data_c.items()}
        data_w["threshold"] = new_thr[win]
        self.store.add(cells_w, data_w)
        # :707-712 stats
        old_obj = cur["objective"]
        old_obj[~occ] = 0.0
        old_obj = old_obj[can][win]
        new_sum = self.objective_sum + np.sum(data_w["objective"] - old_obj)
        best = cells_w[np.argmax(data_w["objective"])]
        self._stats_update(new_sum, best)
        return info

    def add_single(self, solution, objective, measures, **extras):
        """Scalar rule, _grid_archive.py:716-848: t' = t(1-lr) + f lr."""
        sol = np.asarray(solution, dtype=self.store.fields["solution"].dtype)
        obj = np.asarray(objective, dtype=self.dtype)
        meas = np.asarray(measures, dtype=self.store.fields["measures"].dtype)
        if not np.isfinite(obj) or not np.all(np.isfinite(meas)):
            raise ValueError("non-finite")
        cell = self.index_fn(meas[None])[0]
        occ, cur = self.store.retrieve([cell])
        occ = occ[0]
        if occ:
            t = cur["threshold"][0]
        else:
            t = np.asarray(0.0, self.dtype) if self.threshold_min == -np.inf else self.threshold_min
        status = np.int32(0)
        is_new = (not occ) and self.threshold_min < obj
        improve = occ and t < obj
        if is_new or improve:
            status = np.int32(1) if improve else np.int32(2)
            row = {"solution": sol[None], "objective": obj[None], "measures": meas[None]}
            row.update({k: np.asarray(v)[None] for k, v in extras.items()})
            row["threshold"] = np.asarray(t * (1.0 - self.learning_rate) + obj * self.learning_rate)[None]
            self.store.add(np.asarray([cell], dtype=np.int32), row)
            old = cur["objective"][0] if occ else np.asarray(0.0, dtype=self.dtype)
            self._stats_update(self.objective_sum + obj - old, cell)
        return {"status": status, "value": obj - t}

    def sample_elites(self, n):  # _grid_archive.py:912-924
        if len(self) == 0:
            raise IndexError("No elements in archive.")
        r = self.rng.choice(len(self), size=n, replace=True)
        sel = self.store.occupied_list[r]
        return self.store.retrieve(sel)[1]


def make_grid_oracle(*, solution_dim, dims, ranges, epsilon=1e-6, dtype=np.float64, **kw):
    """GridArchive constructor arithmetic, _grid_archive.py:122-170."""
    dims = np.array(dims, dtype=np.int32)
    lo = np.array([r[0] for r in ranges], dtype=dtype)
    hi = np.array([r[1] for r in ranges], dtype=dtype)
    interval = hi - lo
    eps = np.asarray(epsilon, dtype=dtype)
    return OracleArchive(
        solution_dim=solution_dim,
        measure_dim=len(dims),
        cells=int(np.prod(dims)),
        index_fn=lambda m: grid_index_of(m, dims, lo, interval, eps),
        dtype=dtype,
        **kw,
    )


def make_cvt_oracle(*, solution_dim, centroids, method="scipy_kd_tree", workers=1, dtype=np.float64, **kw):
    centroids = np.asarray(centroids, dtype=dtype)
    index = CentroidIndex(centroids, method=method, workers=workers)
    return OracleArchive(
        solution_dim=solution_dim,
        measure_dim=centroids.shape[1],
        cells=len(centroids),
        index_fn=index.index_of,
        dtype=dtype,
        **kw,
    )


# --------------------------------------------------------------------------------------
# A15: rankers   (ribs/emitters/rankers.py:136-398)
# --------------------------------------------------------------------------------------
def rank_one_key(values):
    """`np.flip(np.argsort(v))` (rankers.py:150-152, :354)."""
    return np.flip(np.argsort(values)), values


def rank_two_stage(status, values):
    """`np.flip(np.lexsort((values, status)))` (rankers.py:183-203)."""
    rv = np.stack((status, values), axis=-1)
    return np.flip(np.lexsort(np.flip(rv, axis=-1).T)), rv


# --------------------------------------------------------------------------------------
# ES common   (weights: ribs/emitters/opt/_cma_es.py:239-247)
# --------------------------------------------------------------------------------------
def es_weights(mu):
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w = w / np.sum(w)
    mueff = np.sum(w) ** 2 / np.sum(w**2)
    return w, mueff


def _flat_fitness(rv):
    return len(rv) >= 2 and np.linalg.norm(rv[0] - rv[-1]) < 1e-12


class OracleCMAES:
    """A12: full-covariance CMA-ES (ribs/emitters/opt/_cma_es.py:85-328)."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64,
                 lower_bounds=-np.inf, upper_bounds=np.inf):
        n = solution_dim
        self.n = n
        self.batch_size = 4 + int(3 * np.log(n)) if batch_size is None else batch_size
        self.sigma0 = sigma0
        self.dtype = dtype
        self.lower_bounds = np.asarray(lower_bounds, dtype=dtype)
        self.upper_bounds = np.asarray(upper_bounds, dtype=dtype)
        self.rng = np.random.default_rng(seed)
        *_, c1, cmu = self._params(self.batch_size // 2)
        self.lazy_gap_evals = 0.5 * n * self.batch_size * (c1 + cmu) ** -1 / n**2  # :130-139

    def _params(self, mu):  # :239-259
        n = self.n
        w, mueff = es_weights(mu)
        cc = (4 + mueff / n) / (n + 4 + 2 * mueff / n)
        cs = (mueff + 2) / (n + mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + mueff)
        cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((n + 2) ** 2 + mueff))
        return w, mueff, cc, cs, c1, cmu

    def reset(self, x0):  # :149-159
        n, dt = self.n, self.dtype
        self.current_eval = 0
        self.sigma = self.sigma0
        self.mean = np.array(x0, dt)
        self.pc = np.zeros(n, dtype=dt)
        self.ps = np.zeros(n, dtype=dt)
        self.C = np.eye(n, dtype=dt)
        self.B = np.eye(n, dtype=dt)
        self.eigvals = np.ones(n, dtype=dt)
        self.cond = 1
        self.invsqrt = np.eye(n, dtype=dt)
        self.updated_eval = 0

    def _refresh(self):  # DecompMatrix.update_eigensystem :51-82
        if self.current_eval <= self.updated_eval + self.lazy_gap_evals:
            return False
        self.C = np.maximum(self.C, self.C.T)
        lam, B = np.linalg.eigh(self.C)
        self.eigvals = np.abs(lam).real.astype(self.dtype)
        self.B = B.real.astype(self.dtype)
        self.cond = np.max(self.eigvals) / np.min(self.eigvals)
        inv = (self.B * (1 / np.sqrt(self.eigvals))) @ self.B.T
        self.invsqrt = np.maximum(inv, inv.T)
        self.updated_eval = self.current_eval
        return True

    def ask(self, z=None):
        """:199-237. `z` (lambda x n, unit normal) may be injected; it is scaled by sigma
        exactly as `rng.normal(0, sigma)` scales its unit draws."""
        lam = self.batch_size
        self._refresh()
        T = self.B * np.sqrt(self.eigvals)
        sol = np.empty((lam, self.n), dtype=self.dtype)
        remaining = np.arange(lam)
        first = True
        while len(remaining) > 0:
            if z is not None and first:
                u = (0.0 + self.sigma * np.asarray(z)).astype(self.dtype)
            else:
                u = self.rng.normal(0.0, self.sigma, (len(remaining), self.n)).astype(self.dtype)
            first = False
            new = (T @ u.T).T + self.mean[None]
            oob = np.logical_or(new < self.lower_bounds, new > self.upper_bounds)
            sol[remaining] = new
            remaining = remaining[np.any(oob, axis=1)]
        self.solutions = sol
        return sol

    def tell(self, ranking_indices, num_parents):  # :274-328
        n = self.n
        self.current_eval += len(self.solutions[ranking_indices])
        if num_parents == 0:
            return
        parents = self.solutions[ranking_indices][:num_parents]
        w, mueff, cc, cs, c1, cmu = self._params(num_parents)
        damps = 1 + 2 * max(0, np.sqrt((mueff - 1) / (n + 1)) - 1) + cs
        old = self.mean
        self.mean = np.sum(parents * w[:, None], axis=0)
        y = self.mean - old
        zz = self.invsqrt @ y
        self.ps = (1 - cs) * self.ps + (np.sqrt(cs * (2 - cs) * mueff) / self.sigma) * zz
        left = np.sum(np.square(self.ps)) / n / (1 - (1 - cs) ** (2 * self.current_eval / self.batch_size))
        hsig = 1 if left < 2 + 4.0 / (n + 1) else 0
        self.pc = (1 - cc) * self.pc + hsig * np.sqrt(cc * (2 - cc) * mueff) * y
        ys = parents - old[None]
        rank_mu = np.einsum("ki,kj", ys * w[:, None], ys)
        c1a = c1 * (1 - (1 - hsig**2) * cc * (2 - cc))
        # :261-270 -- rank-one term carries c1 twice.
        self.C = self.C * (1 - c1a - cmu * np.sum(w)) + (c1 * np.outer(self.pc, self.pc)) * c1 \
            + rank_mu * cmu / (self.sigma**2)
        ssq = np.sum(np.square(self.ps))
        self.sigma *= np.exp(min(1, (cs / damps) * (ssq / n - 1) / 2))

    def check_stop(self, rv):  # :161-179
        if self.cond > 1e14:
            return True
        if self.sigma * np.sqrt(np.max(self.eigvals)) < 1e-11:
            return True
        return bool(_flat_fitness(rv))


class OracleSepCMAES:
    """A13: separable CMA-ES (ribs/emitters/opt/_sep_cma_es.py:55-311)."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64,
                 lower_bounds=-np.inf, upper_bounds=np.inf):
        self.n = solution_dim
        self.batch_size = 4 + int(3 * np.log(self.n)) if batch_size is None else batch_size
        self.sigma0 = sigma0
        self.dtype = dtype
        self.lower_bounds = np.asarray(lower_bounds, dtype=dtype)
        self.upper_bounds = np.asarray(upper_bounds, dtype=dtype)
        self.rng = np.random.default_rng(seed)

    def _params(self, mu):  # :203-231
        n = self.n
        w, mueff = es_weights(mu)
        cc = (1 + 1 / n + mueff / n) / (n**0.5 + 1 / n + 2 * mueff / n)
        cs = (mueff + 2) / (n + mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + mueff)
        c1s = c1 * (1.0 / (n + 2.0 * np.sqrt(n) + float(mueff) / n))
        cmus = min(1 - c1s, (0 + mueff + 1.0 / mueff - 2) / (n + 4 * np.sqrt(n) + mueff / 2.0))
        return w, mueff, cc, cs, c1s, cmus

    def reset(self, x0):
        self.current_eval = 0
        self.sigma = self.sigma0
        self.mean = np.array(x0, self.dtype)
        self.pc = np.zeros(self.n, dtype=self.dtype)
        self.ps = np.zeros(self.n, dtype=self.dtype)
        self.C = np.ones(self.n, dtype=self.dtype)

    def ask(self, z=None):  # :155-191
        lam = self.batch_size
        tv = np.sqrt(self.C)
        sol = np.empty((lam, self.n), dtype=self.dtype)
        remaining = np.arange(lam)
        first = True
        while len(remaining) > 0:
            if z is not None and first:
                u = (0.0 + self.sigma * np.asarray(z)).astype(self.dtype)
            else:
                u = self.rng.normal(0.0, self.sigma, (len(remaining), self.n)).astype(self.dtype)
            first = False
            new = tv[None] * u + self.mean[None]
            oob = np.logical_or(new < self.lower_bounds, new > self.upper_bounds)
            sol[remaining] = new
            remaining = remaining[np.any(oob, axis=1)]
        self.solutions = sol
        return sol

    def tell(self, ranking_indices, num_parents):  # :257-311
        n = self.n
        self.current_eval += len(self.solutions[ranking_indices])
        if num_parents == 0:
            return
        parents = self.solutions[ranking_indices][:num_parents]
        w, mueff, cc, cs, c1, cmu = self._params(num_parents)
        damps = 1 + 2 * max(0, np.sqrt((mueff - 1) / (n + 1)) - 1) + cs
        old = self.mean
        self.mean = np.sum(parents * w[:, None], axis=0)
        y = self.mean - old
        zz = (1 / np.sqrt(self.C)) * y
        self.ps = (1 - cs) * self.ps + (np.sqrt(cs * (2 - cs) * mueff) / self.sigma) * zz
        left = np.sum(np.square(self.ps)) / n / (1 - (1 - cs) ** (2 * self.current_eval / self.batch_size))
        hsig = 1 if left < 2 + 4.0 / (n + 1) else 0
        self.pc = (1 - cc) * self.pc + hsig * np.sqrt(cc * (2 - cc) * mueff) * y
        ys = parents - old[None]
        rank_mu = np.sum((ys * w[:, None]) * ys, axis=0)
        c1a = c1 * (1 - (1 - hsig**2) * cc * (2 - cc))
        self.C = self.C * (1 - c1a - cmu * np.sum(w)) + (c1 * self.pc**2) * c1 + rank_mu * cmu / (self.sigma**2)
        ssq = np.sum(np.square(self.ps))
        self.sigma *= np.exp(min(1, (cs / damps) * (ssq / n - 1) / 2))

    def check_stop(self, rv):  # :121-138
        if np.max(self.C) / np.min(self.C) > 1e14:
            return True
        if self.sigma * np.sqrt(max(self.C)) < 1e-11:
            return True
        return bool(_flat_fitness(rv))


class OracleLMMAES:
    """A14: LM-MA-ES (ribs/emitters/opt/_lm_ma_es.py:24-239)."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64,
                 lower_bounds=-np.inf, upper_bounds=np.inf, n_vectors=None):
        n = solution_dim
        self.n = n
        self.batch_size = 4 + int(3 * np.log(n)) if batch_size is None else batch_size
        if self.batch_size > n:
            raise ValueError("batch_size is greater than solution_dim")
        self.sigma0 = sigma0
        self.dtype = dtype
        self.lower_bounds = np.asarray(lower_bounds, dtype=dtype)
        self.upper_bounds = np.asarray(upper_bounds, dtype=dtype)
        self.rng = np.random.default_rng(seed)
        self.n_vectors = self.batch_size if n_vectors is None else n_vectors
        self.csigma = 2 * self.batch_size / n  # :84
        with np.errstate(over="ignore"):
            self.cd = 1 / (1.5 ** np.arange(self.n_vectors) * n)  # :86
            self.cc = self.batch_size / (4.0 ** np.arange(self.n_vectors) * n)  # :88-90

    def reset(self, x0):  # :103-110
        self.current_gens = 0
        self.sigma = self.sigma0
        self.mean = np.array(x0, self.dtype)
        self.ps = np.zeros(self.n, dtype=self.dtype)
        self.m = np.zeros((self.n_vectors, self.n))

    def ask(self, z=None):  # :126-194
        lam = self.batch_size
        sol = np.empty((lam, self.n), dtype=self.dtype)
        zs = np.empty((lam, self.n), dtype=self.dtype)
        remaining = np.arange(lam)
        first = True
        itrs = min(self.current_gens, self.n_vectors)
        while len(remaining) > 0:
            if z is not None and first:
                zz = np.asarray(z, dtype=np.float64)
            else:
                zz = self.rng.standard_normal((len(remaining), self.n))
            first = False
            zs[remaining] = zz
            d = zz
            for j in range(itrs):
                d = (1 - self.cd[j]) * d + self.cd[j] * self.m[j][None] * (d @ self.m[j])[:, None]
            new = self.mean[None] + self.sigma * d
            oob = np.logical_or(new < self.lower_bounds, new > self.upper_bounds)
            sol[remaining] = new
            remaining = remaining[np.any(oob, axis=1)]
        self.solutions = sol
        self.solution_z = zs
        return sol

    def tell(self, ranking_indices, num_parents):  # :209-239
        self.current_gens += 1
        if num_parents == 0:
            return
        w, mueff = es_weights(num_parents)
        parents = self.solutions[ranking_indices][:num_parents]
        zp = self.solution_z[ranking_indices][:num_parents]
        zbar = np.sum(w[:, None] * zp, axis=0)
        self.mean = np.sum(w[:, None] * parents, axis=0)
        cs = self.csigma
        self.ps = (1 - cs) * self.ps + np.sqrt(mueff * cs * (2 - cs)) * zbar
        cc = self.cc[:, None]
        self.m = (1 - cc) * self.m + np.sqrt(mueff * cc * (2 - cc)) * zbar[None]
        self.sigma *= np.exp(cs / 2 * (np.sum(self.ps**2) / self.n - 1))

    def check_stop(self, rv):  # :112-124
        if self.sigma < 1e-12:
            return True
        return bool(_flat_fitness(rv))


class OracleOpenAIES:
    """ribs/emitters/opt/_openai_es.py:49-208 with AdamOpt (ribs/emitters/opt/_adam_opt.py:60-120) inlined; the
    unit normals are injectable (`ask(z)`: the (batch/2, n) half under mirror sampling, else (batch, n))."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, mirror_sampling=True,
                 lr=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, l2_coeff=0.0):
        self.n = solution_dim
        self.batch_size = 4 + int(3 * np.log(solution_dim)) if batch_size is None else batch_size
        if batch_size is None and self.batch_size % 2 != 0:  # :100-102
            self.batch_size += 1
        if self.batch_size <= 1:
            raise ValueError("Batch size of 1 is not supported")
        if mirror_sampling and self.batch_size % 2 != 0:
            raise ValueError("If using mirror sampling, batch_size must be an even number.")
        self.sigma0, self.dtype, self.mirror = sigma0, dtype, mirror_sampling
        self.lr, self.beta1, self.beta2, self.eps, self.l2 = lr, beta1, beta2, epsilon, l2_coeff
        self.rng = np.random.default_rng(seed)

    def reset(self, x0):  # :120-123, AdamOpt.reset
        self.theta = np.asarray(x0, dtype=np.float64).copy()
        self.m = np.zeros_like(self.theta)
        self.v = np.zeros_like(self.theta)
        self.t = 0
        self.last_update_ratio = np.inf
        self.noise = None

    @property
    def mean(self):
        return self.theta

    def ask(self, z=None):  # :141-178 (unbounded)
        lam = self.batch_size
        if self.mirror:
            half = self.rng.standard_normal((lam // 2, self.n)) if z is None else np.asarray(z, dtype=np.float64)
            self.noise = np.concatenate((half, -half))
        else:
            self.noise = self.rng.standard_normal((lam, self.n)) if z is None else np.asarray(z, dtype=np.float64)
        self.solutions = self.theta[None] + self.sigma0 * self.noise
        return self.solutions

    def tell(self, ranking_indices, num_parents=None):  # :180-208
        lam = self.batch_size
        ranks = np.empty(lam, dtype=np.int32)
        ranks[np.asarray(ranking_indices)[::-1]] = np.arange(lam)
        ranks = (ranks / (lam - 1)) - 0.5
        if self.mirror:
            half = lam // 2
            gradient = np.sum(self.noise[:half] * (ranks[:half] - ranks[half:])[:, None], axis=0)
            gradient /= half * self.sigma0
        else:
            gradient = np.sum(self.noise * ranks[:, None], axis=0)
            gradient /= lam * self.sigma0
        prev = self.theta.copy()
        g = -gradient  # AdamOpt.step
        g += self.l2 * self.theta
        self.t += 1
        a = self.lr * np.sqrt(1 - self.beta2**self.t) / (1 - self.beta1**self.t)
        self.m = self.beta1 * self.m + (1 - self.beta1) * g
        self.v = self.beta2 * self.v + (1 - self.beta2) * (g * g)
        self.theta = self.theta + (-a * self.m / (np.sqrt(self.v) + self.eps))
        self.last_update_ratio = np.linalg.norm(self.theta - prev) / np.linalg.norm(self.theta)

    def check_stop(self, rv):  # :125-139
        if self.last_update_ratio < 1e-9:
            return True
        return bool(_flat_fitness(rv))


ORACLE_ES = {"cma_es": OracleCMAES, "sep_cma_es": OracleSepCMAES, "lm_ma_es": OracleLMMAES,
             "openai_es": OracleOpenAIES}


# ---------------------------------------------------------------------------------------
# MAP-Elites variation emitters (next rows of the scope table)
# ---------------------------------------------------------------------------------------
class OracleGaussianEmitter:
    """ribs/emitters/_gaussian_emitter.py:139-166 with the unit normals injectable:
    rng.normal(scale=sigma, size) == sigma * rng.standard_normal(size) on the same stream."""

    def __init__(self, archive, sigma, x0, batch_size, lower=-np.inf, upper=np.inf, seed=None, dtype=np.float64):
        self.archive, self.dtype = archive, dtype
        self.sigma = np.asarray(sigma, dtype=dtype)
        self.x0 = np.array(x0, dtype=dtype)
        self.batch_size = batch_size
        self.lower, self.upper = lower, upper
        self.rng = np.random.default_rng(seed)

    def ask(self, z=None):
        B = self.batch_size
        if len(self.archive) == 0:
            parents = np.repeat(self.x0[None], repeats=B, axis=0)  # :158
        else:
            parents = self.archive.sample_elites(B)["solution"]  # :160
        shape = (B, len(self.x0))
        z = self.rng.standard_normal(shape) if z is None else np.asarray(z, dtype=np.float64).reshape(shape)
        noise = (self.sigma.astype(np.float64) * z).astype(self.dtype)  # :162-165
        return np.clip(parents + noise, self.lower, self.upper)  # :166


class OracleIsoLineEmitter:
    """ribs/emitters/_iso_line_emitter.py:156-201."""

    def __init__(self, archive, iso_sigma, line_sigma, x0, batch_size, lower=-np.inf, upper=np.inf, seed=None,
                 dtype=np.float64):
        self.archive, self.dtype = archive, dtype
        self.iso_sigma = np.asarray(iso_sigma, dtype=dtype)
        self.line_sigma = np.asarray(line_sigma, dtype=dtype)
        self.x0 = np.array(x0, dtype=dtype)
        self.batch_size = batch_size
        self.lower, self.upper = lower, upper
        self.rng = np.random.default_rng(seed)

    def ask(self, z=None, line_z=None):
        B, n = self.batch_size, len(self.x0)
        if len(self.archive) == 0:
            parents = np.repeat(self.x0[None], repeats=2 * B, axis=0)  # :182
        else:
            parents = self.archive.sample_elites(2 * B)["solution"]  # :184
        parents = parents.reshape(2, B, n)  # :186
        elites, directions = parents[0], parents[1] - parents[0]  # :187-188
        z = self.rng.standard_normal((B, n)) if z is None else np.asarray(z, dtype=np.float64).reshape(B, n)
        iso = (float(self.iso_sigma) * z).astype(self.dtype)  # :190-193
        lz = self.rng.standard_normal((B, 1)) if line_z is None else np.asarray(line_z, dtype=np.float64).reshape(B, 1)
        line = (float(self.line_sigma) * lz).astype(self.dtype)  # :194-198
        return np.clip(elites + iso + line * directions, self.lower, self.upper)  # :200-201
