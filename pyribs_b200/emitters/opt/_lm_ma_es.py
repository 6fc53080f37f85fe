"""LM-MA-ES on the device (reference ribs/emitters/opt/_lm_ma_es.py:24-239)."""
from __future__ import annotations

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200.emitters.opt._es_base import DeviceES, es_weights, stream_ptr


class LMMAEvolutionStrategy(DeviceES):
    """Limited-memory matrix adaptation ES: `n_vectors` direction vectors replace the covariance."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, lower_bounds=-np.inf,
                 upper_bounds=np.inf, n_vectors=None, device=None):
        super().__init__(sigma0, solution_dim, batch_size, seed, dtype, lower_bounds, upper_bounds, device)
        n = self.solution_dim
        if self.batch_size > n:
            raise ValueError(f"batch_size ({self.batch_size}) is greater than solution_dim ({n})")
        self.n_vectors = self.batch_size if n_vectors is None else int(n_vectors)
        self.csigma = 2 * self.batch_size / n  # _lm_ma_es.py:84
        with np.errstate(over="ignore"):
            self.cd = 1 / (1.5 ** np.arange(self.n_vectors) * n)  # :86
            self.cc = self.batch_size / (4.0 ** np.arange(self.n_vectors) * n)  # :88-90
        f64 = dict(dtype=torch.float64, device=self._device)
        self._cd_d = torch.from_numpy(self.cd).to(self._device)
        self._cc_d = torch.from_numpy(self.cc).to(self._device)
        self._mean = torch.zeros(n, **f64)
        self._ps = torch.zeros(n, **f64)
        self._m = torch.zeros((self.n_vectors, n), **f64)
        self._z_mean = torch.empty(n, **f64)
        self._ask_ws = None
        self.current_gens = None

    def reset(self, x0):
        self._cancel_prefetch()
        with torch.cuda.stream(self._es_stream()):
            self._reset(x0)

    def _reset(self, x0):
        self.current_gens = 0
        self._mean.copy_(torch.from_numpy(np.array(x0, dtype=np.float64)))
        self._ps.zero_()
        self._m.zero_()
        self._status.zero_()
        self._status[0] = float(self.sigma0)

    @property
    def mean(self):
        return self._to_host(self._mean)

    @property
    def ps(self):
        return self._to_host(self._ps)

    @property
    def m(self):
        return self._to_host(self._m)

    _zig_normals = True
    _fused_rng = True  # the ask entry point fills z itself (tell averages the parents' z, :224-226)

    def _sample_rows(self, z, rows, row_ids):
        n = self.solution_dim
        itrs = min(self.current_gens, self.n_vectors)
        need = int(self._L.qd_lmma_ask_workspace_doubles(rows, n, itrs))
        if self._ask_ws is None or self._ask_ws.numel() < need:
            self._ask_ws = torch.empty(need, dtype=torch.float64, device=self._device)
        ids = None if row_ids is None else row_ids.data_ptr()
        if z is not None:
            _lib.check(self._L.qd_lmma_ask(
                z.data_ptr(), rows, n, ids, self._mean.data_ptr(), self._status.data_ptr(), self._m.data_ptr(),
                self._cd_d.data_ptr(), itrs, self._solutions.data_ptr(), self._ask_ws.data_ptr(), stream_ptr()))
            return
        # unit normals generated block by block inside the call (the block's z is still in L2 when the product and
        # the output pass read it); the stream is the one `_normals(rows)` would have produced
        znew = torch.empty((rows, n), dtype=torch.float64, device=self._device)
        _lib.check(self._L.qd_lmma_ask_rng(
            rows, n, ids, self._mean.data_ptr(), self._status.data_ptr(), self._m.data_ptr(), self._cd_d.data_ptr(),
            itrs, self._seed, self._rng_offset, znew.data_ptr(), self._solutions.data_ptr(), self._ask_ws.data_ptr(),
            stream_ptr()))
        self._rng_offset += (rows * n + 3) // 4
        if row_ids is None:
            self._z = znew
        else:
            self._z[row_ids.long()] = znew

    def tell(self, ranking_indices, ranking_values, num_parents):
        """_lm_ma_es.py:209-239."""
        n = self.solution_dim
        self.current_gens += 1
        num_parents = int(num_parents)
        if num_parents == 0:
            return
        w, mueff, w_d = self._weights_device(num_parents)
        with torch.cuda.device(self._device), torch.cuda.stream(self._es_stream()):
            rank_d = self._rank_to_device(ranking_indices, num_parents)
            self._weighted(self._z, rank_d, w_d, num_parents, None, self._z_mean, None)
            self._weighted(self._solutions, rank_d, w_d, num_parents, None, self._new_mean, None)
            _lib.check(self._L.qd_lmma_tell(
                n, self.n_vectors, self._new_mean.data_ptr(), self._z_mean.data_ptr(), self._mean.data_ptr(),
                self._ps.data_ptr(), self._m.data_ptr(), self._cc_d.data_ptr(), float(self.csigma), float(mueff),
                self._status.data_ptr(), self._ws.data_ptr(), stream_ptr()))

    def check_stop(self, ranking_values):
        """_lm_ma_es.py:112-124."""
        if self.sigma < 1e-12:
            return True
        return bool(self._flat(ranking_values))
