"""sep-CMA-ES on the device (reference ribs/emitters/opt/_sep_cma_es.py:55-311)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200.emitters.opt._es_base import DeviceES, es_weights, stream_ptr


class SeparableCMAEvolutionStrategy(DeviceES):
    """Diagonal-covariance CMA-ES; every step is an HBM-bound stream over (batch, n)."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, lower_bounds=-np.inf,
                 upper_bounds=np.inf, device=None):
        super().__init__(sigma0, solution_dim, batch_size, seed, dtype, lower_bounds, upper_bounds, device)
        self.current_eval = None
        n = self.solution_dim
        f64 = dict(dtype=torch.float64, device=self._device)
        self._mean = torch.zeros(n, **f64)
        self._pc = torch.zeros(n, **f64)
        self._ps = torch.zeros(n, **f64)
        self._C = torch.ones(n, **f64)
        self._rank_mu = torch.empty(n, **f64)

    def _calc_strat_params(self, mu):
        """_sep_cma_es.py:203-231."""
        n = self.solution_dim
        w, mueff, _ = self._weights_device(mu)
        cc = (1 + 1 / n + mueff / n) / (n**0.5 + 1 / n + 2 * mueff / n)
        cs = (mueff + 2) / (n + mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + mueff)
        c1_sep = c1 * (1.0 / (n + 2.0 * np.sqrt(n) + float(mueff) / n))
        cmu_sep = min(1 - c1_sep, (0 + mueff + 1.0 / mueff - 2) / (n + 4 * np.sqrt(n) + mueff / 2.0))
        return w, mueff, cc, cs, c1_sep, cmu_sep

    def reset(self, x0):
        self._cancel_prefetch()
        with torch.cuda.stream(self._es_stream()):
            self._reset(x0)

    def _reset(self, x0):
        self.current_eval = 0
        self._mean.copy_(torch.from_numpy(np.array(x0, dtype=np.float64)))
        self._pc.zero_()
        self._ps.zero_()
        self._C.fill_(1.0)
        self._status.zero_()
        self._status[0] = float(self.sigma0)
        self._status[2] = 1.0
        self._status[3] = 1.0

    # state views (NumPy copies) used by tests and check_stop
    @property
    def mean(self):
        return self._to_host(self._mean)

    @property
    def pc(self):
        return self._to_host(self._pc)

    @property
    def ps(self):
        return self._to_host(self._ps)

    @property
    def cov(self):
        return self._to_host(self._C)

    _zig_normals = True
    _fused_rng = True  # tell re-reads the solutions only (:281-303), so z never has to exist in HBM

    def _sample_rows(self, z, rows, row_ids):
        if z is None:  # generator fused into the transform: X is the only HBM traffic of the ask
            n = self.solution_dim
            _lib.check(self._L.qd_sep_ask_rng(
                rows, n, None if row_ids is None else row_ids.data_ptr(), self._mean.data_ptr(), self._C.data_ptr(),
                self._status.data_ptr(), self._seed, self._rng_offset, self._solutions.data_ptr(), None, stream_ptr()))
            self._rng_offset += rows * ((n + 3) // 4)
            return
        _lib.check(self._L.qd_sep_ask(
            z.data_ptr(), rows, self.solution_dim, None if row_ids is None else row_ids.data_ptr(),
            self._mean.data_ptr(), self._C.data_ptr(), self._status.data_ptr(), self._solutions.data_ptr(),
            stream_ptr()))

    def tell(self, ranking_indices, ranking_values, num_parents):
        """_sep_cma_es.py:257-311."""
        n = self.solution_dim
        self.current_eval += len(ranking_indices)
        num_parents = int(num_parents)
        if num_parents == 0:
            return
        w, mueff, cc, cs, c1, cmu = self._calc_strat_params(num_parents)
        damps = 1 + 2 * max(0, np.sqrt((mueff - 1) / (n + 1)) - 1) + cs
        with torch.cuda.device(self._device), torch.cuda.stream(self._es_stream()):
            rank_d = self._rank_to_device(ranking_indices, num_parents)
            w_d = self._weights_device(num_parents)[2]
            self._weighted(self._solutions, rank_d, w_d, num_parents, self._mean, self._new_mean, self._rank_mu)
            sc = self._scalars(cc=cc, cs=cs, c1=c1, cmu=cmu, mueff=mueff, damps=damps, wsum=np.sum(w),
                               hsig_exp=2 * self.current_eval / self.batch_size)
            _lib.check(self._L.qd_sep_tell(
                n, self._new_mean.data_ptr(), self._rank_mu.data_ptr(), self._mean.data_ptr(), self._pc.data_ptr(),
                self._ps.data_ptr(), self._C.data_ptr(), self._status.data_ptr(), C.byref(sc), self._ws.data_ptr(),
                stream_ptr()))

    def check_stop(self, ranking_values):
        """_sep_cma_es.py:121-138 (condition number, sigma * sqrt(max C), flat fitness)."""
        st = self._read_status()
        sigma, cmax, cmin = st[0], st[2], st[3]
        if cmax / cmin > 1e14:
            return True
        if sigma * np.sqrt(cmax) < 1e-11:
            return True
        return bool(self._flat(ranking_values))
