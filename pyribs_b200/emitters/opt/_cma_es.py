"""CMA-ES on the device (reference ribs/emitters/opt/_cma_es.py:25-328)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200.emitters.opt._es_base import DeviceES, es_weights, stream_ptr


class CMAEvolutionStrategy(DeviceES):
    """Full-covariance CMA-ES. Sampling (batch x n x n) and the rank-mu update (n x mu x n) are
    hand-written fp64 GEMMs with fused gathers / epilogues; the lazy eigendecomposition uses
    cuSOLVER through torch.linalg.eigh exactly where the reference calls LAPACK
    (_cma_es.py:70)."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, lower_bounds=-np.inf,
                 upper_bounds=np.inf, device=None):
        super().__init__(sigma0, solution_dim, batch_size, seed, dtype, lower_bounds, upper_bounds, device)
        n = self.solution_dim
        *_, c1, cmu = self._calc_strat_params(self.batch_size // 2)
        self.lazy_gap_evals = 0.5 * n * self.batch_size * (c1 + cmu) ** -1 / n**2  # _cma_es.py:130-139
        self.current_eval = None
        f64 = dict(dtype=torch.float64, device=self._device)
        self._mean = torch.zeros(n, **f64)
        self._pc = torch.zeros(n, **f64)
        self._ps = torch.zeros(n, **f64)
        self._C = torch.eye(n, **f64)
        self._Cnext = torch.empty((n, n), **f64)
        self._B = torch.eye(n, **f64)
        self._eig = torch.ones(n, **f64)
        self._T = torch.eye(n, **f64)
        self._invsqrt = torch.eye(n, **f64)
        self.condition_number = 1.0
        self.updated_eval = 0
        self._eig_max = 1.0

    def _calc_strat_params(self, mu):
        """_cma_es.py:239-259."""
        n = self.solution_dim
        w, mueff = es_weights(mu)
        cc = (4 + mueff / n) / (n + 4 + 2 * mueff / n)
        cs = (mueff + 2) / (n + mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + mueff)
        cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((n + 2) ** 2 + mueff))
        return w, mueff, cc, cs, c1, cmu

    def reset(self, x0):
        self._cancel_prefetch()
        with torch.cuda.stream(self._es_stream()):
            self._reset(x0)

    def _reset(self, x0):
        n = self.solution_dim
        self.current_eval = 0
        self._mean.copy_(torch.from_numpy(np.array(x0, dtype=np.float64)))
        self._pc.zero_()
        self._ps.zero_()
        eye = torch.eye(n, dtype=torch.float64, device=self._device)
        self._C.copy_(eye)
        self._B.copy_(eye)
        self._T.copy_(eye)
        self._invsqrt.copy_(eye)
        self._eig.fill_(1.0)
        self.condition_number = 1.0
        self.updated_eval = 0
        self._eig_max = 1.0
        self._status.zero_()
        self._status[0] = float(self.sigma0)
        self._status[2] = 1.0
        self._status[3] = 1.0

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

    @property
    def invsqrt(self):
        return self._to_host(self._invsqrt)

    @property
    def eigenvalues(self):
        return self._to_host(self._eig)

    def set_eigensystem(self, eigenvalues, eigenbasis):
        """Parity hook: installs (Lambda, B) computed elsewhere (eigenvector signs / the basis of
        degenerate subspaces differ between LAPACK and cuSOLVER, and the sample m + B sqrt(L) z
        depends on them)."""
        self._cancel_prefetch()
        with torch.cuda.stream(self._es_stream()):
            self._L.qd_sym_max(self.solution_dim, self._C.data_ptr(), stream_ptr())  # C = max(C, C^T), :67
            self._eig.copy_(torch.from_numpy(np.asarray(eigenvalues, dtype=np.float64)))
            self._B.copy_(torch.from_numpy(np.ascontiguousarray(eigenbasis, dtype=np.float64)))
            self._finish_eigensystem()
        self.updated_eval = self.current_eval

    def _finish_eigensystem(self):
        n = self.solution_dim
        _lib.check(self._L.qd_cma_transform(n, self._B.data_ptr(), self._eig.data_ptr(), self._T.data_ptr(),
                                            stream_ptr()))
        _lib.check(self._L.qd_cma_invsqrt(n, self._B.data_ptr(), self._eig.data_ptr(), self._invsqrt.data_ptr(),
                                          stream_ptr()))
        # eigenvalue extrema stay on the device (status[2], status[3]); check_stop reads them later
        self._status[2] = self._eig.max()
        self._status[3] = self._eig.min()

    def update_eigensystem(self, force=False):
        """DecompMatrix.update_eigensystem, _cma_es.py:51-82."""
        if not force and self.current_eval <= self.updated_eval + self.lazy_gap_evals:
            return False
        if self._pool is not None:
            raise RuntimeError("the eigensystem of a pooled strategy is refreshed by its pool")
        if force:
            self._cancel_prefetch()
            with torch.cuda.stream(self._es_stream()):
                return self._refresh_now()
        return self._refresh_now()

    def _refresh_now(self):
        n = self.solution_dim
        _lib.check(self._L.qd_sym_max(n, self._C.data_ptr(), stream_ptr()))
        if n <= 150:  # own Jacobi solver: no host synchronisation, so refreshes of different emitters overlap
            _lib.check(self._L.qd_sym_eigh_batched(n, 1, self._C.data_ptr(), self._eig.data_ptr(), self._B.data_ptr(),
                                                   None, stream_ptr()))
            self._eig.abs_()
        else:
            lam, B = torch.linalg.eigh(self._C)
            self._eig.copy_(lam.abs())
            self._B.copy_(B)
        self._finish_eigensystem()
        self.updated_eval = self.current_eval
        return True

    def _prepare_ask(self, lam):
        self.update_eigensystem()

    def _sample_rows(self, z, rows, row_ids):
        _lib.check(self._L.qd_cma_ask(
            z.data_ptr(), rows, self.solution_dim, None if row_ids is None else row_ids.data_ptr(),
            self._mean.data_ptr(), self._T.data_ptr(), self._status.data_ptr(), self._solutions.data_ptr(),
            stream_ptr()))

    def tell(self, ranking_indices, ranking_values, num_parents):
        """_cma_es.py:274-328."""
        if self._pool is not None:
            return self._pool.tell_many([self._slot], [ranking_indices], [num_parents])
        n = self.solution_dim
        self.current_eval += len(ranking_indices)
        num_parents = int(num_parents)
        if num_parents == 0:
            return
        w, mueff, cc, cs, c1, cmu = self._calc_strat_params(num_parents)
        damps = 1 + 2 * max(0, np.sqrt((mueff - 1) / (n + 1)) - 1) + cs
        with torch.cuda.device(self._device), torch.cuda.stream(self._es_stream()):
            rank_d = self._rank_to_device(ranking_indices, num_parents)
            w_d = torch.from_numpy(w).to(self._device)
            self._weighted(self._solutions, rank_d, w_d, num_parents, None, self._new_mean, None)
            sc = self._scalars(cc=cc, cs=cs, c1=c1, cmu=cmu, mueff=mueff, damps=damps, wsum=np.sum(w),
                               hsig_exp=2 * self.current_eval / self.batch_size)
            _lib.check(self._L.qd_cma_tell(
                n, self._solutions.data_ptr(), rank_d.data_ptr(), w_d.data_ptr(), num_parents,
                self._new_mean.data_ptr(), self._invsqrt.data_ptr(), self._mean.data_ptr(), self._pc.data_ptr(),
                self._ps.data_ptr(), self._C.data_ptr(), self._Cnext.data_ptr(), self._status.data_ptr(), C.byref(sc),
                self._ws.data_ptr(), stream_ptr()))
            self._C, self._Cnext = self._Cnext, self._C

    def check_stop(self, ranking_values):
        """_cma_es.py:161-179."""
        st = self._read_status()
        self.condition_number = st[2] / st[3]
        if self.condition_number > 1e14:
            return True
        if st[0] * np.sqrt(st[2]) < 1e-11:
            return True
        return bool(self._flat(ranking_values))
