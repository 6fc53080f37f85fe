"""OpenAI-ES on the device (reference ribs/emitters/opt/_openai_es.py:25-208, AdamOpt _adam_opt.py:60-120;
SURVEY.md section 8(f) rank 4).

ask: theta + sigma0 * noise, with mirror sampling noise = [z; -z] -- one kernel, the ziggurat generator fused in
(qd_mirror_ask_rng / qd_sep_ask_rng with a unit scale), or qd_sep_ask when z is injected. tell: rank-normalised
weights on the host (batch_size numbers), the gradient estimate as a weighted sum of the noise rows on the device
(qd_weighted_rows), then the Adam step and the update ratio in one kernel (qd_openai_tell). Adam's state (theta, m,
v) lives in HBM; nothing returns to the host between ask and tell.
"""
from __future__ import annotations

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200.emitters.opt._es_base import DeviceES, stream_ptr


class OpenAIEvolutionStrategy(DeviceES):
    _zig_normals = True
    _fused_rng = True

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, lower_bounds=-np.inf,
                 upper_bounds=np.inf, mirror_sampling=True, lr=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8,
                 l2_coeff=0.0, device=None):
        if batch_size is None:  # _openai_es.py:61-63, :100-102: an even default for mirror sampling
            batch_size = 4 + int(3 * np.log(solution_dim))
            batch_size += batch_size % 2
        super().__init__(sigma0, solution_dim, batch_size, seed, dtype, lower_bounds, upper_bounds, device)
        if mirror_sampling and self._bounded:
            raise ValueError("Bounds are currently not supported when using mirror_sampling in OpenAI-ES; see "
                             "OpenAIEvolutionStrategy.ask() for more info.")
        if self.batch_size <= 1:
            raise ValueError("Batch size of 1 is not supported because rank normalization does not work with batch "
                             "size of 1.")
        if mirror_sampling and self.batch_size % 2 != 0:
            raise ValueError("If using mirror sampling, batch_size must be an even number.")
        self.mirror_sampling = bool(mirror_sampling)
        self._lr, self._beta1, self._beta2, self._eps, self._l2 = (float(lr), float(beta1), float(beta2),
                                                                   float(epsilon), float(l2_coeff))
        n = self.solution_dim
        f64 = dict(dtype=torch.float64, device=self._device)
        self._theta = torch.zeros(n, **f64)
        self._m = torch.zeros(n, **f64)
        self._v = torch.zeros(n, **f64)
        self._ones = torch.ones(n, **f64)
        self._grad = torch.empty(n, **f64)
        self._t = 0
        self._arange = None
        self.last_update_ratio = None

    def reset(self, x0):
        self._cancel_prefetch()
        with torch.cuda.stream(self._es_stream()):
            self._theta.copy_(torch.from_numpy(np.array(x0, dtype=np.float64)))
            self._m.zero_()
            self._v.zero_()
            self._status.zero_()
            self._status[0] = float(self.sigma0)
            self._status[4] = float("inf")
        self._t = 0
        self.last_update_ratio = np.inf

    # state views (NumPy copies)
    @property
    def mean(self):
        return self._to_host(self._theta)

    @property
    def theta(self):
        return self._to_host(self._theta)

    @property
    def noise(self):
        """The (batch_size, n) noise of the last ask (mirror sampling: [z; -z])."""
        if self._z is None:
            return None
        z = self._to_host(self._z)
        return np.concatenate((z, -z)) if self.mirror_sampling else z

    def _as_z(self, z, rows):  # injected normals: the (batch/2, n) half under mirror sampling
        if z is None or not self.mirror_sampling:
            return super()._as_z(z, rows)
        return super()._as_z(z, rows // 2)

    def _sample_rows(self, z, rows, row_ids):
        n, L = self.solution_dim, self._L
        X, st = self._solutions, self._status
        if self.mirror_sampling:
            half = rows // 2
            if z is None:
                zbuf = torch.empty((half, n), dtype=torch.float64, device=self._device)
                _lib.check(L.qd_mirror_ask_rng(half, n, self._theta.data_ptr(), self._ones.data_ptr(), st.data_ptr(),
                                               self._seed, self._rng_offset, X.data_ptr(), zbuf.data_ptr(), stream_ptr()))
                self._rng_offset += half * ((n + 3) // 4)
                self._z = zbuf
                return
            if self._arange is None or self._arange.numel() != half:
                self._arange = torch.arange(half, 2 * half, dtype=torch.int32, device=self._device)
            for zz, ids in ((z, None), (-z, self._arange)):  # theta + sigma0 * z and theta + sigma0 * (-z)
                _lib.check(L.qd_sep_ask(zz.data_ptr(), half, n, None if ids is None else ids.data_ptr(),
                                        self._theta.data_ptr(), self._ones.data_ptr(), st.data_ptr(), X.data_ptr(),
                                        stream_ptr()))
            return
        ids = None if row_ids is None else row_ids.data_ptr()
        if z is None:
            zbuf = torch.empty((rows, n), dtype=torch.float64, device=self._device)
            _lib.check(L.qd_sep_ask_rng(rows, n, ids, self._theta.data_ptr(), self._ones.data_ptr(), st.data_ptr(),
                                        self._seed, self._rng_offset, X.data_ptr(), zbuf.data_ptr(), stream_ptr()))
            self._rng_offset += rows * ((n + 3) // 4)
            if row_ids is None:
                self._z = zbuf
            else:
                self._z[row_ids.long()] = zbuf
            return
        _lib.check(L.qd_sep_ask(z.data_ptr(), rows, n, ids, self._theta.data_ptr(), self._ones.data_ptr(),
                                st.data_ptr(), X.data_ptr(), stream_ptr()))

    def tell(self, ranking_indices, ranking_values, num_parents):
        """_openai_es.py:180-208."""
        lam, n = self.batch_size, self.solution_dim
        idx = (ranking_indices.cpu().numpy() if isinstance(ranking_indices, torch.Tensor)
               else np.asarray(ranking_indices))
        ranks = np.empty(lam, dtype=np.int32)
        ranks[idx[::-1]] = np.arange(lam)  # ranks[i] = rank of noise[i], increasing
        ranks = (ranks / (lam - 1)) - 0.5
        if self.mirror_sampling:
            rows = lam // 2
            w = ranks[:rows] - ranks[rows:]
        else:
            rows, w = lam, ranks
        self._t += 1
        a = self._lr * np.sqrt(1 - self._beta2**self._t) / (1 - self._beta1**self._t)
        with torch.cuda.device(self._device), torch.cuda.stream(self._es_stream()):
            if self._rank_rows is None or self._rank_rows.numel() != rows:
                self._rank_rows = torch.arange(rows, dtype=torch.int64, device=self._device)
            w_d = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)).to(self._device)
            self._weighted(self._z, self._rank_rows, w_d, rows, None, self._grad, None)
            _lib.check(self._L.qd_openai_tell(
                n, self._grad.data_ptr(), float(rows * self.sigma0), self._theta.data_ptr(), self._m.data_ptr(),
                self._v.data_ptr(), float(a), self._beta1, self._beta2, self._eps, self._l2, self._status.data_ptr(),
                self._ws.data_ptr(), stream_ptr()))
        self.last_update_ratio = None  # read from the device status block on demand

    _rank_rows = None

    def check_stop(self, ranking_values):
        """_openai_es.py:125-139."""
        ratio = self.last_update_ratio
        if ratio is None:
            ratio = self.last_update_ratio = float(self._read_status()[4])
        if ratio < 1e-9:
            return True
        return bool(self._flat(ranking_values))
