"""Shared plumbing of the device-resident evolution strategies.

API contract: ribs/emitters/opt/_evolution_strategy_base.py:25-118 -- constructor kwargs
`(sigma0, solution_dim, batch_size, seed, dtype, lower_bounds, upper_bounds, **es_kwargs)`,
attributes `batch_size, sigma0, solution_dim, dtype`, methods `reset / ask / tell /
check_stop`.

Randomness: the reference draws from NumPy's PCG64, which a GPU cannot replay. Two modes:
  * default -- unit normals come from the library's Philox4x32-10 kernel (counter-based, so
    a run is reproducible for a given seed);
  * parity  -- `ask(z=...)` accepts the (batch, n) unit normals explicitly (tests inject the
    same z into the CPU oracle).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import arr_readonly

BOUNDS_SAMPLING_THRESHOLD = 100
BOUNDS_WARNING = (
    "During bounds handling, this ES resampled at least "
    f"{BOUNDS_SAMPLING_THRESHOLD} times. This may indicate that your solution space bounds are too tight."
)


def stream_ptr():
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def es_weights(mu):
    """ribs/emitters/opt/_cma_es.py:241-246."""
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w = w / np.sum(w)
    mueff = np.sum(w) ** 2 / np.sum(w**2)
    return w, mueff


class DeviceES:
    """Base: device buffers, RNG, bounds resampling loop, weighted parent reduction."""

    def __init__(self, sigma0, solution_dim, batch_size=None, seed=None, dtype=np.float64, lower_bounds=-np.inf,
                 upper_bounds=np.inf, device=None):
        if not torch.cuda.is_available():
            raise _lib.QdError("pyribs_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = _lib.lib()
        self.batch_size = 4 + int(3 * np.log(solution_dim)) if batch_size is None else int(batch_size)
        self.sigma0 = sigma0
        self.solution_dim = int(solution_dim)
        self.dtype = dtype
        self.lower_bounds = np.asarray(lower_bounds, dtype=dtype)
        self.upper_bounds = np.asarray(upper_bounds, dtype=dtype)
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ss = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
        self._seed = int(ss.generate_state(1, dtype=np.uint64)[0])
        self._rng_offset = 0
        n = self.solution_dim
        self._bounded = bool(np.any(np.isfinite(self.lower_bounds)) or np.any(np.isfinite(self.upper_bounds)))
        if self._bounded:
            lo = np.broadcast_to(self.lower_bounds.astype(np.float64), (n,)).copy()
            hi = np.broadcast_to(self.upper_bounds.astype(np.float64), (n,)).copy()
            self._lo_d = torch.from_numpy(lo).to(self._device)
            self._hi_d = torch.from_numpy(hi).to(self._device)
        f64 = dict(dtype=torch.float64, device=self._device)
        self._status = torch.zeros(8, **f64)
        self._ws = torch.zeros(int(self._L.qd_es_workspace_doubles(n)), **f64)
        self._new_mean = torch.empty(n, **f64)
        self._partial = None
        self._solutions = None
        self._z = None
        self._stream = None
        self._ask_event = None
        self._host_buf = None
        self._host_bufs = [None, None]  # two page-locked buffers alternate: the array ask() returned stays intact
        self._host_turn = 0             # until the ask after the next one
        self._pending = None
        self._host_ready = False
        self._rng_mark = 0
        self._pool = None  # set when a CMAPool adopts this strategy (emitters/_emitter_pool.py)
        self._slot = -1

    # ---- helpers ------------------------------------------------------------------------------
    # Strategies whose ask is an HBM stream over (batch, n) draw from the ziggurat generator (one Philox
    # counter per four normals); the GEMM-bound CMA-ES keeps the Box-Muller stream its pooled form shares.
    _zig_normals = False
    # ... and those that do not need z afterwards sample inside the ask kernel (`_sample_rows(None, ...)`).
    _fused_rng = False

    def _normals(self, rows):
        z = torch.empty((rows, self.solution_dim), dtype=torch.float64, device=self._device)
        cnt = z.numel()
        if self._zig_normals:
            _lib.check(self._L.qd_fill_normal_zig(z.data_ptr(), cnt, self._seed, self._rng_offset, stream_ptr()))
            self._rng_offset += (cnt + 3) // 4
        else:
            _lib.check(self._L.qd_fill_normal(z.data_ptr(), cnt, self._seed, self._rng_offset, stream_ptr()))
            self._rng_offset += (cnt + 1) // 2
        return z

    def _as_z(self, z, rows):
        if z is None:
            return self._normals(rows)
        if not isinstance(z, torch.Tensor):
            z = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64))
        z = z.to(device=self._device, dtype=torch.float64).contiguous()
        assert tuple(z.shape) == (rows, self.solution_dim), "z must be (rows, solution_dim)"
        return z

    def _sample_rows(self, z, rows, row_ids):
        """Writes rows of self._solutions from unit normals `z`; implemented by subclasses."""
        raise NotImplementedError

    # ---- ask: enqueue (own stream) + wait -------------------------------------------------------
    def _es_stream(self):
        """Every ES owns a CUDA stream: the kernels of different emitters (in particular the
        single-CTA eigensolver of a covariance refresh) overlap on the GPU."""
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self._device)
            self._ask_event = torch.cuda.Event()
        # Entered from the caller's stream, the ES stream first gets behind it: the state tensors were initialised
        # there, and rankings / injected normals handed in as CUDA tensors are produced there. (Inside a
        # `with torch.cuda.stream(es_stream)` block the two are the same stream and nothing is recorded.)
        cur = torch.cuda.current_stream(self._device)
        if cur != self._stream:
            self._stream.wait_stream(cur)
        return self._stream

    def _enqueue_host_copy(self):
        lam, n = self._solutions.shape
        with torch.cuda.device(self._device), torch.cuda.stream(self._es_stream()):
            self._host_turn ^= 1
            buf = self._host_bufs[self._host_turn]
            if buf is None or tuple(buf.shape) != (lam, n):
                buf = self._host_bufs[self._host_turn] = torch.empty((lam, n), dtype=torch.float64).pin_memory()
            self._host_buf = buf
            self._host_buf.copy_(self._solutions, non_blocking=True)
            self._ask_event.record(self._stream)
        self._host_ready = True

    def _enqueue_ask(self, lam, z, want_host=True):
        """Enqueues sampling of `lam` solutions (+ optionally the device->host copy) on the ES stream."""
        n = self.solution_dim
        stream = self._es_stream()
        with torch.cuda.device(self._device), torch.cuda.stream(stream):
            self._solutions = torch.empty((lam, n), dtype=torch.float64, device=self._device)
            self._prepare_ask(lam)
            zz = None if (z is None and self._fused_rng) else self._as_z(z, lam)
            self._z = zz
            self._sample_rows(zz, lam, None)
            if self._bounded:  # resampling loop of the reference (needs the host to see the violations)
                remaining = torch.arange(lam, dtype=torch.int32, device=self._device)
                itrs = 0
                while True:
                    oob = torch.empty(remaining.numel(), dtype=torch.int32, device=self._device)
                    _lib.check(self._L.qd_es_out_of_bounds(
                        self._solutions.data_ptr(), remaining.numel(), n, remaining.data_ptr(), self._lo_d.data_ptr(),
                        self._hi_d.data_ptr(), oob.data_ptr(), stream_ptr()))
                    remaining = remaining[oob.bool()].contiguous()
                    itrs += 1
                    if remaining.numel() == 0:
                        break
                    if itrs > BOUNDS_SAMPLING_THRESHOLD:
                        warnings.warn(BOUNDS_WARNING, stacklevel=3)
                    znew = None if self._fused_rng else self._normals(remaining.numel())
                    if self._z is not None and znew is not None:
                        self._z[remaining.long()] = znew
                    self._sample_rows(znew, remaining.numel(), remaining)
            self._ask_event.record(stream)
        self._host_ready = False
        if want_host:
            self._enqueue_host_copy()
        self._pending = lam

    def prefetch_ask(self):
        """Starts the next default `ask()` right away (called by the emitter at the end of `tell`), so
        that by the time the scheduler asks, the samples of all emitters have been produced
        concurrently. Identical results: `ask` is a pure function of the ES state and the RNG
        counter, and nothing changes either between this call and the next `ask`."""
        if self._pending is None and not self._bounded and self._pool is None:
            self._rng_mark = self._rng_offset
            # small batches also start their device->host copy now; big ones wait to see what ask() wants
            self._enqueue_ask(self.batch_size, None, want_host=self.batch_size * self.solution_dim * 8 <= (8 << 20))

    def _cancel_prefetch(self):
        if self._pending is not None:
            self._ask_event.synchronize()
            self._rng_offset = self._rng_mark
            self._pending = None

    def ask(self, batch_size=None, z=None, return_device=False):
        """Samples `batch_size` solutions (reference ask loops, e.g. _cma_es.py:199-237).

        Returns a read-only NumPy array by default (drop-in); `return_device=True` returns the
        CUDA tensor that `tell` will re-read, without a device->host copy.
        """
        lam = self.batch_size if batch_size is None else int(batch_size)
        if self._pool is not None:
            if z is not None or lam != self.batch_size:
                raise ValueError("a pooled strategy samples batch_size solutions from its own stream")
            return self._pool.ask_one(self._slot, return_device)
        if self._pending is not None and (z is not None or lam != self._pending):
            self._cancel_prefetch()
        if self._pending is None:
            self._enqueue_ask(lam, z, want_host=not return_device)
        self._pending = None
        if return_device:
            torch.cuda.current_stream().wait_event(self._ask_event)
            return self._solutions
        if not self._host_ready:
            self._enqueue_host_copy()
        self._ask_event.synchronize()
        # Like the reference (`arr_readonly(self._solutions)`, _cma_es.py:237) the result is a read-only view of a
        # buffer the strategy owns, valid until it is sampled into again (here: the second ask from now) -- copying
        # 328 MB out of the page-locked buffer cost 4x the device->host transfer at configs[3].
        out = self._host_buf.numpy()
        if np.dtype(self.dtype) != np.float64:
            out = out.astype(self.dtype)
        return arr_readonly(out, view=True)

    def _prepare_ask(self, lam):
        pass

    # ---- checkpoint / resume -------------------------------------------------------------------------
    _ES_TRANSIENT = ("_L", "_stream", "_ask_event", "_host_buf", "_host_bufs", "_partial", "_w_cache", "_pending",
                     "_host_ready")

    def __getstate__(self):
        if self._pool is not None:
            raise TypeError("a strategy adopted by a pooled Scheduler cannot be pickled on its own; pickle the scheduler")
        self._cancel_prefetch()
        if self._stream is not None:
            self._stream.synchronize()
        d = {}
        for k, v in self.__dict__.items():
            if k in self._ES_TRANSIENT:
                continue
            d[k] = ("__cuda__", v.cpu().numpy()) if isinstance(v, torch.Tensor) and v.is_cuda else v
        return d

    def __setstate__(self, d):
        self._L = _lib.lib()
        dev = d["_device"]
        for k, v in d.items():
            if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], str) and v[0] == "__cuda__":
                v = torch.from_numpy(v[1]).to(dev)
            self.__dict__[k] = v
        self._stream = self._ask_event = self._host_buf = self._partial = self._pending = None
        self._host_bufs = [None, None]
        self._host_turn = 0
        self._host_ready = False

    def _weighted(self, X, rank_d, w_d, mu, old_mean, out_mean, out_sq):
        n = self.solution_dim
        need = int(self._L.qd_wsum_partial_bytes(mu, n)) // 8
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=torch.float64, device=self._device)
        _lib.check(self._L.qd_weighted_rows(
            X.data_ptr(), n, rank_d.data_ptr(), w_d.data_ptr(), mu, None if old_mean is None else old_mean.data_ptr(),
            None if out_mean is None else out_mean.data_ptr(), None if out_sq is None else out_sq.data_ptr(),
            self._partial.data_ptr(), stream_ptr()))

    def _weights_device(self, mu):
        """(w, mueff, w on the device) of `es_weights(mu)`, cached: the parent count rarely changes between
        generations, and a pageable host->device copy per tell costs more than the kernels it feeds."""
        hit = self.__dict__.get("_w_cache")
        if hit is None or hit[0] != mu:
            w, mueff = es_weights(mu)
            hit = (mu, w, mueff, torch.from_numpy(w).to(self._device))
            self._w_cache = hit
        return hit[1], hit[2], hit[3]

    def _rank_to_device(self, ranking_indices, num_parents):
        if isinstance(ranking_indices, torch.Tensor):
            r = ranking_indices[:num_parents].to(device=self._device, dtype=torch.int64).contiguous()
        else:
            r = torch.from_numpy(np.array(np.asarray(ranking_indices)[:num_parents], dtype=np.int64, copy=True)).to(
                self._device)
        return r

    def _to_host(self, t):
        with torch.cuda.stream(self._es_stream()):
            return t.cpu().numpy()

    def _read_status(self):
        """Device status block -> NumPy (ordered after everything enqueued on the ES stream)."""
        with torch.cuda.stream(self._es_stream()):
            return self._status.cpu().numpy()

    @property
    def sigma(self):
        return float(self._read_status()[0])

    @staticmethod
    def _flat(ranking_values):
        rv = np.asarray(ranking_values)
        return len(rv) >= 2 and np.linalg.norm(rv[0] - rv[-1]) < 1e-12

    def _scalars(self, **kw):
        sc = _lib.CmaScalars()
        for k, v in kw.items():
            setattr(sc, k, float(v))
        return sc
