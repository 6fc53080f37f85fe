"""GaussianEmitter and IsoLineEmitter on the device.

Reference: ribs/emitters/_gaussian_emitter.py:16-170, ribs/emitters/_iso_line_emitter.py:16-205, bounds
handling ribs/emitters/_emitter_base.py:47-123. Same constructors, `ask` / `tell`, properties and errors.
The parents never leave the GPU: `archive.sample_elite_cells` draws the cells with the archive's generator
exactly as `sample_elites` would, and one fused kernel (csrc/emit.cu qd_emit_variation) gathers the parents'
rows from the archive's solution field, adds the noise and clips to the bounds.

Randomness as for the evolution strategies: by default the unit normals come from the library's Philox
stream (seeded with `seed`); `ask(z=..., line_z=...)` accepts them explicitly (parity runs inject the normals
the reference's generator produced).
"""
from __future__ import annotations

import numbers

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import arr_readonly, check_batch_shape, check_shape, torch_dtype
from pyribs_b200.archives._device_store import current_stream_ptr
from pyribs_b200.emitters._evolution_strategy_emitter import _process_bounds


class _VariationEmitter:
    def __init__(self, archive, *, x0, initial_solutions, bounds, lower_bounds, upper_bounds, batch_size, seed):
        self._archive = archive
        self._solution_dim = archive.solution_dim
        sdt = archive.dtypes["solution"]
        if np.dtype(sdt) not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("solution dtype must be float32 or float64 on the GPU path")
        if bounds is not None and (lower_bounds is not None or upper_bounds is not None):
            raise ValueError(
                "Cannot specify both bounds and lower_bounds/upper_bounds; either specify bounds or specify "
                "lower_bounds/upper_bounds.")
        if bounds is not None:
            self._lower_bounds, self._upper_bounds = _process_bounds(bounds, self._solution_dim, sdt)
        else:
            self._lower_bounds = (np.full(self._solution_dim, -np.inf, dtype=sdt) if lower_bounds is None else
                                  np.asarray(lower_bounds, dtype=sdt))
            self._upper_bounds = (np.full(self._solution_dim, np.inf, dtype=sdt) if upper_bounds is None else
                                  np.asarray(upper_bounds, dtype=sdt))
            check_shape(self._lower_bounds, "lower_bounds", self._solution_dim, "solution_dim")
            check_shape(self._upper_bounds, "upper_bounds", self._solution_dim, "solution_dim")
        self._rng = np.random.default_rng(seed)
        self._batch_size = batch_size
        self._x0 = None
        self._initial_solutions = None
        self._noise_shape = ((batch_size, self._solution_dim) if isinstance(self._solution_dim, numbers.Integral)
                             else (batch_size, *self._solution_dim))
        if x0 is None and initial_solutions is None:
            raise ValueError("Either x0 or initial_solutions must be provided.")
        if x0 is not None and initial_solutions is not None:
            raise ValueError("x0 and initial_solutions cannot both be provided.")
        if x0 is not None:
            self._x0 = np.array(x0, dtype=sdt)
            check_shape(self._x0, "x0", archive.solution_dim, "archive.solution_dim")
        else:
            self._initial_solutions = np.asarray(initial_solutions, dtype=sdt)
            check_batch_shape(self._initial_solutions, "initial_solutions", archive.solution_dim,
                              "archive.solution_dim")
        # device side
        self._device = archive.device
        self._n = int(np.prod(self._noise_shape[1:]))
        ss = seed if isinstance(seed, np.random.SeedSequence) else np.random.SeedSequence(seed)
        self._seed = int(ss.generate_state(1, dtype=np.uint64)[0])
        self._rng_offset = 0
        self._dt_code = _lib.QD_F64 if np.dtype(sdt) == np.float64 else _lib.QD_F32
        td = torch_dtype(sdt)
        self._lo_d = torch.from_numpy(np.ascontiguousarray(self._lower_bounds).reshape(-1)).to(self._device)
        self._hi_d = torch.from_numpy(np.ascontiguousarray(self._upper_bounds).reshape(-1)).to(self._device)
        self._x0_d = None if self._x0 is None else torch.from_numpy(self._x0.reshape(-1).copy()).to(self._device)
        self._td = td

    # ---- properties (EmitterBase / the two emitters) ---------------------------------------------
    @property
    def archive(self):
        return self._archive

    @property
    def solution_dim(self):
        return self._solution_dim

    @property
    def lower_bounds(self):
        return self._lower_bounds

    @property
    def upper_bounds(self):
        return self._upper_bounds

    @property
    def x0(self):
        return self._x0

    @property
    def initial_solutions(self):
        return self._initial_solutions

    @property
    def batch_size(self):
        return self._batch_size

    def _clip(self, solutions):
        return np.clip(solutions, self.lower_bounds, self.upper_bounds)

    def tell(self, solution, objective, measures, add_info, **fields):
        """These emitters keep no state: EmitterBase.tell does nothing (_emitter_base.py:160-190)."""

    # ---- shared ask ---------------------------------------------------------------------------------
    def _normals(self, count):
        z = torch.empty(count, dtype=torch.float64, device=self._device)
        _lib.check(_lib.lib().qd_fill_normal(z.data_ptr(), count, self._seed, self._rng_offset, current_stream_ptr()))
        self._rng_offset += (count + 1) // 2
        return z

    def _as_normals(self, z, shape):
        count = int(np.prod(shape))
        if z is None:
            return self._normals(count)
        if not isinstance(z, torch.Tensor):
            z = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64))
        z = z.to(device=self._device, dtype=torch.float64).contiguous()
        assert z.numel() == count, f"injected normals must have {count} elements"
        return z

    def _vary(self, n_parents, sigma, line_sigma, z, line_z, return_device):
        B, n = self._batch_size, self._n
        arch = self._archive
        with torch.cuda.device(self._device):
            if arch.empty:
                pa = pb = None
            else:
                cells = arch.sample_elite_cells(n_parents * B)  # same draw as sample_elites (:160 / :184)
                cd = torch.from_numpy(cells).to(self._device)
                pa, pb = cd[:B], (cd[B:] if n_parents == 2 else None)
            zz = self._as_normals(z, (B, n))
            lz = self._as_normals(line_z, (B,)) if n_parents == 2 else None
            sig = torch.from_numpy(np.ascontiguousarray(sigma).reshape(-1)).to(self._device)
            X = torch.empty((B, n), dtype=self._td, device=self._device)
            arch._store.join_row_copy()  # a pipelined add() may still be writing the solution field
            sol = arch._store._fields["solution"]
            _lib.check(_lib.lib().qd_emit_variation(
                sol.data_ptr(), self._dt_code, n, B, None if pa is None else pa.data_ptr(),
                None if pb is None else pb.data_ptr(), None if self._x0_d is None else self._x0_d.data_ptr(),
                zz.data_ptr(), None if lz is None else lz.data_ptr(), sig.data_ptr(), int(sig.numel() > 1),
                float(line_sigma), self._lo_d.data_ptr(), self._hi_d.data_ptr(), X.data_ptr(), current_stream_ptr()))
        X = X.view(self._noise_shape)
        return X if return_device else X.cpu().numpy()


class GaussianEmitter(_VariationEmitter):
    """Adds Gaussian noise to elites drawn from the archive (ribs/emitters/_gaussian_emitter.py)."""

    def __init__(self, archive, *, sigma, x0=None, initial_solutions=None, bounds=None, lower_bounds=None,
                 upper_bounds=None, batch_size=64, seed=None):
        super().__init__(archive, x0=x0, initial_solutions=initial_solutions, bounds=bounds, lower_bounds=lower_bounds,
                         upper_bounds=upper_bounds, batch_size=batch_size, seed=seed)
        self._sigma = np.asarray(sigma, dtype=archive.dtypes["solution"])

    @property
    def sigma(self):
        return self._sigma

    def ask(self, z=None, return_device=False):
        """_gaussian_emitter.py:139-166."""
        if self.archive.empty and self._initial_solutions is not None:
            return self._clip(self._initial_solutions)
        return self._vary(1, self._sigma, 0.0, z, None, return_device)


class IsoLineEmitter(_VariationEmitter):
    """Iso+Line operator of Vassiliades and Mouret (ribs/emitters/_iso_line_emitter.py)."""

    def __init__(self, archive, *, iso_sigma=0.01, line_sigma=0.2, x0=None, initial_solutions=None, bounds=None,
                 lower_bounds=None, upper_bounds=None, batch_size=64, seed=None):
        super().__init__(archive, x0=x0, initial_solutions=initial_solutions, bounds=bounds, lower_bounds=lower_bounds,
                         upper_bounds=upper_bounds, batch_size=batch_size, seed=seed)
        sdt = archive.dtypes["solution"]
        self._iso_sigma = np.asarray(iso_sigma, dtype=sdt)
        self._line_sigma = np.asarray(line_sigma, dtype=sdt)

    @property
    def iso_sigma(self):
        return self._iso_sigma

    @property
    def line_sigma(self):
        return self._line_sigma

    def ask(self, z=None, line_z=None, return_device=False):
        """_iso_line_emitter.py:156-201."""
        if self.archive.empty and self._initial_solutions is not None:
            return self._clip(self._initial_solutions)
        return self._vary(2, self._iso_sigma, float(self._line_sigma), z, line_z, return_device)
