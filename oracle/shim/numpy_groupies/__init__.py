"""Test-infrastructure shim for `numpy_groupies.aggregate_nb` (absent from this image).

The reference uses exactly three reductions (`_grid_archive.py:495,499,693`):
`len`, `sum`, `argmax`. Semantics restated from numpy_groupies' numba backend
(pyproject.toml:35 pins `>=0.9.16`, no lockfile): output length max(group_idx)+1, one
forward loop over the batch, `sum` accumulates in batch order in the input dtype,
`argmax` replaces only on strict `best < val` so the first occurrence wins, untouched
groups get `fill_value`.
"""
import numba as nb
import numpy as np


@nb.njit(cache=False)
def _len(group_idx, size):
    out = np.zeros(size, dtype=np.int64)
    for i in range(group_idx.shape[0]):
        out[group_idx[i]] += 1
    return out


@nb.njit(cache=False)
def _sum(group_idx, a, size, fill):
    out = np.zeros(size, dtype=a.dtype)
    seen = np.zeros(size, dtype=np.bool_)
    for i in range(group_idx.shape[0]):
        out[group_idx[i]] += a[i]
        seen[group_idx[i]] = True
    for j in range(size):
        if not seen[j]:
            out[j] = fill
    return out


@nb.njit(cache=False)
def _argmax(group_idx, a, size, fill):
    out = np.full(size, fill, dtype=np.int64)
    best = np.empty(size, dtype=a.dtype)
    seen = np.zeros(size, dtype=np.bool_)
    for i in range(group_idx.shape[0]):
        g = group_idx[i]
        if not seen[g]:
            seen[g] = True
            best[g] = a[i]
            out[g] = i
        elif best[g] < a[i]:
            best[g] = a[i]
            out[g] = i
    return out


def aggregate_nb(group_idx, a, func="sum", fill_value=0, **_kwargs):
    group_idx = np.asarray(group_idx)
    if group_idx.size == 0:
        raise ValueError("group_idx is empty")
    size = int(group_idx.max()) + 1
    if func == "len":
        out = _len(group_idx, size)
        if fill_value != 0:
            out = np.where(out == 0, fill_value, out)
        return out
    a = np.asarray(a)
    if a.ndim == 0:
        a = np.broadcast_to(a, group_idx.shape).copy()
    if func == "sum":
        return _sum(group_idx, a, size, fill_value)
    if func == "argmax":
        return _argmax(group_idx, a, size, fill_value)
    raise NotImplementedError(func)


aggregate = aggregate_nb
