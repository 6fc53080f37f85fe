"""Host-side arithmetic of the owner-computes (cell-sharded) store: which cells a rank owns and how the ranks'
answers combine into the archive-wide one. Pure NumPy + torch.distributed collectives on whatever backend the
process group has (NCCL in production, gloo in tests/test_shard_store_cpu.py) -- no CUDA in here.

Reference semantics being reassembled: ArchiveBase stats (`ribs/archives/_grid_archive.py:325-360`), `data()` /
`retrieve()` of `ArrayStore` (`ribs/archives/_array_store.py:330-485`)."""
from __future__ import annotations

import numpy as np


def shard_cell_range(cells: int, world: int, rank: int):
    """Contiguous cell range [lo, hi) owned by `rank` (equal shares, the last ranks may hold less)."""
    per = -(-int(cells) // int(world))
    lo = min(rank * per, int(cells))
    return lo, min(lo + per, int(cells))


def merge_shard_stats(rows, odt):
    """rows[r] = (elites, objective sum, best objective or -inf, has-best) of rank r, in rank order. Returns
    (elites, objective sum accumulated in rank = cell order in the objective dtype, best objective or None,
    owner = lowest rank holding the best objective or -1)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 4)
    osum = np.asarray(0.0, dtype=odt)
    for v in rows[:, 1]:
        osum = np.asarray(osum + np.asarray(v, dtype=odt), dtype=odt)
    has = rows[:, 3] > 0
    if not has.any():
        return int(rows[:, 0].sum()), float(osum), None, -1
    best = np.where(has, rows[:, 2], -np.inf)
    return int(rows[:, 0].sum()), float(osum), float(best.max()), int(np.argmax(best))


def stitch_shard_data(parts, want=None):
    """parts[r] = rank r's `data()` dict with archive-wide cell indices; the shards one after the other."""
    keys = want or list(parts[0])
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in keys}


def stitch_shard_retrieve(cells, parts):
    """parts[r] = (mask of the cells rank r owns, occupied & mask, rank r's retrieved rows): every row comes from
    the rank that owns its cell."""
    cells = np.asarray(cells)
    occupied = np.zeros(len(cells), dtype=bool)
    out = {k: np.array(v, copy=True) for k, v in parts[0][2].items()}
    for mine, occ, data in parts:
        occupied[mine] = occ[mine]
        for k in out:
            out[k][mine] = data[k][mine]
    out["index"] = cells.astype(np.int32)
    return occupied, out


def all_gather_rows(mine, world, group=None, device=None):
    """COLLECTIVE: every rank contributes four doubles, every rank gets the (world, 4) table."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(x) for x in mine], dtype=torch.float64, device=device)
    out = torch.empty(4 * world, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.view(world, 4).cpu().numpy()
