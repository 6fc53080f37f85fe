"""CVTArchive on the B200 (API of ribs/archives/_cvt_archive.py:137-1119)."""
from __future__ import annotations

import numbers
import os
from pathlib import Path

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import check_batch_shape
from pyribs_b200.archives._archive_base import DeviceArchive
from pyribs_b200.archives._device_store import current_stream_ptr

_DEPRECATED = {
    "cells": "`cells` is deprecated in pyribs 0.9.0. Please pass in `centroids` instead.",
    "custom_centroids": "`custom_centroids` is deprecated in pyribs 0.9.0. Please pass in `centroids` instead.",
    "centroid_method": "`centroid_method` is deprecated in pyribs 0.9.0. Please generate centroids and pass them in instead.",
    "use_kd_tree": "`use_kd_tree` is deprecated in pyribs 0.9.0. Please use `nearest_neighbors` instead.",
    "ckdtree_kwargs": "`ckdtree_kwargs` is deprecated in pyribs 0.9.0. Please use `kdtree_kwargs` instead.",
}


def merge_sharded_argmin(idx, d_local, group, mask_losers):
    """Combines per-rank (exact distance, global index) results of a centroid-sharded search.

    all_reduce(MIN) over the distances, every rank masks its index unless it holds the global
    minimum (`mask_losers(d_local, d_min, idx)` sets losers to INT32_MAX), all_reduce(MIN) over
    the indices: the lowest index among exact ties wins, as np.argmin over the unsharded
    centroid array would. Works on any backend (NCCL on the GPU path,
    gloo in the CPU tests)."""
    import torch.distributed as dist

    d_min = d_local.clone()
    dist.all_reduce(d_min, op=dist.ReduceOp.MIN, group=group)
    mask_losers(d_local, d_min, idx)
    dist.all_reduce(idx, op=dist.ReduceOp.MIN, group=group)
    return idx, d_min


def shard_bounds(n_centroids, world, rank):
    """Contiguous centroid range [lo, hi) owned by `rank`."""
    per = -(-n_centroids // world)
    lo = min(rank * per, n_centroids)
    return lo, min(lo + per, n_centroids)


class CVTArchive(DeviceArchive):
    """Centroidal-Voronoi archive: a solution belongs to the cell of its nearest centroid.

    `nearest_neighbors` accepts the reference's names; "scipy_kd_tree" (default) and
    "brute_force" both run the exact GPU search (fp64 squared distances, lowest centroid
    index on exact ties -- the brute-force semantics of _cvt_archive.py:626-632, which the
    KD-tree reproduces on tie-free data). "sklearn_nn" (arbitrary metrics) has no GPU
    implementation and raises.

    Multi-GPU: with `shard_centroids=True` and an initialised torch.distributed NCCL
    process group, rank r searches centroids [r*C/R, (r+1)*C/R) and the ranks combine with
    an all-reduce(min) over the exact distances followed by an all-reduce(min) over the
    candidate indices; the store is replicated.
    """

    def __init__(self, *, solution_dim, centroids, ranges, learning_rate=None, threshold_min=-np.inf,
                 qd_score_offset=0.0, seed=None, solution_dtype=None, objective_dtype=None, measures_dtype=None,
                 dtype=None, extra_fields=None, samples=100_000, k_means_kwargs=None,
                 nearest_neighbors="scipy_kd_tree", kdtree_kwargs=None, kdtree_query_kwargs=None, chunk_size=None,
                 sklearn_nn_kwargs=None, cells=None, custom_centroids=None, centroid_method=None, use_kd_tree=None,
                 ckdtree_kwargs=None, device=None, search_method="auto", shard_centroids=False, data_parallel=False,
                 process_group=None, shard_store=False):
        for name, val in (("cells", cells), ("custom_centroids", custom_centroids),
                          ("centroid_method", centroid_method), ("use_kd_tree", use_kd_tree),
                          ("ckdtree_kwargs", ckdtree_kwargs)):
            if val is not None:
                raise ValueError(_DEPRECATED[name])
        if nearest_neighbors not in ("scipy_kd_tree", "brute_force", "sklearn_nn"):
            raise ValueError(f"Unknown value `{nearest_neighbors}` for nearest_neighbors.")
        if nearest_neighbors == "sklearn_nn":
            raise NotImplementedError(
                "nearest_neighbors='sklearn_nn' (arbitrary sklearn metrics) has no GPU implementation and this "
                "package has no CPU fallback; use 'scipy_kd_tree' or 'brute_force' (Euclidean).")
        # The search is exact Euclidean nearest neighbour. Options that would change the answer of the reference's
        # KDTree.query (_cvt_archive.py:605-609) are refused rather than silently ignored; options that only tune
        # the tree (leafsize, workers, ...) have no counterpart here and are accepted.
        q = dict(kdtree_query_kwargs or {})
        if q.get("p", 2) != 2 or q.get("eps", 0) != 0 or q.get("k", 1) != 1 \
                or q.get("distance_upper_bound", np.inf) != np.inf:
            raise NotImplementedError(
                "kdtree_query_kwargs that change the result (p != 2, eps > 0, k != 1, distance_upper_bound) are not "
                "supported by the GPU search, which returns the exact Euclidean nearest centroid.")
        if kdtree_kwargs and kdtree_kwargs.get("boxsize") is not None:
            raise NotImplementedError("periodic KD-trees (kdtree_kwargs['boxsize']) are not supported by the GPU search.")
        self._nearest_neighbors = nearest_neighbors
        from pyribs_b200.archives._archive_base import parse_all_dtypes

        _, _, mdt = parse_all_dtypes(dtype, solution_dtype, objective_dtype, measures_dtype)
        if isinstance(centroids, numbers.Integral):  # _cvt_archive.py:351-360: k-means over uniform samples
            from pyribs_b200.archives._k_means import k_means_centroids

            centroids, _ = k_means_centroids(centroids=centroids, ranges=ranges, samples=samples, dtype=mdt, seed=seed,
                                             k_means_kwargs=k_means_kwargs, device=device)
        if isinstance(centroids, (str, Path)):
            cen = np.load(centroids).astype(mdt)
        else:
            cen = np.asarray(centroids, dtype=mdt)
        check_batch_shape(cen, "centroids", len(ranges), "measure_dim", batch_name="num_centroids")
        self._centroids = cen
        DeviceArchive.__init__(
            self, solution_dim=solution_dim, measure_dim=len(ranges), cells=len(cen), learning_rate=learning_rate,
            threshold_min=threshold_min, qd_score_offset=qd_score_offset, seed=seed, solution_dtype=solution_dtype,
            objective_dtype=objective_dtype, measures_dtype=measures_dtype, dtype=dtype, extra_fields=extra_fields,
            device=device, data_parallel=data_parallel, process_group=process_group, shard_store=shard_store)
        lo, hi = zip(*ranges)
        self._lower_bounds = np.array(lo, dtype=mdt)
        self._upper_bounds = np.array(hi, dtype=mdt)
        self._interval_size = self._upper_bounds - self._lower_bounds
        self._meas_code = _lib.QD_F64 if np.dtype(mdt) == np.float64 else _lib.QD_F32
        search_method = os.environ.get("QD_CVT_METHOD", search_method)
        if search_method not in ("auto", "tensor", "scan", "bucket", "tree"):
            raise ValueError("search_method must be 'auto', 'tensor', 'scan', 'bucket' or 'tree'")
        if search_method == "bucket" and self._measure_dim > 3:
            raise ValueError("search_method='bucket' supports measure_dim <= 3")
        self._search_method = search_method
        # centroid shard owned by this rank
        self._pg = process_group
        self._world = 1
        self._rank = 0
        if shard_centroids:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise RuntimeError("shard_centroids=True needs an initialised torch.distributed process group")
            self._world = dist.get_world_size(process_group)
            self._rank = dist.get_rank(process_group)
        self._shard_lo, self._shard_hi = shard_bounds(len(cen), self._world, self._rank)
        self._geom_cache = {}
        self._force_scan = False
        self._rebuild_device_side()

    def _transient_extra(self):
        return ("_centroids_dev", "_prepared", "_buckets", "_tree", "_cvt_scratch", "_geom_cache", "_pg")

    def _rebuild_device_side(self):
        """Centroid shard + search structure on the device (construction and unpickling)."""
        self.__dict__.setdefault("_pg", None)
        self._geom_cache = {}
        cen, search_method = self._centroids, self._search_method
        shard = np.ascontiguousarray(cen[self._shard_lo:self._shard_hi])
        with torch.cuda.device(self._device):
            self._centroids_dev = torch.from_numpy(shard).to(self._device) if len(shard) else None
            self._prepared = None
            self._buckets = None
            self._tree = None
            if len(shard) and (search_method == "tree" or (search_method == "auto" and 3 < self._measure_dim <= 10)):
                # mid measure_dim: bounding-box tree in KD order (csrc/cvt_tree.cu), the analogue of the reference's
                # KD-tree where a bucket grid no longer works. Measured on uniform data (tools/bench_extras.py):
                # 8-D 1M x 1M 9.7 ms against 89 ms exhaustive; from ~12-D on boxes stop pruning (16-D: 14x slower
                # than the contraction) and the exhaustive tensor-core search below takes over
                nbytes = int(_lib.lib().qd_cvt_tree_bytes(len(shard), self._measure_dim))
                self._tree = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
                _lib.check(_lib.lib().qd_cvt_tree_prepare(self._centroids_dev.data_ptr(), self._meas_code, len(shard),
                                                          self._measure_dim, self._tree.data_ptr(),
                                                          current_stream_ptr()))
            elif len(shard) and self._measure_dim <= 3 and search_method in ("auto", "bucket"):
                # low measure_dim: bucket grid, the analogue of the KD-tree the reference builds here
                nbytes = int(_lib.lib().qd_cvt_bucket_bytes(len(shard), self._measure_dim))
                self._buckets = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
                _lib.check(_lib.lib().qd_cvt_bucket_prepare(self._centroids_dev.data_ptr(), self._meas_code,
                                                            len(shard), self._measure_dim,
                                                            self._buckets.data_ptr(), current_stream_ptr()))
            elif len(shard):
                nbytes = _lib.lib().qd_cvt_prepared_bytes(len(shard), self._measure_dim)
                self._prepared = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self._device)
                _lib.check(_lib.lib().qd_cvt_prepare(self._centroids_dev.data_ptr(), self._meas_code, len(shard),
                                                     self._measure_dim, self._prepared.data_ptr(),
                                                     current_stream_ptr()))
        self._cvt_scratch = None

    @property
    def centroids(self):
        return self._centroids

    @property
    def lower_bounds(self):
        return self._lower_bounds

    @property
    def upper_bounds(self):
        return self._upper_bounds

    @property
    def interval_size(self):
        return self._interval_size

    def same_geometry(self, other):
        import weakref

        hit = self._geom_cache.get(id(other))
        if hit is None or hit[0]() is not other:  # (an id can be reused by a new object after garbage collection)
            same = bool(
                type(other) is type(self) and self._device == other._device
                and self._centroids.shape == other._centroids.shape and self._centroids.dtype == other._centroids.dtype
                and getattr(other, "_world", 1) == 1 and self._world == 1
                and np.array_equal(self._centroids, other._centroids))
            if len(self._geom_cache) > 16:
                self._geom_cache.clear()
            hit = self._geom_cache[id(other)] = (weakref.ref(other), same)
        return hit[1]

    def _host_measures_ok(self):
        return self._buckets is not None and not self._force_scan and self._search_method != "scan"

    def _pick_method(self, B, C):
        if self._force_scan or self._search_method == "scan":
            return 1
        if self._buckets is not None:
            return 2
        if self._tree is not None:
            return 3
        if self._search_method == "tensor":
            return 0
        return 0 if B * C >= (1 << 26) else 1

    # ---- A3 ---------------------------------------------------------------------------------
    def _small_index_ok(self, B):
        # (the tensor-core search stages the queries with TMA: device memory only)
        return self._world == 1 and self._pick_method(B, self._shard_hi - self._shard_lo) != 0

    def _local_search(self, measures, flags_ptr, want_dist, out=None):
        B = int(measures.shape[0])
        Cn = self._shard_hi - self._shard_lo
        idx = out if out is not None else torch.empty(B, dtype=torch.int32, device=self._device)
        dist = torch.empty(B, dtype=torch.float64, device=self._device) if want_dist else None
        if Cn == 0:
            idx.fill_(2**31 - 1)
            if dist is not None:
                dist.fill_(float("inf"))
            return idx, dist
        method = self._pick_method(B, Cn)
        # the tensor-core filter can (rarely) overflow its candidate list, which is only visible in the
        # flags of the finished call: such insertions are waited for even when given CUDA tensors
        self._needs_flag_check = method == 0
        if method == 0:
            need = int(_lib.lib().qd_cvt_scratch_bytes(B, Cn, self._measure_dim))
        elif method == 3:
            need = int(_lib.lib().qd_cvt_tree_scratch_bytes(B, Cn))
        else:
            need = 256
        if self._cvt_scratch is None or self._cvt_scratch.numel() < need:
            self._cvt_scratch = torch.empty(max(need, 256), dtype=torch.uint8, device=self._device)
        prepared = self._buckets if method == 2 else (self._tree if method == 3 else self._prepared)
        _lib.check(_lib.lib().qd_cvt_index_of(
            measures.data_ptr(), self._meas_code, B, self._measure_dim, self._centroids_dev.data_ptr(), Cn,
            prepared.data_ptr() if prepared is not None else None, self._shard_lo, method, idx.data_ptr(),
            dist.data_ptr() if dist is not None else None, flags_ptr, self._cvt_scratch.data_ptr(),
            self._cvt_scratch.numel(), current_stream_ptr()))
        return idx, dist

    def _index_of_device(self, measures: torch.Tensor, flags_ptr: int, out=None) -> torch.Tensor:
        if self._world == 1:
            return self._local_search(measures, flags_ptr, False, out)[0]
        idx, d_local = self._local_search(measures, flags_ptr, True)

        def mask(d_loc, d_min, ix):
            _lib.check(_lib.lib().qd_cvt_mask_losers(d_loc.data_ptr(), d_min.data_ptr(), ix.data_ptr(),
                                                     int(ix.numel()), current_stream_ptr()))

        idx, _ = merge_sharded_argmin(idx, d_local, self._pg, mask)
        return idx
