"""k_means_centroids on the device.

Reference: ribs/archives/_cvt_archive.py:36-134 -- uniform samples over the measure ranges, clustered by
`sklearn.cluster.k_means(samples, centroids, n_init=1, init="random", algorithm="lloyd", random_state=seed)`.
Same signature, same sample array (the reference's generator calls are repeated verbatim), same Lloyd iteration
and stopping rule as scikit-learn's `_kmeans_single_lloyd` (strict convergence of the labels, or total squared
centre shift <= tol * mean per-feature variance; empty clusters move to the samples farthest from their centres).
The assignment step is the archive's own exact nearest-centroid search (bucket grid / tcgen05 filter + fp64
rescoring), the update step csrc/kmeans.cu (order-independent exact sums, so a seeded run is reproducible).
scikit-learn's random initial centres come from a different generator, so the centroids are not bit-comparable
with the reference's; tests compare the quantisation error against scikit-learn's on the same samples.
"""
from __future__ import annotations

import numbers

import numpy as np
import torch

from pyribs_b200 import _lib
from pyribs_b200._utils import check_batch_shape
from pyribs_b200.archives._device_store import current_stream_ptr


def k_means_centroids(*, centroids, ranges, samples=100_000, dtype=np.float64, seed=None, k_means_kwargs=None,
                      device=None):
    """Returns (centroids (k, measure_dim), samples (num_samples, measure_dim)), both of `dtype`."""
    from pyribs_b200.archives._cvt_archive import CVTArchive

    measure_dim = len(ranges)
    lo, hi = (np.array(b, dtype=dtype) for b in zip(*ranges))
    kw = {} if k_means_kwargs is None else dict(k_means_kwargs)
    n_init = kw.pop("n_init", 1)
    init = kw.pop("init", "random")
    algorithm = kw.pop("algorithm", "lloyd")
    random_state = kw.pop("random_state", seed)
    max_iter = int(kw.pop("max_iter", 300))
    tol = float(kw.pop("tol", 1e-4))
    kw.pop("verbose", None), kw.pop("copy_x", None), kw.pop("return_n_iter", None)
    if kw:
        raise TypeError(f"k_means_kwargs not supported on the GPU path: {sorted(kw)}")
    if algorithm != "lloyd" or n_init not in (1, "auto"):
        raise NotImplementedError("the GPU k-means runs one Lloyd clustering (algorithm='lloyd', n_init=1)")
    if isinstance(samples, numbers.Integral):  # _cvt_archive.py:104-111
        rng = np.random.default_rng(seed)
        samples = rng.uniform(lo, hi, size=(samples, measure_dim)).astype(dtype)
    else:
        samples = np.asarray(samples, dtype=dtype)
        check_batch_shape(samples, "samples", measure_dim, "measure_dim", batch_name="num_samples")
    if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("dtype must be float32 or float64")
    n, k = len(samples), int(centroids)
    if n < k:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    if isinstance(init, str):
        if init != "random":
            raise NotImplementedError("init must be 'random' or an array of initial centres on the GPU path")
        rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        cen = samples[rs.choice(n, size=k, replace=False)].copy()  # sklearn _init_centroids, init="random"
    else:
        cen = np.array(init, dtype=dtype)
        check_batch_shape(cen, "init", measure_dim, "measure_dim", batch_name="n_clusters")
        if len(cen) != k:
            raise ValueError(f"The shape of the initial centers {cen.shape} does not match the number of clusters {k}.")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    L = _lib.lib()
    code = _lib.QD_F64 if np.dtype(dtype) == np.float64 else _lib.QD_F32
    tdt = torch.float64 if code == _lib.QD_F64 else torch.float32
    with torch.cuda.device(dev):
        X = torch.from_numpy(np.ascontiguousarray(samples)).to(dev)
        absmax = float(np.abs(samples).max()) if n else 0.0
        threshold = float(np.mean(np.var(samples.astype(np.float64), axis=0))) * tol  # sklearn _tolerance
        bins = torch.empty(int(L.qd_kmeans_workspace_doubles(k, measure_dim)), dtype=torch.float64, device=dev)
        count = torch.empty(k, dtype=torch.int32, device=dev)
        sums = torch.empty((k, measure_dim), dtype=torch.float64, device=dev)
        shift = torch.empty(k, dtype=torch.float64, device=dev)
        new_c = torch.empty((k, measure_dim), dtype=tdt, device=dev)
        search_ranges = [(float(a), float(b)) for a, b in zip(lo, hi)]
        labels_old = None
        strict = False

        def assign(c_np):
            index = CVTArchive(solution_dim=1, centroids=c_np, ranges=search_ranges, dtype=dtype, device=dev)
            return index.index_of(X)

        for _ in range(max_iter):
            labels = assign(cen)
            old_c = torch.from_numpy(cen).to(dev)
            _lib.check(L.qd_kmeans_update(X.data_ptr(), code, n, measure_dim, labels.data_ptr(), k, absmax,
                                          bins.data_ptr(), count.data_ptr(), old_c.data_ptr(), new_c.data_ptr(),
                                          sums.data_ptr(), shift.data_ptr(), current_stream_ptr()))
            cnt = count.cpu().numpy().astype(np.int64)
            if np.any(cnt == 0):
                # sklearn _relocate_empty_clusters_dense: every empty cluster takes the sample farthest from its
                # centre (and that sample leaves its old cluster)
                empty = np.flatnonzero(cnt == 0)
                d2 = ((X.double() - old_c.double()[labels.long()]) ** 2).sum(1).cpu().numpy()
                far = np.argpartition(d2, -len(empty))[:-len(empty) - 1:-1]
                s_h, lab_h = sums.cpu().numpy(), labels.cpu().numpy()
                for new_id, far_idx in zip(empty, far):
                    old_id = lab_h[far_idx]
                    s_h[old_id] -= samples[far_idx]
                    s_h[new_id] = samples[far_idx]
                    cnt[new_id] = 1
                    cnt[old_id] -= 1
                new_np = (s_h / np.maximum(cnt, 1)[:, None]).astype(dtype)
                new_np[cnt == 0] = cen[cnt == 0]
                total_shift = float(((new_np.astype(np.float64) - cen.astype(np.float64)) ** 2).sum())
            else:
                new_np = new_c.cpu().numpy().copy()
                total_shift = float(shift.sum().item())
            cen = new_np
            if labels_old is not None and torch.equal(labels, labels_old):
                strict = True
                break
            if total_shift <= threshold:
                break
            labels_old = labels
    if len(np.unique(cen, axis=0)) < k:  # _cvt_archive.py:124-131
        raise RuntimeError(
            f"While generating the CVT, k-means clustering found {len(np.unique(cen, axis=0))} centroids, but this "
            f"archive needs {k} cells. This most likely happened because there are too few samples and/or too many "
            "cells.")
    del strict
    return cen, samples
