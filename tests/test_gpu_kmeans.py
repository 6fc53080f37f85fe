"""k_means_centroids on the device (reference ribs/archives/_cvt_archive.py:36-134, tests/archives/
k_means_centroids_test.py) and CVTArchive(centroids=<int>)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["float64", "float32"])
def test_basic(dtype):
    """tests/archives/k_means_centroids_test.py:9-28: shapes, dtypes, bounds; and the samples are the reference's
    (same generator calls): uniform(lower, upper, (n, d)) of default_rng(seed)."""
    from pyribs_b200.archives import k_means_centroids

    centroids, samples = k_means_centroids(centroids=100, ranges=[(-1, 1), (-1, 1)], samples=1000, dtype=dtype, seed=42)
    assert centroids.shape == (100, 2) and centroids.dtype == dtype
    assert np.all(centroids >= [-1, -1]) and np.all(centroids <= [1, 1])
    assert samples.shape == (1000, 2) and samples.dtype == dtype
    np.testing.assert_array_equal(samples, np.random.default_rng(42).uniform([-1, -1], [1, 1], (1000, 2)).astype(dtype))
    assert len(np.unique(centroids, axis=0)) == 100
    again, _ = k_means_centroids(centroids=100, ranges=[(-1, 1), (-1, 1)], samples=1000, dtype=dtype, seed=42)
    np.testing.assert_array_equal(centroids, again)  # exact sums: a seeded run is reproducible bit for bit


def test_samples_bad_shape():
    from pyribs_b200.archives import k_means_centroids

    with pytest.raises(ValueError, match=r"Expected samples to be an array with shape .*"):
        k_means_centroids(centroids=100, ranges=[(-1, 1), (-1, 1)], samples=np.zeros((1000, 3)), dtype=np.float64, seed=42)
    with pytest.raises(ValueError):
        k_means_centroids(centroids=100, ranges=[(-1, 1), (-1, 1)], samples=50, seed=1)


@pytest.mark.parametrize("k,n,d", [(500, 20_000, 2), (300, 6_000, 5), (64, 70, 3)])
def test_quality_matches_sklearn(k, n, d):
    """Both are Lloyd runs from random initial centres, so the centroids differ, but the quantisation error
    (inertia) of the result must be as good as scikit-learn's on the same samples, and every centroid must be the
    mean of the samples nearest to it (a fixed point of the iteration)."""
    from sklearn.cluster import k_means

    from pyribs_b200.archives import k_means_centroids

    ranges = [(-1.0, 2.0)] * d
    cen, samples = k_means_centroids(centroids=k, ranges=ranges, samples=n, seed=7)

    def inertia(c):
        d2 = ((samples[:, None, :] - c[None]) ** 2).sum(-1) if n * k < 5e6 else None
        if d2 is None:
            from scipy.spatial import cKDTree
            return float((cKDTree(c).query(samples)[0] ** 2).sum())
        return float(d2.min(1).sum())

    ref = k_means(samples, k, n_init=1, init="random", algorithm="lloyd", random_state=7)[0]
    mine, theirs = inertia(cen), inertia(ref)
    # (with hardly more samples than centres the outcome is dominated by the random initial centres)
    assert mine < theirs * (1.04 if n >= 10 * k else 1.25), (mine, theirs)
    if n < 10 * k:
        return
    from scipy.spatial import cKDTree
    lab = cKDTree(cen).query(samples)[1]
    means = np.array([samples[lab == c].mean(0) if np.any(lab == c) else cen[c] for c in range(k)])
    # converged by tolerance, not necessarily to the exact fixed point: centres sit close to their cell means
    assert np.sqrt(((means - cen) ** 2).sum(1)).max() < 0.05 * 3.0


def test_cvt_archive_with_centroid_count():
    """CVTArchive(centroids=<int>) clusters uniform samples itself (_cvt_archive.py:351-360), like the reference's
    `cvt_map_elites` configuration (examples/sphere.py)."""
    from pyribs_b200.archives import CVTArchive, k_means_centroids

    a = CVTArchive(solution_dim=3, centroids=200, ranges=[(-1, 1), (-2, 2)], samples=5000, seed=11)
    expect, _ = k_means_centroids(centroids=200, ranges=[(-1, 1), (-2, 2)], samples=5000, seed=11)
    np.testing.assert_array_equal(a.centroids, expect)
    assert a.cells == 200
    rng = np.random.default_rng(0)
    info = a.add(rng.standard_normal((1000, 3)), rng.standard_normal(1000), rng.uniform(-1, 1, (1000, 2)) * [1, 2])
    assert info["status"].shape == (1000,) and 50 < len(a) <= 200
    with pytest.raises(TypeError):
        k_means_centroids(centroids=10, ranges=[(-1, 1)], samples=100, k_means_kwargs={"bogus": 1})
