"""Constructor contracts that are decided on the host before any device work (so they run without a GPU):
deprecated CVTArchive kwargs raise as in the reference (ribs/archives/_cvt_archive.py:299-324); options the GPU
search cannot honour are refused instead of silently ignored."""
import numpy as np
import pytest


def _cvt(**kw):
    from pyribs_b200.archives import CVTArchive

    return CVTArchive(solution_dim=3, centroids=np.zeros((4, 2)), ranges=[(-1, 1)] * 2, **kw)


@pytest.mark.parametrize("name", ["cells", "custom_centroids", "centroid_method", "use_kd_tree", "ckdtree_kwargs"])
def test_deprecated_cvt_kwargs_raise(name):
    with pytest.raises(ValueError, match="deprecated in pyribs 0.9.0"):
        _cvt(**{name: 1})


def test_unknown_nearest_neighbors():
    with pytest.raises(ValueError, match="Unknown value"):
        _cvt(nearest_neighbors="annoy")
    with pytest.raises(NotImplementedError):
        _cvt(nearest_neighbors="sklearn_nn")


@pytest.mark.parametrize("q", [{"p": 1}, {"eps": 0.1}, {"k": 2}, {"distance_upper_bound": 0.5}])
def test_result_changing_kdtree_query_options_are_refused(q):
    """tests/archives/cvt_archive_test.py:194-208 uses p=1 (Manhattan): the GPU search is Euclidean only."""
    with pytest.raises(NotImplementedError):
        _cvt(kdtree_query_kwargs=q)
    with pytest.raises(NotImplementedError):
        _cvt(kdtree_kwargs={"boxsize": 2.0})
