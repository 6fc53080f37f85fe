"""CategoricalArchive on the device against trajectories of the reference's CategoricalArchive
(tests/golden/categorical.npz, oracle/gen_golden.py gen_categorical) and the reference's own unit tests
(tests/archives/categorical_archive_test.py)."""
import numpy as np
import pytest

from golden_util import assert_close, load_golden, steps_of

pytestmark = pytest.mark.gpu
CATS = [["A", "B", "C"], ["One", "Two", "Three", "Four"]]


def test_properties_and_index_of():
    """categorical_archive_test.py:11-42, :78-130."""
    from pyribs_b200.archives import CategoricalArchive

    a = CategoricalArchive(solution_dim=(), categories=CATS)
    assert (a.dims == [3, 4]).all() and a.categories == CATS and a.cells == 12
    assert (a.index_of([["A", "One"], ["B", "Two"], ["C", "Four"]]) == [0, 5, 11]).all()
    assert a.index_of_single(["C", "One"]) == 8
    with pytest.raises(KeyError):
        a.index_of([["Z", "One"]])
    with pytest.raises(ValueError):
        a.index_of([["A", "One", "x"]])
    f32 = CategoricalArchive(solution_dim=(), categories=CATS, dtype=np.float32)
    assert f32.dtypes["solution"] == np.float32 and f32.dtypes["objective"] == np.float32
    assert f32.dtypes["measures"] == np.object_ and f32.dtypes["threshold"] == np.float32
    assert f32.dtypes["index"] == np.int32
    with pytest.raises(ValueError, match=r"dtype cannot be used at the same time as .*"):
        CategoricalArchive(solution_dim=(), categories=CATS, dtype=np.float32, measures_dtype=object)


@pytest.mark.parametrize("tag", ["me", "mae"])
def test_trajectory_matches_reference(tag):
    from pyribs_b200.archives import CategoricalArchive

    case = load_golden("categorical")[tag]
    kw = {} if tag == "me" else {"learning_rate": 0.2, "threshold_min": -1.0}
    a = CategoricalArchive(solution_dim=3, categories=CATS, seed=9, **kw)
    for j, s in enumerate(steps_of(case)):
        m = s["measures"].astype(object)
        np.testing.assert_array_equal(a.index_of(m), s["index_of"])
        info = a.add(s["solution"], s["objective"], m)
        np.testing.assert_array_equal(info["status"], s["status"], err_msg=f"{tag} s{j} status")
        assert_close(info["value"], s["value"], 1e-6, f"{tag} s{j} value")
        d = a.data()
        order = np.argsort(d["index"])
        np.testing.assert_array_equal(d["index"][order], s["index"])
        np.testing.assert_array_equal(d["solution"][order], s["d_solution"])
        np.testing.assert_array_equal(d["objective"][order], s["d_objective"])
        assert_close(d["threshold"][order], s["d_threshold"], 1e-6, f"{tag} s{j} threshold")
        assert d["measures"].dtype == object
        np.testing.assert_array_equal(d["measures"][order].astype("U5"), s["d_measures"])
        assert len(a) == int(s["n"])
        assert_close(a.stats.qd_score, s["qd_score"], 1e-6, f"{tag} s{j} qd")
        assert_close(a.stats.obj_max, s["obj_max"], 1e-12, f"{tag} s{j} obj_max")
        np.testing.assert_array_equal(a.best_elite["solution"], s["best_solution"])
        np.testing.assert_array_equal(a.best_elite["measures"].astype("U5"), s["best_measures"])
    single = a.add_single(np.ones(3), 100.0, ["B", "Four"])
    assert int(single["status"]) == int(case["single"]["status"])
    assert_close(single["value"], case["single"]["value"], 1e-6, "single value")
    np.testing.assert_array_equal(a.best_elite["measures"].astype("U5"), case["single"]["best_measures"])
    # read side: retrieve (occupied and not), sample_elites, iteration, tuple / single-field data
    occ, got = a.retrieve([["B", "Four"], ["A", "One"]])
    assert occ[0] and list(got["measures"][0]) == ["B", "Four"] and got["objective"][0] == 100.0
    if not occ[1]:
        assert got["measures"][1][0] is None and np.isnan(got["objective"][1])
    o1, g1 = a.retrieve_single(["B", "Four"])
    assert o1 and list(g1["measures"]) == ["B", "Four"]
    e = a.sample_elites(6)
    assert e["measures"].shape == (6, 2) and all(v[0] in CATS[0] and v[1] in CATS[1] for v in e["measures"])
    rows = list(a)
    assert len(rows) == len(a) and all(r["measures"][0] in CATS[0] for r in rows)
    names = a.data("measures")
    assert names.dtype == object and names.shape == (len(a), 2)
    tup = a.data(["objective", "measures"], return_type="tuple")
    assert tup[1].dtype == object and len(tup[0]) == len(a)
