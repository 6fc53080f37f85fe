"""Host-side rankers vs the reference's golden orderings (CPU)."""
import numpy as np

from golden_util import load_golden


class _Archive:
    lower_bounds = np.array([-1.0, -2.0])
    upper_bounds = np.array([1.0, 2.0])


def test_rankers_match_reference_orderings():
    from pyribs_b200.emitters.rankers import _get_ranker

    g = load_golden("rankers")
    for tag, c in g.items():
        data = {"objective": c["objective"], "measures": c["measures"]}
        info = {"status": c["status"], "value": c["value"]}
        for name in ("imp", "2imp", "obj", "2obj", "rd", "2rd"):
            rk = _get_ranker(name, 11)
            rk.reset(None, _Archive())
            if name in ("rd", "2rd"):
                np.testing.assert_allclose(rk.target_measure_dir, c[name]["dir"], rtol=0, atol=0)
            idx, vals = rk.rank(None, _Archive(), data, info)
            np.testing.assert_array_equal(idx, c[name]["indices"], err_msg=f"{tag} {name}")
            np.testing.assert_array_equal(vals, c[name]["values"], err_msg=f"{tag} {name}")
