"""Helpers shared by the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    """Returns the nested dict that oracle/gen_golden.py flattened into `name`.npz."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    root = {}
    for key in z.files:
        parts = key.split("/")
        d = root
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = z[key]
    return root


def steps_of(case):
    return [case[k] for k in sorted((k for k in case if k[0] == "s" and k[1:].isdigit()), key=lambda s: int(s[1:]))]


def gens_of(case):
    return [case[k] for k in sorted((k for k in case if k[0] == "g" and k[1:].isdigit()), key=lambda s: int(s[1:]))]


def assert_close(a, b, rtol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = np.maximum(np.abs(b), 1e-300)
    bad = np.abs(a - b) > rtol * np.maximum(scale, np.max(np.abs(b)) * 1e-3 if b.size else 0)
    assert not np.any(bad), f"{what}: {np.sum(bad)} mismatches, max abs err {np.max(np.abs(a - b))}"


def check_archive_dump(got, want, rtol, tag=""):
    """`got`/`want`: dicts with occupied, occupied_list, n, field_*, stats, best_elite."""
    n = int(want["n"])
    assert int(got["n"]) == n, f"{tag}: n_occupied {got['n']} vs {n}"
    np.testing.assert_array_equal(got["occupied"], want["occupied"], err_msg=f"{tag} occupied")
    np.testing.assert_array_equal(
        np.asarray(got["occupied_list"])[:n], np.asarray(want["occupied_list"])[:n], err_msg=f"{tag} occupied_list"
    )
    occ = np.asarray(want["occupied"])
    for key, w in want.items():
        if not key.startswith("field_"):
            continue
        g = np.asarray(got[key])
        if key in ("field_threshold",):
            assert_close(g[occ], w[occ], rtol, f"{tag} {key}")
        else:  # elite rows are copies of inputs: exact
            np.testing.assert_array_equal(g[occ], w[occ], err_msg=f"{tag} {key}")
    for k, w in want["stats"].items():
        g = got["stats"][k]
        if np.isnan(w):
            assert g is None or np.isnan(g), f"{tag} stats.{k}"
        else:
            assert_close(g, w, rtol, f"{tag} stats.{k}")
    if "best_elite" in want:
        for k, w in want["best_elite"].items():
            g = got["best_elite"][k]
            if k == "threshold":
                assert_close(g, w, rtol, f"{tag} best_elite.{k}")
            else:
                np.testing.assert_array_equal(np.asarray(g), w, err_msg=f"{tag} best_elite.{k}")
