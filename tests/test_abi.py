"""CPU checks of the boundary: the library loads and exports every symbol the header declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__  # noqa: F401  (ensures the library is built)
    from pyribs_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/qd_b200.h but not exported"
    assert set(_lib.PROTOTYPES) == set(names), set(_lib.PROTOTYPES) ^ set(names)
    assert _lib.lib().qd_abi_version() == 1


def test_struct_layouts_match_header():
    from pyribs_b200 import _lib

    assert ctypes.sizeof(_lib.AddResult) == 56
    assert _lib.lib().qd_insert_scratch_bytes(1000) > 1000 * 40


def test_product_does_not_import_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "pyribs_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+(oracle|qd_oracle)", src, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_no_gpu_means_loud_failure():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyribs_b200 import _lib
    from pyribs_b200.archives import GridArchive

    with pytest.raises(_lib.QdError):
        GridArchive(solution_dim=3, dims=[10, 20], ranges=[(-1, 1), (-2, 2)])
