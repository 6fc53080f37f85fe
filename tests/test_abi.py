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
    assert _lib.lib().qd_abi_version() == 6


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors in pyribs_b200/_lib.py against the header as gcc lays it out."""
    import subprocess

    from pyribs_b200 import _lib

    fields = {"qd_add_result": [f[0] for f in _lib.AddResult._fields_],
              "qd_store": [f[0] for f in _lib.Store._fields_],
              "qd_cma_scalars": [f[0] for f in _lib.CmaScalars._fields_]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qd_b200.h"', 'int main(void) {']
    for st, names in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for n in names:
            lines.append(f'printf("{st}.{n} %zu\\n", offsetof({st}, {n}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for st, cls in (("qd_add_result", _lib.AddResult), ("qd_store", _lib.Store), ("qd_cma_scalars", _lib.CmaScalars)):
        assert int(got[st]) == ctypes.sizeof(cls), st
        for n in fields[st]:
            assert int(got[f"{st}.{n}"]) == getattr(cls, n).offset, f"{st}.{n}"
    assert _lib.lib().qd_insert_scratch_bytes(1000) >= 1000 * 40


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
