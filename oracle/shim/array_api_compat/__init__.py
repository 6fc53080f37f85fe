"""Test-infrastructure shim: re-exports SciPy's vendored array_api_compat.

The reference (`/root/reference/ribs/_utils.py:9-11`, `_array_store.py:15-19`) imports
`array_api_compat`, which is not installed in this image. SciPy vendors the same
package (v1.14) under `scipy._external.array_api_compat`. This shim exists only so the
UNMODIFIED reference can be imported here to pin the oracle and generate goldens.
It is never imported by the product package.
"""
import sys as _sys
import importlib as _importlib

from scipy._external import array_api_compat as _aac
from scipy._external.array_api_compat import *  # noqa: F401,F403
from scipy._external.array_api_compat import (  # noqa: F401
    array_namespace,
    is_cupy_array,
    is_numpy_array,
    is_torch_array,
)

__version__ = getattr(_aac, "__version__", "vendored")

# Make `import array_api_compat.numpy as np_compat` / `.torch` resolve to the vendored
# sub-packages.
for _name in ("numpy", "torch", "common", "cupy"):
    try:
        _mod = _importlib.import_module(f"scipy._external.array_api_compat.{_name}")
    except Exception:  # pragma: no cover - optional backends
        continue
    _sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
