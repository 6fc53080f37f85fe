"""Host-side validation for the archive / emitter boundary.

Mirrors the error contract of the reference's `ribs/_utils.py` (validate_batch :109-214,
validate_single :217-238, check_batch_shape :34-57, check_shape :60-78): same exception
types, same conditions. Finite checks on `objective` / `measures` of a batch are NOT done
here -- the kernels fold them into their first pass and report through a device flag, see
DeviceArchive.add.
"""
from __future__ import annotations

import numbers

import numpy as np
import torch

NP_TO_TORCH = {
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
    np.dtype(np.int32): torch.int32,
    np.dtype(np.int64): torch.int64,
    np.dtype(np.int16): torch.int16,
    np.dtype(np.int8): torch.int8,
    np.dtype(np.uint8): torch.uint8,
    np.dtype(np.bool_): torch.bool,
    np.dtype(np.float16): torch.float16,
}


def torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        return dt
    dt = np.dtype(dt)
    if dt not in NP_TO_TORCH:
        raise TypeError(
            f"dtype {dt} cannot live in device memory; only numeric fields are supported "
            "(object-dtype fields have no GPU representation and there is no CPU fallback)"
        )
    return NP_TO_TORCH[dt]


def is_device_tensor(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def shape_of(x):
    return tuple(x.shape)


def as_dim_tuple(dim):
    return (int(dim),) if isinstance(dim, numbers.Integral) else tuple(int(v) for v in dim)


def check_batch_shape(array, array_name, dim, dim_name, extra_msg="", batch_name="batch_size"):
    dim = as_dim_tuple(dim)
    if array.ndim != 1 + len(dim) or shape_of(array)[1:] != dim:
        dim_str = ", ".join(map(str, dim))
        raise ValueError(
            f"Expected {array_name} to be an array with shape ({batch_name}, {dim_str}) (i.e. shape "
            f"({batch_name}, {dim_name})) but it had shape {shape_of(array)}.{extra_msg}"
        )


def check_shape(array, array_name, dim, dim_name, extra_msg=""):
    dim = as_dim_tuple(dim)
    if array.ndim != len(dim) or shape_of(array) != dim:
        comma = "," if len(dim) == 1 else ""
        raise ValueError(
            f"Expected {array_name} to be an array with shape {dim} (i.e. shape ({dim_name}{comma})) "
            f"but it had shape {shape_of(array)}.{extra_msg}"
        )


def check_is_1d(array, array_name, extra_msg=""):
    if array.ndim != 1:
        raise ValueError(f"Expected {array_name} to be a 1D array but it had shape {shape_of(array)}.{extra_msg}")


def check_solution_batch_dim(array, array_name, batch_size, is_1d=False, extra_msg=""):
    if array.ndim == 0 or array.shape[0] != batch_size:
        raise ValueError(
            f"{array_name} does not match the batch dimension of solution -- since solution has shape "
            f"({batch_size}, ..), {array_name} should have shape ({batch_size},{'' if is_1d else ' ..'}), "
            f"but it has shape {shape_of(array)}.{extra_msg}"
        )


def check_finite(x, name):
    ok = bool(torch.isfinite(x).all()) if isinstance(x, torch.Tensor) else bool(np.all(np.isfinite(x)))
    if not ok:
        if np.isscalar(x) or getattr(x, "ndim", 1) == 0:
            raise ValueError(f"{name} must be finite (infinity and NaN values are not supported).")
        raise ValueError(f"All elements of {name} must be finite (infinity and NaN values are not supported).")


def as_array(x, dtype=None):
    """np.asarray for host inputs; device tensors pass through (cast when asked)."""
    if isinstance(x, torch.Tensor):
        if dtype is not None and x.dtype != torch_dtype(dtype):
            x = x.to(torch_dtype(dtype))
        return x
    return np.asarray(x, dtype=dtype)


def arr_readonly(arr, view=False):
    if isinstance(arr, np.ndarray):
        out = arr.view() if view else arr
        out.flags.writeable = False
        return out
    return arr


def validate_batch(archive, data, add_info=None):
    """Shape / batch-size validation of add() and tell() arguments (ribs/_utils.py:109-214)."""
    data["solution"] = as_array(data["solution"], archive.dtypes["solution"])
    check_batch_shape(data["solution"], "solution", archive.solution_dim, "solution_dim", "")
    batch_size = data["solution"].shape[0]
    for name, arr in data.items():
        if name == "solution":
            continue
        if name == "objective":
            if arr is None:
                raise ValueError("objective cannot be None")
            arr = as_array(arr, archive.dtypes["objective"])
            check_is_1d(arr, "objective", "")
            check_solution_batch_dim(arr, "objective", batch_size, is_1d=True)
        elif name == "measures":
            arr = as_array(arr, archive.dtypes["measures"])
            check_batch_shape(arr, "measures", archive.measure_dim, "measure_dim", "")
            check_solution_batch_dim(arr, "measures", batch_size, is_1d=False)
        else:
            arr = as_array(arr)
            check_solution_batch_dim(arr, name, batch_size, is_1d=False)
        data[name] = arr
    if add_info is None:
        return data
    for name, arr in add_info.items():
        arr = as_array(arr)
        if name in ("status", "value"):
            check_is_1d(arr, name, "")
            check_solution_batch_dim(arr, name, batch_size, is_1d=True)
            if name == "status":
                check_finite(arr, "status")
        else:
            check_solution_batch_dim(arr, name, batch_size, is_1d=False)
        add_info[name] = arr
    return data, add_info


def validate_single(archive, data):
    data["solution"] = as_array(data["solution"], archive.dtypes["solution"])
    check_shape(data["solution"], "solution", archive.solution_dim, "solution_dim")
    if data["objective"] is None:
        raise ValueError("objective cannot be None")
    data["objective"] = as_array(data["objective"], archive.dtypes["objective"])
    check_finite(data["objective"], "objective")
    data["measures"] = as_array(data["measures"], archive.dtypes["measures"])
    check_shape(data["measures"], "measures", archive.measure_dim, "measure_dim")
    check_finite(data["measures"], "measures")
    return data
