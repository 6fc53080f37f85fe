"""ctypes binding of the C ABI in include/qd_b200.h. There is no CPU fallback: if the
library is missing or fails to load, importing the product classes raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libqd_b200.so")

QD_F32, QD_F64 = 0, 1
QD_MAX_FIELDS = 16
FLAG_NONFINITE_MEASURES = 1
FLAG_NONFINITE_OBJECTIVE = 2
FLAG_CVT_OVERFLOW = 4
INSERT_WHOLE, INSERT_CLASSIFY, INSERT_COMMIT, INSERT_ROWS, INSERT_COPY, INSERT_BUF1 = 0, 1, 2, 3, 4, 0x100


class QdError(RuntimeError):
    pass


class AddResult(C.Structure):
    _fields_ = [
        ("n_insertable", C.c_int32),
        ("n_new", C.c_int32),
        ("best_cell", C.c_int32),
        ("flags", C.c_uint32),
        ("best_objective", C.c_double),
        ("delta_sum", C.c_double * 3),
        ("batch_absmax", C.c_double),
        ("n_occupied", C.c_int64),
        ("objective_sum", C.c_double),
        ("obj_max", C.c_double),
        ("has_obj_max", C.c_int32),
        ("last_failed_flags", C.c_uint32),
        ("n_commits", C.c_uint64),
        ("best_serial", C.c_uint64),
        ("n_failed", C.c_uint64),
        ("n_winners", C.c_uint64),
    ]


class Store(C.Structure):
    _fields_ = [
        ("capacity", C.c_int64),
        ("obj_dtype", C.c_int),
        ("occupied", C.c_void_p),
        ("occupied_list", C.c_void_p),
        ("objective", C.c_void_p),
        ("threshold", C.c_void_p),
        ("n_fields", C.c_int),
        ("field_dst", C.c_void_p * QD_MAX_FIELDS),
        ("field_row_bytes", C.c_int64 * QD_MAX_FIELDS),
        ("scratch", C.c_void_p),
        ("best_snapshot", C.c_void_p),
    ]


class CmaPool(C.Structure):
    _fields_ = [("E", C.c_int64), ("n", C.c_int64), ("lam", C.c_int64)] + [
        (k, C.c_void_p) for k in ("mean", "pc", "ps", "C", "B", "eig", "T", "invsqrt", "status", "new_mean", "ws", "X", "z")
    ] + [("ws_stride", C.c_int64)]


class CmaScalars(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("cc", "cs", "c1", "cmu", "mueff", "damps", "wsum", "hsig_exp")]


_vp, _i64, _i32, _int, _dbl, _sz, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_double, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); mirrors include/qd_b200.h one to one.
PROTOTYPES = {
    "qd_abi_version": (_int, []),
    "qd_last_error": (C.c_char_p, []),
    "qd_launch_count": (_i64, []),
    "qd_stream_synchronize": (_int, [_vp]),
    "qd_memcpy_async": (_int, [_vp, _vp, _sz, _int, _vp]),
    "qd_event_record": (_int, [_vp, _vp]),
    "qd_stream_wait_event": (_int, [_vp, _vp]),
    "qd_grid_index_of": (_int, [_vp, _int, _i64, _int, C.POINTER(_i32), C.POINTER(_dbl), C.POINTER(_dbl), _dbl,
                                _vp, _vp, _vp]),
    "qd_cvt_prepared_bytes": (_sz, [_i64, _int]),
    "qd_cvt_prepare": (_int, [_vp, _int, _i64, _int, _vp, _vp]),
    "qd_cvt_bucket_bytes": (_sz, [_i64, _int]),
    "qd_cvt_bucket_prepare": (_int, [_vp, _int, _i64, _int, _vp, _vp]),
    "qd_cvt_tree_bytes": (_sz, [_i64, _int]),
    "qd_cvt_tree_scratch_bytes": (_sz, [_i64, _i64]),
    "qd_cvt_tree_prepare": (_int, [_vp, _int, _i64, _int, _vp, _vp]),
    "qd_cvt_tree_build_host": (_int, [_vp, _int, _i64, _int, _vp]),
    "qd_cvt_scratch_bytes": (_sz, [_i64, _i64, _int]),
    "qd_cvt_index_of": (_int, [_vp, _int, _i64, _int, _vp, _i64, _vp, _i32, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qd_cvt_debug_buffers": (None, [_vp, _vp]),
    "qd_cvt_mask_losers": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "qd_insert_scratch_bytes": (_sz, [_i64]),
    "qd_insert_scratch_init": (_int, [_vp, _i64, _vp]),
    "qd_insert_set_result_mirror": (_int, [_vp, _i64, _vp, _vp]),
    "qd_insert_flags_ptr": (_vp, [_vp, _i64]),
    "qd_flags_or": (_int, [_vp, _int, _vp, _vp]),
    "qd_insert_phase_times": (_int, [_vp, _i64, C.POINTER(_u64), _vp]),
    "qd_best_snapshot_bytes": (_sz, [C.POINTER(Store)]),
    "qd_store_clear": (_int, [C.POINTER(Store), _vp]),
    "qd_store_set_state": (_int, [C.POINTER(Store), _i64, _dbl, _int, _dbl, _dbl, _vp]),
    "qd_store_read_state": (_int, [C.POINTER(Store), _vp, _vp]),
    "qd_dp_push": (_int, [_int, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_vp), _int, _int, _vp]),
    "qd_insert": (_int, [C.POINTER(Store), _i64, _vp, _vp, C.POINTER(_vp), _dbl, _dbl, _vp, _vp, _vp, _int, _vp]),
    "qd_insert_sharded_rows": (_int, [C.POINTER(Store), _i64, _vp, _vp, C.POINTER(_vp), _vp, _i64, _dbl, _dbl, _vp, _vp,
                                      _vp, _int, _vp]),
    "qd_insert_owned_cells": (_int, [C.POINTER(Store), _i64, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _dbl, _dbl, _vp, _vp]),
    "qd_grid_insert": (_int, [C.POINTER(Store), _i64, _vp, _int, _int, C.POINTER(_i32), C.POINTER(_dbl),
                              C.POINTER(_dbl), _dbl, _vp, C.POINTER(_vp), _dbl, _dbl, _vp, _vp, _vp, _vp, _int,
                              _vp]),
    "qd_store_gather": (_int, [_i64, _vp, _int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _vp, _vp, _vp]),
    "qd_fill_normal": (_int, [_vp, _i64, _u64, _u64, _vp]),
    "qd_es_out_of_bounds": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "qd_wsum_partial_bytes": (_sz, [_i64, _i64]),
    "qd_weighted_rows": (_int, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "qd_es_workspace_doubles": (_sz, [_i64]),
    "qd_sep_ask": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qd_sep_tell": (_int, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(CmaScalars), _vp, _vp]),
    "qd_sym_max": (_int, [_i64, _vp, _vp]),
    "qd_sym_eigh_batched": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "qd_emit_variation": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _int, _dbl, _vp, _vp, _vp, _vp]),
    "qd_cma_pool_refresh": (_int, [C.POINTER(CmaPool), _vp, _int, _vp]),
    "qd_cma_pool_ask": (_int, [C.POINTER(CmaPool), _vp, _vp, _vp]),
    "qd_cma_pool_tell": (_int, [C.POINTER(CmaPool), _vp, _int, _vp, _vp, _vp, _vp, _vp]),
    "qd_cma_transform": (_int, [_i64, _vp, _vp, _vp, _vp]),
    "qd_cma_invsqrt": (_int, [_i64, _vp, _vp, _vp, _vp]),
    "qd_cma_ask": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qd_cma_tell": (_int, [_i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           C.POINTER(CmaScalars), _vp, _vp]),
    "qd_lmma_ask": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "qd_lmma_ask_rng": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _u64, _u64, _vp, _vp, _vp, _vp]),
    "qd_mirror_ask_rng": (_int, [_i64, _i64, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "qd_openai_tell": (_int, [_i64, _vp, _dbl, _vp, _vp, _vp, _dbl, _dbl, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
    "qd_kmeans_workspace_doubles": (C.c_size_t, [_i64, _i64]),
    "qd_kmeans_update": (_int, [_vp, _int, _i64, _i64, _vp, _i64, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qd_lmma_set_block_bytes": (_i64, [_i64]),
    "qd_lmma_ask_workspace_doubles": (C.c_size_t, [_i64, _i64, _i64]),
    "qd_fill_normal_zig": (_int, [_vp, _i64, _u64, _u64, _vp]),
    "qd_sep_ask_rng": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _vp]),
    "qd_lmma_tell": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Loads (once) and returns the shared library; raises QdError when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QdError(
                f"{LIB_PATH} not found: build it with `python -m pyribs_b200.build` "
                "(there is no CPU fallback for this package)"
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise QdError(f"qd_b200 error {rc}: {lib().qd_last_error().decode()}")


def launch_count() -> int:
    return int(lib().qd_launch_count())
