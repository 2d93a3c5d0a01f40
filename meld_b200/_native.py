"""ctypes binding of libmeld_b200.so (the C-ABI declared in include/meld_b200.h).

PyTorch tensors are only the device-memory container: every call passes raw
``data_ptr()`` values and the current CUDA stream.  There is no CPU fallback --
if the shared library is missing or there is no CUDA device the product path
raises immediately.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmeld_b200.so")

_lib = None

FLAG_KEEP_KNN_KERNEL = 1
FLAG_SIMT_SEARCH = 2

# name -> (restype, argtypes); mirrors include/meld_b200.h one to one.
_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
_pi64, _pint, _pdbl = C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_double)
SIGNATURES = {
    "meld_b200_version": (C.c_int, []),
    "meld_b200_last_error": (C.c_char_p, []),
    "meld_b200_launch_count": (C.c_int64, []),
    "meld_b200_sync_count": (C.c_int64, []),
    "meld_b200_device_info": (C.c_int, [_pint, _pint, _pint]),
    "meld_b200_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "meld_b200_knn_graph_build": (C.c_int, [_vp, _i64, _i64, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _vp, C.POINTER(_vp)]),
    "meld_b200_dense_graph_build": (C.c_int, [_vp, _i64, _i64, _i32, _dbl, _dbl, _dbl, _i32, _vp, C.POINTER(_vp)]),
    "meld_b200_knn_candidates": (C.c_int, [_vp, _i64, _i64, _i32, _dbl, _dbl, _dbl, _i64, _i64, _i32, _vp, C.POINTER(_vp)]),
    "meld_b200_cands_info": (C.c_int, [_vp, _pi64, _pi64, _pint, _pi64]),
    "meld_b200_cands_export": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "meld_b200_cands_destroy": (C.c_int, [_vp]),
    "meld_b200_graph_from_candidates": (C.c_int, [_i64, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _dbl, _dbl, _dbl, _dbl, _i32, _vp, C.POINTER(_vp)]),
    "meld_b200_stage2_begin": (C.c_int, [_vp, _vp, _pi64, _i32, _i32, _dbl, _dbl, _dbl, _dbl, _vp, C.POINTER(_vp), _pi64]),
    "meld_b200_stage2_records": (C.c_int, [_vp, _vp, _vp]),
    "meld_b200_stage2_assemble": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "meld_b200_stage2_finish": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "meld_b200_stage2_destroy": (C.c_int, [_vp]),
    "meld_b200_debug_candidate_search": (C.c_int, [_vp, _i64, _i64, _i32, _dbl, _dbl, _dbl, _i32, _vp, _vp, _vp, _pi64]),
    "meld_b200_graph_from_csr": (C.c_int, [_i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "meld_b200_graph_info": (C.c_int, [_vp, _pi64, _pi64, _pi64, _pi64]),
    "meld_b200_graph_export_csr": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "meld_b200_graph_permutation": (C.c_int, [_vp, _vp, _pint, _vp]),
    "meld_b200_graph_knn_kernel_nnz": (C.c_int, [_vp, _pi64]),
    "meld_b200_graph_export_knn_kernel": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "meld_b200_graph_build_stats": (C.c_int, [_vp, _pi64]),
    "meld_b200_graph_build_times": (C.c_int, [_vp, _pdbl]),
    "meld_b200_graph_destroy": (C.c_int, [_vp]),
    "meld_b200_estimate_lmax": (C.c_int, [_vp, _i32, _dbl, _vp, _pdbl, _pint]),
    "meld_b200_cheby_filter": (C.c_int, [_vp, _dbl, _pdbl, _i32, _vp, _i32, _vp, _vp]),
    "meld_b200_cheby_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _dbl, _dbl, _dbl, _dbl, _dbl, _i32, _vp]),
    "meld_b200_cheby_sweep": (C.c_int, [_vp, _dbl, _pdbl, _i32, _i32, _vp, _i32, _vp, _vp]),
    "meld_b200_graph_permute_signal": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "meld_b200_graph_row_slice": (C.c_int, [_vp, _i64, _i64, _vp, C.POINTER(_vp)]),
    "meld_b200_dist_create": (C.c_int, [_i32, _i32, _i64, _i32, _vp, C.POINTER(_vp)]),
    "meld_b200_dist_handle_bytes": (C.c_int, []),
    "meld_b200_dist_export": (C.c_int, [_vp, _vp]),
    "meld_b200_dist_connect": (C.c_int, [_vp, _vp]),
    "meld_b200_dist_connect_local": (C.c_int, [_vp, C.POINTER(_vp), _i32]),
    "meld_b200_graph_mark_columns": (C.c_int, [_vp, _vp, _vp]),
    "meld_b200_graph_set_halo": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "meld_b200_dist_error": (C.c_int, [_vp, _pint]),
    "meld_b200_dist_destroy": (C.c_int, [_vp]),
    "meld_b200_cheby_filter_dist": (C.c_int, [_vp, _vp, _dbl, _pdbl, _i32, _vp, _i32, _vp, _vp]),
    "meld_b200_estimate_lmax_dist": (C.c_int, [_vp, _vp, _i32, _dbl, _vp, _pdbl, _pint]),
    "meld_b200_release_workspace": (C.c_int, []),
    "meld_b200_indicator_matrix": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "meld_b200_l1_normalize_rows": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
}


class NativeError(RuntimeError):
    """A libmeld_b200 call returned a negative status."""


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # the library is git-ignored: a fresh checkout builds it on first use when nvcc is there
        import shutil

        if shutil.which(os.environ.get("NVCC", "nvcc")) and not os.environ.get("MELD_B200_NO_AUTOBUILD"):
            from . import build as _build

            _build.build(verbose=False)
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "libmeld_b200.so not found at {}: build it with `python -m meld_b200.build` "
            "(there is no CPU fallback)".format(LIB_PATH)
        )
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().meld_b200_last_error()
        raise NativeError("{} failed (status {}): {}".format(what, rc, msg.decode() if msg else ""))


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise NativeError("meld_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def current_stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def from_numpy_readonly(arr):
    """``torch.from_numpy`` for arrays the engine only reads.  ``DataFrame.values`` is read-only under pandas'
    copy-on-write and torch warns about tensors over non-writable memory; nothing here writes to its inputs."""
    import warnings

    import torch

    if arr.flags.writeable:
        return torch.from_numpy(arr)
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message="The given NumPy array is not writable", category=UserWarning)
        return torch.from_numpy(arr)


def ptr(t):
    """Raw device pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def set_tuning(**kw):
    for k, v in kw.items():
        check(lib().meld_b200_set_tuning(k.encode(), int(v)), "set_tuning")


def device_info():
    sm, maj, mnr = C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib().meld_b200_device_info(C.byref(sm), C.byref(maj), C.byref(mnr)), "device_info")
    return sm.value, maj.value, mnr.value
