"""ctypes view of include/uwbgo.h (the C ABI of the batched window solver).

The structures mirror the header field for field.  `load_library()` loads the in-tree
`libuwbgo.so` built by `localization_b200/csrc/Makefile`; there is no fallback of any kind:
a missing library is an ImportError, a missing GPU makes `uwbgo_create` fail.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 5
SHARED_ANCHORS = 1
DIAG_INFO = 2

EDGE_RANGE_ANCHOR = 0
EDGE_RANGE_POSE = 1
EDGE_PRIOR = 2
EDGE_SE3 = 3

CHI2_STRIDE = 4
STATUS_STRIDE = 4
FLAG_CHOL_FAIL = 1
FLAG_TERMINATED = 2
FLAG_NONFINITE = 4

E_INVALID, E_TOPOLOGY, E_CUDA, E_NODEVICE, E_NOMEM = -1, -2, -3, -4, -5

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class CTopology(C.Structure):
    _fields_ = [
        ("n_poses", C.c_int32), ("n_anchors", C.c_int32), ("n_antennas", C.c_int32),
        ("n_edges", C.c_int32),
        ("edge_kind", _pi), ("edge_a", _pi), ("edge_b", _pi), ("edge_ant", _pi),
        ("edge_robust", _pi), ("edge_ant_b", _pi),
    ]


_pf = C.POINTER(C.c_float)


class CRangeMsgs(C.Structure):
    _fields_ = [
        ("distance", _pf), ("distance_err", _pf), ("dt_anchor", _pd), ("dt_pose", _pd), ("v_max", C.c_double),
    ]


class CBatch(C.Structure):
    _fields_ = [
        ("n_windows", C.c_int64),
        ("pose_t", _pd), ("pose_R", _pd), ("oplus_count", _pi), ("anchors", _pd),
        ("ant_offsets", _pd), ("range_d", _pd), ("range_info", _pd),
        ("prior_Z", _pd), ("prior_info", _pd), ("se3_Z", _pd), ("se3_info", _pd),
        ("range_msgs", C.POINTER(CRangeMsgs)), ("shared", C.c_int32), ("reserved", C.c_int32),
    ]


class CConfig(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_int32), ("max_trials", C.c_int32),
        ("orthogonalize_after", C.c_int32), ("reserved", C.c_int32),
        ("tau", C.c_double), ("good_step_lower", C.c_double), ("good_step_upper", C.c_double),
        ("kernel_delta", C.c_double), ("jacobian_delta", C.c_double),
    ]


class CResult(C.Structure):
    _fields_ = [
        ("pose_t", _pd), ("pose_R", _pd), ("oplus_count", _pi), ("chi2", _pd), ("status", _pi),
        ("edge_chi2", _pd), ("marginal", _pd), ("marginal_ok", _pi),
    ]


# every symbol include/uwbgo.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "uwbgo_abi_version": (C.c_int, []),
    "uwbgo_config_default": (None, [C.POINTER(CConfig)]),
    "uwbgo_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "uwbgo_destroy": (None, [_vp]),
    "uwbgo_last_error": (C.c_char_p, []),
    "uwbgo_set_pipeline": (C.c_int, [_vp, C.c_int64, C.c_int]),
    "uwbgo_set_window_path": (C.c_int, [_vp, C.c_int64]),
    "uwbgo_host_alloc": (_vp, [C.c_size_t]),
    "uwbgo_host_free": (None, [_vp]),
    "uwbgo_solve_batch": (C.c_int, [_vp, C.POINTER(CTopology), C.POINTER(CBatch),
                                    C.POINTER(CConfig), C.POINTER(CResult)]),
    "uwbgo_solve_batch_device": (C.c_int, [_vp, C.POINTER(CTopology), C.POINTER(CBatch),
                                           C.POINTER(CConfig), C.POINTER(CResult), _vp]),
    "uwbgo_linearize_batch": (C.c_int, [_vp, C.POINTER(CTopology), C.POINTER(CBatch),
                                        C.POINTER(CConfig), _pd, _pd, _pd, _pd]),
    "uwbgo_linearize_batch_device": (C.c_int, [_vp, C.POINTER(CTopology), C.POINTER(CBatch),
                                               C.POINTER(CConfig), _pd, _pd, _pd, _pd, _vp]),
    "uwbgo_factor_solve_batch": (C.c_int, [_vp, C.c_int32, C.c_int64, _pd, _pd, _pd, _pd, _pd, _pi]),
    "uwbgo_factor_solve_batch_device": (C.c_int, [_vp, C.c_int32, C.c_int64, _pd, _pd, _pd, _pd,
                                                  _pd, _pi, _vp]),
    "uwbgo_stream_create": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, _pd, C.c_double, C.POINTER(CConfig),
                                      C.POINTER(_vp)]),
    "uwbgo_stream_destroy": (None, [_vp]),
    "uwbgo_stream_load": (C.c_int, [_vp, _pd, _pi, C.POINTER(C.c_float), C.POINTER(C.c_float), _pd]),
    "uwbgo_stream_step": (C.c_int, [_vp, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float), _pd, _pd, _pd, _pi]),
    "uwbgo_stream_load_robots": (C.c_int, [_vp, _pd, _pi, C.POINTER(C.c_float), C.POINTER(C.c_float), _pd]),
    "uwbgo_stream_step_robots": (C.c_int, [_vp, _pi, C.POINTER(C.c_float), C.POINTER(C.c_float), _pd, _pd, _pd, _pi]),
    "uwbgo_stream_set_outlier_gate": (C.c_int, [_vp, C.c_double]),
    "uwbgo_stream_read": (C.c_int, [_vp, _pd]),
    "uwbgo_launch_count": (C.c_int64, [_vp]),
    "uwbgo_last_path": (C.c_int, [_vp]),
    "uwbgo_set_profiling": (C.c_int, [_vp, C.c_int]),
    "uwbgo_last_kernel_ms": (C.c_double, [_vp]),
    "uwbgo_mean_kernel_ms": (C.c_double, [_vp, C.c_int]),
    "uwbgo_measure_fp64_peak": (C.c_double, [_vp, _pd]),
    "uwbgo_selftest_math": (C.c_int, [_vp, C.c_uint64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libuwbgo.so")
_lib = None


def load_library(path: str | None = None):
    """Load libuwbgo.so and type its entry points.  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("UWBGO_LIB") or LIB_PATH  # UWBGO_LIB: developer override (A/B builds)
    if not os.path.exists(p):
        raise ImportError(
            f"{p} not found: build it with `make -C localization_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.uwbgo_abi_version() != ABI_VERSION:
        raise ImportError("libuwbgo.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib
