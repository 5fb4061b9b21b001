"""TEST INFRASTRUCTURE, NOT PRODUCT CODE — ctypes wrapper of oracle/libuwbgo_oracle.so, the
plain-C CPU restatement of the reference's hot path (see oracle/uwbgo_oracle.c for the
reference file:line citations and the "parity unpinned" statement).

The structures of include/uwbgo.h are reused (imported from localization_b200._ffi /
.graph): the oracle speaks the same ABI so that tests feed both sides the same buffers."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from localization_b200 import _ffi
from localization_b200.graph import Batch, Config, Result, Topology

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuwbgo_oracle.so")
_lib = None


def build(force: bool = False):
    src = os.path.join(_HERE, "uwbgo_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        lib.uwbgo_oracle_solve_batch.restype = C.c_int
        lib.uwbgo_oracle_solve_batch.argtypes = [C.POINTER(_ffi.CTopology), C.POINTER(_ffi.CBatch),
                                                 C.POINTER(_ffi.CConfig), C.POINTER(_ffi.CResult),
                                                 pd, C.c_int]
        lib.uwbgo_oracle_linearize_batch.restype = C.c_int
        lib.uwbgo_oracle_linearize_batch.argtypes = [C.POINTER(_ffi.CTopology), C.POINTER(_ffi.CBatch),
                                                     C.POINTER(_ffi.CConfig), pd, pd, pd, pd, C.c_int]
        lib.uwbgo_oracle_factor_solve_batch.restype = C.c_int
        lib.uwbgo_oracle_factor_solve_batch.argtypes = [C.c_int32, C.c_int64, pd, pd, pd, pd, pd, pi]
        lib.uwbgo_oracle_log.restype = C.c_double
        lib.uwbgo_oracle_log.argtypes = [C.c_double]
        lib.uwbgo_oracle_quat_to_R.argtypes = [pd, pd]
        lib.uwbgo_oracle_R_to_quat.argtypes = [pd, pd]
        _lib = lib
    return _lib


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def solve(topo: Topology, batch: Batch, cfg: Config | None = None, n_threads: int = 0,
          trace: bool = False, edge_chi2: bool = False, marginals: bool = False) -> Result:
    lib = load()
    cfg = cfg or Config()
    batch.check(topo)
    W, N = batch.n_windows, topo.n_poses
    res = Result.empty(W, N, topo.n_edges if edge_chi2 else None, marginals)
    tr = np.zeros((W, max(cfg.max_iterations, 1), 4)) if trace else None
    t, b, c, r = topo.c_struct(), batch.c_struct(), cfg.c_struct(), res.c_struct()
    nt = n_threads or (os.cpu_count() or 1)
    rc = lib.uwbgo_oracle_solve_batch(C.byref(t), C.byref(b), C.byref(c), C.byref(r),
                                      _pd(tr) if trace else None, nt)
    if rc != 0:
        raise RuntimeError(f"oracle solve failed: {rc}")
    res.trace = tr
    return res


def linearize(topo: Topology, batch: Batch, cfg: Config | None = None, n_threads: int = 0):
    lib = load()
    cfg = cfg or Config()
    batch.check(topo)
    W, N = batch.n_windows, topo.n_poses
    Hd = np.zeros((W, N, 6, 6))
    Ho = np.zeros((W, max(N - 1, 1), 6, 6))
    bb = np.zeros((W, N, 6))
    chi = np.zeros((W, 2))
    t, b, c = topo.c_struct(), batch.c_struct(), cfg.c_struct()
    nt = n_threads or (os.cpu_count() or 1)
    rc = lib.uwbgo_oracle_linearize_batch(C.byref(t), C.byref(b), C.byref(c), _pd(Hd), _pd(Ho),
                                          _pd(bb), _pd(chi), nt)
    if rc != 0:
        raise RuntimeError(f"oracle linearize failed: {rc}")
    return Hd, Ho[:, :N - 1], bb, chi


def factor_solve(H_diag, H_off, b, lam):
    lib = load()
    H_diag = np.ascontiguousarray(H_diag, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    lam = np.ascontiguousarray(lam, np.float64)
    W, N = b.shape[0], b.shape[1]
    H_off = np.ascontiguousarray(H_off, np.float64) if N > 1 else np.zeros(1)
    x = np.zeros((W, N, 6))
    ok = np.zeros(W, np.int32)
    rc = lib.uwbgo_oracle_factor_solve_batch(N, W, _pd(H_diag), _pd(H_off), _pd(b), _pd(lam), _pd(x),
                                             ok.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError(f"oracle factor_solve failed: {rc}")
    return x, ok


def det_log(x: float) -> float:
    return float(load().uwbgo_oracle_log(float(x)))
