"""Solver: one GPU context of libuwbgo.so (replaces the solver chain the reference builds in
Localization::Localization, src/localization/localization.cpp:44-52, and the call
optimizer.initializeOptimization(); optimizer.optimize(iteration_max) at :168-170)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .graph import Batch, Config, Result, Topology


class UwbgoError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"uwbgo error {code}: {text}")
        self.code = code


class Solver:
    def __init__(self, device: int = 0):
        self._lib = _ffi.load_library()
        h = C.c_void_p()
        rc = self._lib.uwbgo_create(device, C.byref(h))
        if rc != 0:
            raise UwbgoError(rc, self._lib.uwbgo_last_error().decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.uwbgo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise UwbgoError(rc, self._lib.uwbgo_last_error().decode())

    # -- host-array entry points ------------------------------------------------------------
    def solve(self, topo: Topology, batch: Batch, cfg: Config | None = None,
              out: Result | None = None, edge_chi2: bool = False, marginals: bool = False) -> Result:
        """edge_chi2 / marginals: the extras of uwbgo_result (per-edge chi2 of the last trial, covariance block of the
        newest pose)"""
        cfg = cfg or Config()
        batch.check(topo)
        res = out or Result.empty(batch.n_windows, topo.n_poses, topo.n_edges if edge_chi2 else None, marginals)
        t, b, c, r = topo.c_struct(), batch.c_struct(), cfg.c_struct(), res.c_struct()
        self._check(self._lib.uwbgo_solve_batch(self._h, C.byref(t), C.byref(b), C.byref(c), C.byref(r)))
        return res

    def linearize(self, topo: Topology, batch: Batch, cfg: Config | None = None):
        """One computeActiveErrors + buildSystem.  Returns H_diag [W][N][6][6],
        H_off [W][N-1][6][6], b [W][N][6], chi2 [W][2] = (plain, robust)."""
        cfg = cfg or Config()
        batch.check(topo)
        W, N = batch.n_windows, topo.n_poses
        Hd = np.zeros((W, N, 6, 6))
        Ho = np.zeros((W, max(N - 1, 1), 6, 6))
        bb = np.zeros((W, N, 6))
        chi = np.zeros((W, 2))
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        t, b, c = topo.c_struct(), batch.c_struct(), cfg.c_struct()
        self._check(self._lib.uwbgo_linearize_batch(self._h, C.byref(t), C.byref(b), C.byref(c),
                                                    pd(Hd), pd(Ho), pd(bb), pd(chi)))
        return Hd, Ho[:, :N - 1], bb, chi

    def factor_solve(self, H_diag, H_off, b, lam):
        """(H + lambda I) x = b for W block-tridiagonal systems.  Returns x [W][N][6], ok [W]."""
        H_diag = np.ascontiguousarray(H_diag, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        lam = np.ascontiguousarray(lam, np.float64)
        W, N = b.shape[0], b.shape[1]
        H_off = np.ascontiguousarray(H_off, np.float64) if N > 1 else np.zeros(1)
        x = np.zeros((W, N, 6))
        ok = np.zeros(W, np.int32)
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self._lib.uwbgo_factor_solve_batch(
            self._h, N, W, pd(H_diag), pd(H_off), pd(b), pd(lam), pd(x),
            ok.ctypes.data_as(C.POINTER(C.c_int32))))
        return x, ok

    # -- device-pointer entry points (pointers as ints, e.g. torch.Tensor.data_ptr()) ---------
    def solve_device(self, topo: Topology, cbatch: _ffi.CBatch, cfg: Config, cres: _ffi.CResult,
                     stream: int = 0):
        t, c = topo.c_struct(), cfg.c_struct()
        self._check(self._lib.uwbgo_solve_batch_device(
            self._h, C.byref(t), C.byref(cbatch), C.byref(c), C.byref(cres), C.c_void_p(stream)))

    def linearize_device(self, topo: Topology, cbatch: _ffi.CBatch, cfg: Config, H_diag: int,
                         H_off: int, b: int, chi2: int, stream: int = 0):
        t, c = topo.c_struct(), cfg.c_struct()
        cast = lambda p: C.cast(C.c_void_p(p), C.POINTER(C.c_double))
        self._check(self._lib.uwbgo_linearize_batch_device(
            self._h, C.byref(t), C.byref(cbatch), C.byref(c), cast(H_diag), cast(H_off), cast(b),
            cast(chi2), C.c_void_p(stream)))

    def factor_solve_device(self, N: int, W: int, H_diag: int, H_off: int, b: int, lam: int, x: int,
                            ok: int, stream: int = 0):
        cast = lambda p: C.cast(C.c_void_p(p), C.POINTER(C.c_double))
        self._check(self._lib.uwbgo_factor_solve_batch_device(
            self._h, N, W, cast(H_diag), cast(H_off), cast(b), cast(lam), cast(x),
            C.cast(C.c_void_p(ok), C.POINTER(C.c_int32)), C.c_void_p(stream)))

    # -- introspection / tuning ---------------------------------------------------------------
    def set_pipeline(self, windows_per_chunk: int, n_lanes: int):
        self._check(self._lib.uwbgo_set_pipeline(self._h, windows_per_chunk, n_lanes))

    def set_window_path(self, max_windows: int):
        """batches of up to max_windows windows take the one-CTA-per-window kernel (0 = never, < 0 = default)"""
        self._check(self._lib.uwbgo_set_window_path(self._h, max_windows))

    @property
    def launch_count(self) -> int:
        return int(self._lib.uwbgo_launch_count(self._h))

    @property
    def last_path(self) -> int:
        return int(self._lib.uwbgo_last_path(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.uwbgo_set_profiling(self._h, 1 if on else 0))

    def last_kernel_ms(self) -> float:
        return float(self._lib.uwbgo_last_kernel_ms(self._h))

    def mean_kernel_ms(self, last_n: int) -> float:
        return float(self._lib.uwbgo_mean_kernel_ms(self._h, int(last_n)))

    def measure_fp64_peak(self):
        ms = C.c_double(0.0)
        flops = self._lib.uwbgo_measure_fp64_peak(self._h, C.byref(ms))
        return float(flops), float(ms.value)


    def selftest_math(self, n_operands: int, mode: int = 0, seed: int = 1):
        """uwbgo_selftest_math: (compared, mismatching, flagged) counts."""
        counts = (C.c_int64 * 3)()
        self._check(self._lib.uwbgo_selftest_math(self._h, seed, n_operands, mode, counts))
        return int(counts[0]), int(counts[1]), int(counts[2])


class _PinnedOwner:
    def __init__(self, lib, ptr):
        self._lib, self._ptr = lib, ptr

    def __del__(self):
        try:
            self._lib.uwbgo_host_free(self._ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array over page-locked memory from uwbgo_host_alloc; freed with the array."""
    lib = _ffi.load_library()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = max(count * dtype.itemsize, 1)
    p = lib.uwbgo_host_alloc(n)
    if not p:
        raise MemoryError("uwbgo_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    buf._owner = _PinnedOwner(lib, p)  # the array's base keeps buf, buf keeps the allocation
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
