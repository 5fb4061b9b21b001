"""ResidentFleet: ctypes wrapper of uwbgo_stream (include/uwbgo.h, "resident fleet") -- the sliding windows of W
robots stay on the device; a step sends one range message per robot and returns the newest poses.  Mirrors
Localization::addRangeEdge -> solve() (reference localization.cpp:297-376) on the Robot ring (robot.cpp:75-110)."""
import ctypes as C

import numpy as np

from . import _ffi
from .graph import Config, Topology
from .solver import Solver, UwbgoError, pinned_empty


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


FLAG_REJECTED = 8  # UWBGO_FLAG_REJECTED


class ResidentFleet:
    def __init__(self, solver: Solver, n_poses: int, anchors: np.ndarray, n_windows: int, v_max: float, cfg: Config):
        self._lib = _ffi.load_library()
        self._solver = solver  # keeps the context alive
        self.N, self.W = int(n_poses), int(n_windows)
        self.anchors = np.ascontiguousarray(anchors, dtype=np.float64).reshape(-1, 3)
        self.A = self.anchors.shape[0]
        self._h = C.c_void_p()
        c = cfg.c_struct()
        rc = self._lib.uwbgo_stream_create(solver._h, self.N, self.A, self.W, _pd(self.anchors), float(v_max), C.byref(c),
                                           C.byref(self._h))
        if rc != 0:
            raise UwbgoError(f"uwbgo_stream_create failed: {rc}")
        self._newest = pinned_empty((self.W, 3))
        self._chi2 = pinned_empty((self.W, 4))
        self._status = pinned_empty((self.W, 4), np.int32)

    def _check(self, rc, what):
        if rc != 0:
            raise UwbgoError(f"{what} failed: {rc}: {self._lib.uwbgo_last_error().decode()}")

    def load(self, pose_t, anchor_of_pose, distance, distance_err, dt):
        pose_t = np.ascontiguousarray(pose_t, dtype=np.float64)
        aop = np.ascontiguousarray(anchor_of_pose, dtype=np.int32)
        d = np.ascontiguousarray(distance, dtype=np.float32)
        e = np.ascontiguousarray(distance_err, dtype=np.float32)
        dt = np.ascontiguousarray(dt, dtype=np.float64)
        assert pose_t.shape == (self.W, self.N, 3) and aop.shape in ((self.N,), (self.W, self.N))
        assert d.shape == (self.W, self.N) and e.shape == (self.W, self.N) and dt.shape == (self.W, self.N - 1)
        # anchor_of_pose [N]: one anchor sequence for the fleet; [W][N]: one per robot
        fn, what = ((self._lib.uwbgo_stream_load, "uwbgo_stream_load") if aop.ndim == 1 else
                    (self._lib.uwbgo_stream_load_robots, "uwbgo_stream_load_robots"))
        self._check(fn(self._h, _pd(pose_t), aop.ctypes.data_as(C.POINTER(C.c_int32)), _pf(d), _pf(e), _pd(dt)), what)

    def step(self, anchor, distance, distance_err, dt):
        """one range message per robot from `anchor` (an int: the same anchor for the fleet; an array [W]: one
        per robot, for a fleet loaded with anchor_of_pose [W][N]); returns (newest pose [W][3], chi2 [W][4],
        status [W][4]) -- views of buffers the next step overwrites"""
        d = np.ascontiguousarray(distance, dtype=np.float32)
        e = np.ascontiguousarray(distance_err, dtype=np.float32)
        dt = np.ascontiguousarray(dt, dtype=np.float64)
        assert d.shape == (self.W,) and e.shape == (self.W,) and dt.shape == (self.W,)
        if np.ndim(anchor) == 0:
            self._check(self._lib.uwbgo_stream_step(self._h, int(anchor), _pf(d), _pf(e), _pd(dt), _pd(self._newest),
                                                    _pd(self._chi2), self._status.ctypes.data_as(C.POINTER(C.c_int32))),
                        "uwbgo_stream_step")
        else:
            a = np.ascontiguousarray(anchor, dtype=np.int32)
            assert a.shape == (self.W,)
            self._check(self._lib.uwbgo_stream_step_robots(self._h, a.ctypes.data_as(C.POINTER(C.c_int32)), _pf(d), _pf(e),
                                                           _pd(dt), _pd(self._newest), _pd(self._chi2),
                                                           self._status.ctypes.data_as(C.POINTER(C.c_int32))),
                        "uwbgo_stream_step_robots")
        return self._newest, self._chi2, self._status

    def set_outlier_gate(self, distance_outlier: float):
        """robot/distance_outlier of addRangeEdge (localization.cpp:305-313); negative = off.  Per-robot fleets only:
        a refused robot keeps its window and reports status[2] == FLAG_REJECTED"""
        self._check(self._lib.uwbgo_stream_set_outlier_gate(self._h, float(distance_outlier)), "uwbgo_stream_set_outlier_gate")

    def read(self) -> np.ndarray:
        out = np.empty((self.W, self.N, 3))
        self._check(self._lib.uwbgo_stream_read(self._h, _pd(out)), "uwbgo_stream_read")
        return out

    def close(self):
        if self._h:
            self._lib.uwbgo_stream_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def chain_topology(n_poses: int, n_anchors: int, anchor_of_pose) -> Topology:
    """the window addRangeEdge builds with the given anchor per pose (what uwbgo_stream solves in a step)"""
    from .graph import EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE
    edges = []
    for k in range(n_poses):
        edges.append((EDGE_RANGE_ANCHOR, k, int(anchor_of_pose[k]), 0, 1))
        if k > 0:
            edges.append((EDGE_RANGE_POSE, k - 1, k, 0, 1))
    return Topology.from_edges(n_poses, n_anchors, 0, edges)
