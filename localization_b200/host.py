"""ctypes wrapper of libuwbgo_host.so (include/uwbgo_host.h): fleets of ROS-free Localization
instances (the reference's class, src/localization/localization.h:99-128) whose solve() calls are
batched into uwbgo_solve_batch."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from .solver import Solver

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libuwbgo_host.so")
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
SOLVE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_ffi.CTopology), C.POINTER(_ffi.CBatch),
                       C.POINTER(_ffi.CConfig), C.POINTER(_ffi.CResult))


class CLocParams(C.Structure):
    _fields_ = [
        ("trajectory_length", C.c_int32), ("maximum_iteration", C.c_int32),
        ("maximum_velocity", C.c_double), ("distance_outlier", C.c_double),
        ("minimum_optimize_error", C.c_double),
        ("n_nodes", C.c_int32), ("n_antennas", C.c_int32),
        ("nodes_id", _pi), ("nodes_pos", _pd), ("antenna_offset", _pd),
        ("publish_range", C.c_int32), ("publish_pose", C.c_int32), ("publish_twist", C.c_int32),
        ("publish_lidar", C.c_int32), ("publish_imu", C.c_int32), ("reserved", C.c_int32),
        ("filename_prefix", C.c_char_p), ("filename_suffix", C.c_char_p),
    ]


_u32 = C.c_uint32
HOST_SYMBOLS = {
    "uwbgo_fleet_create": (C.c_void_p, [C.c_void_p]),
    "uwbgo_fleet_create_with_solver": (C.c_void_p, [SOLVE_FN, C.c_void_p]),
    "uwbgo_fleet_destroy": (None, [C.c_void_p]),
    "uwbgo_fleet_add": (C.c_int, [C.c_void_p, C.POINTER(CLocParams)]),
    "uwbgo_fleet_size": (C.c_int, [C.c_void_p]),
    "uwbgo_fleet_flush": (C.c_int, [C.c_void_p]),
    "uwbgo_fleet_add_range": (C.c_int, [C.c_void_p, C.c_int, _u32, _u32, _u32, C.c_char_p, C.c_int, C.c_int,
                                        C.c_float, C.c_float, C.c_int]),
    "uwbgo_fleet_add_imu": (C.c_int, [C.c_void_p, C.c_int, _u32, _u32, _u32, C.c_char_p, _pd, _pd]),
    "uwbgo_fleet_add_lidar": (C.c_int, [C.c_void_p, C.c_int, _u32, _u32, _u32, C.c_char_p, C.c_double]),
    "uwbgo_fleet_add_twist": (C.c_int, [C.c_void_p, C.c_int, _u32, _u32, _u32, C.c_char_p, _pd, _pd, _pd]),
    "uwbgo_fleet_add_pose": (C.c_int, [C.c_void_p, C.c_int, _u32, _u32, _u32, C.c_char_p, _pd, _pd, _pd]),
    "uwbgo_fleet_add_range_each": (C.c_int, [C.c_void_p, _u32, _u32, _u32, C.c_char_p, C.c_int, C.c_int,
                                             C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]),
    "uwbgo_fleet_add_imu_each": (C.c_int, [C.c_void_p, _u32, _u32, _u32, C.c_char_p, _pd, _pd]),
    "uwbgo_fleet_add_typed_range_edge": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                                   C.c_double, C.c_int, C.c_int, C.c_int]),
    "uwbgo_fleet_solve": (C.c_int, [C.c_void_p, C.c_int]),
    "uwbgo_fleet_published_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "uwbgo_fleet_published": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _pd, _pd, _pd]),
    "uwbgo_fleet_published_all": (C.c_int, [C.c_void_p, C.c_int, _pd, _pd, _pd]),
    "uwbgo_fleet_stats": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "uwbgo_fleet_last_solve": (C.c_int, [C.c_void_p, C.c_int, _pd, _pi]),
    "uwbgo_fleet_last_error": (C.c_char_p, [C.c_void_p, C.c_int]),
    "uwbgo_fleet_window_poses": (C.c_int, [C.c_void_p, C.c_int, _pd, C.c_int]),
}
_lib = None


def load_host_library():
    global _lib
    if _lib is None:
        _ffi.load_library()  # libuwbgo.so first: libuwbgo_host.so links against it
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `make -C localization_b200/host`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in HOST_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


@dataclass
class LocParams:
    """ROS parameters of the reference node (cfg/*.yaml keys, localization.cpp:58-159)."""
    trajectory_length: int = 10
    maximum_iteration: int = 20
    maximum_velocity: float = 1.0
    distance_outlier: float = 1.0
    minimum_optimize_error: float = 1000.0
    nodes_id: list = field(default_factory=list)     # /uwb/nodesId: anchors..., self last
    nodes_pos: list = field(default_factory=list)    # /uwb/nodesPos, 3 per node
    antenna_offset: list = field(default_factory=list)
    publish_range: bool = False
    publish_pose: bool = False
    publish_twist: bool = False
    publish_lidar: bool = False
    publish_imu: bool = False
    filename_prefix: str | None = None
    filename_suffix: str | None = None

    @staticmethod
    def from_yaml(path: str, **overrides) -> "LocParams":
        """cfg/*.yaml of the reference (robot/..., optimizer/..., publish_flag/...)."""
        import yaml
        with open(path) as f:
            y = yaml.safe_load(f)
        r, o, pf = y.get("robot", {}), y.get("optimizer", {}), y.get("publish_flag", {})
        p = LocParams(trajectory_length=int(r.get("trajectory_length", 10)),
                      maximum_velocity=float(r.get("maximum_velocity", 1.0)),
                      distance_outlier=float(r.get("distance_outlier", 1.0)),
                      maximum_iteration=int(o.get("maximum_iteration", 20)),
                      minimum_optimize_error=float(o.get("minimum_optimize_error", 1000.0)),
                      publish_range=bool(pf.get("range", False)), publish_pose=bool(pf.get("pose", False)),
                      publish_twist=bool(pf.get("twist", False)), publish_lidar=bool(pf.get("lidar", False)),
                      publish_imu=bool(pf.get("imu", False)))
        for k, v in overrides.items():
            setattr(p, k, v)
        return p


class Fleet:
    """Localization instances advanced in lockstep; flush() = one uwbgo_solve_batch per structure."""

    def __init__(self, solver: Solver | None = None, solve_fn=None):
        self._lib = load_host_library()
        self._keep = []
        if solve_fn is not None:  # test infrastructure: any callable with the uwbgo_solve_batch contract
            self._cb = SOLVE_FN(solve_fn)
            self._h = self._lib.uwbgo_fleet_create_with_solver(self._cb, None)
        else:
            if solver is None:
                raise ValueError("Fleet needs a Solver (GPU context)")
            self._solver = solver
            self._h = self._lib.uwbgo_fleet_create(solver._h)
        if not self._h:
            raise RuntimeError("uwbgo_fleet_create failed")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.uwbgo_fleet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, p: LocParams) -> int:
        ids = np.asarray(p.nodes_id, np.int32)
        pos = np.asarray(p.nodes_pos, np.float64).reshape(-1)
        ant = np.asarray(p.antenna_offset, np.float64).reshape(-1)
        c = CLocParams()
        c.trajectory_length, c.maximum_iteration = p.trajectory_length, p.maximum_iteration
        c.maximum_velocity, c.distance_outlier = p.maximum_velocity, p.distance_outlier
        c.minimum_optimize_error = p.minimum_optimize_error
        c.n_nodes, c.n_antennas = len(ids), len(ant) // 3
        c.nodes_id, c.nodes_pos = ids.ctypes.data_as(_pi), pos.ctypes.data_as(_pd)
        c.antenna_offset = ant.ctypes.data_as(_pd) if len(ant) else None
        c.publish_range, c.publish_pose, c.publish_twist = int(p.publish_range), int(p.publish_pose), int(p.publish_twist)
        c.publish_lidar, c.publish_imu = int(p.publish_lidar), int(p.publish_imu)
        c.filename_prefix = p.filename_prefix.encode() if p.filename_prefix else None
        c.filename_suffix = p.filename_suffix.encode() if p.filename_suffix else None
        idx = self._lib.uwbgo_fleet_add(self._h, C.byref(c))
        if idx < 0:
            raise RuntimeError(f"uwbgo_fleet_add failed: {idx}")
        return idx

    def __len__(self):
        return self._lib.uwbgo_fleet_size(self._h)

    def flush(self):
        rc = self._lib.uwbgo_fleet_flush(self._h)
        if rc != 0:
            raise RuntimeError(f"fleet flush failed ({rc}): {self.last_error(0)}")

    # messages ---------------------------------------------------------------------------------
    def add_range(self, member, seq, sec, nsec, frame_id, requester, responder, distance, distance_err, antenna):
        self._lib.uwbgo_fleet_add_range(self._h, member, seq, sec, nsec, frame_id.encode(), requester, responder,
                                        float(distance), float(distance_err), antenna)

    def add_range_each(self, seq, sec, nsec, frame_id, requester, responder, distance, distance_err, antenna):
        d = np.ascontiguousarray(distance, np.float32)
        e = np.ascontiguousarray(distance_err, np.float32)
        assert len(d) == len(self) and len(e) == len(self)
        self._lib.uwbgo_fleet_add_range_each(self._h, seq, sec, nsec, frame_id.encode(), requester, responder,
                                             d.ctypes.data_as(C.POINTER(C.c_float)),
                                             e.ctypes.data_as(C.POINTER(C.c_float)), antenna)

    def add_imu(self, member, seq, sec, nsec, frame_id, q_xyzw, cov9):
        q = np.ascontiguousarray(q_xyzw, np.float64)
        c = np.ascontiguousarray(cov9, np.float64)
        self._lib.uwbgo_fleet_add_imu(self._h, member, seq, sec, nsec, frame_id.encode(), q.ctypes.data_as(_pd),
                                      c.ctypes.data_as(_pd))

    def add_imu_each(self, seq, sec, nsec, frame_id, q_xyzw, cov9):
        q = np.ascontiguousarray(q_xyzw, np.float64).reshape(len(self), 4)
        c = np.ascontiguousarray(cov9, np.float64)
        self._lib.uwbgo_fleet_add_imu_each(self._h, seq, sec, nsec, frame_id.encode(), q.ctypes.data_as(_pd),
                                           c.ctypes.data_as(_pd))

    def add_lidar(self, member, seq, sec, nsec, frame_id, z):
        self._lib.uwbgo_fleet_add_lidar(self._h, member, seq, sec, nsec, frame_id.encode(), float(z))

    def add_twist(self, member, seq, sec, nsec, frame_id, linear, angular, cov36):
        a = [np.ascontiguousarray(x, np.float64) for x in (linear, angular, cov36)]
        self._lib.uwbgo_fleet_add_twist(self._h, member, seq, sec, nsec, frame_id.encode(),
                                        *[x.ctypes.data_as(_pd) for x in a])

    def add_pose(self, member, seq, sec, nsec, frame_id, position, q_xyzw, cov36):
        a = [np.ascontiguousarray(x, np.float64) for x in (position, q_xyzw, cov36)]
        self._lib.uwbgo_fleet_add_pose(self._h, member, seq, sec, nsec, frame_id.encode(),
                                       *[x.ctypes.data_as(_pd) for x in a])

    EDGE_RANGE, EDGE_RANGE_OFFSET = 0, 1

    def add_typed_range_edge(self, member, edge_class, from_age, to_age=-1, to_anchor=-1, measurement=0.0,
                             information=1.0, off_from=0, off_to=0, cauchy=True):
        """EdgeSE3Range (setVertexOffset) / EdgeSE3RangeOffset (setParameterId) between existing vertices;
        raises on an offset the solve path cannot carry"""
        rc = self._lib.uwbgo_fleet_add_typed_range_edge(self._h, member, edge_class, from_age, to_age, to_anchor,
                                                        float(measurement), float(information), off_from, off_to,
                                                        int(cauchy))
        if rc != 0:
            raise ValueError(self.last_error(member))

    def solve(self, member):
        self._lib.uwbgo_fleet_solve(self._h, member)

    # results ----------------------------------------------------------------------------------
    def published(self, member=0):
        """(realtime [n][8], optimized [n][8], error [n]); rows = stamp x y z qx qy qz qw"""
        n = self._lib.uwbgo_fleet_published_count(self._h, member)
        rt, op, er = np.zeros((n, 8)), np.zeros((n, 8)), np.zeros(n)
        if n:
            self._lib.uwbgo_fleet_published_all(self._h, member, rt.ctypes.data_as(_pd), op.ctypes.data_as(_pd),
                                                er.ctypes.data_as(_pd))
        return rt, op, er

    def stats(self, member=0):
        s = (C.c_int64 * 7)()
        self._lib.uwbgo_fleet_stats(self._h, member, s)
        return dict(zip(("solves", "rejected", "skipped", "errors", "fleet_windows", "fleet_batches", "settled_alone"),
                        list(s)))

    def last_solve(self, member=0):
        chi2, st = np.zeros(4), np.zeros(4, np.int32)
        self._lib.uwbgo_fleet_last_solve(self._h, member, chi2.ctypes.data_as(_pd), st.ctypes.data_as(_pi))
        return chi2, st

    def last_error(self, member=0) -> str:
        return self._lib.uwbgo_fleet_last_error(self._h, member).decode()

    def window_poses(self, member=0, capacity=1024):
        buf = np.zeros((capacity, 3))
        n = self._lib.uwbgo_fleet_window_poses(self._h, member, buf.ctypes.data_as(_pd), capacity)
        return buf[:max(n, 0)].copy()
