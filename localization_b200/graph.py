"""Windows of the range graph as plain arrays: what the g2o graph of the reference holds at the
moment `Localization::solve()` runs (reference src/localization/localization.cpp:164-192).

Topology = graph structure shared by every window of a batch (edges in g2o insertion order);
Batch    = per-window numbers, window-major, FP64/int32, C-contiguous;
Config   = g2o's Levenberg-Marquardt constants;  Result = estimates, chi2 and status.
These classes only hold and validate data and build the ctypes structures of include/uwbgo.h.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from ._ffi import EDGE_PRIOR, EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE, EDGE_SE3


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pd(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


@dataclass
class Topology:
    """Edge list of one window in g2o insertion order (the order g2o accumulates H and b)."""
    n_poses: int
    n_anchors: int
    n_antennas: int
    edge_kind: np.ndarray
    edge_a: np.ndarray
    edge_b: np.ndarray
    edge_ant: np.ndarray
    edge_robust: np.ndarray
    edge_ant_b: np.ndarray | None = None  # vertex-1 antenna numbers (offset[1] / pidTo); None = all 0

    def __post_init__(self):
        self.edge_kind = _i32(self.edge_kind)
        self.edge_a = _i32(self.edge_a)
        self.edge_b = _i32(self.edge_b)
        self.edge_ant = _i32(self.edge_ant)
        self.edge_robust = _i32(self.edge_robust)
        n = len(self.edge_kind)
        if self.edge_ant_b is not None:
            self.edge_ant_b = _i32(self.edge_ant_b)
            if len(self.edge_ant_b) != n:
                raise ValueError("edge arrays differ in length")
            if not self.edge_ant_b.any():
                self.edge_ant_b = None
        for a in (self.edge_a, self.edge_b, self.edge_ant, self.edge_robust):
            if len(a) != n:
                raise ValueError("edge arrays differ in length")

    @property
    def n_edges(self) -> int:
        return len(self.edge_kind)

    def counts(self):
        k = self.edge_kind
        er = int(np.sum((k == EDGE_RANGE_ANCHOR) | (k == EDGE_RANGE_POSE)))
        return er, int(np.sum(k == EDGE_PRIOR)), int(np.sum(k == EDGE_SE3))

    def c_struct(self) -> _ffi.CTopology:
        t = _ffi.CTopology()
        t.n_poses, t.n_anchors, t.n_antennas, t.n_edges = (
            self.n_poses, self.n_anchors, self.n_antennas, self.n_edges)
        t.edge_kind, t.edge_a, t.edge_b = _pi(self.edge_kind), _pi(self.edge_a), _pi(self.edge_b)
        t.edge_ant, t.edge_robust = _pi(self.edge_ant), _pi(self.edge_robust)
        t.edge_ant_b = _pi(self.edge_ant_b)
        return t

    # -- the window shapes Localization builds -------------------------------------------------
    @staticmethod
    def from_edges(n_poses, n_anchors, n_antennas, edges) -> "Topology":
        """edges: iterable of (kind, a, b, ant, robust) or (kind, a, b, ant, robust, ant_b)."""
        edges = [tuple(x) + (0,) * (6 - len(x)) for x in edges]
        e = np.asarray(edges, dtype=np.int32).reshape(-1, 6)
        return Topology(n_poses, n_anchors, n_antennas, e[:, 0], e[:, 1], e[:, 2], e[:, 3], e[:, 4], e[:, 5])

    @staticmethod
    def uwb_chain(n_poses: int, n_anchors: int, antennas: int = 0, imu: bool = False,
                  lidar: bool = False) -> "Topology":
        """Window produced by Localization::addRangeEdge (localization.cpp:297-376) with a new
        vertex per range message: per pose k an anchor range edge (anchor k mod A, antenna
        1 + k mod K when K > 0) followed by the zero-length trajectory edge (k-1, k)
        (localization.cpp:331-340); optional EdgeSE3Prior edges per pose except the newest, in
        the order the callbacks insert them: lidar (localization.cpp:462-496) then IMU
        (localization.cpp:499-535)."""
        edges = []
        for k in range(n_poses):
            ant = 1 + k % antennas if antennas > 0 else 0
            edges.append((EDGE_RANGE_ANCHOR, k, k % n_anchors, ant, 1))
            if k > 0:
                edges.append((EDGE_RANGE_POSE, k - 1, k, 0, 1))
            if k < n_poses - 1:
                if lidar:
                    edges.append((EDGE_PRIOR, k, 0, 0, 0))
                if imu:
                    edges.append((EDGE_PRIOR, k, 0, 0, 0))
        return Topology.from_edges(n_poses, n_anchors, antennas, edges)

    @staticmethod
    def uwb_twist(n_poses: int, n_anchors: int, antennas: int = 0) -> "Topology":
        """uwb_twist window: vertices are created by Localization::addTwistEdge
        (localization.cpp:438-459: EdgeSE3(prev, new), Cauchy); range messages that arrive
        between twists take the else-branch of addRangeEdge (localization.cpp:348-357): one
        anchor range edge on the newest pose, no new vertex."""
        edges = []
        for k in range(n_poses):
            if k > 0:
                edges.append((EDGE_SE3, k - 1, k, 0, 1))
            ant = 1 + k % antennas if antennas > 0 else 0
            edges.append((EDGE_RANGE_ANCHOR, k, k % n_anchors, ant, 1))
        return Topology.from_edges(n_poses, n_anchors, antennas, edges)


    @staticmethod
    def uwb_pose(n_poses: int, n_anchors: int, keyframe_len: int = 4, antennas: int = 0) -> "Topology":
        """uwb_pose window: vertices are created by Localization::addPoseEdge
        (localization.cpp:254-290): EdgeSE3(key_vertex, new) with Cauchy kernel, where key_vertex only
        moves when the keyframe (frame_id) changes, i.e. a star per keyframe, not a chain; range
        messages in between take the merged-covariance branch (localization.cpp:348-357)."""
        edges = []
        for k in range(n_poses):
            if k > 0:
                edges.append((EDGE_SE3, ((k - 1) // keyframe_len) * keyframe_len, k, 0, 1))
            ant = 1 + k % antennas if antennas > 0 else 0
            edges.append((EDGE_RANGE_ANCHOR, k, k % n_anchors, ant, 1))
        return Topology.from_edges(n_poses, n_anchors, antennas, edges)

    def parents(self) -> np.ndarray:
        """the one older neighbour of every pose (-1: none)"""
        par = np.full(self.n_poses, -1, np.int32)
        for k, a, b in zip(self.edge_kind, self.edge_a, self.edge_b):
            if k in (EDGE_RANGE_POSE, EDGE_SE3):
                par[b] = a
        return par


@dataclass
class RangeMsgs:
    """Compact form of the range data (uwbgo_range_msgs): message fields instead of edge parameters; the
    device builds measurement and information as Localization::addRangeEdge does
    (localization.cpp:316-319,331,338,350)."""
    distance: np.ndarray                    # [W][Era] float32
    distance_err: np.ndarray                # [W][Era] float32
    dt_pose: np.ndarray | None = None       # [W][Erp] float64
    dt_anchor: np.ndarray | None = None     # [W][Era] float64, merged-covariance branch
    v_max: float = 1.0

    def __post_init__(self):
        self.distance = np.ascontiguousarray(self.distance, np.float32)
        self.distance_err = np.ascontiguousarray(self.distance_err, np.float32)
        if self.dt_pose is not None:
            self.dt_pose = _f64(self.dt_pose)
        if self.dt_anchor is not None:
            self.dt_anchor = _f64(self.dt_anchor)

    def slice(self, lo, hi):
        s = lambda a: None if a is None else a[lo:hi]
        return RangeMsgs(s(self.distance), s(self.distance_err), s(self.dt_pose), s(self.dt_anchor), self.v_max)

    def expand(self, topo: "Topology"):
        """(range_d, range_info) [W][Er] in FP64: the numpy statement of the device-side expansion (x * x for the
        squares, (v_max * dt) / 3, one IEEE division)"""
        W = self.distance.shape[0]
        kinds = [k for k in topo.edge_kind if k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE)]
        rd, ri = np.zeros((W, len(kinds))), np.zeros((W, len(kinds)))
        ka = kp = 0
        for s_, k in enumerate(kinds):
            if k == EDGE_RANGE_ANCHOR:
                e = self.distance_err[:, ka].astype(np.float64)
                cov = e * e
                if self.dt_anchor is not None:
                    m = (self.v_max * self.dt_anchor[:, ka]) / 3.0
                    cov = cov + m * m
                rd[:, s_] = self.distance[:, ka].astype(np.float64)
                ka += 1
            else:
                m = (self.v_max * self.dt_pose[:, kp]) / 3.0
                cov = m * m
                kp += 1
            ri[:, s_] = 1.0 / cov
        return rd, ri

    def c_struct(self) -> _ffi.CRangeMsgs:
        m = _ffi.CRangeMsgs()
        pf = lambda a: None if a is None else a.ctypes.data_as(_ffi._pf)
        m.distance, m.distance_err = pf(self.distance), pf(self.distance_err)
        m.dt_anchor, m.dt_pose, m.v_max = _pd(self.dt_anchor), _pd(self.dt_pose), float(self.v_max)
        return m

    @property
    def nbytes(self):
        return sum(a.nbytes for a in (self.distance, self.distance_err, self.dt_pose, self.dt_anchor) if a is not None)


@dataclass
class Batch:
    """Per-window numbers of W windows sharing one Topology (see include/uwbgo.h)."""
    pose_t: np.ndarray                      # [W][N][3]
    anchors: np.ndarray | None = None       # [W][A][3]
    range_d: np.ndarray | None = None       # [W][Er]
    range_info: np.ndarray | None = None    # [W][Er]
    pose_R: np.ndarray | None = None        # [W][N][3][3]; None = identity
    oplus_count: np.ndarray | None = None   # [W][N]
    ant_offsets: np.ndarray | None = None   # [K][3]
    prior_Z: np.ndarray | None = None       # [W][Ep][12]  (R row-major 9, t 3)
    prior_info: np.ndarray | None = None    # [W][Ep][6][6]
    se3_Z: np.ndarray | None = None         # [W][Es][12]
    se3_info: np.ndarray | None = None      # [W][Es][6][6]
    range_msgs: RangeMsgs | None = None     # compact form, instead of range_d / range_info
    shared_anchors: bool = False            # anchors is [A][3], one constellation for all windows
    info_diag: bool = False                 # prior_info / se3_info are [W][E*][6]: the diagonals (UWBGO_DIAG_INFO)

    def __post_init__(self):
        self.pose_t = _f64(self.pose_t)
        for name in ("anchors", "range_d", "range_info", "pose_R", "ant_offsets", "prior_Z",
                     "prior_info", "se3_Z", "se3_info"):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, _f64(v))
        if self.oplus_count is not None:
            self.oplus_count = _i32(self.oplus_count)

    @property
    def n_windows(self) -> int:
        return self.pose_t.shape[0]

    def check(self, topo: Topology):
        W, N = self.n_windows, topo.n_poses
        er, ep, es = topo.counts()

        def need(name, arr, size):
            if size == 0:
                return
            if arr is None or arr.size != size:
                raise ValueError(f"{name}: expected {size} elements, got "
                                 f"{None if arr is None else arr.size}")
        if self.pose_t.size != W * N * 3:
            raise ValueError("pose_t must be [W][N][3]")
        if self.pose_R is not None and self.pose_R.size != W * N * 9:
            raise ValueError("pose_R must be [W][N][3][3]")
        if self.oplus_count is not None and self.oplus_count.size != W * N:
            raise ValueError("oplus_count must be [W][N]")
        need("anchors", self.anchors, (1 if self.shared_anchors else W) * topo.n_anchors * 3)
        need("ant_offsets", self.ant_offsets, topo.n_antennas * 3)
        if self.range_msgs is not None:
            if self.range_d is not None or self.range_info is not None:
                raise ValueError("range_msgs replaces range_d / range_info")
            era = int(np.sum(topo.edge_kind == EDGE_RANGE_ANCHOR))
            erp = int(np.sum(topo.edge_kind == EDGE_RANGE_POSE))
            m = self.range_msgs
            need("range_msgs.distance", m.distance, W * era)
            need("range_msgs.distance_err", m.distance_err, W * era)
            need("range_msgs.dt_pose", m.dt_pose, W * erp)
            if m.dt_anchor is not None:
                need("range_msgs.dt_anchor", m.dt_anchor, W * era)
        else:
            need("range_d", self.range_d, W * er)
            need("range_info", self.range_info, W * er)
        need("prior_Z", self.prior_Z, W * ep * 12)
        need("prior_info", self.prior_info, W * ep * (6 if self.info_diag else 36))
        need("se3_Z", self.se3_Z, W * es * 12)
        need("se3_info", self.se3_info, W * es * (6 if self.info_diag else 36))

    def slice(self, lo: int, hi: int) -> "Batch":
        def s(a):
            return None if a is None else a[lo:hi]
        return Batch(pose_t=s(self.pose_t), anchors=self.anchors if self.shared_anchors else s(self.anchors),
                     range_msgs=None if self.range_msgs is None else self.range_msgs.slice(lo, hi),
                     shared_anchors=self.shared_anchors, info_diag=self.info_diag, range_d=s(self.range_d),
                     range_info=s(self.range_info), pose_R=s(self.pose_R),
                     oplus_count=s(self.oplus_count), ant_offsets=self.ant_offsets,
                     prior_Z=s(self.prior_Z), prior_info=s(self.prior_info), se3_Z=s(self.se3_Z),
                     se3_info=s(self.se3_info))

    def c_struct(self) -> _ffi.CBatch:
        b = _ffi.CBatch()
        b.n_windows = self.n_windows
        b.pose_t, b.pose_R, b.oplus_count = _pd(self.pose_t), _pd(self.pose_R), _pi(self.oplus_count)
        b.anchors, b.ant_offsets = _pd(self.anchors), _pd(self.ant_offsets)
        b.range_d, b.range_info = _pd(self.range_d), _pd(self.range_info)
        b.prior_Z, b.prior_info = _pd(self.prior_Z), _pd(self.prior_info)
        b.se3_Z, b.se3_info = _pd(self.se3_Z), _pd(self.se3_info)
        if self.range_msgs is not None:
            self._msgs_c = self.range_msgs.c_struct()   # kept alive with the Batch
            b.range_msgs = C.pointer(self._msgs_c)
        b.shared = (_ffi.SHARED_ANCHORS if self.shared_anchors else 0) | (_ffi.DIAG_INFO if self.info_diag else 0)
        return b

    def expanded(self, topo: Topology) -> "Batch":
        """the same windows with range_d / range_info written out, per-window anchors and full information matrices"""
        rd, ri = (self.range_d, self.range_info) if self.range_msgs is None else self.range_msgs.expand(topo)
        anchors = self.anchors
        if self.shared_anchors:
            anchors = np.ascontiguousarray(np.broadcast_to(self.anchors.reshape(1, -1, 3),
                                                           (self.n_windows, topo.n_anchors, 3)))
        def full(a):
            if a is None or not self.info_diag:
                return a
            d = a.reshape(a.shape[0], -1, 6)
            m = np.zeros(d.shape + (6,))
            m[..., np.arange(6), np.arange(6)] = d
            return m
        return Batch(pose_t=self.pose_t, anchors=anchors, range_d=rd, range_info=ri, pose_R=self.pose_R,
                     oplus_count=self.oplus_count, ant_offsets=self.ant_offsets, prior_Z=self.prior_Z,
                     prior_info=full(self.prior_info), se3_Z=self.se3_Z, se3_info=full(self.se3_info))

    def with_info_diag(self) -> "Batch":
        """the same windows with the information matrices of the 6-D edges passed as their diagonals; raises if a
        matrix has a non-zero (or -0.0) entry off the diagonal"""
        def diag(a):
            if a is None:
                return None
            m = a.reshape(a.shape[0], -1, 6, 6)
            off = m.copy()
            off[..., np.arange(6), np.arange(6)] = 0.0
            if np.any(off.view(np.int64) != 0):
                raise ValueError("information matrix with entries off the diagonal")
            return np.ascontiguousarray(m[..., np.arange(6), np.arange(6)])
        import dataclasses
        return dataclasses.replace(self, prior_info=diag(self.prior_info), se3_info=diag(self.se3_info), info_diag=True)


@dataclass
class Config:
    """g2o OptimizationAlgorithmLevenberg / VertexSE3 / RobustKernelCauchy constants; the
    defaults are g2o's at the commit the reference pins (README.md:26-33), max_iterations is
    optimizer/maximum_iteration (localization.cpp:65)."""
    max_iterations: int = 20
    max_trials: int = 10
    orthogonalize_after: int = 1000
    tau: float = 1e-5
    good_step_lower: float = 1.0 / 3.0
    good_step_upper: float = 2.0 / 3.0
    kernel_delta: float = 1.0
    jacobian_delta: float = 1e-9

    def c_struct(self) -> _ffi.CConfig:
        c = _ffi.CConfig()
        c.max_iterations, c.max_trials = self.max_iterations, self.max_trials
        c.orthogonalize_after, c.reserved = self.orthogonalize_after, 0
        c.tau, c.good_step_lower, c.good_step_upper = self.tau, self.good_step_lower, self.good_step_upper
        c.kernel_delta, c.jacobian_delta = self.kernel_delta, self.jacobian_delta
        return c


@dataclass
class Result:
    pose_t: np.ndarray        # [W][N][3]
    pose_R: np.ndarray        # [W][N][3][3]
    oplus_count: np.ndarray   # [W][N]
    chi2: np.ndarray          # [W][4] plain, robust, g2o-stale, final lambda
    status: np.ndarray        # [W][4] iterations, trials, flags, trials of last iteration
    trace: np.ndarray | None = field(default=None)
    edge_chi2: np.ndarray | None = field(default=None)    # [W][E]  per-edge chi2 of the last trial (sums to chi2[:, 2])
    marginal: np.ndarray | None = field(default=None)     # [W][6][6] covariance block of the newest pose
    marginal_ok: np.ndarray | None = field(default=None)  # [W]

    @staticmethod
    def empty(W: int, N: int, n_edges: int | None = None, marginals: bool = False) -> "Result":
        r = Result(np.zeros((W, N, 3)), np.zeros((W, N, 3, 3)), np.zeros((W, N), np.int32),
                   np.zeros((W, 4)), np.zeros((W, 4), np.int32))
        if n_edges is not None:
            r.edge_chi2 = np.zeros((W, n_edges))
        if marginals:
            r.marginal, r.marginal_ok = np.zeros((W, 6, 6)), np.zeros(W, np.int32)
        return r

    def c_struct(self) -> _ffi.CResult:
        r = _ffi.CResult()
        r.pose_t, r.pose_R, r.oplus_count = _pd(self.pose_t), _pd(self.pose_R), _pi(self.oplus_count)
        r.chi2, r.status = _pd(self.chi2), _pi(self.status)
        r.edge_chi2, r.marginal, r.marginal_ok = _pd(self.edge_chi2), _pd(self.marginal), _pi(self.marginal_ok)
        return r
