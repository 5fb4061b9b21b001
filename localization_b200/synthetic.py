"""Synthetic window generators for the BASELINE.json configurations.

Data only: numpy draws in a fixed order from default_rng(seed), FP64, so the CPU checker and the GPU are
fed identical buffers.  Edge parameters follow the reference's factories:
  range information = 1 / distance_err^2            (localization.cpp:318,331,608-627)
  trajectory edge   = EdgeSE3Range(prev, new), measurement 0, information
                      1 / (v_max * dt / 3)^2        (localization.cpp:319,338)
  newest pose       = copy of its predecessor's estimate (robot.cpp:90)
  IMU prior         = EdgeSE3Prior, information diag(0,0,0,1/c,1/c,1/c), pose rotation
                      overwritten by the IMU rotation (localization.cpp:507-525)
  lidar prior       = EdgeSE3Prior, information (2,2) = 1/0.05, pose z overwritten
                      (localization.cpp:470-486)
  twist edge        = EdgeSE3(prev, new), Z = (RPY(w dt), v dt), information
                      (Sigma dt^2)^-1               (localization.cpp:560-605)
"""
from __future__ import annotations

import numpy as np

from .graph import Batch, Topology

SEED_C3 = 20260101
IMU_ORIENTATION_COV = 4.592449e-06  # bag/data_example.bag /imu/data orientation_covariance[0]
ANTENNA_OFFSETS = np.array([[0.20, 0.0, 0.0], [-0.10, 0.17, 0.0], [-0.10, -0.17, 0.0]])


def _anchors(rng, W, A):
    if A == 8:
        base = np.array([[sx * 3.0, sy * 3.0, z] for z in (0.5, 2.0) for sx in (-1, 1) for sy in (-1, 1)])
    elif A == 16:
        xs = np.linspace(-3.0, 3.0, 4)
        base = np.array([[x, y, z] for z in (0.5, 2.0) for y in (-3.0, 3.0) for x in xs])
    else:
        ang = 2 * np.pi * np.arange(A) / max(A, 1)
        base = np.stack([3.0 * np.cos(ang), 3.0 * np.sin(ang), 0.5 + 1.5 * (np.arange(A) % 2)], 1)
    return base[None] + rng.uniform(-0.25, 0.25, size=(W, A, 3))


def _trajectory(rng, W, N):
    p0 = np.stack([rng.uniform(-1.5, 1.5, W), rng.uniform(-1.5, 1.5, W), rng.uniform(1.0, 1.4, W)], 1)
    v = rng.normal(0.0, 0.3, size=(W, 3))
    dv = rng.normal(0.0, 0.05, size=(W, N, 3))
    dt = rng.uniform(0.025, 0.040, size=(W, N))
    p = np.empty((W, N, 3))
    vel = np.empty((W, N, 3))
    cur = p0
    for i in range(N):
        sp = np.linalg.norm(v, axis=1, keepdims=True)
        v = np.where(sp > 0.7, v * (0.7 / np.maximum(sp, 1e-300)), v)
        p[:, i] = cur
        vel[:, i] = v
        cur = cur + v * dt[:, i:i + 1]
        v = v + dv[:, i]
    return p, vel, dt


def _ranges(rng, truth_points, anchors, A, with_err=False):
    """d = float32(|p_i - a_(i mod A)| + N(0, 0.05^2)); distance_err in {0.055 (75 %), 0.024}."""
    W, N, _ = truth_points.shape
    idx = np.arange(N) % A
    a = anchors[:, idx, :]
    d = np.linalg.norm(truth_points - a, axis=2) + rng.normal(0.0, 0.05, size=(W, N))
    d = d.astype(np.float32).astype(np.float64)
    err = np.where(rng.uniform(size=(W, N)) < 0.75, np.float32(0.055), np.float32(0.024)).astype(np.float64)
    if with_err:
        return d, 1.0 / (err * err), err
    return d, 1.0 / (err * err)


def uwb_only(W: int, N: int = 50, A: int = 8, v_max: float = 5.0, seed: int = SEED_C3, compact: bool = False,
             shared_anchors: bool = False):
    """C3 / C5: UWB-only chain windows.  Returns (Topology, Batch, truth [W][N][3]).
    compact: the range data as message fields (RangeMsgs: float32 distance / distance_err, the stamp
    differences dt) instead of the edge parameters -- the same numbers, the expansion is exact.
    shared_anchors: one anchor constellation for the whole batch (a fleet in one anchor field) instead of a
    jittered one per window."""
    rng = np.random.default_rng(seed)
    anchors = _anchors(rng, W, A)
    if shared_anchors:
        anchors = np.ascontiguousarray(np.broadcast_to(anchors[:1], anchors.shape))
    p, _, dt = _trajectory(rng, W, N)
    d, info, err = _ranges(rng, p, anchors, A, with_err=True)
    init = p + rng.normal(0.0, 0.1, size=(W, N, 3))
    if N > 1:
        init[:, N - 1] = init[:, N - 2]
    topo = Topology.uwb_chain(N, A)
    if compact:
        from .graph import RangeMsgs
        msgs = RangeMsgs(distance=d.astype(np.float32), distance_err=err.astype(np.float32),
                         dt_pose=np.ascontiguousarray(dt[:, :N - 1]), v_max=v_max)
        return topo, Batch(pose_t=init, anchors=anchors[0] if shared_anchors else anchors, range_msgs=msgs,
                           shared_anchors=shared_anchors), p
    er = topo.counts()[0]
    rd = np.zeros((W, er))
    ri = np.zeros((W, er))
    slot = 0
    for k in range(N):  # same order as Topology.uwb_chain
        rd[:, slot] = d[:, k]
        ri[:, slot] = info[:, k]
        slot += 1
        if k > 0:
            rd[:, slot] = 0.0
            s = v_max * dt[:, k - 1] / 3.0
            ri[:, slot] = 1.0 / (s * s)
            slot += 1
    return topo, Batch(pose_t=init, anchors=anchors, range_d=rd, range_info=ri), p


def _yaw_R(yaw):
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.zeros(yaw.shape + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 1, 0], R[..., 1, 1], R[..., 2, 2] = c, -s, s, c, 1.0
    return R


def _rpy_R(r, p, y):
    """tf::Matrix3x3::setRPY(roll, pitch, yaw) = Rz(yaw) Ry(pitch) Rx(roll)"""
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    R = np.empty(r.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def uwb_imu_lidar(W: int, N: int = 20, A: int = 8, v_max: float = 5.0, seed: int = SEED_C3 + 4,
                  antennas: int = 3, lidar: bool = True):
    """C4a (and the C2 window shape with lidar=False, antennas=0): range chain with antenna
    offsets plus lidar-z and IMU-rotation EdgeSE3Prior edges on poses 0..N-2."""
    rng = np.random.default_rng(seed)
    anchors = _anchors(rng, W, A)
    p, _, dt = _trajectory(rng, W, N)
    yaw = rng.uniform(-np.pi, np.pi, size=(W, 1)) + np.cumsum(rng.normal(0.0, 0.02, size=(W, N)), axis=1)
    R_true = _yaw_R(yaw)
    K = antennas
    off = ANTENNA_OFFSETS[:K] if K > 0 else None
    if K > 0:
        ant = np.arange(N) % K
        pts = p + np.einsum("wnij,nj->wni", R_true, off[ant])
    else:
        pts = p
    d, info = _ranges(rng, pts, anchors, A)
    init_t = p + rng.normal(0.0, 0.1, size=(W, N, 3))
    # IMU: measured rotation = truth * small rotation noise; written into the pose estimate
    noise = rng.normal(0.0, 2e-3, size=(W, N, 3))
    R_imu = R_true @ _rpy_R(noise[..., 0], noise[..., 1], noise[..., 2])
    z_lidar = p[..., 2] + rng.normal(0.0, 0.02, size=(W, N))
    init_R = R_imu.copy()
    if lidar:
        init_t[:, :N - 1, 2] = z_lidar[:, :N - 1]
    if N > 1:  # newest vertex copies its predecessor (robot.cpp:90)
        init_t[:, N - 1] = init_t[:, N - 2]
        init_R[:, N - 1] = init_R[:, N - 2]
    topo = Topology.uwb_chain(N, A, antennas=K, imu=True, lidar=lidar)
    er, ep, _ = topo.counts()
    rd, ri = np.zeros((W, er)), np.zeros((W, er))
    pZ, pI = np.zeros((W, ep, 12)), np.zeros((W, ep, 6, 6))
    rs = ps = 0
    for k in range(N):
        rd[:, rs], ri[:, rs] = d[:, k], info[:, k]
        rs += 1
        if k > 0:
            s = v_max * dt[:, k - 1] / 3.0
            ri[:, rs] = 1.0 / (s * s)
            rs += 1
        if k < N - 1:
            if lidar:  # measurement = the pose at insertion time with z written in
                pZ[:, ps, :9] = init_R[:, k].reshape(W, 9)
                pZ[:, ps, 9:] = init_t[:, k]
                pI[:, ps, 2, 2] = 1.0 / 0.05
                ps += 1
            pZ[:, ps, :9] = init_R[:, k].reshape(W, 9)
            pZ[:, ps, 9:] = init_t[:, k]
            for j in range(3):
                pI[:, ps, 3 + j, 3 + j] = 1.0 / IMU_ORIENTATION_COV
            ps += 1
    # estimates drift away from the priors' measurement before the solve (later range updates)
    init_t = init_t + rng.normal(0.0, 0.02, size=(W, N, 3))
    if N > 1:
        init_t[:, N - 1] = init_t[:, N - 2]
    batch = Batch(pose_t=init_t, pose_R=init_R, anchors=anchors, range_d=rd, range_info=ri,
                  ant_offsets=off, prior_Z=pZ, prior_info=pI)
    return topo, batch, p


def uwb_twist(W: int, N: int = 15, A: int = 8, v_max: float = 1.0, seed: int = SEED_C3 + 5,
              antennas: int = 3):
    """C4b: vertices chained by twist EdgeSE3 edges, anchor ranges through the merged-covariance
    branch (localization.cpp:348-357)."""
    rng = np.random.default_rng(seed)
    anchors = _anchors(rng, W, A)
    p, vel, dt = _trajectory(rng, W, N)
    yaw0 = rng.uniform(-np.pi, np.pi, size=(W, 1))
    wz = rng.normal(0.0, 0.3, size=(W, N))
    yaw = yaw0 + np.concatenate([np.zeros((W, 1)), np.cumsum(wz[:, :-1] * dt[:, :-1], axis=1)], axis=1)
    R_true = _yaw_R(yaw)
    K = antennas
    off = ANTENNA_OFFSETS[:K] if K > 0 else None
    if K > 0:
        ant = np.arange(N) % K
        pts = p + np.einsum("wnij,nj->wni", R_true, off[ant])
    else:
        pts = p
    d, info = _ranges(rng, pts, anchors, A)
    sig = np.array([0.05, 0.05, 0.05, 0.02, 0.02, 0.02])
    topo = Topology.uwb_twist(N, A, antennas=K)
    er, _, es = topo.counts()
    rd, ri = np.zeros((W, er)), np.zeros((W, er))
    sZ, sI = np.zeros((W, es, 12)), np.zeros((W, es, 6, 6))
    for k in range(N):
        if k > 0:
            dtk = dt[:, k - 1]
            v_body = np.einsum("wji,wj->wi", R_true[:, k - 1], vel[:, k - 1]) + rng.normal(0, sig[0], (W, 3))
            w_body = np.stack([np.zeros(W), np.zeros(W), wz[:, k - 1]], 1) + rng.normal(0, sig[3], (W, 3))
            Rz = _rpy_R(w_body[:, 0] * dtk, w_body[:, 1] * dtk, w_body[:, 2] * dtk)
            sZ[:, k - 1, :9] = Rz.reshape(W, 9)
            sZ[:, k - 1, 9:] = v_body * dtk[:, None]
            cov = (sig * sig)[None, :] * (dtk * dtk)[:, None]
            for j in range(6):
                sI[:, k - 1, j, j] = 1.0 / cov[:, j]
        rd[:, k] = d[:, k]
        cov_motion = (v_max * dt[:, max(k - 1, 0)] / 3.0) ** 2
        ri[:, k] = 1.0 / (1.0 / info[:, k] + cov_motion)
    init_t = p + rng.normal(0.0, 0.1, size=(W, N, 3))
    nz = rng.normal(0.0, 0.02, size=(W, N, 3))
    init_R = R_true @ _rpy_R(nz[..., 0], nz[..., 1], nz[..., 2])
    if N > 1:
        init_t[:, N - 1] = init_t[:, N - 2]
        init_R[:, N - 1] = init_R[:, N - 2]
    batch = Batch(pose_t=init_t, pose_R=init_R, anchors=anchors, range_d=rd, range_info=ri,
                  ant_offsets=off, se3_Z=sZ, se3_info=sI)
    return topo, batch, p


def uwb_pose(W: int, N: int = 24, A: int = 8, keyframe_len: int = 4, seed: int = SEED_C3 + 6,
             antennas: int = 3, v_max: float = 0.5):
    """Keyframe-relative pose edges (cfg/uwb_pose.yaml): EdgeSE3(key vertex, new vertex) stars plus
    merged-covariance anchor ranges.  Z = key^-1 * pose + noise, information = Sigma^-1
    (localization.cpp:271-277)."""
    rng = np.random.default_rng(seed)
    anchors = _anchors(rng, W, A)
    p, vel, dt = _trajectory(rng, W, N)
    yaw = rng.uniform(-np.pi, np.pi, size=(W, 1)) + np.cumsum(rng.normal(0.0, 0.03, size=(W, N)), axis=1)
    R_true = _yaw_R(yaw)
    K = antennas
    off = ANTENNA_OFFSETS[:K] if K > 0 else None
    pts = p + np.einsum("wnij,nj->wni", R_true, off[np.arange(N) % K]) if K > 0 else p
    d, info = _ranges(rng, pts, anchors, A)
    topo = Topology.uwb_pose(N, A, keyframe_len, antennas=K)
    par = topo.parents()
    er, _, es = topo.counts()
    rd, ri = np.zeros((W, er)), np.zeros((W, er))
    sZ, sI = np.zeros((W, es, 12)), np.zeros((W, es, 6, 6))
    sig = np.array([0.02, 0.02, 0.02, 0.01, 0.01, 0.01])
    for k in range(N):
        if k > 0:
            a = par[k]
            Rrel = np.swapaxes(R_true[:, a], 1, 2) @ R_true[:, k]
            trel = np.einsum("wji,wj->wi", R_true[:, a], p[:, k] - p[:, a])
            nz = rng.normal(0, sig[3], (W, 3))
            sZ[:, k - 1, :9] = (Rrel @ _rpy_R(nz[:, 0], nz[:, 1], nz[:, 2])).reshape(W, 9)
            sZ[:, k - 1, 9:] = trel + rng.normal(0, sig[0], (W, 3))
            M = rng.normal(0, 0.1, (W, 6, 6))
            cov = np.einsum("i,wij,j->wij", sig, np.eye(6)[None] + 0.1 * (M + np.swapaxes(M, 1, 2)), sig)
            sI[:, k - 1] = np.linalg.inv(cov)
        rd[:, k] = d[:, k]
        ri[:, k] = 1.0 / (1.0 / info[:, k] + (v_max * dt[:, max(k - 1, 0)] / 3.0) ** 2)
    init_t = p + rng.normal(0.0, 0.05, size=(W, N, 3))
    nz = rng.normal(0.0, 0.02, size=(W, N, 3))
    init_R = R_true @ _rpy_R(nz[..., 0], nz[..., 1], nz[..., 2])
    if N > 1:
        init_t[:, N - 1] = init_t[:, N - 2]
        init_R[:, N - 1] = init_R[:, N - 2]
    batch = Batch(pose_t=init_t, pose_R=init_R, anchors=anchors, range_d=rd, range_info=ri,
                  ant_offsets=off, se3_Z=sZ, se3_info=sI)
    return topo, batch, p


def with_vertex1_offsets(topo: Topology, seed: int = 0) -> Topology:
    """The same window with antenna offsets on vertex 1 of its range edges as well
    (EdgeSE3Range::setVertexOffset(1, .) / EdgeSE3RangeOffset pidTo; the reference's own factories
    never set them, localization.cpp:333): random antenna numbers 0..K on every range edge."""
    rng = np.random.default_rng(seed)
    is_range = topo.edge_kind <= 1
    ant_b = np.where(is_range, rng.integers(0, topo.n_antennas + 1, size=topo.n_edges), 0)
    return Topology(topo.n_poses, topo.n_anchors, topo.n_antennas, topo.edge_kind, topo.edge_a, topo.edge_b,
                    topo.edge_ant, topo.edge_robust, ant_b)
