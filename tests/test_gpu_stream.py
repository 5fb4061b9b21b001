"""uwbgo_stream (resident fleet): the windows stay on the device, a step sends one range message per robot.  Every
step must give the bits of the oracle on the window the reference would hold -- Localization::addRangeEdge
(localization.cpp:297-376): new vertex = copy of the newest estimate, range edge to the anchor, zero-length
trajectory edge to the previous vertex, oldest vertex dropped (robot.cpp:75-110) -- built here by shifting the
arrays on the host."""
import numpy as np
import pytest

from localization_b200 import Batch, Config
from localization_b200.graph import RangeMsgs
from localization_b200.stream import ResidentFleet, chain_topology
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,N,A,steps", [(200, 12, 4, 9), (33, 5, 3, 7), (1000, 50, 8, 3)])
def test_resident_fleet_matches_shifted_windows(_gpu_solver, W, N, A, steps):
    rng = np.random.default_rng(100 + N)
    v_max = 5.0
    cfg = Config(max_iterations=10)
    anchors = rng.uniform(-6.0, 6.0, (A, 3)) + np.array([0.0, 0.0, 2.0])
    T = N + steps
    vel = rng.normal(0.0, 0.6, (W, 1, 3))
    truth = rng.uniform(-3.0, 3.0, (W, 1, 3)) + np.cumsum(np.broadcast_to(vel, (W, T, 3)) * 0.04 + rng.normal(0, 0.01, (W, T, 3)), axis=1)
    aop_all = np.arange(T) % A
    d_all = (np.linalg.norm(truth - anchors[aop_all][None], axis=2) + rng.normal(0, 0.05, (W, T))).astype(np.float32)
    e_all = np.where(rng.uniform(size=(W, T)) < 0.7, np.float32(0.055), np.float32(0.024)).astype(np.float32)
    dt_all = rng.uniform(0.02, 0.06, (W, T))                     # dt_all[:, k]: stamp of pose k minus stamp of pose k - 1
    pose = truth[:, :N] + rng.normal(0, 0.1, (W, N, 3))
    d, e, dt, aop = d_all[:, :N].copy(), e_all[:, :N].copy(), dt_all[:, 1:N].copy(), list(aop_all[:N])
    fleet = ResidentFleet(_gpu_solver, N, anchors, W, v_max, cfg)
    try:
        fleet.load(pose, aop, d, e, dt)
        assert np.array_equal(fleet.read(), pose)
        for s in range(steps):
            k = N + s
            a = int(aop_all[k])
            pose_in = np.concatenate([pose[:, 1:], pose[:, -1:]], axis=1)
            d = np.concatenate([d[:, 1:], d_all[:, k:k + 1]], axis=1)
            e = np.concatenate([e[:, 1:], e_all[:, k:k + 1]], axis=1)
            dt = np.concatenate([dt[:, 1:], dt_all[:, k:k + 1]], axis=1)
            aop = aop[1:] + [a]
            topo = chain_topology(N, A, aop)
            batch = Batch(pose_t=pose_in, anchors=anchors, shared_anchors=True,
                          range_msgs=RangeMsgs(distance=d, distance_err=e, dt_pose=np.ascontiguousarray(dt), v_max=v_max))
            ref = oracle.solve(topo, batch, cfg)
            newest, chi2, status = fleet.step(a, d_all[:, k], e_all[:, k], dt_all[:, k])
            assert _gpu_solver.last_path == (3 if W <= 592 else 2)   # small fleets: WINDOW kernels; else the CHAIN kernel
            assert np.array_equal(newest, ref.pose_t[:, -1]), (s, np.abs(newest - ref.pose_t[:, -1]).max())
            assert np.array_equal(chi2, ref.chi2) and np.array_equal(status, ref.status)
            pose = ref.pose_t
        assert np.array_equal(fleet.read(), pose)
    finally:
        fleet.close()


@pytest.mark.parametrize("W,N,A,steps,window", [(200, 12, 4, 6, True), (200, 12, 4, 3, False), (33, 5, 3, 5, True),
                                                (1000, 50, 8, 2, True)])
def test_resident_fleet_with_one_anchor_sequence_per_robot(_gpu_solver, W, N, A, steps, window):
    """_load_robots / _step_robots: every robot ranges its own anchor in a step (UwbRange::responder_id,
    localization.cpp:305-306,331).  The oracle solves each robot's window with the anchor ids of ITS edges: all
    windows at once through window-private anchor rows, and a sample of them one by one on the topology the
    reference would hold (anchor ids in the edge list, one shared constellation) -- the same bits either way.
    Fleets of up to 592 robots hand the solve the expanded window-major form and take the WINDOW kernels (one CTA
    per robot) unless that path is switched off; larger ones the tile kernels on the message fields."""
    _gpu_solver.set_window_path(-1 if window else 0)
    rng = np.random.default_rng(300 + N)
    v_max = 5.0
    cfg = Config(max_iterations=10)
    anchors = rng.uniform(-6.0, 6.0, (A, 3)) + np.array([0.0, 0.0, 2.0])
    T = N + steps
    vel = rng.normal(0.0, 0.6, (W, 1, 3))
    truth = rng.uniform(-3.0, 3.0, (W, 1, 3)) + np.cumsum(np.broadcast_to(vel, (W, T, 3)) * 0.04 + rng.normal(0, 0.01, (W, T, 3)), axis=1)
    aop_all = rng.integers(0, A, (W, T)).astype(np.int32)
    d_all = (np.linalg.norm(truth - anchors[aop_all], axis=2) + rng.normal(0, 0.05, (W, T))).astype(np.float32)
    e_all = np.where(rng.uniform(size=(W, T)) < 0.7, np.float32(0.055), np.float32(0.024)).astype(np.float32)
    dt_all = rng.uniform(0.02, 0.06, (W, T))
    pose = truth[:, :N] + rng.normal(0, 0.1, (W, N, 3))
    d, e, dt, aop = d_all[:, :N].copy(), e_all[:, :N].copy(), dt_all[:, 1:N].copy(), aop_all[:, :N].copy()
    topo_rows = chain_topology(N, N, np.arange(N))               # pose k reads row k of its window's anchors
    fleet = ResidentFleet(_gpu_solver, N, anchors, W, v_max, cfg)
    try:
        fleet.load(pose, aop, d, e, dt)
        assert np.array_equal(fleet.read(), pose)
        with pytest.raises(Exception):
            fleet.step(0, d_all[:, N], e_all[:, N], dt_all[:, N])   # loaded per robot: a fleet-wide anchor is refused
        for s in range(steps):
            k = N + s
            pose_in = np.concatenate([pose[:, 1:], pose[:, -1:]], axis=1)
            d = np.concatenate([d[:, 1:], d_all[:, k:k + 1]], axis=1)
            e = np.concatenate([e[:, 1:], e_all[:, k:k + 1]], axis=1)
            dt = np.concatenate([dt[:, 1:], dt_all[:, k:k + 1]], axis=1)
            aop = np.concatenate([aop[:, 1:], aop_all[:, k:k + 1]], axis=1)
            msgs = RangeMsgs(distance=d, distance_err=e, dt_pose=np.ascontiguousarray(dt), v_max=v_max)
            ref = oracle.solve(topo_rows, Batch(pose_t=pose_in, anchors=np.ascontiguousarray(anchors[aop]), range_msgs=msgs), cfg)
            for w in range(0, W, max(1, W // 5)):
                one = oracle.solve(chain_topology(N, A, aop[w]),
                                   Batch(pose_t=pose_in[w:w + 1], anchors=anchors, shared_anchors=True,
                                         range_msgs=RangeMsgs(distance=d[w:w + 1], distance_err=e[w:w + 1],
                                                              dt_pose=np.ascontiguousarray(dt[w:w + 1]), v_max=v_max)), cfg)
                assert np.array_equal(one.pose_t[0], ref.pose_t[w]) and np.array_equal(one.chi2[0], ref.chi2[w])
            newest, chi2, status = fleet.step(aop_all[:, k], d_all[:, k], e_all[:, k], dt_all[:, k])
            assert _gpu_solver.last_path == (3 if window and W <= 592 else 2)   # WINDOW kernels / straight-line CHAIN kernel
            assert np.array_equal(newest, ref.pose_t[:, -1]), (s, np.abs(newest - ref.pose_t[:, -1]).max())
            assert np.array_equal(chi2, ref.chi2) and np.array_equal(status, ref.status)
            pose = ref.pose_t
        assert np.array_equal(fleet.read(), pose)
        bad = aop_all[:, -1].copy()
        bad[W // 2] = A
        with pytest.raises(Exception):
            fleet.step(bad, d_all[:, -1], e_all[:, -1], dt_all[:, -1])   # one id out of range: nothing is queued
        assert np.array_equal(fleet.read(), pose)
        # a stream can be reloaded the other way
        fleet.load(pose, list(np.arange(N) % A), d, e, dt)
        with pytest.raises(Exception):
            fleet.step(aop_all[:, -1], d_all[:, -1], e_all[:, -1], dt_all[:, -1])
        fleet.step(0, d_all[:, -1], e_all[:, -1], dt_all[:, -1])
    finally:
        fleet.close()
        _gpu_solver.set_window_path(-1)


@pytest.mark.parametrize("W,N,A,steps", [(300, 12, 4, 8), (64, 50, 8, 3)])
def test_resident_fleet_outlier_gate(_gpu_solver, W, N, A, steps):
    """the outlier gate of addRangeEdge (localization.cpp:305-313): a robot whose message is refused keeps its window
    (estimates, message fields, anchor ids), nothing is solved for it; the others step as usual.  Host restatement
    of the gate: Eigen's norm of the 3-vector, |estimate - float32 range| > distance_outlier."""
    from localization_b200.stream import FLAG_REJECTED
    rng = np.random.default_rng(500 + N)
    v_max, outlier = 5.0, 1.0
    cfg = Config(max_iterations=10)
    anchors = rng.uniform(-6.0, 6.0, (A, 3)) + np.array([0.0, 0.0, 2.0])
    T = N + steps
    vel = rng.normal(0.0, 0.6, (W, 1, 3))
    truth = rng.uniform(-3.0, 3.0, (W, 1, 3)) + np.cumsum(np.broadcast_to(vel, (W, T, 3)) * 0.04 + rng.normal(0, 0.01, (W, T, 3)), axis=1)
    aop_all = rng.integers(0, A, (W, T)).astype(np.int32)
    d_all = np.linalg.norm(truth - anchors[aop_all], axis=2) + rng.normal(0, 0.05, (W, T))
    d_all[:, N:] += np.where(rng.uniform(size=(W, steps)) < 0.2, rng.choice([-2.5, 2.5, 0.9, 1.1], (W, steps)), 0.0)  # multipath spikes
    d_all = d_all.astype(np.float32)
    e_all = np.full((W, T), 0.055, np.float32)
    dt_all = rng.uniform(0.02, 0.06, (W, T))
    pose = truth[:, :N] + rng.normal(0, 0.1, (W, N, 3))
    d, e, dt, aop = d_all[:, :N].copy(), e_all[:, :N].copy(), dt_all[:, 1:N].copy(), aop_all[:, :N].copy()
    chi2_prev = np.zeros((W, 4))
    topo_rows = chain_topology(N, N, np.arange(N))
    fleet = ResidentFleet(_gpu_solver, N, anchors, W, v_max, cfg)
    n_rejected = 0
    try:
        fleet.load(pose, aop, d, e, dt)
        fleet.set_outlier_gate(outlier)
        for s in range(steps):
            k = N + s
            diff = pose[:, -1] - anchors[aop_all[:, k]]
            est = np.sqrt(diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1] + diff[:, 2] * diff[:, 2])
            rej = np.abs(est - d_all[:, k].astype(np.float64)) > outlier
            n_rejected += int(rej.sum())
            acc = ~rej
            pose_in = np.concatenate([pose[:, 1:], pose[:, -1:]], axis=1)
            d_s = np.concatenate([d[:, 1:], d_all[:, k:k + 1]], axis=1)
            e_s = np.concatenate([e[:, 1:], e_all[:, k:k + 1]], axis=1)
            dt_s = np.concatenate([dt[:, 1:], dt_all[:, k:k + 1]], axis=1)
            aop_s = np.concatenate([aop[:, 1:], aop_all[:, k:k + 1]], axis=1)
            ref = oracle.solve(topo_rows, Batch(pose_t=pose_in, anchors=np.ascontiguousarray(anchors[aop_s]),
                                                range_msgs=RangeMsgs(distance=d_s, distance_err=e_s,
                                                                     dt_pose=np.ascontiguousarray(dt_s), v_max=v_max)), cfg)
            m = acc[:, None]
            pose = np.where(acc[:, None, None], ref.pose_t, pose)
            d, e, dt, aop = np.where(m, d_s, d), np.where(m, e_s, e), np.where(m, dt_s, dt), np.where(m, aop_s, aop)
            chi2_exp = np.where(m, ref.chi2, chi2_prev)
            status_exp = np.where(m, ref.status, np.array([0, 0, FLAG_REJECTED, 0], np.int32)[None])
            newest, chi2, status = fleet.step(aop_all[:, k], d_all[:, k], e_all[:, k], dt_all[:, k])
            assert np.array_equal(status[:, 2] == FLAG_REJECTED, rej)
            assert np.array_equal(newest, pose[:, -1]), (s, np.abs(newest - pose[:, -1]).max())
            assert np.array_equal(chi2, chi2_exp) and np.array_equal(status, status_exp)
            assert np.array_equal(fleet.read(), pose)
            chi2_prev = chi2_exp
        assert 0 < n_rejected < W * steps
        # the gate off again: every message is taken
        fleet.set_outlier_gate(-1.0)
        _, _, status = fleet.step(aop_all[:, -1], d_all[:, -1] + np.float32(5.0), e_all[:, -1], dt_all[:, -1])
        assert not (status[:, 2] & FLAG_REJECTED).any()
        # fleet-wide anchor sequences cannot take the gate
        fleet.load(pose, list(np.arange(N) % A), d, e, dt)
        fleet.set_outlier_gate(outlier)
        with pytest.raises(Exception):
            fleet.step(0, d_all[:, -1], e_all[:, -1], dt_all[:, -1])
    finally:
        fleet.close()


def test_resident_fleet_rejects_bad_arguments(_gpu_solver):
    cfg = Config(max_iterations=3)
    anchors = np.zeros((4, 3))
    fleet = ResidentFleet(_gpu_solver, 6, anchors, 8, 5.0, cfg)
    try:
        z = np.zeros(8, np.float32)
        with pytest.raises(Exception):
            fleet.step(0, z, z, np.zeros(8))                     # nothing loaded yet
        fleet.load(np.zeros((8, 6, 3)), [0, 1, 2, 3, 0, 1], np.ones((8, 6), np.float32), np.ones((8, 6), np.float32), np.ones((8, 5)))
        with pytest.raises(Exception):
            fleet.step(4, z, z, np.zeros(8))                     # anchor out of range
    finally:
        fleet.close()
