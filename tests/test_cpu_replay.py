"""CPU tests of the host layer (ROS-free Localization / Robot, rows f1-f2 of SURVEY §8): the example
recording replayed through the C++ host library with the CPU oracle plugged in as the solver (the
product backend needs a GPU), window shapes, gates, log format, ATE against Vicon."""
import ctypes as C
import os

import numpy as np
import pytest

from localization_b200.host import Fleet, LocParams
from localization_b200.tools import ate
from localization_b200.tools.replay import load_messages, node_params, replay
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
MSGS = os.path.join(HERE, "golden", "bag_example_msgs.npz")
REF_CFG = "/root/reference/cfg"

# cfg/uwb_only.yaml and cfg/uwb_imu.yaml values (committed here: /root/reference is absent on the GPU box)
UWB_ONLY = dict(trajectory_length=10, maximum_velocity=5.0, distance_outlier=1.0, maximum_iteration=10,
                minimum_optimize_error=2000.0, publish_range=True)
UWB_IMU = dict(trajectory_length=12, maximum_velocity=3.0, distance_outlier=3.0, maximum_iteration=10,
               minimum_optimize_error=1000.0, publish_range=True, publish_imu=False)


def oracle_backend(record=None):
    lib = oracle.load()

    def fn(_user, topo, batch, cfg, out):
        if record is not None:
            t, b = topo.contents, batch.contents
            record.append((t.n_poses, t.n_anchors, t.n_edges, int(b.n_windows), bool(b.pose_R)))
        return lib.uwbgo_oracle_solve_batch(topo, batch, cfg, out, None, 1)
    return fn


def test_cfg_values_match_reference_yaml():
    if not os.path.isdir(REF_CFG):
        pytest.skip("reference tree not present")
    p = LocParams.from_yaml(os.path.join(REF_CFG, "uwb_only.yaml"))
    assert {k: getattr(p, k) for k in UWB_ONLY} == UWB_ONLY
    p = LocParams.from_yaml(os.path.join(REF_CFG, "uwb_imu.yaml"))
    assert {k: getattr(p, k) for k in UWB_IMU} == UWB_IMU


def test_bag_fixture_matches_bag():
    """the committed message fixture is what the rosbag reader extracts from the reference's bag"""
    bag = "/root/reference/bag/data_example.bag"
    if not os.path.exists(bag):
        pytest.skip("reference tree not present")
    from localization_b200.tools.rosbag_reader import load_example_bag
    b, m = load_example_bag(bag), load_messages(MSGS)
    assert len(b.uwb_stamp) == 1444 and len(b.imu_sec) == 4514 and len(b.vicon_stamp) == 1965
    assert np.array_equal(b.uwb_distance.astype(np.float32), m["uwb_distance"])
    assert np.array_equal(b.imu_quat_xyzw, m["imu_quat_xyzw"])
    assert set(b.uwb_frame) == {"uwb"} and set(b.imu_frame) == {"imu_link"}
    # anchors: SURVEY §4.3
    loc = {int(r): tuple(l) for r, l in zip(b.uwb_responder, b.uwb_responder_location)}
    assert loc == {100: (3.0, -3.0, 0.58), 101: (3.0, 3.0, 1.97), 102: (-3.0, 3.0, 0.54), 103: (-3.0, -3.0, 1.76)}


def test_uwb_only_replay(tmp_path):
    """BASELINE configs[0]: data_example.bag through cfg/uwb_only.yaml, single robot"""
    msgs = load_messages(MSGS)
    rec = []
    fleet = Fleet(solve_fn=oracle_backend(rec))
    prm = LocParams(**UWB_ONLY, filename_prefix=str(tmp_path / "run"), filename_suffix="_t.txt")
    replay(msgs, prm, fleet, members=1)
    st = fleet.stats(0)
    # 1444 ranges, solve once the window has filled: SURVEY §4.3 (1,434 solves, 10 poses, 19 edges)
    assert st["solves"] == 1434 and st["errors"] == 0 and st["skipped"] == 0
    assert all(r == (10, 4, 19, 1, False) for r in rec)   # identity rotations -> pose_R NULL -> fast path
    rt, op, err = fleet.published(0)
    assert len(rt) == 1434 - st["rejected"] or len(rt) == 1434
    assert (err < 2000).all() and np.median(err) < 10
    gt_t, gt_p = msgs["vicon_stamp"], msgs["vicon_pos"]
    res = ate.evaluate_ate(gt_t, gt_p, rt[:, 0], rt[:, 1:4])
    assert res["pairs"] > 1000 and res["rmse"] < 0.30, res          # decimetre level (0.18 m raw 3-D)
    assert res["rmse_xy_raw"] < 0.12, res                           # 6.5 cm xy in the survey probe
    res_opt = ate.evaluate_ate(gt_t, gt_p, op[:, 0], op[:, 1:4])
    assert res_opt["rmse"] < 0.30
    # TUM log: 3 '#' header lines, then "stamp x y z qx qy qz qw" per publish
    fleet.close()
    lines = open(str(tmp_path / "run_realtime_t.txt")).read().splitlines()
    assert lines[0] == "# iteration_max:10" and lines[1] == "# trajectory_length:10"
    body = [l for l in lines if not l.startswith("#")]
    assert len(body) == len(rt) and len(body[0].split()) == 8
    s, x = ate.read_trajectory(str(tmp_path / "run_realtime_t.txt"))
    assert np.allclose(s, rt[:, 0]) and np.allclose(x, rt[:, 1:4], atol=1e-5)
    # the destructor appends the newer half of the final window to the optimized log (localization.cpp:705-716)
    s2, _ = ate.read_trajectory(str(tmp_path / "run_optimized_t.txt"))
    assert len(s2) == len(op) + 5


def test_uwb_imu_replay_window_shape():
    """BASELINE configs[1] window shape: 12 poses, 12 + 11 range-type edges, 11 IMU priors"""
    msgs = load_messages(MSGS)
    rec = []
    fleet = Fleet(solve_fn=oracle_backend(rec))
    replay(msgs, LocParams(**UWB_IMU), fleet, members=1, use_imu=True, max_ranges=200)
    st = fleet.stats(0)
    assert st["solves"] == 200 - 12 and st["errors"] == 0
    assert all(r == (12, 4, 34, 1, True) for r in rec[5:])           # 23 range + 11 prior edges, rotations present
    rt, _, err = fleet.published(0)
    assert len(rt) > 150 and (err < 1000).all()
    res = ate.evaluate_ate(msgs["vicon_stamp"], msgs["vicon_pos"], rt[:, 0], rt[:, 1:4])
    assert res["rmse"] < 0.5, res


def test_fleet_batches_by_structure():
    """M replicas in lockstep -> one batch of M windows per flush; replicas with identical inputs
    stay identical, a perturbed replica diverges"""
    msgs = load_messages(MSGS)
    rec = []
    fleet = Fleet(solve_fn=oracle_backend(rec))
    noise = np.zeros((3, len(msgs["uwb_distance"])))
    noise[2] = np.random.default_rng(0).normal(0, 0.05, noise.shape[1])
    replay(msgs, LocParams(**UWB_ONLY), fleet, members=3, range_noise=noise, max_ranges=60)
    assert all(r[3] == 3 for r in rec) and len(rec) == 50
    a, b, c = (fleet.published(i)[0] for i in range(3))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert fleet.stats(0)["fleet_windows"] == 150 and fleet.stats(0)["fleet_batches"] == 50


def test_outlier_gate_and_chi2_gate():
    msgs = load_messages(MSGS)
    fleet = Fleet(solve_fn=oracle_backend())
    p = node_params(msgs, LocParams(**{**UWB_ONLY, "minimum_optimize_error": 1e-9}))
    fleet.add(p)
    for i in range(30):
        d = float(msgs["uwb_distance"][i]) + (5.0 if i == 20 else 0.0)   # a 5 m outlier after warm-up
        fleet.add_range(0, int(msgs["uwb_seq"][i]), int(msgs["uwb_sec"][i]), int(msgs["uwb_nsec"][i]), "uwb",
                        200, int(msgs["uwb_responder"][i]), d, float(msgs["uwb_distance_err"][i]), 1)
        fleet.flush()
    st = fleet.stats(0)
    assert st["rejected"] == 1                       # |estimate - d| > distance_outlier (localization.cpp:308)
    assert st["solves"] == 30 - 10 - 1
    assert st["skipped"] == st["solves"]             # every error >= minimum_optimize_error -> publish skipped
    assert len(fleet.published(0)[0]) == 0


def test_pose_edges_to_a_key_vertex():
    """addPoseEdge ties every new pose of a keyframe to one key vertex (localization.cpp:258-267): a
    star, not a chain; solved by the forest elimination.  cfg/uwb_pose.yaml style: pose + range."""
    msgs = load_messages(MSGS)
    rec = []
    fleet = Fleet(solve_fn=oracle_backend(rec))
    p = node_params(msgs, LocParams(**{**UWB_ONLY, "trajectory_length": 8, "publish_range": False,
                                      "publish_pose": True, "maximum_iteration": 10}))
    fleet.add(p)
    cov = (np.eye(6) * 1e-4).reshape(-1)
    rng = np.random.default_rng(0)
    n = 0
    for k in range(20):
        frame = f"kf{k // 5}"                       # keyframe changes every 5 poses
        rel = np.array([0.05 * (k % 5 + 1), 0.0, 0.0]) + rng.normal(0, 0.003, 3)
        fleet.add_pose(0, k, 100 + k, 0, frame, rel, [0, 0, 0, 1], cov)
        fleet.flush()
        i = k % len(msgs["uwb_distance"])
        fleet.add_range(0, k, 100 + k, 5000, "other", 200, int(msgs["uwb_responder"][i]),
                        float(msgs["uwb_distance"][i]), 0.5, 0)   # merged-covariance branch: no new vertex
        n += 1
    st = fleet.stats(0)
    assert st["errors"] == 0 and st["solves"] == 20, fleet.last_error(0)
    assert rec[-1][0] == 8 and rec[-1][4]           # full window of 8 poses, rotations present
    assert len(fleet.published(0)[0]) > 0


def oracle_backend_recording_offsets(rec):
    lib = oracle.load()

    def fn(_user, topo, batch, cfg, out):
        t = topo.contents
        E = t.n_edges
        rec.append(([t.edge_kind[e] for e in range(E)], [t.edge_ant[e] for e in range(E)],
                    [t.edge_ant_b[e] for e in range(E)] if t.edge_ant_b else None, int(batch.contents.n_windows)))
        return lib.uwbgo_oracle_solve_batch(topo, batch, cfg, out, None, 1)
    return fn


def test_typed_edges_carry_both_offsets_to_the_solver():
    """EdgeSE3Range::setVertexOffset(0/1) and EdgeSE3RangeOffset::setParameterId(0/1) reach the C ABI as
    edge_ant / edge_ant_b; nothing is dropped on the way (types_edge_se3range.cpp:99-114,
    types_edge_se3range_offset.cpp:126-149)"""
    from typed_edge_scenario import drive
    rec = []
    fleet = Fleet(solve_fn=oracle_backend_recording_offsets(rec))
    out = drive(fleet, MSGS, members=1, rounds=2)
    assert fleet.stats(0)["errors"] == 0, fleet.last_error(0)
    kinds, ant, ant_b, _ = rec[-1]
    pairs = [(a, b) for k, a, b in zip(kinds, ant, ant_b) if k in (0, 1)]
    for want in ((2, 1), (0, 3), (1, 3), (0, 2)):
        assert want in pairs, (want, pairs)
    assert any(k == 2 for k in kinds)                # IMU priors: the window has rotations
    chi2, status = out[-1][0][1], out[-1][0][2]
    assert np.isfinite(chi2).all() and status[0] >= 1


def test_typed_edge_with_an_offset_the_solver_cannot_carry_is_refused():
    msgs = load_messages(MSGS)
    fleet = Fleet(solve_fn=oracle_backend())
    fleet.add(node_params(msgs, LocParams(**UWB_ONLY, antenna_offset=[0.1, 0.0, 0.0])))
    with pytest.raises(ValueError, match="antenna"):
        fleet.add_typed_range_edge(0, Fleet.EDGE_RANGE, from_age=0, to_age=1, off_from=2)      # not in the table
    with pytest.raises(ValueError, match="newer pose"):
        fleet.add_typed_range_edge(0, Fleet.EDGE_RANGE_OFFSET, from_age=3, to_age=2, off_from=1)
    with pytest.raises(ValueError, match="anchor"):
        fleet.add_typed_range_edge(0, Fleet.EDGE_RANGE, from_age=0, to_anchor=555)
    assert fleet.stats(0)["errors"] == 3


def test_deferred_mode_settles_a_queued_window_before_the_graph_changes():
    """range -> twist with no flush in between: the twist evicts the oldest vertex of the queued window.
    The queued window is solved first (on its own), so the stream equals the flush-every-message stream."""
    msgs = load_messages(MSGS)
    cov = (np.eye(6) * 1e-2).reshape(-1)

    def run(flush_every):
        fleet = Fleet(solve_fn=oracle_backend())
        fleet.add(node_params(msgs, LocParams(**{**UWB_ONLY, "trajectory_length": 6})))
        for i in range(40):
            fleet.add_range(0, i, int(msgs["uwb_sec"][i]), int(msgs["uwb_nsec"][i]), "uwb", 200,
                            int(msgs["uwb_responder"][i]), float(msgs["uwb_distance"][i]),
                            float(msgs["uwb_distance_err"][i]), 0)
            if flush_every:
                fleet.flush()
            fleet.add_twist(0, i, int(msgs["uwb_sec"][i]), int(msgs["uwb_nsec"][i]) + 2000, "odom",
                            [0.1, 0.0, 0.0], [0.0, 0.0, 0.01], cov)
            if i % 7 == 0:
                fleet.solve(0)      # a second solve while one may be queued
                if flush_every:
                    fleet.flush()
        fleet.flush()
        return fleet

    a, b = run(True), run(False)
    assert a.stats(0)["errors"] == 0 and b.stats(0)["errors"] == 0
    assert a.stats(0)["solves"] == b.stats(0)["solves"] > 30
    assert b.stats(0)["settled_alone"] > 25 and a.stats(0)["settled_alone"] == 0
    for x, y in zip(a.published(0), b.published(0)):
        assert np.array_equal(x, y)
    assert np.array_equal(a.window_poses(0), b.window_poses(0))


def test_fleet_members_with_different_antenna_tables_or_iteration_limits_do_not_share_a_batch():
    msgs = load_messages(MSGS)
    rec = []
    fleet = Fleet(solve_fn=oracle_backend_recording_offsets(rec))
    base = {**UWB_ONLY, "trajectory_length": 5}
    fleet.add(node_params(msgs, LocParams(**base, antenna_offset=[0.1, 0.0, 0.0])))
    fleet.add(node_params(msgs, LocParams(**base, antenna_offset=[0.1, 0.0, 0.0])))
    fleet.add(node_params(msgs, LocParams(**base, antenna_offset=[0.2, 0.0, 0.0])))
    fleet.add(node_params(msgs, LocParams(**{**base, "maximum_iteration": 3}, antenna_offset=[0.1, 0.0, 0.0])))
    for i in range(8):
        for mem in range(4):
            fleet.add_range(mem, i, int(msgs["uwb_sec"][i]), int(msgs["uwb_nsec"][i]), "uwb", 200,
                            int(msgs["uwb_responder"][i]), float(msgs["uwb_distance"][i]),
                            float(msgs["uwb_distance_err"][i]), 1)
        n0 = len(rec)
        fleet.flush()
        if len(rec) > n0:
            assert sorted(r[3] for r in rec[n0:]) == [1, 1, 2]     # {0, 1} together, 2 and 3 alone
    assert len(rec) >= 6
    assert fleet.last_solve(3)[1][0] <= 3 < fleet.last_solve(0)[1][0]
