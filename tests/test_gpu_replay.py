"""GPU tests of the caller of the hot path: the example recording replayed through the ROS-free
Localization fleet with the CUDA library as the solver, against the same replay with the CPU
oracle as the solver.  Because both solvers execute the same IEEE operation sequence, even the
free-running warm-started streams must agree bit for bit (SURVEY Appendix B shows that any
rounding difference would grow to centimetres through the warm start)."""
import os

import numpy as np
import pytest

from localization_b200.host import Fleet, LocParams
from localization_b200.tools import ate
from localization_b200.tools.replay import load_messages, replay
from test_cpu_replay import MSGS, UWB_IMU, UWB_ONLY, oracle_backend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,use_imu", [(UWB_ONLY, False), (UWB_IMU, True)])
def test_single_robot_stream_matches_oracle(solver, cfg, use_imu):
    """BASELINE configs[0] and [1]: one robot, one window per range message"""
    msgs = load_messages(MSGS)
    gpu = replay(msgs, LocParams(**cfg), Fleet(solver=solver), members=1, use_imu=use_imu)
    cpu = replay(msgs, LocParams(**cfg), Fleet(solve_fn=oracle_backend()), members=1, use_imu=use_imu)
    assert gpu.stats(0)["solves"] == cpu.stats(0)["solves"] >= 1432 and gpu.stats(0)["errors"] == 0
    for a, b in zip(gpu.published(0), cpu.published(0)):
        assert np.array_equal(a, b)
    rt = gpu.published(0)[0]
    res = ate.evaluate_ate(msgs["vicon_stamp"], msgs["vicon_pos"], rt[:, 0], rt[:, 1:4])
    assert res["rmse"] < 0.08, res        # 4.9 cm (uwb_only) / 5.6 cm (uwb_imu) Horn-aligned ATE vs Vicon


def test_monte_carlo_fleet_matches_oracle(solver):
    """many replicas of the recording with perturbed ranges, advanced in lockstep: one batch of M
    windows per range message through uwbgo_solve_batch"""
    msgs = load_messages(MSGS)
    M, R = 256, 80
    noise = np.random.default_rng(3).normal(0, 0.05, (M, len(msgs["uwb_distance"])))
    gpu = replay(msgs, LocParams(**UWB_ONLY), Fleet(solver=solver), members=M, range_noise=noise, max_ranges=R)
    cpu = replay(msgs, LocParams(**UWB_ONLY), Fleet(solve_fn=oracle_backend()), members=M, range_noise=noise,
                 max_ranges=R)
    st = gpu.stats(0)
    assert st["fleet_batches"] == R - 10 and st["fleet_windows"] == M * (R - 10)
    for i in (0, 1, M // 2, M - 1):
        for a, b in zip(gpu.published(i), cpu.published(i)):
            assert np.array_equal(a, b)
    assert not np.array_equal(gpu.published(0)[0], gpu.published(1)[0])


def test_windows_built_through_the_edge_classes_match_oracle(solver):
    """EdgeSE3Range::setVertexOffset(0 / 1, ...) and EdgeSE3RangeOffset::setParameterId(0 / 1, ...)
    (reference src/types/types_edge_se3range.cpp:99-114, types_edge_se3range_offset.cpp:126-149) on
    windows with rotations: the GPU solve of what the type API builds equals the oracle's, bit for bit"""
    from typed_edge_scenario import drive
    M = 5
    noise = np.random.default_rng(11).normal(0, 0.03, (M, 64))
    gpu_fleet, cpu_fleet = Fleet(solver=solver), Fleet(solve_fn=oracle_backend())
    gpu = drive(gpu_fleet, MSGS, members=M, rounds=4, noise=noise)
    cpu = drive(cpu_fleet, MSGS, members=M, rounds=4, noise=noise)
    assert gpu_fleet.stats(0)["errors"] == 0 == cpu_fleet.stats(0)["errors"], gpu_fleet.last_error(0)
    assert gpu_fleet.stats(0)["fleet_windows"] == 4 * M
    for rg, rc in zip(gpu, cpu):
        for (pg, cg, sg), (pc, cc, sc) in zip(rg, rc):
            assert np.array_equal(pg, pc) and np.array_equal(cg, cc) and np.array_equal(sg, sc)
    assert not np.array_equal(gpu[-1][0][0], gpu[-1][1][0])          # the members differ
    # and the lever arms matter: the same stream without vertex-1 offsets gives other poses
    import typed_edge_scenario as sc
    saved = sc.ANTENNAS
    try:
        sc.ANTENNAS = saved[:6] + [0.0, 0.0, 0.0]
        other = drive(Fleet(solve_fn=oracle_backend()), MSGS, members=1, rounds=4, noise=noise)
    finally:
        sc.ANTENNAS = saved
    assert not np.array_equal(other[-1][0][0], cpu[-1][0][0])
