"""Scenario shared by the CPU and GPU tests of the type API: windows whose extra range edges are built
through EdgeSE3Range::setVertexOffset(0 / 1, ...) (reference src/types/types_edge_se3range.cpp:99-114)
and EdgeSE3RangeOffset::setParameterId(0 / 1, ...) (types_edge_se3range_offset.cpp:61-79,126-149)."""
import numpy as np

from localization_b200.host import Fleet, LocParams
from localization_b200.tools.replay import load_messages, node_params

ANTENNAS = [0.21, -0.04, 0.02, -0.18, 0.07, 0.05, 0.0, 0.3, -0.1]   # three lever arms, body frame
BASE = dict(trajectory_length=8, maximum_velocity=3.0, distance_outlier=3.0, maximum_iteration=10,
            minimum_optimize_error=1e9, publish_range=False)


def drive(fleet: Fleet, msgs_path: str, members: int = 1, rounds: int = 4, noise=None):
    """ranges + IMU orientations fill the window (non-identity rotations, so lever arms matter), then
    every round adds typed edges between vertices of the window and solves"""
    m = load_messages(msgs_path)
    p = node_params(m, LocParams(**BASE, antenna_offset=ANTENNAS))
    for _ in range(members):
        fleet.add(p)
    cov = m["imu_orientation_cov"]
    k_imu = 0
    n_msg = 0
    out = []
    for r in range(rounds):
        for _ in range(3 if r else 10):
            i = n_msg
            n_msg += 1
            for mem in range(members):
                d = float(m["uwb_distance"][i]) + (0.0 if noise is None else float(noise[mem, i]))
                fleet.add_range(mem, int(m["uwb_seq"][i]), int(m["uwb_sec"][i]), int(m["uwb_nsec"][i]), "uwb",
                                int(m["uwb_requester"][i]), int(m["uwb_responder"][i]), d,
                                float(m["uwb_distance_err"][i]), 1 + i % 3)
                q = m["imu_quat_xyzw"][(k_imu * 37) % len(m["imu_quat_xyzw"])]
                fleet.add_imu(mem, i, int(m["uwb_sec"][i]), int(m["uwb_nsec"][i]) + 1000, "imu_link", q, cov)
            k_imu += 1
        for mem in range(members):
            # EDGE_RANGE, both offsets by value: pose (antenna 2) -- anchor 101 (antenna 1 on the anchor side)
            fleet.add_typed_range_edge(mem, Fleet.EDGE_RANGE, from_age=3, to_anchor=101, measurement=4.1 + 0.1 * r,
                                       information=25.0, off_from=2, off_to=1)
            # EDGE_RANGE between two poses, vertex-1 offset only
            fleet.add_typed_range_edge(mem, Fleet.EDGE_RANGE, from_age=5, to_age=6, measurement=0.05,
                                       information=40.0, off_from=0, off_to=3, cauchy=False)
            # EDGE_RANGE_OFFSET, parameter ids on both sides
            fleet.add_typed_range_edge(mem, Fleet.EDGE_RANGE_OFFSET, from_age=6, to_age=7, measurement=0.02,
                                       information=30.0, off_from=1, off_to=3)
            # EDGE_RANGE_OFFSET to an anchor, pidTo only
            fleet.add_typed_range_edge(mem, Fleet.EDGE_RANGE_OFFSET, from_age=7, to_anchor=103,
                                       measurement=3.3 + 0.05 * r, information=16.0, off_from=0, off_to=2)
            fleet.solve(mem)
        fleet.flush()
        out.append([(fleet.window_poses(mem), *fleet.last_solve(mem)) for mem in range(members)])
    return out
