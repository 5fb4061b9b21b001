"""Replay of the reference's example recording (bag/data_example.bag) through fleets of ROS-free
Localization instances: the caller of the hot path, end to end (BASELINE.json configs[0], [1]).

Anchors come from the messages' responder_location; the self initial position and
antenna offsets are not in the reference tree (uwb_driver/cfg/anchor.yaml is external) and are
chosen here: first Vicon position, one antenna with zero lever arm."""
from __future__ import annotations

import numpy as np

from ..host import Fleet, LocParams


def load_messages(npz_path: str):
    return dict(np.load(npz_path, allow_pickle=False))


def bag_to_npz(bag_path: str, npz_path: str):
    """Extract the fields the callbacks read into a small npz (committed as a test fixture)."""
    from .rosbag_reader import load_example_bag
    b = load_example_bag(bag_path)
    np.savez_compressed(
        npz_path,
        uwb_seq=b.uwb_seq, uwb_sec=b.uwb_sec, uwb_nsec=b.uwb_nsec,
        uwb_requester=b.uwb_requester, uwb_responder=b.uwb_responder,
        uwb_distance=b.uwb_distance.astype(np.float32), uwb_distance_err=b.uwb_distance_err.astype(np.float32),
        uwb_antenna=b.uwb_antenna, uwb_responder_location=b.uwb_responder_location,
        imu_seq=b.imu_seq, imu_sec=b.imu_sec, imu_nsec=b.imu_nsec,
        imu_quat_xyzw=b.imu_quat_xyzw, imu_orientation_cov=b.imu_orientation_cov[0],
        vicon_stamp=b.vicon_stamp, vicon_pos=b.vicon_pos, vicon_quat_xyzw=b.vicon_quat_xyzw)


def node_params(msgs, base: LocParams) -> LocParams:
    """/uwb/nodesId, /uwb/nodesPos from the recording: anchors 100..103 at their
    responder_location, self (200) last at the first Vicon position."""
    ids, pos = [], []
    for r in sorted(set(int(x) for x in msgs["uwb_responder"])):
        k = int(np.argmax(msgs["uwb_responder"] == r))
        ids.append(r)
        pos += list(msgs["uwb_responder_location"][k])
    ids.append(int(msgs["uwb_requester"][0]))
    pos += list(msgs["vicon_pos"][0])
    p = LocParams(**{**base.__dict__})
    p.nodes_id, p.nodes_pos = ids, pos
    if not p.antenna_offset:
        p.antenna_offset = [0.0, 0.0, 0.0]
    return p


def merged_events(msgs, use_imu: bool):
    """(kind, index) in header-stamp order; kind 0 = range, 1 = imu (rosbag play order)"""
    tu = msgs["uwb_sec"].astype(np.float64) + 1e-9 * msgs["uwb_nsec"]
    ev = [(t, 0, i) for i, t in enumerate(tu)]
    if use_imu:
        ti = msgs["imu_sec"].astype(np.float64) + 1e-9 * msgs["imu_nsec"]
        ev += [(t, 1, i) for i, t in enumerate(ti)]
    ev.sort()
    return ev


def replay(msgs, params: LocParams, fleet: Fleet, members: int = 1, use_imu: bool = False,
           range_noise: np.ndarray | None = None, max_ranges: int | None = None):
    """Feed the recording to `members` Localization instances in lockstep.  range_noise, if given,
    is [members][n_ranges] metres added to the recorded distances (Monte-Carlo replay).
    Returns the fleet (results via fleet.published(i))."""
    p = node_params(msgs, params)
    for _ in range(members):
        fleet.add(p)
    n_r = 0
    for _t, kind, i in merged_events(msgs, use_imu):
        if kind == 0:
            if max_ranges is not None and n_r >= max_ranges:
                break
            d = np.full(members, msgs["uwb_distance"][i], np.float32)
            if range_noise is not None:
                d = (d.astype(np.float64) + range_noise[:, i]).astype(np.float32)
            e = np.full(members, msgs["uwb_distance_err"][i], np.float32)
            fleet.add_range_each(int(msgs["uwb_seq"][i]), int(msgs["uwb_sec"][i]), int(msgs["uwb_nsec"][i]), "uwb",
                                 int(msgs["uwb_requester"][i]), int(msgs["uwb_responder"][i]), d, e,
                                 int(msgs["uwb_antenna"][i]))
            n_r += 1
            fleet.flush()
        else:
            q = np.tile(msgs["imu_quat_xyzw"][i], (members, 1))
            fleet.add_imu_each(int(msgs["imu_seq"][i]), int(msgs["imu_sec"][i]), int(msgs["imu_nsec"][i]), "imu_link",
                               q, msgs["imu_orientation_cov"])
            if params.publish_imu:
                fleet.flush()
    return fleet
