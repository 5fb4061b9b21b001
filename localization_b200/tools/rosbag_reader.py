"""Minimal ROS-free reader for rosbag v2.0 files with uncompressed chunks — enough for the
reference's fixture bag/data_example.bag (replaces the rosbag dependency of the reference's
script/bag_to_txt.py:52-65 and script/bag_to_csv.py).

Record = <u32 header_len><header fields: <u32 len>name=value ...><u32 data_len><data>.
Ops: 0x03 bag header, 0x05 chunk, 0x07 connection, 0x02 message, 0x04/0x06 index (skipped).
Message bodies are little-endian packed ROS serialisation."""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np


def _fields(buf: bytes) -> dict:
    out, i = {}, 0
    while i < len(buf):
        (n,) = struct.unpack_from("<I", buf, i)
        i += 4
        name, _, val = buf[i:i + n].partition(b"=")
        out[name.decode()] = val
        i += n
    return out


def _records(buf: bytes, pos: int = 0):
    while pos + 4 <= len(buf):
        (hl,) = struct.unpack_from("<I", buf, pos)
        hdr = _fields(buf[pos + 4:pos + 4 + hl])
        pos += 4 + hl
        (dl,) = struct.unpack_from("<I", buf, pos)
        data = buf[pos + 4:pos + 4 + dl]
        pos += 4 + dl
        yield hdr, data


def read_bag(path: str):
    """Yield (topic, msg_type, receive_time_sec, body_bytes) in file order."""
    raw = open(path, "rb").read()
    magic = b"#ROSBAG V2.0\n"
    if not raw.startswith(magic):
        raise ValueError("not a rosbag v2.0 file")
    conns = {}
    out = []

    def handle(hdr, data):
        op = hdr["op"][0]
        if op == 0x07:
            cid = struct.unpack("<I", hdr["conn"])[0]
            info = _fields(data)
            conns[cid] = (hdr["topic"].decode(), info["type"].decode())
        elif op == 0x02:
            cid = struct.unpack("<I", hdr["conn"])[0]
            sec, nsec = struct.unpack("<II", hdr["time"])
            topic, typ = conns[cid]
            out.append((topic, typ, sec + nsec * 1e-9, data))

    for hdr, data in _records(raw, len(magic)):
        op = hdr["op"][0]
        if op == 0x05:
            if hdr.get("compression", b"none") != b"none":
                raise ValueError("compressed chunks are not supported")
            for h2, d2 in _records(data):
                handle(h2, d2)
        else:
            handle(hdr, data)
    return out


def _header(body: bytes):
    seq, sec, nsec, n = struct.unpack_from("<IIII", body, 0)
    frame = body[16:16 + n].decode()
    return seq, sec, nsec, frame, 16 + n


@dataclass
class BagData:
    """Arrays extracted from the example bag."""
    uwb_stamp: np.ndarray      # [M] header stamp (s, float64) and exact (sec, nsec)
    uwb_sec: np.ndarray
    uwb_nsec: np.ndarray
    uwb_seq: np.ndarray
    uwb_frame: list
    uwb_requester: np.ndarray
    uwb_responder: np.ndarray
    uwb_distance: np.ndarray       # float32 values widened to float64
    uwb_distance_err: np.ndarray
    uwb_antenna: np.ndarray
    uwb_responder_location: np.ndarray  # [M][3]
    imu_sec: np.ndarray
    imu_nsec: np.ndarray
    imu_seq: np.ndarray
    imu_frame: list
    imu_quat_xyzw: np.ndarray      # [K][4]
    imu_orientation_cov: np.ndarray  # [K][9]
    vicon_stamp: np.ndarray
    vicon_pos: np.ndarray          # [V][3]
    vicon_quat_xyzw: np.ndarray    # [V][4]


def load_example_bag(path: str) -> BagData:
    u, im, vi = [], [], []
    for topic, typ, _t, body in read_bag(path):
        seq, sec, nsec, frame, o = _header(body)
        if typ == "uwb_driver/UwbRange":
            rq, rqi, rs, rsi, l1, l2, noise, vpeak, d, derr, ddot, ddoterr, ant, sw, ut = struct.unpack_from(
                "<BBBBHHHHffffBHI", body, o)
            loc = struct.unpack_from("<3d", body, o + struct.calcsize("<BBBBHHHHffffBHI"))
            u.append((seq, sec, nsec, frame, rq, rs, d, derr, ant, loc))
        elif typ == "sensor_msgs/Imu":
            q = struct.unpack_from("<4d", body, o)
            cov = struct.unpack_from("<9d", body, o + 32)
            im.append((seq, sec, nsec, frame, q, cov))
        elif typ.endswith("viconPoseMsg"):
            p = struct.unpack_from("<3d", body, o)
            q = struct.unpack_from("<4d", body, o + 24)
            vi.append((sec + nsec * 1e-9, p, q))
    f32 = lambda x: np.asarray(x, np.float32).astype(np.float64)
    return BagData(
        uwb_stamp=np.array([a[1] + a[2] * 1e-9 for a in u]), uwb_sec=np.array([a[1] for a in u], np.int64),
        uwb_nsec=np.array([a[2] for a in u], np.int64), uwb_seq=np.array([a[0] for a in u], np.int64),
        uwb_frame=[a[3] for a in u], uwb_requester=np.array([a[4] for a in u], np.int32),
        uwb_responder=np.array([a[5] for a in u], np.int32), uwb_distance=f32([a[6] for a in u]),
        uwb_distance_err=f32([a[7] for a in u]), uwb_antenna=np.array([a[8] for a in u], np.int32),
        uwb_responder_location=np.array([a[9] for a in u]),
        imu_sec=np.array([a[1] for a in im], np.int64), imu_nsec=np.array([a[2] for a in im], np.int64),
        imu_seq=np.array([a[0] for a in im], np.int64), imu_frame=[a[3] for a in im],
        imu_quat_xyzw=np.array([a[4] for a in im]), imu_orientation_cov=np.array([a[5] for a in im]),
        vicon_stamp=np.array([a[0] for a in vi]), vicon_pos=np.array([a[1] for a in vi]),
        vicon_quat_xyzw=np.array([a[2] for a in vi]))
