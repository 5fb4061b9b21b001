"""g2o text interchange for windows: the on-disk format either side of the hot path.

The reference registers its two edge types with g2o's factory as EDGE_RANGE and EDGE_RANGE_OFFSET
(src/types/types_edge_se3range.cpp:39, types_edge_se3range_offset.cpp:39) with the payloads
    EDGE_RANGE        id0 id1  meas info                    (types_edge_se3range.cpp:43-64)
    EDGE_RANGE_OFFSET id0 id1  pidFrom pidTo  meas info      (types_edge_se3range_offset.cpp:55-90)
next to g2o's stock
    VERTEX_SE3:QUAT   id  x y z qx qy qz qw          FIX id
    PARAMS_SE3OFFSET  pid x y z qx qy qz qw
    EDGE_SE3:QUAT     id0 id1  x y z qx qy qz qw  <21 upper-triangular information entries>
    EDGE_SE3_PRIOR    id  pid  x y z qx qy qz qw  <21 upper-triangular information entries>
so a window written here can be loaded by a real g2o build (with the reference's types linked) and
solved there — the only route to ever pin parity against the true reference.

EDGE_RANGE cannot carry the per-vertex antenna offset (setVertexOffset is not serialised by the
reference), so range edges with an antenna are written as EDGE_RANGE_OFFSET with
PARAMS_SE3OFFSET ids 0 (identity, localization.cpp:54-56) and k (antenna k).  g2o's text format has
no robust kernels; they are kept in '#' comment lines ("# ROBUST <edge index>"), which g2o skips.
Vertex ids follow the reference: pose slot*300 + ID (robot.cpp:43,94), anchors by node id."""
from __future__ import annotations

import numpy as np

from .._ffi import EDGE_PRIOR, EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE, EDGE_SE3
from ..graph import Batch, Topology

_UP = [(r, c) for r in range(6) for c in range(r, 6)]


def _quat(R):
    """rotation matrix -> (x, y, z, w), w >= 0 (Eigen's Quaterniond(Matrix3d))"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i], q[3] = 0.25 * s, (R[k, j] - R[j, k]) / s
        q[j], q[k] = (R[j, i] + R[i, j]) / s, (R[k, i] + R[i, k]) / s
    return -q if q[3] < 0 else q


def _rot(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _f(v):
    return " ".join(repr(float(x)) for x in np.ravel(v))


def write_window(path: str, topo: Topology, batch: Batch, w: int = 0, self_id: int = 200,
                 anchor_ids=None):
    """One window of a batch as a g2o text file."""
    N, A = topo.n_poses, topo.n_anchors
    anchor_ids = list(anchor_ids) if anchor_ids is not None else [100 + k for k in range(A)]
    pid = lambda i: i * 300 + self_id
    R = batch.pose_R[w].reshape(N, 3, 3) if batch.pose_R is not None else np.tile(np.eye(3), (N, 1, 1))
    t = batch.pose_t[w].reshape(N, 3)
    lines = ["PARAMS_SE3OFFSET 0 0 0 0 0 0 0 1"]
    for k in range(topo.n_antennas):
        lines.append(f"PARAMS_SE3OFFSET {k + 1} {_f(batch.ant_offsets[k])} 0 0 0 1")
    for a in range(A):
        lines.append(f"VERTEX_SE3:QUAT {anchor_ids[a]} {_f(batch.anchors[w].reshape(A, 3)[a])} 0 0 0 1")
        lines.append(f"FIX {anchor_ids[a]}")
    for i in range(N):
        lines.append(f"VERTEX_SE3:QUAT {pid(i)} {_f(t[i])} {_f(_quat(R[i]))}")
    sr = sp = ss = 0
    for e in range(topo.n_edges):
        k, a, b, ant = topo.edge_kind[e], topo.edge_a[e], topo.edge_b[e], topo.edge_ant[e]
        ant_b = 0 if topo.edge_ant_b is None else topo.edge_ant_b[e]
        if topo.edge_robust[e]:
            lines.append(f"# ROBUST {e}")
        if k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE):
            v1 = anchor_ids[b] if k == EDGE_RANGE_ANCHOR else pid(b)
            d, info = batch.range_d[w].reshape(-1)[sr], batch.range_info[w].reshape(-1)[sr]
            sr += 1
            if ant > 0 or ant_b > 0:
                lines.append(f"EDGE_RANGE_OFFSET {pid(a)} {v1} {ant} {ant_b} {_f(d)} {_f(info)}")
            else:
                lines.append(f"EDGE_RANGE {pid(a)} {v1} {_f(d)} {_f(info)}")
        elif k == EDGE_PRIOR:
            Z, I = batch.prior_Z[w].reshape(-1, 12)[sp], batch.prior_info[w].reshape(-1, 6, 6)[sp]
            sp += 1
            lines.append(f"EDGE_SE3_PRIOR {pid(a)} 0 {_f(Z[9:])} {_f(_quat(Z[:9].reshape(3, 3)))} "
                         f"{_f([I[r, c] for r, c in _UP])}")
        else:
            Z, I = batch.se3_Z[w].reshape(-1, 12)[ss], batch.se3_info[w].reshape(-1, 6, 6)[ss]
            ss += 1
            lines.append(f"EDGE_SE3:QUAT {pid(a)} {pid(b)} {_f(Z[9:])} {_f(_quat(Z[:9].reshape(3, 3)))} "
                         f"{_f([I[r, c] for r, c in _UP])}")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def read_window(path: str):
    """Inverse of write_window (any g2o file with these tags): returns (Topology, Batch of 1 window).
    Poses are the non-fixed vertices in ascending id order (ring-slot order of the reference)."""
    verts, fixed, params, edges, robust = {}, set(), {}, [], set()
    with open(path) as f:
        for line in f:
            v = line.split()
            if not v:
                continue
            if v[0] == "#":
                if len(v) >= 3 and v[1] == "ROBUST":
                    robust.add(int(v[2]))
                continue
            tag, x = v[0], v[1:]
            if tag == "VERTEX_SE3:QUAT":
                verts[int(x[0])] = np.array(x[1:8], float)
            elif tag == "FIX":
                fixed.add(int(x[0]))
            elif tag == "PARAMS_SE3OFFSET":
                params[int(x[0])] = np.array(x[1:8], float)
            elif tag == "EDGE_RANGE":
                edges.append(("range", int(x[0]), int(x[1]), 0, float(x[2]), float(x[3]), 0))
            elif tag == "EDGE_RANGE_OFFSET":
                edges.append(("range", int(x[0]), int(x[1]), int(x[2]), float(x[4]), float(x[5]), int(x[3])))
            elif tag in ("EDGE_SE3:QUAT", "EDGE_SE3_PRIOR"):
                vals = np.array(x[2:], float)
                I = np.zeros((6, 6))
                for (r, c), val in zip(_UP, vals[7:28]):
                    I[r, c] = I[c, r] = val
                Z = np.concatenate([_rot(vals[3:7]).reshape(9), vals[:3]])
                if tag == "EDGE_SE3:QUAT":
                    edges.append(("se3", int(x[0]), int(x[1]), Z, I))
                else:
                    edges.append(("prior", int(x[0]), Z, I))
    pose_ids = sorted(i for i in verts if i not in fixed)
    anchor_ids = sorted(i for i in verts if i in fixed)
    pidx = {v: k for k, v in enumerate(pose_ids)}
    aidx = {v: k for k, v in enumerate(anchor_ids)}
    n_ant = max([p for p in params if p > 0], default=0)
    te, rd, ri, pZ, pI, sZ, sI = [], [], [], [], [], [], []
    for e, ed in enumerate(edges):
        rb = 1 if e in robust else 0
        if ed[0] == "range":
            _, v0, v1, ant, d, info, ant_b = ed
            if v1 in aidx:
                te.append((EDGE_RANGE_ANCHOR, pidx[v0], aidx[v1], ant, rb, ant_b))
            else:
                te.append((EDGE_RANGE_POSE, pidx[v0], pidx[v1], ant, rb, ant_b))
            rd.append(d)
            ri.append(info)
        elif ed[0] == "prior":
            te.append((EDGE_PRIOR, pidx[ed[1]], 0, 0, rb))
            pZ.append(ed[2])
            pI.append(ed[3])
        else:
            te.append((EDGE_SE3, pidx[ed[1]], pidx[ed[2]], 0, rb))
            sZ.append(ed[3])
            sI.append(ed[4])
    topo = Topology.from_edges(len(pose_ids), len(anchor_ids), n_ant, te)
    P = np.array([verts[i] for i in pose_ids])
    R = np.array([_rot(p[3:7]) for p in P])
    arr = lambda a, shape: np.array(a, float).reshape((1,) + shape) if len(a) else None
    batch = Batch(pose_t=P[None, :, :3], pose_R=None if np.array_equal(R, np.tile(np.eye(3), (len(P), 1, 1))) else R[None],
                  anchors=np.array([verts[i][:3] for i in anchor_ids])[None] if anchor_ids else None,
                  range_d=arr(rd, (len(rd),)), range_info=arr(ri, (len(ri),)),
                  ant_offsets=np.array([params[k + 1][:3] for k in range(n_ant)]) if n_ant else None,
                  prior_Z=arr(pZ, (len(pZ), 12)), prior_info=arr(pI, (len(pI), 6, 6)),
                  se3_Z=arr(sZ, (len(sZ), 12)), se3_info=arr(sI, (len(sI), 6, 6)))
    return topo, batch
