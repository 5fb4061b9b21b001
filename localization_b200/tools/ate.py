"""py3 absolute-trajectory-error tooling: timestamp association and Horn-aligned RMSE, the
semantics of the reference's script/associate.py:71-101 and script/evaluate_ate.py:47-79,129-162
(which are Python 2), on numpy arrays and TUM-format files ("stamp x y z qx qy qz qw")."""
from __future__ import annotations

import numpy as np


def read_trajectory(path: str):
    """TUM file -> (stamps [n], xyz [n][3]); '#' lines skipped; ',' and tabs accepted."""
    stamps, xyz = [], []
    with open(path) as f:
        for line in f:
            line = line.replace(",", " ").replace("\t", " ").strip()
            if not line or line[0] == "#":
                continue
            v = line.split()
            if len(v) > 3:
                stamps.append(float(v[0]))
                xyz.append([float(v[1]), float(v[2]), float(v[3])])
    return np.array(stamps), np.array(xyz).reshape(-1, 3)


def associate(first: np.ndarray, second: np.ndarray, offset: float = 0.0, max_difference: float = 0.02):
    """Greedy closest-stamp matching (best differences first, each stamp used once).
    Returns index pairs sorted by first stamp."""
    order2 = np.argsort(second)
    s2 = second[order2] + offset
    cand = []
    for i, a in enumerate(first):
        lo, hi = np.searchsorted(s2, a - max_difference), np.searchsorted(s2, a + max_difference)
        for j in range(lo, hi):
            d = abs(a - s2[j])
            if d < max_difference:
                cand.append((d, i, int(order2[j])))
    cand.sort()
    used1, used2, out = set(), set(), []
    for _d, i, j in cand:
        if i not in used1 and j not in used2:
            used1.add(i)
            used2.add(j)
            out.append((i, j))
    out.sort(key=lambda p: first[p[0]])
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def align(model: np.ndarray, data: np.ndarray):
    """Horn's closed-form alignment of model onto data (both [n][3]).
    Returns rot [3][3], trans [3], per-point translational error [n]."""
    mm, dm = model.mean(0), data.mean(0)
    Wm = (model - mm).T @ (data - dm)
    U, _, Vh = np.linalg.svd(Wm.T)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vh) < 0:
        S[2, 2] = -1
    rot = U @ S @ Vh
    trans = dm - rot @ mm
    err = (rot @ model.T).T + trans - data
    return rot, trans, np.sqrt((err * err).sum(1))


def evaluate_ate(gt_stamps, gt_xyz, est_stamps, est_xyz, offset=0.0, max_difference=0.02, do_align=True):
    """RMSE / mean / median / max of the translational error after association (+ alignment)."""
    m = associate(gt_stamps, est_stamps, offset, max_difference)
    if len(m) < 2:
        raise ValueError("Couldn't find matching timestamp pairs between groundtruth and estimated trajectory")
    gt, est = gt_xyz[m[:, 0]], est_xyz[m[:, 1]]
    if do_align:
        _, _, err = align(est, gt)
    else:
        d = est - gt
        err = np.sqrt((d * d).sum(1))
    return {"pairs": int(len(m)), "rmse": float(np.sqrt((err * err).mean())), "mean": float(err.mean()),
            "median": float(np.median(err)), "max": float(err.max()),
            "rmse_xy_raw": float(np.sqrt((((est - gt)[:, :2]) ** 2).sum(1).mean()))}
