"""Regenerates tests/golden/oracle_golden.npz from the CPU oracle (oracle/uwbgo_oracle.c).
The reference ships no golden vectors and its g2o cannot be built offline, so these vectors pin
the ORACLE (regressions, compiler changes), not the reference: parity with g2o stays unpinned.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from localization_b200 import Config, synthetic  # noqa: E402
from oracle import oracle  # noqa: E402

cases = {"uwb_only": (synthetic.uwb_only, dict(W=8, N=10, A=4, seed=101), 10),
         "uwb_imu_lidar": (synthetic.uwb_imu_lidar, dict(W=6, N=8, A=4, seed=102), 20),
         "uwb_twist": (synthetic.uwb_twist, dict(W=6, N=7, A=4, seed=103), 12),
         "uwb_pose": (synthetic.uwb_pose, dict(W=6, N=11, A=4, keyframe_len=4, seed=104), 10)}
out = {}
for name, (make, kw, iters) in cases.items():
    topo, batch, _ = make(**kw)
    r = oracle.solve(topo, batch, Config(max_iterations=iters))
    Hd, Ho, b, chi = oracle.linearize(topo, batch, Config())
    out[f"{name}_in_pose_t"] = batch.pose_t
    out[f"{name}_pose_t"], out[f"{name}_pose_R"] = r.pose_t, r.pose_R
    out[f"{name}_chi2"], out[f"{name}_status"] = r.chi2, r.status
    out[f"{name}_Hd"], out[f"{name}_b"] = Hd, b
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
print("wrote", len(out), "arrays")
