"""One C1-shaped window (uwb_only, 10 poses) per uwbgo_solve_batch call through the WINDOW path (driver for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from localization_b200 import Config, Solver, synthetic
from oracle import oracle
s = Solver(0)
topo, batch, _ = synthetic.uwb_only(1, 10, 4, seed=3)
cfg = Config(max_iterations=10)
for _ in range(4):
    got = s.solve(topo, batch, cfg)
assert s.last_path == 3
ref = oracle.solve(topo, batch, cfg)
assert np.array_equal(got.pose_t, ref.pose_t) and np.array_equal(got.chi2, ref.chi2)
print("window path ok, trials", got.status[0, 1])
