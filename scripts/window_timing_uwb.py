import os, sys
sys.path.insert(0, '/root/repo')
from localization_b200 import Config, Solver, synthetic
s = Solver(0)
s.set_window_path(1 << 30)
topo, batch, _ = synthetic.uwb_only(1, 10, 4, seed=3)
for _ in range(2):
    s.solve(topo, batch, Config(max_iterations=10))
