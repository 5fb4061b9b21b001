import os, sys
sys.path.insert(0, '/root/repo')
from localization_b200 import Config, Solver, synthetic
s = Solver(0)
s.set_window_path(1 << 30)
topo, batch, _ = synthetic.uwb_imu_lidar(1, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=11)
for _ in range(2):
    s.solve(topo, batch, Config(max_iterations=10))
