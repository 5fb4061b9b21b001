"""WINDOW path against the tile kernels by batch size (where should uwbgo_set_window_path's default sit?):
device-resident-free host calls (uwbgo_solve_batch), microseconds per call, C1 / C2 / C4a-shaped windows."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic

s = Solver(0)
cases = [
    ("uwb_only N=10 (C1, 3x3 blocks)", lambda W: synthetic.uwb_only(W, 10, 4, seed=3), 10),
    ("uwb_only N=50 (3x3 blocks)", lambda W: synthetic.uwb_only(W, 50, 8, seed=4), 10),
    ("uwb_imu N=12 (C2, block-diagonal)", lambda W: synthetic.uwb_imu_lidar(W, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=11), 10),
    ("uwb_imu_lidar N=20 3 antennas (6x6)", lambda W: synthetic.uwb_imu_lidar(W, 20, 8, seed=2), 20),
]

def timed(topo, batch, cfg, reps):
    for _ in range(3):
        s.solve(topo, batch, cfg)
    t0 = time.perf_counter()
    for _ in range(reps):
        s.solve(topo, batch, cfg)
    return (time.perf_counter() - t0) / reps * 1e6

for name, mk, it in cases:
    print(name)
    for W in (32, 148, 296, 444, 592, 888, 1184, 2368, 4736):
        topo, batch, _ = mk(W)
        cfg = Config(max_iterations=it)
        s.set_window_path(1 << 30)
        tw = timed(topo, batch, cfg, 20)
        pw = s.last_path
        s.set_window_path(0)
        tt = timed(topo, batch, cfg, 20)
        print(f"   W={W:5d}  window path {tw:9.1f} us (path {pw})   tile kernels {tt:9.1f} us (path {s.last_path})   "
              f"{'WINDOW' if tw < tt else 'tile'}", flush=True)
