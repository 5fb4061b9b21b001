"""Latency of one window per call through the host API (BASELINE configs[0]/[1] shape) and
throughput at the C5 shape (N = 200, A = 16)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic
s = Solver(0)
for name, mk, it in (("uwb_only N=10 (C1 shape)", lambda: synthetic.uwb_only(1, 10, 4), 10),
                     ("uwb_imu N=12 (C2 shape)", lambda: synthetic.uwb_imu_lidar(1, 12, 4, antennas=0, lidar=False), 10)):
    topo, batch, _ = mk()
    cfg = Config(max_iterations=it)
    for _ in range(5):
        s.solve(topo, batch, cfg)
    t0 = time.perf_counter()
    for _ in range(50):
        s.solve(topo, batch, cfg)
    print(f"{name}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per single-window solve (host API, path {s.last_path})")
W = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
topo, batch, _ = synthetic.uwb_only(W, 200, 16, seed=5)
cfg = Config(max_iterations=10)
s.solve(topo, batch.slice(0, 4096), cfg)
t0 = time.perf_counter()
r = s.solve(topo, batch, cfg)
dt = time.perf_counter() - t0
print(f"C5 shape: {W} windows x N=200, A=16 through the host API (pageable numpy): {dt*1e3:.1f} ms -> {W/dt/1e3:.1f} K windows/s; trials mean {r.status[:,1].mean():.2f}")
