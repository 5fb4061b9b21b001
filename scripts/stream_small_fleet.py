"""Step latency of small resident fleets (uwbgo_stream_step_robots): WINDOW kernels (one CTA per robot) against the
tile kernels, microseconds per step, host arrays in and out."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic
from localization_b200.stream import ResidentFleet

s = Solver(0)
for N, A in ((10, 4), (50, 8)):
    for W in (1, 8, 64, 148, 592):
        topo, b, _ = synthetic.uwb_only(W, N, A, seed=7, compact=True, shared_anchors=True)
        m, cfg = b.range_msgs, Config(max_iterations=10)
        aop = ((np.arange(N)[None, :] + np.arange(W)[:, None]) % A).astype(np.int32)
        out = []
        for limit in (-1, 0):
            s.set_window_path(limit)
            fleet = ResidentFleet(s, N, b.anchors, W, m.v_max, cfg)
            fleet.load(b.pose_t, aop, m.distance, m.distance_err, m.dt_pose)
            def step(k):
                fleet.step(((N + k + np.arange(W)) % A).astype(np.int32), m.distance[:, k % N], m.distance_err[:, k % N], m.dt_pose[:, k % (N - 1)])
            for k in range(5):
                step(k)
            t0 = time.perf_counter()
            for k in range(50):
                step(5 + k)
            out.append(((time.perf_counter() - t0) / 50 * 1e6, s.last_path))
            fleet.close()
        print(f"N={N:3d} robots={W:4d}  WINDOW kernels {out[0][0]:8.1f} us/step (path {out[0][1]})   tile kernels {out[1][0]:8.1f} us/step (path {out[1][1]})", flush=True)
s.set_window_path(-1)
