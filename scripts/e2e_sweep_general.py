"""Developer loop: end-to-end (pinned host buffers) timing of uwbgo_solve_batch on the C4 shapes vs pipeline settings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Batch, Config, Result, Solver, synthetic
from localization_b200.solver import pinned_empty
ARR = ("pose_t", "pose_R", "anchors", "range_d", "range_info", "prior_Z", "prior_info", "se3_Z", "se3_info")
for name, make, N, iters in (("c4a", synthetic.uwb_imu_lidar, 20, 20), ("c4b", synthetic.uwb_twist, 15, 12)):
    W = 8192
    topo, batch, _ = make(W, N, 8)
    cfg = Config(max_iterations=iters)
    s = Solver(0)
    hb = Batch(pose_t=batch.pose_t, ant_offsets=batch.ant_offsets)
    for k in ARR:
        v = getattr(batch, k)
        if v is not None:
            a = pinned_empty(v.shape); a[...] = v; setattr(hb, k, a)
    res = Result(pinned_empty((W, N, 3)), pinned_empty((W, N, 3, 3)), None, pinned_empty((W, 4)), pinned_empty((W, 4), np.int32))
    for chunk, lanes in [(8192, 1), (4096, 2), (2048, 4), (1024, 8), (512, 8), (256, 8)]:
        s.set_pipeline(chunk, lanes)
        for _ in range(2): s.solve(topo, hb, cfg, out=res)
        t0 = time.perf_counter()
        for _ in range(4): s.solve(topo, hb, cfg, out=res)
        dt = (time.perf_counter() - t0) / 4
        print(f"{name} chunk={chunk:6d} lanes={lanes}  {dt*1e3:7.2f} ms/step  {W/dt/1e6:.3f} M windows/s", flush=True)
    s.close()
