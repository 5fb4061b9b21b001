"""Developer loop: end-to-end (host buffers) timing of uwbgo_solve_batch vs pipeline settings."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Batch, Config, Result, Solver, synthetic, _ffi
from localization_b200.solver import pinned_empty
W, N, A = 65536, 50, 8
COMPACT = "--expanded" not in sys.argv
topo, batch, _ = synthetic.uwb_only(W, N, A, compact=COMPACT, shared_anchors=COMPACT)
cfg = Config(max_iterations=10)
s = Solver(0)
hb = Batch(pose_t=batch.pose_t, anchors=batch.anchors, range_d=batch.range_d, range_info=batch.range_info,
           shared_anchors=batch.shared_anchors)
for k in ("pose_t", "anchors", "range_d", "range_info"):
    if getattr(batch, k) is not None:
        a = pinned_empty(getattr(batch, k).shape); a[...] = getattr(batch, k); setattr(hb, k, a)
if batch.range_msgs is not None:
    from localization_b200.graph import RangeMsgs
    m = batch.range_msgs
    pin = lambda x, dt: None if x is None else np.copyto(pinned_empty(x.shape, dt), x) or None
    def pinned(x, dt):
        if x is None: return None
        a = pinned_empty(x.shape, dt); a[...] = x; return a
    hb.range_msgs = RangeMsgs.__new__(RangeMsgs)
    hb.range_msgs.distance, hb.range_msgs.distance_err = pinned(m.distance, np.float32), pinned(m.distance_err, np.float32)
    hb.range_msgs.dt_pose, hb.range_msgs.dt_anchor, hb.range_msgs.v_max = pinned(m.dt_pose, np.float64), None, m.v_max
res = Result(pinned_empty((W, N, 3)), None, None, pinned_empty((W, 4)), pinned_empty((W, 4), np.int32))
def run():
    s.solve(topo, hb, cfg, out=res)
settings = [(16384, 4), (8192, 8), (10944, 6), (13120, 5), (21856, 3), (32768, 2), (65536, 1), (4096, 8), (6144, 8), (5472, 8), (8192, 4), (12288, 8)]
for chunk, lanes in settings:
    s.set_pipeline(chunk, lanes)
    for _ in range(2): run()
    t0 = time.perf_counter()
    for _ in range(4): run()
    dt = (time.perf_counter() - t0) / 4
    print(f"chunk={chunk:6d} lanes={lanes}  {dt*1e3:7.2f} ms/step  {W/dt/1e6:.2f} M windows/s", flush=True)
