"""Developer loop: per-chunk timeline of one uwbgo_solve_batch call (UWBGO_PIPE_TRACE) per pipeline setting."""
import ctypes as C, os, sys, time
os.environ['UWBGO_PIPE_TRACE'] = '1'
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
settings = [(8192, 8), (12288, 8), (4096, 8)] if len(sys.argv) < 2 or not sys.argv[1][0].isdigit() else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:] if a[0].isdigit()]
for chunk, lanes in settings:
    s.set_pipeline(chunk, lanes)
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(2)
    os.dup2(devnull, 2)
    for _ in range(3): run()
    nb = int(os.environ.get("UWBGO_TRACE_BARRIER", "0"))
    if nb:  # several processes, one per GPU: start the traced calls together and keep every GPU busy around them
        open(f"/tmp/uwbgo_ready_{os.environ['UWBGO_TRACE_RANK']}", "w").close()
        while sum(os.path.exists(f"/tmp/uwbgo_ready_{k}") for k in range(nb)) < nb:
            time.sleep(0.001)
        for _ in range(20): run()
    os.dup2(saved, 2)
    t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
    if nb:
        os.dup2(devnull, 2)
        for _ in range(10): run()
        os.dup2(saved, 2)
    print(f"chunk={chunk:6d} lanes={lanes}  {dt*1e3:7.2f} ms wall (traced call)", file=sys.stderr, flush=True)
