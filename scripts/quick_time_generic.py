"""like quick_time.py but on a topology that is NOT the standard chain (extra anchor edge on pose 0)
so the table-driven FAST kernel runs"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from localization_b200 import Config, Solver, synthetic, _ffi, Topology, Batch
W, N, A = 65536, 50, 8
dev = torch.device("cuda", 0)
topo, batch, _ = synthetic.uwb_only(W, N, A)
e = np.stack([topo.edge_kind, topo.edge_a, topo.edge_b, topo.edge_ant, topo.edge_robust], 1)
e = np.concatenate([e, [[0, 0, 1, 0, 1]]])
topo = Topology.from_edges(N, A, 0, e)
batch.range_d = np.concatenate([batch.range_d, batch.range_d[:, :1]], 1).copy()
batch.range_info = np.concatenate([batch.range_info, batch.range_info[:, :1]], 1).copy()
cfg = Config(max_iterations=10)
s = Solver(0)
cb = _ffi.CBatch(); cb.n_windows = W
keep = {}
for k in ("pose_t", "anchors", "range_d", "range_info"):
    keep[k] = torch.from_numpy(getattr(batch, k)).to(dev)
    setattr(cb, k, C.cast(C.c_void_p(keep[k].data_ptr()), C.POINTER(C.c_double)))
pose = torch.empty((W, N, 3), dtype=torch.float64, device=dev)
chi2 = torch.empty((W, 4), dtype=torch.float64, device=dev)
status = torch.empty((W, 4), dtype=torch.int32, device=dev)
cr = _ffi.CResult()
cr.pose_t = C.cast(C.c_void_p(pose.data_ptr()), C.POINTER(C.c_double))
cr.chi2 = C.cast(C.c_void_p(chi2.data_ptr()), C.POINTER(C.c_double))
cr.status = C.cast(C.c_void_p(status.data_ptr()), C.POINTER(C.c_int32))
s.set_profiling(True)
st = torch.cuda.current_stream(dev).cuda_stream
ms = []
for _ in range(5):
    s.solve_device(topo, cb, cfg, cr, st)
    ms.append(s.last_kernel_ms())
torch.cuda.synchronize()
print(f"{os.environ.get('UWBGO_LIB','default'):40s} generic fast path={s.last_path} kernel_ms median={np.median(ms[2:]):.3f}")
s.close()
