"""Time the device-resident fused LM kernel on C3 (developer loop): prints kernel ms."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from localization_b200 import Config, Solver, synthetic, _ffi
W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
A = int(sys.argv[3]) if len(sys.argv) > 3 else 8
dev = torch.device("cuda", 0)
topo, batch, _ = synthetic.uwb_only(W, N, A)
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
for _ in range(6):
    s.solve_device(topo, cb, cfg, cr, st)
    ms.append(s.last_kernel_ms())
torch.cuda.synchronize()
import hashlib
h = hashlib.sha1(pose.cpu().numpy().tobytes() + chi2.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"{os.environ.get('UWBGO_LIB','default'):40s} W={W} N={N} kernel_ms median={np.median(ms[2:]):.3f} min={min(ms):.3f}  -> {W/np.median(ms[2:])/1e3:.2f} M windows/s  hash={h}")
s.close()
