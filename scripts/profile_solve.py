"""Small driver for ncu captures: a few device-resident solves of the C3 workload (or a stage)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from localization_b200 import Config, Solver, synthetic, _ffi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=65536)
ap.add_argument("--poses", type=int, default=50)
ap.add_argument("--anchors", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--kind", default="uwb_only", choices=["uwb_only", "imu_lidar", "twist"])
ap.add_argument("--stage", default="solve", choices=["solve", "linearize"])
ap.add_argument("--window-path", type=int, default=-1, help="force the WINDOW path for batches up to this many windows")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.kind == "uwb_only":
    topo, batch, _ = synthetic.uwb_only(a.windows, a.poses, a.anchors)
elif a.kind == "imu_lidar":
    topo, batch, _ = synthetic.uwb_imu_lidar(a.windows, a.poses, a.anchors)
else:
    topo, batch, _ = synthetic.uwb_twist(a.windows, a.poses, a.anchors)
cfg = Config(max_iterations=a.iters)
s = Solver(0)
if a.window_path >= 0:
    s.set_window_path(a.window_path)
W, N = batch.n_windows, topo.n_poses
keep = {}
cb = _ffi.CBatch()
cb.n_windows = W
for k in ("pose_t", "pose_R", "anchors", "range_d", "range_info", "prior_Z", "prior_info", "se3_Z", "se3_info"):
    v = getattr(batch, k)
    if v is not None:
        keep[k] = torch.from_numpy(v).to(dev)
        setattr(cb, k, C.cast(C.c_void_p(keep[k].data_ptr()), C.POINTER(C.c_double)))
if batch.ant_offsets is not None:
    cb.ant_offsets = batch.ant_offsets.ctypes.data_as(C.POINTER(C.c_double))
pose = torch.empty((W, N, 3), dtype=torch.float64, device=dev)
chi2 = torch.empty((W, 4), dtype=torch.float64, device=dev)
status = torch.empty((W, 4), dtype=torch.int32, device=dev)
cr = _ffi.CResult()
cr.pose_t = C.cast(C.c_void_p(pose.data_ptr()), C.POINTER(C.c_double))
cr.chi2 = C.cast(C.c_void_p(chi2.data_ptr()), C.POINTER(C.c_double))
cr.status = C.cast(C.c_void_p(status.data_ptr()), C.POINTER(C.c_int32))
s.set_profiling(True)
st = torch.cuda.current_stream(dev).cuda_stream
if a.stage == "solve":
    for _ in range(a.reps):
        s.solve_device(topo, cb, cfg, cr, st)
        print("kernel ms", s.last_kernel_ms())
else:
    Hd = torch.empty((W, N, 36), dtype=torch.float64, device=dev)
    Ho = torch.empty((W, max(N - 1, 1), 36), dtype=torch.float64, device=dev)
    b = torch.empty((W, N, 6), dtype=torch.float64, device=dev)
    chi = torch.empty((W, 2), dtype=torch.float64, device=dev)
    for _ in range(a.reps):
        s.linearize_device(topo, cb, cfg, Hd.data_ptr(), Ho.data_ptr(), b.data_ptr(), chi.data_ptr(), st)
        print("kernel ms", s.last_kernel_ms())
torch.cuda.synchronize()
print("trials mean", status[:, 1].double().mean().item() if a.stage == "solve" else "-")
import hashlib  # noqa: E402
print("digest", hashlib.sha1(pose.cpu().numpy().tobytes() + chi2.cpu().numpy().tobytes()).hexdigest()[:12])
s.close()
