"""WINDOW path (one CTA per window) against the CPU oracle and the tile kernels on a spread of
topologies, plus single-call latency through the host API next to the oracle on one host thread."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from localization_b200 import Config, Solver, synthetic
from oracle import oracle

def same(a, b):
    return all(np.array_equal(getattr(a, f), getattr(b, f)) for f in ("pose_t", "pose_R", "chi2", "status", "oplus_count"))

s = Solver(0)
cases = [
    ("uwb_only N=10", lambda W: synthetic.uwb_only(W, 10, 4, seed=3), 10),
    ("uwb_only N=50", lambda W: synthetic.uwb_only(W, 50, 8, seed=4), 10),
    ("uwb_imu N=12 (C2)", lambda W: synthetic.uwb_imu_lidar(W, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=11), 10),
    ("uwb_imu_lidar N=20 3 antennas", lambda W: synthetic.uwb_imu_lidar(W, 20, 8, seed=2), 20),
    ("uwb_twist N=15", lambda W: synthetic.uwb_twist(W, 15, 8, seed=5), 12),
    ("uwb_pose N=24 K=4", lambda W: synthetic.uwb_pose(W, 24, 8, keyframe_len=4, seed=6), 10),
]
bad = 0
for name, mk, it in cases:
    for W in (1, 33):
        topo, batch, _ = mk(W)
        cfg = Config(max_iterations=it)
        s.set_window_path(1 << 30)
        got = s.solve(topo, batch, cfg)
        path = s.last_path
        ref = oracle.solve(topo, batch, cfg, n_threads=8)
        s.set_window_path(0)
        tile = s.solve(topo, batch, cfg)
        ok = same(got, ref)
        bad += (not ok) or path != 3
        print(f"{name:32s} W={W:3d} path {path} window==oracle {ok} tile==oracle {same(tile, ref)} "
              f"max|dt| {np.abs(got.pose_t - ref.pose_t).max():.3g} trials {ref.status[:, 1].mean():.1f}", flush=True)
        if not ok:
            print("   status got", got.status[0], "ref", ref.status[0], "chi2 got", got.chi2[0], "ref", ref.chi2[0])
# counters near the re-orthogonalisation, vertex-1 offsets
topo, batch, _ = synthetic.uwb_imu_lidar(16, 12, 8, seed=5)
topo = synthetic.with_vertex1_offsets(topo, seed=3)
batch.oplus_count = np.random.default_rng(1).integers(985, 1001, size=(16, 12)).astype(np.int32)
cfg = Config(max_iterations=8)
s.set_window_path(1 << 30)
got = s.solve(topo, batch, cfg)
ok = same(got, oracle.solve(topo, batch, cfg, n_threads=8))
bad += not ok
print("counters + vertex-1 offsets: window==oracle", ok, "path", s.last_path)

print("\nlatency of ONE window per call through the host API (uwbgo_solve_batch, numpy arrays):")
for name, mk, it in cases:
    topo, batch, _ = mk(1)
    cfg = Config(max_iterations=it)
    out = {}
    for mode, wm in (("window", 1 << 30), ("tile", 0)):
        s.set_window_path(wm)
        for _ in range(5):
            s.solve(topo, batch, cfg)
        t0 = time.perf_counter()
        for _ in range(50):
            r = s.solve(topo, batch, cfg)
        out[mode] = (time.perf_counter() - t0) / 50 * 1e6
    for _ in range(3):
        oracle.solve(topo, batch, cfg, n_threads=1)
    t0 = time.perf_counter()
    for _ in range(20):
        oracle.solve(topo, batch, cfg, n_threads=1)
    cpu = (time.perf_counter() - t0) / 20 * 1e6
    print(f"{name:32s} window {out['window']:8.1f} us   tile {out['tile']:8.1f} us   oracle 1 thread {cpu:8.1f} us   trials {r.status[0, 1]}", flush=True)
print("FAILURES", bad)
sys.exit(1 if bad else 0)
