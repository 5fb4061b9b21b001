#!/usr/bin/env python
"""bench.py — LM-solved UWB windows per second (BASELINE.json metric) on N B200s of one node.

A step = one pass of the hot path (optimizer.initializeOptimization() + optimize(10), i.e.
uwbgo_solve_batch) over one batch of synthetic C3 windows: 65,536 UWB-only windows per GPU,
8 anchors, 50-pose window, 10 LM iterations.  Windows are independent, so
ranks shard them with no data-path collective (weak scaling: 65,536 windows per GPU); the one
NCCL gather of final poses to rank 0 that north_star allows is inside the timed region.

  value  windows/s with the batch already resident in HBM (uwbgo_solve_batch_device)
  e2e    windows/s through the host-pointer C-ABI call (uwbgo_solve_batch) with pinned HOST
         buffers: host->device and device->host copies inside the timed region
  --impl reference   the CPU restatement of the reference's g2o LM (oracle/) on all host cores,
                     the same 65,536 C3 windows per step
  --scaling strong   65,536 windows in TOTAL, split over the ranks (the configuration north_star quotes its
                     target on); at N > 1 the default weak-scaling line carries the strong numbers as well
  --workload c1|c2   BASELINE configs[0] / [1]: ONE window per solve() call (Localization::addRangeEdge ->
                     solve(), localization.cpp:371-375): microseconds per call of uwbgo_solve_batch next to the
                     oracle on one host thread, for batches of 1, 8 and 32 windows
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOWS_PER_GPU = 65536
N_POSES, N_ANCHORS, LM_ITERS = 50, 8, 10
METRIC = "LM-solved UWB windows/sec"
UNIT = "windows/s"
WORKLOAD = ("C3: synthetic UWB-only windows, 65536 per GPU, 8 anchors (one constellation for the fleet), 50-pose window, "
            "10 LM iterations, 99 edges (50 anchor-range + 49 trajectory), Cauchy kernel; range data passed as message "
            "fields (float32 distance / distance_err, stamp differences), edge parameters built on the device")
# algorithmic bytes per window of the fused solve (inputs 4,584 + outputs 2,824)
ALGO_BYTES_SOLVE = 7408
ALGO_BYTES_LINEARIZE = 35496


NOTE_C3 = ("7.4 KB of algorithmic bytes per window against ~1.1 Mflop of FP64 per window: by the algorithmic "
                        "bytes the fused solve is far from the HBM roof (SURVEY 8d).  What it really moves is its scratch: "
                        "the per-pose substitution records and estimates are streamed out and back once per LM trial "
                        "(traffic, from ncu), and at the measured kernel time that stream runs at traffic_rate_gbs = "
                        "traffic_frac of the HBM peak -- the kernel is bounded by its own scratch traffic and by FP64 "
                        "issue latency at the same time (DESIGN.md section 4)")
NOTE_GENERAL = ("6x6 windows (rotations, antenna offsets, EdgeSE3Prior / EdgeSE3): by the algorithmic bytes the fused solve is "
                "far from the HBM roof; the ITEM kernel (one CTA of 8 warps per tile of 32 windows at 128 registers, two tiles "
                "per SM, the LM trial cut into per-edge / per-pose items drawn from a queue, the serial elimination on two "
                "warps) is bounded by L2 / FP64 latency with 16 warps per SM, 8,192 windows being 1.7 tiles per SM "
                "(DESIGN.md sections 3.2 and 4.2); traffic = linearisation, H and L records, information matrices and "
                "spills streamed once per LM trial")


def _info_diag(made):
    topo, batch, extra = made
    return topo, batch.with_info_diag(), extra


ARRAYS = ("pose_t", "pose_R", "anchors", "range_d", "range_info", "prior_Z", "prior_info", "se3_Z", "se3_info")
# name -> (description, default windows per GPU, generator, LM iterations, SURVEY algorithmic bytes per window or None)
WORKLOADS = {
    "c3": (WORKLOAD, WINDOWS_PER_GPU,
           lambda syn, W, seed: syn.uwb_only(W, N_POSES, N_ANCHORS, seed=seed, compact=True, shared_anchors=True),
           LM_ITERS, ALGO_BYTES_SOLVE),
    # the same windows with the range data expanded on the host and per-window (jittered) anchors: round 1's input form
    "c3x": (WORKLOAD.split(";")[0].replace(" (one constellation for the fleet)", "") + "; expanded input form "
            "(range_d / range_info doubles, per-window anchors)", WINDOWS_PER_GPU,
            lambda syn, W, seed: syn.uwb_only(W, N_POSES, N_ANCHORS, seed=seed), LM_ITERS, ALGO_BYTES_SOLVE),
    "c4a": ("C4a: synthetic uwb_imu_lidar windows, 8 anchors, 20-pose window, 3 antennas with lever arms, IMU + lidar "
            "EdgeSE3Prior on every pose but the newest, 20 LM iterations; information matrices of the priors passed as "
            "their diagonals (UWBGO_DIAG_INFO: what Localization builds is diagonal), rebuilt on the device", 8192,
            lambda syn, W, seed: _info_diag(syn.uwb_imu_lidar(W, 20, 8, seed=seed)), 20, None),
    "c4b": ("C4b: synthetic uwb_twist windows, 8 anchors, 15-pose window, twist EdgeSE3 chain, merged-covariance "
            "ranges with 3 antennas, 12 LM iterations; information matrices of the twist edges passed as their "
            "diagonals (UWBGO_DIAG_INFO), rebuilt on the device", 8192,
            lambda syn, W, seed: _info_diag(syn.uwb_twist(W, 15, 8, seed=seed)), 12, None),
    # the same windows with full 6x6 information matrices on the wire (the input form of the first half of round 2)
    "c4ax": ("C4a with full 6x6 information matrices on the wire", 8192,
             lambda syn, W, seed: syn.uwb_imu_lidar(W, 20, 8, seed=seed), 20, None),
    "c4bx": ("C4b with full 6x6 information matrices on the wire", 8192,
             lambda syn, W, seed: syn.uwb_twist(W, 15, 8, seed=seed), 12, None),
    "c5": ("C5: synthetic UWB-only Monte-Carlo windows, 16 anchors, 200-pose window, 10 LM iterations "
           "(1,048,576 windows over 8 GPUs = 131,072 per GPU)", 131072,
           lambda syn, W, seed: syn.uwb_only(W, 200, 16, seed=seed, compact=True, shared_anchors=True), 10, 29200),
}


# BASELINE configs[0] / [1]: the window one range message of bag/data_example.bag makes solve() optimise
# (cfg/uwb_only.yaml: trajectory_length 10, 10 iterations; cfg/uwb_imu.yaml: 12 poses, IMU prior on every pose
# but the newest, maximum_velocity 3, 10 iterations)
LATENCY_WORKLOADS = {
    "c1": ("C1 shape: one uwb_only window per call (10 poses, 4 anchors, 19 range edges, 10 LM iterations)",
           lambda syn, W, seed: syn.uwb_only(W, 10, 4, seed=seed), 10),
    "c2": ("C2 shape: one uwb_imu window per call (12 poses, 4 anchors, 23 range edges + 11 IMU EdgeSE3Prior, "
           "10 LM iterations)",
           lambda syn, W, seed: syn.uwb_imu_lidar(W, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=seed), 10),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(index):
    """Several ranks on one host: run this rank (and first-touch its pinned host buffers) on the CPUs NVML reports
    as closest to its GPU, so the end-to-end leg's host<->device copies do not cross sockets.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({min(cpus)}-{max(cpus)})"
    except Exception as e:  # noqa: BLE001
        return f"unavailable: {type(e).__name__}"
    return None


def cpu_baseline(topo, batch, cfg, budget_s=12.0, name="C3"):
    """oracle (CPU restatement of the reference's g2o LM) on a bounded sample, all host threads"""
    from oracle import oracle
    cores = os.cpu_count() or 1
    probe = min(batch.n_windows, 64 * cores)
    t0 = time.perf_counter()
    oracle.solve(topo, batch.slice(0, probe), cfg, n_threads=cores)
    dt = time.perf_counter() - t0
    n = int(min(batch.n_windows, max(probe, probe * budget_s / max(dt, 1e-6))))
    t0 = time.perf_counter()
    oracle.solve(topo, batch.slice(0, n), cfg, n_threads=cores)
    dt = time.perf_counter() - t0
    n1 = int(min(batch.n_windows, max(64, n / dt / cores * 1.5)))   # about 1.5 s on one thread
    t0 = time.perf_counter()
    oracle.solve(topo, batch.slice(0, n1), cfg, n_threads=1)
    dt1 = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of the {batch.n_windows} {name} windows, {cores} threads, {dt:.2f} s",
            "single_core": {"value": n1 / dt1, "unit": UNIT, "sample": f"{n1} windows, 1 thread, {dt1:.2f} s"}}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (its g2o/CHOLMOD cannot be built offline; this is
    the g2o-faithful port in oracle/) on all host threads; rank 0 only.  Every step solves the same
    65,536 C3 windows a GPU step solves (shrunk, and said so, only if the run would exceed ~4 minutes)."""
    if rank != 0:
        return
    from localization_b200 import Config, synthetic
    from oracle import oracle
    cores = os.cpu_count() or 1
    cfg = Config(max_iterations=LM_ITERS)
    n = WINDOWS_PER_GPU
    topo, batch, _ = WORKLOADS["c3"][2](synthetic, n, synthetic.SEED_C3)   # the windows rank 0 of the GPU arm solves
    t0 = time.perf_counter()
    oracle.solve(topo, batch.slice(0, 64 * cores), cfg, n_threads=cores)
    rate = 64 * cores / (time.perf_counter() - t0)
    budget = 240.0
    if n * (args.steps + max(args.warmup, 1)) / rate > budget:
        n = int(max(512 * cores, budget * rate / (args.steps + max(args.warmup, 1))))
        batch = batch.slice(0, n)
    for _ in range(max(args.warmup, 1)):
        oracle.solve(topo, batch, cfg, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.solve(topo, batch, cfg, n_threads=cores)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    full = n == WINDOWS_PER_GPU
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "windows_per_gpu": WINDOWS_PER_GPU, "n_poses": N_POSES,
                       "n_anchors": N_ANCHORS, "lm_iterations": LM_ITERS, "windows_per_step": n,
                       "same_windows_as_gpu_step": full},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": (f"all {n} C3 windows of a GPU step" if full else f"{n} of the 65536 C3 windows")
                                       + f" per step x {args.steps} steps, {cores} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_latency(args):
    """--workload c1|c2: one window per solve() call, the reference's own call pattern."""
    import ctypes as C
    import torch
    from localization_b200 import Config, Result, Solver, synthetic, _ffi
    from oracle import oracle
    desc, make, iters = LATENCY_WORKLOADS[args.workload]
    cfg = Config(max_iterations=iters)
    solver = Solver(0)
    olib = oracle.load()
    cores = os.cpu_count() or 1
    calls = max(args.steps, 1) * 20
    sampler = ClockSampler(0)
    sampler.start()
    l0 = solver.launch_count
    table, same_all = {}, True
    solver.set_profiling(True)
    for W in (1, 8, 32):
        topo, batch, _ = make(synthetic, W, synthetic.SEED_C3 + W)
        res, ref = Result.empty(W, topo.n_poses), Result.empty(W, topo.n_poses)
        t_, b_, c_, r_, o_ = topo.c_struct(), batch.c_struct(), cfg.c_struct(), res.c_struct(), ref.c_struct()

        def gpu_call():
            rc = solver._lib.uwbgo_solve_batch(solver._h, C.byref(t_), C.byref(b_), C.byref(c_), C.byref(r_))
            if rc:
                raise RuntimeError(solver._lib.uwbgo_last_error().decode())

        def cpu_call(nt):
            rc = olib.uwbgo_oracle_solve_batch(C.byref(t_), C.byref(b_), C.byref(c_), C.byref(o_), None, nt)
            if rc:
                raise RuntimeError(f"oracle: {rc}")

        def timed(fn, n):
            for _ in range(max(args.warmup, 3)):
                fn()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            return (time.perf_counter() - t0) / n * 1e6

        row = {}
        for mode, wm in (("window_path", -1), ("tile_path", 0)):
            solver.set_window_path(wm)
            row[mode + "_us"] = timed(gpu_call, calls)
            row[mode + "_kernel_us"] = 1e3 * solver.mean_kernel_ms(min(calls, 64))
            row[mode + "_id"] = solver.last_path
        solver.set_window_path(-1)
        gpu_call()
        row["cpu_1_thread_us"] = timed(lambda: cpu_call(1), max(calls // 4, 5))
        if W > 1:
            row["cpu_all_threads_us"] = timed(lambda: cpu_call(min(cores, W)), max(calls // 4, 5))
        same = all(np.array_equal(getattr(res, f), getattr(ref, f)) for f in ("pose_t", "pose_R", "chi2", "status", "oplus_count"))
        row["bit_identical_to_oracle"] = bool(same)
        row["lm_trials_mean"] = float(ref.status[:, 1].mean())
        same_all &= same
        table[str(W)] = row
        if W == 1:
            h2d = sum(getattr(batch, k).nbytes for k in ARRAYS if getattr(batch, k) is not None)
            d2h = res.pose_t.nbytes + res.pose_R.nbytes + res.oplus_count.nbytes + res.chi2.nbytes + res.status.nbytes
    launches = solver.launch_count - l0
    clocks = sampler.stop()
    one = table["1"]
    line = {"metric": "latency of one LM-solved window per solve() call (uwbgo_solve_batch, host buffers in and out)",
            "value": one["window_path_us"], "unit": "us per call", "n_gpus": 1, "steps": calls, "warmup": max(args.warmup, 3),
            "ms_per_step": one["window_path_us"] * 1e-3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "windows_per_call": 1, "lm_iterations": iters,
                       "l2": "one window (a few KB) lives in the shared memory of one SM for the whole call"},
            "e2e": {"value": one["window_path_us"], "unit": "us per call", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "matches_oracle": bool(same_all),
                    "note": "the call IS the end-to-end path: host arrays in, host arrays out, one kernel launch"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "fp64-latency", "kernel": "lm_window_kernel (one CTA per window, state in shared memory)",
                         "kernel_us": one["window_path_kernel_us"], "achieved": None, "peak": None, "unit": None, "frac": None,
                         "traffic": None,
                         "note": "one window is a serial chain: iterations x poses x Cholesky pivots (three per pose on translation-only and block-diagonal windows, six on full 6x6 blocks), each a dependent "
                                 "sqrt -> reciprocal -> multiply -> FMA sequence (~165 cycles of FP64 latency); neither HBM nor "
                                 "the FP64 pipe is loaded by a single window, the measure is microseconds per call next to "
                                 "the CPU (DESIGN.md section 4.3)"},
            "cpu_baseline": {"value": one["cpu_1_thread_us"], "unit": "us per call", "cores": 1, "kind": "port",
                             "sample": f"{max(calls // 4, 5)} calls of the oracle on the same window, one host thread "
                                       "(one window cannot use more)"},
            "batches": table}
    print(json.dumps(line), flush=True)
    solver.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--windows", type=int, default=0, help="windows per GPU (0 = the workload's own)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--stages", action="store_true", help="(default on C3 at N=1) also time the linearise stage kernel")
    ap.add_argument("--no-stages", action="store_true", help="skip the linearise stage timing")
    ap.add_argument("--no-resident", action="store_true", help="skip the resident-fleet (uwbgo_stream) leg")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + sorted(LATENCY_WORKLOADS),
                    help="c3 = BASELINE metric config (default); c1 / c2 = one window per call (latency); "
                         "the others are the remaining BASELINE configs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's windows PER GPU (default); strong: that many windows in total")
    ap.add_argument("--verify-shards", action="store_true",
                    help="N > 1: rank 0 re-solves every other rank's shard and checks the gathered results bit for bit")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload in LATENCY_WORKLOADS:
        if rank == 0:
            run_latency(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from localization_b200 import Config, Solver, synthetic, _ffi
    from localization_b200.solver import pinned_empty

    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl_desc, wl_W, wl_make, wl_iters, wl_bytes = WORKLOADS[args.workload]
    cfg = Config(max_iterations=wl_iters)
    solver = Solver(local)
    solver.set_profiling(True)
    stream = torch.cuda.current_stream(dev)
    import ctypes as C
    from localization_b200 import Batch, Result
    pd = lambda t: C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double))
    pd_np = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(W, seed0):
        """device-resident and end-to-end legs on W windows per rank (rank r solves the windows of seed0 + r)"""
        topo, batch, _ = wl_make(synthetic, W, seed0 + rank)
        N = topo.n_poses
        present = [k for k in ARRAYS if getattr(batch, k) is not None]
        general = batch.pose_R is not None
        msgs = batch.range_msgs
        MSG = (("distance", np.float32), ("distance_err", np.float32), ("dt_anchor", np.float64), ("dt_pose", np.float64))
        pf = lambda t: C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))

        def range_msgs_struct(arrays, as_ptr):
            """uwbgo_range_msgs over `arrays` (name -> tensor / ndarray of the message fields)"""
            cm = _ffi.CRangeMsgs()
            for name, dt in MSG:
                if arrays.get(name) is not None:
                    setattr(cm, name, as_ptr(arrays[name], dt))
            cm.v_max = msgs.v_max
            return cm
        # ---- device-resident leg: results in ONE buffer per rank (poses | chi2 | status | rotations), so the
        # one gather north_star allows carries final poses, chi2 and status words
        d_in = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in present}
        n_t, n_c, n_s, n_r = W * N * 3, W * 4, W * 2, (W * N * 9 if general else 0)   # in doubles; 4 int32 = 2 doubles
        d_out = torch.empty(n_t + n_c + n_s + n_r, dtype=torch.float64, device=dev)
        d_pose = d_out[:n_t].view(W, N, 3)
        d_chi2 = d_out[n_t:n_t + n_c].view(W, 4)
        d_status = d_out[n_t + n_c:n_t + n_c + n_s].view(torch.int32).view(W, 4)
        d_rot = d_out[n_t + n_c + n_s:].view(W, N, 9) if general else None
        gather = [torch.empty_like(d_out) for _ in range(world)] if (world > 1 and rank == 0) else None
        cb = _ffi.CBatch()
        cb.n_windows = W
        for k in present:
            setattr(cb, k, pd(d_in[k]))
        cb.shared = (_ffi.SHARED_ANCHORS if batch.shared_anchors else 0) | (_ffi.DIAG_INFO if batch.info_diag else 0)
        if msgs is not None:
            d_msg = {n: torch.from_numpy(getattr(msgs, n)).to(dev) for n, _ in MSG if getattr(msgs, n) is not None}
            cm_dev = range_msgs_struct(d_msg, lambda t, dt: pf(t) if dt == np.float32 else pd(t))
            cb.range_msgs = C.pointer(cm_dev)
        if batch.ant_offsets is not None:
            cb.ant_offsets = batch.ant_offsets.ctypes.data_as(C.POINTER(C.c_double))
        cr = _ffi.CResult()
        cr.pose_t, cr.chi2 = pd(d_pose), pd(d_chi2)
        if general:
            cr.pose_R = pd(d_rot)
        cr.status = C.cast(C.c_void_p(d_status.data_ptr()), C.POINTER(C.c_int32))

        def step_device():
            solver.solve_device(topo, cb, cfg, cr, stream.cuda_stream)
            if world > 1:
                dist.gather(d_out, gather, dst=0)

        for _ in range(args.warmup):
            step_device()
        barrier()
        l0 = solver.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = solver.launch_count - l0
        # kernel-only duration of the fused LM kernel: the library brackets it with CUDA events on the
        # launching stream in every call; average over the launches of the timed region
        k_ms = solver.mean_kernel_ms(min(args.steps, 64))
        torch.cuda.synchronize(dev)
        out = {"W": W, "N": N, "topo": topo, "batch": batch, "general": general,
               "value": W * world * args.steps / (ms_total * 1e-3), "ms_per_step": ms_total / args.steps,
               "launches": int(launches), "kernel_ms": k_ms, "path": solver.last_path}
        # shards verified on rank 0: every other rank's windows re-solved from their seed, gathered results bit-equal
        if args.verify_shards and world > 1 and rank == 0:
            ok = True
            for r in range(1, world):
                t2, b2, _ = wl_make(synthetic, W, seed0 + r)
                got = solver.solve(t2, b2, cfg)
                g = gather[r].cpu()
                ok &= bool(np.array_equal(g[:n_t].view(W, N, 3).numpy(), got.pose_t))
                ok &= bool(np.array_equal(g[n_t:n_t + n_c].view(W, 4).numpy(), got.chi2))
                ok &= bool(np.array_equal(g[n_t + n_c:n_t + n_c + n_s].view(torch.int32).view(W, 4).numpy(), got.status))
            out["shards_verified"] = ok

        # ---- end-to-end leg: host buffers through uwbgo_solve_batch ----------------------------
        hb = Batch(pose_t=batch.pose_t, ant_offsets=batch.ant_offsets, shared_anchors=batch.shared_anchors,
                   info_diag=batch.info_diag)
        for k in present:
            a = pinned_empty(getattr(batch, k).shape)
            a[...] = getattr(batch, k)
            setattr(hb, k, a)
        h_msg = {}
        if msgs is not None:
            for n, dt in MSG:
                if getattr(msgs, n) is not None:
                    h_msg[n] = pinned_empty(getattr(msgs, n).shape, dt)
                    h_msg[n][...] = getattr(msgs, n)
        hres = Result(pinned_empty((W, N, 3)), pinned_empty((W, N, 3, 3)) if general else None, None,
                      pinned_empty((W, 4)), pinned_empty((W, 4), np.int32))
        h2d = sum(getattr(hb, k).nbytes for k in present) + sum(a.nbytes for a in h_msg.values())
        d2h = hres.pose_t.nbytes + hres.chi2.nbytes + hres.status.nbytes + (hres.pose_R.nbytes if general else 0)
        t_, b_, c_ = topo.c_struct(), hb.c_struct(), cfg.c_struct()
        if msgs is not None:
            cm_host = range_msgs_struct(h_msg, lambda a, dt: a.ctypes.data_as(C.POINTER(C.c_float if dt == np.float32 else C.c_double)))
            b_.range_msgs = C.pointer(cm_host)
        r_ = _ffi.CResult()
        r_.pose_t, r_.chi2 = pd_np(hres.pose_t), pd_np(hres.chi2)
        if general:
            r_.pose_R = pd_np(hres.pose_R)
        r_.status = hres.status.ctypes.data_as(C.POINTER(C.c_int32))

        def step_host():
            rc = solver._lib.uwbgo_solve_batch(solver._h, C.byref(t_), C.byref(b_), C.byref(c_), C.byref(r_))
            if rc:
                raise RuntimeError(solver._lib.uwbgo_last_error().decode())
        for _ in range(args.warmup):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        out.update({"e2e_value": W * world * args.steps / e2e_s, "h2d": h2d, "d2h": d2h,
                    # the two legs must agree bit for bit
                    "same": bool(np.array_equal(hres.pose_t, d_pose.cpu().numpy())
                                 and np.array_equal(hres.chi2, d_chi2.cpu().numpy())
                                 and np.array_equal(hres.status, d_status.cpu().numpy()))})
        # ---- resident fleet (uwbgo_stream): the same robots, their windows kept on the device; a step = one range
        # message per robot in (host arrays), shift + solve on the device, newest pose + chi2 + status out.  UWB-only
        # chain windows in the compact form with one anchor constellation (c3 / c5)
        if msgs is not None and batch.shared_anchors and not general and not args.no_resident:
            from localization_b200.stream import ResidentFleet
            fleet = ResidentFleet(solver, N, batch.anchors, W, msgs.v_max, cfg)
            A = topo.n_anchors
            # the messages of the steps: the fields of the window's own edges, taken round robin (same statistics)
            def pinned_col(a, k, dt):
                c = pinned_empty((W,), dt)
                c[...] = a[:, k]
                return c
            m_d = [pinned_col(msgs.distance, k % N, np.float32) for k in range(8)]
            m_e = [pinned_col(msgs.distance_err, k % N, np.float32) for k in range(8)]
            m_t = [pinned_col(msgs.dt_pose, k % (N - 1), np.float64) for k in range(8)]

            def run_resident(per_robot):
                # per_robot: every robot ranges its own anchor in a step (robot w starts its round robin at w), the
                # anchor id is a fifth message field; else one anchor per step for the whole fleet
                first = (np.arange(W, dtype=np.int64) % A)[:, None] if per_robot else 0
                aop = ((np.arange(N)[None, :] + first) % A).astype(np.int32) if per_robot else np.arange(N) % A
                fleet.load(batch.pose_t, aop, msgs.distance, msgs.distance_err, msgs.dt_pose)
                m_a = [pinned_col(((N + k + first) % A).astype(np.int32), 0, np.int32) for k in range(A)] if per_robot else None
                step_no = [0]

                def step_resident():
                    k = step_no[0]
                    step_no[0] += 1
                    fleet.step(m_a[k % A] if per_robot else (N + k) % A, m_d[k % 8], m_e[k % 8], m_t[k % 8])

                for _ in range(args.warmup):
                    step_resident()
                barrier()
                l1 = solver.launch_count
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    step_resident()
                barrier()
                res_s = max_over_ranks(time.perf_counter() - t0)
                return {"value": W * world * args.steps / res_s, "unit": UNIT, "ms_per_step": res_s / args.steps * 1e3,
                        "h2d_bytes_per_step": W * (20 if per_robot else 16), "d2h_bytes_per_step": W * 72,
                        "kernel_ms": solver.mean_kernel_ms(min(args.steps, 64)),
                        "gpu_launches": int(solver.launch_count - l1) + 2 * args.steps}

            out["resident"] = run_resident(False)
            out["resident"]["what"] = ("uwbgo_stream_step: the windows of the same robots RESIDENT in HBM (the reference's call "
                                       "pattern: one range message per robot -> new vertex = copy of the newest estimate, "
                                       "oldest vertex dropped, solve; localization.cpp:297-376, robot.cpp:75-110); per step "
                                       "one message per robot host -> device (16 B) and the newest pose + chi2 + status "
                                       "device -> host (72 B) inside the timed region, every step a full "
                                       f"{wl_iters}-iteration LM solve of a {N}-pose window; bit-identical to the oracle on "
                                       "host-shifted windows (tests/test_gpu_stream.py)")
            out["resident"]["per_robot_anchors"] = run_resident(True)
            out["resident"]["per_robot_anchors"]["what"] = ("uwbgo_stream_step_robots: the same with one anchor sequence per robot "
                                                            "(the anchor id is a fifth message field, 20 B per robot and step)")
            fleet.close()
        return out

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    W_weak = args.windows or wl_W
    W_strong = max((args.windows or wl_W) // world, 1)
    primary = measure(W_strong if args.scaling == "strong" else W_weak, synthetic.SEED_C3)
    other = None
    if world > 1:   # the other scaling mode rides along (same steps, same timing rules)
        other = measure(W_weak if args.scaling == "strong" else W_strong, synthetic.SEED_C3 + 1000)
    clocks = sampler.stop() if rank == 0 else None   # sampled across all timed regions

    if rank == 0:
        m = primary
        W, N, topo, batch = m["W"], m["N"], m["topo"], m["batch"]
        if wl_bytes is None:   # inputs as passed + estimates, chi2 and status out
            wl_bytes = m["h2d"] // W + N * (96 if m["general"] else 24) + 48
        hbm, how = peaks()
        k_best = m["kernel_ms"] if m["kernel_ms"] and m["kernel_ms"] > 0 else None
        achieved = wl_bytes * W / (k_best * 1e-3) / 1e9 if k_best else None
        traffic, traffic_src, fp64_flop, fp64_pipe = None, None, None, None
        tkey = {"c3": ("lm_chain_tma_kernel", WINDOWS_PER_GPU), "c3x": ("lm_chain_tma_kernel", WINDOWS_PER_GPU),
                "c4a": ("lm_general_items_kernel_c4a", 8192), "c4b": ("lm_general_items_kernel_c4b", 8192),
                "c4ax": ("lm_general_items_kernel_c4a", 8192), "c4bx": ("lm_general_items_kernel_c4b", 8192),
                "c5": ("lm_chain_tma_kernel_c5", 131072)}.get(args.workload)
        for tp in (os.path.join(ROOT, "profiles", "r02_traffic.json"), os.path.join(ROOT, "profiles", "r01_traffic.json")):
            if os.path.exists(tp) and tkey and W == tkey[1]:
                with open(tp) as f:
                    tj = json.load(f).get(tkey[0])
                if tj:
                    traffic, traffic_src = tj["dram_bytes_per_launch"] / 1e9, tj["source"]
                    fp64_flop = tj.get("fp64_flop_per_launch")
                    fp64_pipe = tj.get("fp64_pipe_active_pct")
                    break
        # 6x6 chains: the ITEM kernel (uwbgo_general_items.cu); forests (pose edges to a key vertex) keep the CTA kernel
        kname = {1: "lm_fast_kernel", 2: "lm_chain_tma_kernel", 3: "lm_window_kernel"}.get(
            m["path"], "lm_general_cta_kernel" if getattr(topo, "is_forest", False) else "lm_general_items_kernel")
        fp64, _ = solver.measure_fp64_peak()
        # The fused LM kernel is bounded by FP64 arithmetic (issue latency of dependent chains), not by its
        # algorithmic bytes: frac = FP64 flop executed (ncu, Newton steps of sqrt and division included) / kernel
        # time / the FP64 FMA peak measured by this run's own micro-benchmark (MEASURED_PEAKS.json has no FP64 entry)
        fp64_ach = fp64_flop / (k_best * 1e-3) / 1e12 if (fp64_flop and k_best) else None
        roof = {"bound": "fp64", "kernel": kname + " (fused LM: linearise + assemble + block Cholesky + damping loop)",
                "achieved": fp64_ach, "peak": fp64 / 1e12, "unit": "TFLOP/s",
                "frac": fp64_ach / (fp64 / 1e12) if fp64_ach else None,
                "peak_source": "self-measured FP64 FMA micro-benchmark (uwbgo_measure_fp64_peak), this run",
                "fp64_flop_per_window": fp64_flop / W if fp64_flop else None, "fp64_pipe_active_pct_ncu": fp64_pipe,
                "kernel_ms": k_best,
                "traffic": traffic, "traffic_unit": "GB per launch (dram read+write, ncu)", "traffic_source": traffic_src,
                "hbm": {"algorithmic_bytes_per_window": wl_bytes, "achieved_gbs": achieved, "peak_gbs": hbm,
                        "frac": achieved / hbm if achieved else None, "peak_source": how},
                "note": NOTE_C3 if args.workload in ("c3", "c3x", "c5") else NOTE_GENERAL}
        if traffic and k_best:
            roof["hbm"]["traffic_rate_gbs"] = traffic / (k_best * 1e-3)
            roof["hbm"]["traffic_frac"] = roof["hbm"]["traffic_rate_gbs"] / hbm
        stages = None
        if (args.stages or world == 1) and not args.no_stages and args.workload in ("c3", "c3x") and W == WINDOWS_PER_GPU:
            stages = time_stages(solver, topo, batch.expanded(topo), cfg, dev, hbm)
        # the CPU leg runs at N = 1 only (the contract; at N > 1 this rank is bound to its GPU's CPUs)
        cpu = None if (args.no_cpu or world > 1) else cpu_baseline(topo, batch, cfg, name=args.workload.upper())
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl_desc, "windows_per_gpu": W, "windows_total": W * world, "n_poses": N,
                           "n_anchors": topo.n_anchors, "lm_iterations": wl_iters,
                           "parallelism": f"windows sharded x{world}, no data-path collective"
                                          + ("; one NCCL gather of final poses, chi2 and status words to rank 0 inside the timed step"
                                             if world > 1 else ""),
                           "host_affinity": numa,
                           "l2": f"inputs ({m['h2d'] / 1e6:.0f} MB) and workspace per step exceed the 126 MB L2"
                                 if m["h2d"] > 126e6 else f"inputs {m['h2d'] / 1e6:.0f} MB; the LM workspace is rewritten every trial"},
                "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                        "matches_device_leg": m["same"]},
                "gpu_launches": m["launches"], "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        if "shards_verified" in m:
            line["shards_verified"] = m["shards_verified"]
        if "resident" in m:
            line["e2e_resident"] = m["resident"]
        if other:
            line["strong" if args.scaling == "weak" else "weak"] = {
                "value": other["value"], "unit": UNIT, "ms_per_step": other["ms_per_step"],
                "windows_per_gpu": other["W"], "windows_total": other["W"] * world, "kernel_ms": other["kernel_ms"],
                "e2e": {"value": other["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": other["h2d"],
                        "d2h_bytes_per_step": other["d2h"], "matches_device_leg": other["same"]},
                "gpu_launches": other["launches"],
                **({"e2e_resident": other["resident"]} if "resident" in other else {}),
                "what": ("strong scaling: the workload's windows in TOTAL, split over the ranks (north_star quotes its "
                         "target on 65,536 windows at 8 GPUs)") if args.scaling == "weak"
                        else "weak scaling: the workload's windows per GPU"}
            if "shards_verified" in other:
                line["strong" if args.scaling == "weak" else "weak"]["shards_verified"] = other["shards_verified"]
        if stages:
            line["stages"] = stages
        print(json.dumps(line), flush=True)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def time_stages(solver, topo, batch, cfg, dev, hbm):
    """stage kernels: linearise (HBM roofline) and factor/solve"""
    import ctypes as C
    import torch
    from localization_b200 import _ffi
    W, N = batch.n_windows, topo.n_poses
    d = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("pose_t", "anchors", "range_d", "range_info")}
    Hd = torch.empty((W, N, 36), dtype=torch.float64, device=dev)
    Ho = torch.empty((W, N - 1, 36), dtype=torch.float64, device=dev)
    b = torch.empty((W, N, 6), dtype=torch.float64, device=dev)
    chi = torch.empty((W, 2), dtype=torch.float64, device=dev)
    cb = _ffi.CBatch()
    cb.n_windows = W
    pd = lambda t: C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_double))
    cb.pose_t, cb.anchors, cb.range_d, cb.range_info = pd(d["pose_t"]), pd(d["anchors"]), pd(d["range_d"]), pd(d["range_info"])
    st = torch.cuda.current_stream(dev).cuda_stream
    ms = []
    for _ in range(4):
        solver.linearize_device(topo, cb, cfg, Hd.data_ptr(), Ho.data_ptr(), b.data_ptr(), chi.data_ptr(), st)
        ms.append(solver.last_kernel_ms())
    torch.cuda.synchronize(dev)
    k = float(np.median(ms[1:]))
    ach = ALGO_BYTES_LINEARIZE * W / (k * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and W == WINDOWS_PER_GPU:
        with open(tp) as f:
            traffic = json.load(f).get("linearize_chain_fused_kernel", {}).get("dram_bytes_per_launch")
    return {"linearize": {"kernel_ms": k, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "bound": "hbm",
                          "traffic": traffic / 1e9 if traffic else None,
                          "traffic_unit": "GB per launch (dram read+write, ncu; algorithmic = %.2f GB)" % (ALGO_BYTES_LINEARIZE * W / 1e9),
                          "algorithmic_bytes_per_window": ALGO_BYTES_LINEARIZE,
                          "what": "linearize_chain_fused_kernel (uwbgo_linearize_batch_device): computeActiveErrors + "
                                  "buildSystem of every window, written as full 6x6 blocks in the public window-major "
                                  "layout; one kernel, its DRAM traffic is the algorithmic bytes"}}


if __name__ == "__main__":
    main()
