"""CPU tests of the host layer: the C-ABI library loads and exports every symbol the header
declares, fails loudly without a GPU, topology builders, sharding over a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from localization_b200 import Batch, Config, Topology, _ffi, synthetic
from localization_b200.shard import window_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(name="uwbgo.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef[^;]*\(\s*\*\s*\w+\s*\)[^;]*;", "", src)   # function-pointer typedefs are not symbols
    return sorted(set(re.findall(r"\b(uwbgo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/uwbgo.h but not exported"
    assert sorted(_ffi.SYMBOLS) == names, "ctypes table and header disagree"
    assert _ffi.load_library().uwbgo_abi_version() == _ffi.ABI_VERSION


def test_stream_entry_points_reject_null_arguments():
    """uwbgo_stream_* (resident fleet) check their arguments before touching the device"""
    lib = _ffi.load_library()
    h = ctypes.c_void_p()
    cfg = _ffi.CConfig()
    lib.uwbgo_config_default(ctypes.byref(cfg))
    anchors = (ctypes.c_double * 12)()
    assert lib.uwbgo_stream_create(None, 10, 4, 8, anchors, 5.0, ctypes.byref(cfg), ctypes.byref(h)) == -1   # UWBGO_E_INVALID
    assert not h
    assert lib.uwbgo_stream_step(None, 0, None, None, None, None, None, None) == -1
    assert lib.uwbgo_stream_load(None, None, None, None, None, None) == -1
    assert lib.uwbgo_stream_load_robots(None, None, None, None, None, None) == -1
    assert lib.uwbgo_stream_set_outlier_gate(None, 1.0) == -1
    assert lib.uwbgo_stream_step_robots(None, None, None, None, None, None, None, None) == -1
    assert lib.uwbgo_stream_read(None, None) == -1
    lib.uwbgo_stream_destroy(None)


def test_host_library_exports_every_declared_symbol():
    from localization_b200 import host
    lib = host.load_host_library()
    names = header_functions("uwbgo_host.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/uwbgo_host.h but not exported"
    assert sorted(host.HOST_SYMBOLS) == names, "ctypes table and header disagree"
    assert ctypes.sizeof(host.CLocParams) == 8 + 24 + 8 + 24 + 24 + 16


def test_structs_match_header_layout():
    assert ctypes.sizeof(_ffi.CTopology) == 16 + 6 * 8
    assert ctypes.sizeof(_ffi.CBatch) == 8 + 11 * 8 + 8 + 2 * 4
    assert ctypes.sizeof(_ffi.CRangeMsgs) == 4 * 8 + 8
    assert ctypes.sizeof(_ffi.CConfig) == 16 + 5 * 8
    assert ctypes.sizeof(_ffi.CResult) == 8 * 8
    c = _ffi.CConfig()
    _ffi.load_library().uwbgo_config_default(ctypes.byref(c))
    d = Config()
    assert (c.max_iterations, c.max_trials, c.orthogonalize_after) == (d.max_iterations, d.max_trials, d.orthogonalize_after)
    assert (c.tau, c.good_step_lower, c.good_step_upper, c.kernel_delta, c.jacobian_delta) == (
        d.tau, d.good_step_lower, d.good_step_upper, d.kernel_delta, d.jacobian_delta)


def test_no_cpu_fallback():
    """without a GPU the product refuses to run instead of computing on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from localization_b200 import Solver, UwbgoError
    with pytest.raises(UwbgoError) as ei:
        Solver(0)
    assert ei.value.code == _ffi.E_NODEVICE


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "localization_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"import\s+oracle|from\s+oracle|libuwbgo_oracle|oracle[/\\.]", text), \
                    f"{f} reaches into oracle/"


def test_topology_builders():
    t = Topology.uwb_chain(10, 4)
    assert t.n_edges == 19 and t.counts() == (19, 0, 0)
    # insertion order of Localization::addRangeEdge: anchor edge of the new vertex, then trajectory edge
    assert list(t.edge_kind[:5]) == [0, 0, 1, 0, 1]
    assert list(t.edge_a[:5]) == [0, 1, 0, 2, 1] and list(t.edge_b[:5]) == [0, 1, 1, 2, 2]
    t = Topology.uwb_chain(12, 4, imu=True)
    assert t.counts() == (23, 11, 0)            # 12 + 11 range-type edges, 11 IMU priors
    t = Topology.uwb_twist(15, 8, antennas=3)
    assert t.counts() == (15, 0, 14)
    topo, batch, _ = synthetic.uwb_only(5, 6, 4)
    batch.check(topo)
    with pytest.raises(ValueError):
        Batch(pose_t=np.zeros((5, 6, 3))).check(topo)
    # newest pose is a copy of its predecessor (robot.cpp:90)
    assert np.array_equal(batch.pose_t[:, -1], batch.pose_t[:, -2])
    # trajectory edges carry measurement 0 (localization.cpp:338)
    assert not batch.range_d[:, 2::2].any()


def test_window_range_partitions():
    for W in (0, 1, 7, 65536, 100003):
        for world in (1, 2, 3, 8):
            parts = [window_range(W, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == W
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1


WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["UWBGO_ROOT"])
import numpy as np, torch, torch.distributed as dist
from localization_b200 import Config, synthetic
from localization_b200.shard import shard_batch, gather_to_root
from oracle import oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
topo, batch, _ = synthetic.uwb_only(37, 8, 4, seed=5)       # ragged: 18 + 19 windows
cfg = Config(max_iterations=4)
mine = shard_batch(batch, rank, world)
res = oracle.solve(topo, mine, cfg, n_threads=1)            # CPU stand-in for the per-rank solve
poses = gather_to_root(torch.from_numpy(res.pose_t))
chi2 = gather_to_root(torch.from_numpy(res.chi2))
if rank == 0:
    whole = oracle.solve(topo, batch, cfg, n_threads=1)
    assert np.array_equal(poses.numpy(), whole.pose_t)
    assert np.array_equal(chi2.numpy(), whole.chi2)
    print("SHARD_OK")
else:
    assert poses is None
dist.destroy_process_group()
'''


def test_two_rank_sharding_gloo(tmp_path):
    """N > 1 path on CPU: shard, solve per rank, one gather to rank 0; identical to one rank"""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, UWBGO_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


def test_g2o_text_round_trip(tmp_path):
    """EDGE_RANGE / EDGE_RANGE_OFFSET / VERTEX_SE3:QUAT / EDGE_SE3:QUAT / EDGE_SE3_PRIOR interchange:
    a window written as g2o text and read back solves to the same result (UWB-only: same bits)"""
    from localization_b200.tools import g2o_text
    from oracle import oracle
    topo, batch, _ = synthetic.uwb_only(3, 8, 4, seed=31)
    f = str(tmp_path / "w.g2o")
    g2o_text.write_window(f, topo, batch, w=1)
    text = open(f).read()
    assert "EDGE_RANGE 200 100 " in text and "VERTEX_SE3:QUAT 2300 " in text and "FIX 103" in text
    t2, b2 = g2o_text.read_window(f)
    assert np.array_equal(t2.edge_kind, topo.edge_kind) and np.array_equal(t2.edge_a, topo.edge_a)
    assert np.array_equal(t2.edge_b, topo.edge_b) and np.array_equal(t2.edge_robust, topo.edge_robust)
    cfg = Config(max_iterations=5)
    a, b = oracle.solve(topo, batch.slice(1, 2), cfg), oracle.solve(t2, b2, cfg)
    assert np.array_equal(a.pose_t, b.pose_t) and np.array_equal(a.chi2, b.chi2)   # repr() floats round-trip
    for make, kw in ((synthetic.uwb_imu_lidar, dict(W=2, N=6, A=4, seed=32)), (synthetic.uwb_twist, dict(W=2, N=6, A=4, seed=33)),
                     (synthetic.uwb_pose, dict(W=2, N=9, A=4, seed=34))):
        topo, batch, _ = make(**kw)
        g2o_text.write_window(f, topo, batch, w=0)
        assert "EDGE_RANGE_OFFSET" in open(f).read()
        t2, b2 = g2o_text.read_window(f)
        assert t2.counts() == topo.counts() and np.array_equal(t2.edge_ant, topo.edge_ant)
        a, b = oracle.solve(topo, batch.slice(0, 1), cfg), oracle.solve(t2, b2, cfg)
        # rotations go through quaternions: a 1e-16 perturbation of the inputs, amplified by the numeric
        # Jacobians and the unconverged LM to ~1e-8 m -- still inside the 1e-6 m bar
        assert np.abs(a.pose_t - b.pose_t).max() < 1e-6
        assert np.allclose(a.chi2[:, :2], b.chi2[:, :2], rtol=1e-4)


def test_cpp_type_api():
    """EdgeSE3Range / EdgeSE3RangeOffset / Robot drop-ins (localization_b200/host/test_types.cpp)"""
    host_dir = os.path.join(ROOT, "localization_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host_dir, "test_types"])
    out = subprocess.run([os.path.join(host_dir, "test_types")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "host type API ok" in out.stdout, out.stdout + out.stderr


def test_pin_set_export(tmp_path):
    """tests/export_pin_set.py: windows as g2o text + the oracle's answers, for checking against a real g2o"""
    import json
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(__file__), "export_pin_set.py")
    subprocess.check_call([sys.executable, script, "--out", str(tmp_path), "--windows", "1"])
    exp = json.load(open(tmp_path / "expected.json"))
    assert len(exp) == 5
    from localization_b200.tools import g2o_text
    for f, e in exp.items():
        topo, batch = g2o_text.read_window(str(tmp_path / f))[:2]
        assert len(e["pose_t"]) == topo.n_poses and e["iterations"] >= 1
