"""GPU parity: the CUDA path (through the C ABI of libuwbgo.so) against the CPU oracle on the same
seeded inputs.  The bar (BASELINE.json north_star): poses within 1e-6 m, final chi2 within 1e-9
relative.  Because both sides execute the same IEEE operation sequence the tests ask for more:
bit-identical poses, chi2, lambda and LM status words."""
import numpy as np
import pytest

from localization_b200 import Batch, Config, Topology, synthetic
from localization_b200._ffi import (EDGE_PRIOR, EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE, EDGE_SE3,
                                    FLAG_CHOL_FAIL, FLAG_TERMINATED)
from oracle import oracle

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-6       # BASELINE.json: within 1e-6 m on poses
CHI2_RTOL = 1e-9        # BASELINE.json: within 1e-9 relative on final chi2


def assert_parity(got, ref, exact=True):
    dp = np.abs(got.pose_t - ref.pose_t).max() if got.pose_t.size else 0.0
    assert dp <= POSE_TOL_M, f"pose error {dp}"
    den = np.maximum(np.abs(ref.chi2[:, :2]), 1e-300)
    dc = (np.abs(got.chi2[:, :2] - ref.chi2[:, :2]) / den).max() if got.chi2.size else 0.0
    assert dc <= CHI2_RTOL, f"relative chi2 error {dc}"
    if exact:
        assert np.array_equal(got.status, ref.status)
        assert np.array_equal(got.pose_t, ref.pose_t), f"poses not bit-identical (max diff {dp})"
        assert np.array_equal(got.pose_R, ref.pose_R)
        assert np.array_equal(got.chi2, ref.chi2), f"chi2 not bit-identical (max rel {dc})"
        assert np.array_equal(got.oplus_count, ref.oplus_count)


@pytest.mark.parametrize("W,N,A", [(256, 50, 8), (33, 10, 4), (1, 12, 4), (96, 200, 16)])
def test_uwb_only_fast_path(solver, W, N, A):
    topo, batch, _ = synthetic.uwb_only(W, N, A, seed=7 + W)
    cfg = Config(max_iterations=10)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(1, 2)
    ref = oracle.solve(topo, batch, cfg)
    assert_parity(got, ref)
    assert (ref.status[:, 0] == 10).all()


def test_general_path_equals_fast_path(solver):
    """identity rotations passed explicitly route through the 6x6 kernel: same bits"""
    topo, batch, _ = synthetic.uwb_only(64, 20, 8, seed=3)
    cfg = Config(max_iterations=10)
    fast = solver.solve(topo, batch, cfg)
    assert solver.path_ok(1, 2)
    batch.pose_R = np.tile(np.eye(3), (64, 20, 1, 1))
    gen = solver.solve(topo, batch, cfg)
    assert solver.path_ok(0)
    assert_parity(gen, fast)
    assert_parity(gen, oracle.solve(topo, batch, cfg))


def test_uwb_imu_lidar(solver):
    topo, batch, _ = synthetic.uwb_imu_lidar(128, 20, 8)
    cfg = Config(max_iterations=20)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(0)
    assert_parity(got, oracle.solve(topo, batch, cfg))


def test_prior_information_diagonal_and_dense_tiles(solver):
    """the 6x6 ITEM kernel reads only the diagonal of the prior information matrices of a tile when every one of
    them has exactly +0.0 elsewhere (the matrices Localization builds, localization.cpp:478-479,515-518), literal
    zeros taking the place of the rest: tiles that qualify, tiles that hold one dense matrix, one -0.0 or one NaN
    off the diagonal all give the bits of the dense arithmetic"""
    topo, batch, _ = synthetic.uwb_imu_lidar(160, 12, 6, seed=21)
    cfg = Config(max_iterations=6)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))      # five diagonal tiles
    batch.prior_info[40, 3, 1, 4] = 0.25                                               # tile 1: one dense matrix
    batch.prior_info[40, 3, 4, 1] = 0.25
    batch.prior_info[70, 5, 0, 2] = -0.0                                               # tile 2: a negative zero
    batch.prior_info[130, 1, 5, 0] = np.nan                                            # tile 4: a NaN
    got, ref = solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg)
    ok = np.setdiff1d(np.arange(160), [130])
    assert np.array_equal(got.status, ref.status)
    assert np.array_equal(got.pose_t[ok], ref.pose_t[ok]) and np.array_equal(got.pose_R[ok], ref.pose_R[ok])
    assert np.array_equal(got.chi2[ok], ref.chi2[ok])
    assert np.array_equal(got.pose_t[130], ref.pose_t[130], equal_nan=True)


@pytest.mark.parametrize("jdelta", [1e-9, 1e-3, 3e-8])
def test_rotation_increments_sparse_and_dense(solver, jdelta):
    """the numeric Jacobians of range edges with lever arms perturb the rotation by fromVectorMQT(+-delta e_k); for the
    default delta = 1e-9 those six matrices are I plus one antisymmetric pair and the ITEM kernel forms the perturbed
    rotation from the two products that are not structurally zero; a step whose increments are dense (1e-3: 1 - 2
    delta^2 != 1) takes the dense product.  Identity rotations, rotations with negative zeros and a NaN rotation
    entry all give the oracle's bits"""
    topo, batch, _ = synthetic.uwb_imu_lidar(96, 10, 6, seed=33)
    cfg = Config(max_iterations=5, jacobian_delta=jdelta)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
    batch.pose_R[:32] = np.eye(3)                                   # tile 0: exact zeros and ones
    batch.pose_R[32:64] = np.diag([1.0, -1.0, -1.0])                # tile 1: a half turn about x ...
    batch.pose_R[32:64][batch.pose_R[32:64] == 0.0] = -0.0          # ... with negative zeros off the diagonal
    batch.pose_R[70, 3, 1, 2] = np.nan                              # tile 2: one window with a NaN rotation entry
    got, ref = solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg)
    ok = np.setdiff1d(np.arange(96), [70])
    assert np.array_equal(got.status[ok], ref.status[ok])
    assert np.array_equal(got.pose_t[ok], ref.pose_t[ok]) and np.array_equal(got.pose_R[ok], ref.pose_R[ok])
    assert np.array_equal(got.chi2[ok], ref.chi2[ok])
    assert np.array_equal(got.pose_t[70], ref.pose_t[70], equal_nan=True)


@pytest.mark.parametrize("make,N,iters", [(synthetic.uwb_imu_lidar, 120, 4), (synthetic.uwb_twist, 200, 3)])
def test_long_general_chains(solver, make, N, iters):
    """6x6 chains far longer than the BASELINE shapes (topology tables of the ITEM kernel beyond 48 KB of shared
    memory, hundreds of staged elimination steps), ragged batch"""
    topo, batch, _ = make(45, N, 8, seed=N)
    cfg = Config(max_iterations=iters)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


def test_uwb_imu_c2_shape(solver):
    topo, batch, _ = synthetic.uwb_imu_lidar(64, 12, 4, v_max=3.0, antennas=0, lidar=False, seed=11)
    cfg = Config(max_iterations=10)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


def test_block_diagonal_windows(solver):
    """WINDOW path, block-diagonal elimination (uwbgo_window.cu factor_main_bd): range edges without lever arms plus
    priors whose information has no translation / rotation cross terms -- IMU rotation priors (localization.cpp:515-518)
    and lidar z priors (:478-479) -- solve on the 3x3 diagonal blocks.  A window whose prior information DOES couple
    the two (one symmetric cross term), or holds a -0.0 there, or a NaN measurement, shares the batch: the kernel decides per
    window, every window gives the oracle's bits."""
    cfg = Config(max_iterations=10)
    for lidar, N in ((False, 12), (True, 9)):
        topo, batch, _ = synthetic.uwb_imu_lidar(48, N, 4, v_max=3.0, antennas=0, lidar=lidar, seed=21 + N)
        info = batch.prior_info.reshape(48, -1, 6, 6)
        info[3, 2, 1, 4] = info[3, 2, 4, 1] = 0.37       # couples y translation and pitch: the 6x6 blocks
        info[5, :, 0, 3] = -0.0                          # a negative zero is a zero
        info[7, 1, 2, 5] = info[7, 1, 5, 2] = 1e-300     # tiny, but not zero
        assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
        batch.prior_Z.reshape(48, -1, 12)[9, 0, 10] = np.nan   # a NaN measurement reaches b, not H: the pivots stay positive
        info[11, 0, 3, 3] = np.nan                              # a NaN on the diagonal: no pivot
        info[13, 1, 0, 4] = np.nan                              # a NaN in a cross block: not block-diagonal
        got, ref = solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg)
        ok = np.setdiff1d(np.arange(48), [9, 11, 13])
        assert np.array_equal(got.status, ref.status)
        assert np.array_equal(got.pose_t[ok], ref.pose_t[ok]) and np.array_equal(got.pose_R[ok], ref.pose_R[ok])
        assert np.array_equal(got.chi2[ok], ref.chi2[ok])
        for w in (9, 11, 13):
            assert np.array_equal(got.pose_t[w], ref.pose_t[w], equal_nan=True)
            assert np.array_equal(got.pose_R[w], ref.pose_R[w], equal_nan=True)
            assert np.array_equal(got.chi2[w], ref.chi2[w], equal_nan=True)


def test_uwb_twist(solver):
    topo, batch, _ = synthetic.uwb_twist(128, 15, 8)
    cfg = Config(max_iterations=12)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


@pytest.mark.parametrize("N,K", [(24, 4), (9, 8), (40, 7), (5, 1)])
def test_uwb_pose_key_vertex_stars(solver, N, K):
    """pose edges to a key vertex (localization.cpp:258-267): forest windows, no fill-in"""
    topo, batch, _ = synthetic.uwb_pose(96, N, 8, keyframe_len=K, seed=N)
    cfg = Config(max_iterations=10)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(0)
    assert_parity(got, oracle.solve(topo, batch, cfg))
    Hd, Ho, b, chi = solver.linearize(topo, batch, cfg)
    rHd, rHo, rb, rchi = oracle.linearize(topo, batch, cfg)
    assert np.array_equal(Hd, rHd) and np.array_equal(Ho, rHo) and np.array_equal(b, rb)


def test_two_older_neighbours_are_refused(solver):
    from localization_b200 import UwbgoError
    topo = Topology.from_edges(4, 1, 0, [(EDGE_RANGE_ANCHOR, 0, 0, 0, 1), (EDGE_SE3, 0, 1, 0, 1), (EDGE_SE3, 1, 2, 0, 1),
                                         (EDGE_SE3, 0, 3, 0, 1), (EDGE_SE3, 2, 3, 0, 1)])
    b = Batch(pose_t=np.zeros((1, 4, 3)), anchors=np.ones((1, 1, 3)), range_d=np.ones((1, 1)), range_info=np.ones((1, 1)),
              se3_Z=np.tile(np.r_[np.eye(3).ravel(), 0, 0, 0], (1, 4, 1)), se3_info=np.tile(np.eye(6), (1, 4, 1, 1)))
    with pytest.raises(UwbgoError) as ei:
        solver.solve(topo, b, Config(max_iterations=1))
    assert ei.value.code == -2


def test_oplus_counter_carry_and_reorthogonalisation(solver):
    """VertexSE3::_numOplusCalls carried in close to orthogonalizeAfter: the re-orthogonalisation
    trips inside numeric Jacobians and inside trial updates"""
    topo, batch, _ = synthetic.uwb_imu_lidar(64, 12, 8, seed=5)
    rng = np.random.default_rng(1)
    batch.oplus_count = rng.integers(900, 1001, size=(64, 12)).astype(np.int32)
    cfg = Config(max_iterations=8)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
    cfg = Config(max_iterations=4, orthogonalize_after=7)
    batch.oplus_count = rng.integers(0, 8, size=(64, 12)).astype(np.int32)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
    # fast path only tracks the counter
    topo, batch, _ = synthetic.uwb_only(40, 10, 4, seed=2)
    batch.oplus_count = rng.integers(0, 1001, size=(40, 10)).astype(np.int32)
    cfg = Config(max_iterations=10)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(1, 2)
    assert_parity(got, oracle.solve(topo, batch, cfg))


def test_linearize_bit_exact(solver):
    for make, kw in ((synthetic.uwb_only, dict(W=96, N=50, A=8)),
                     (synthetic.uwb_imu_lidar, dict(W=64)), (synthetic.uwb_twist, dict(W=64))):
        topo, batch, _ = make(**kw)
        cfg = Config()
        Hd, Ho, b, chi = solver.linearize(topo, batch, cfg)
        rHd, rHo, rb, rchi = oracle.linearize(topo, batch, cfg)
        assert np.array_equal(Hd, rHd)
        assert np.array_equal(Ho, rHo)
        assert np.array_equal(b, rb)
        assert np.array_equal(chi, rchi)


def test_linearize_chain_fused_ragged_and_unaligned(solver):
    """the one-kernel linearise stage of CHAIN windows: ragged tiles, 1- and 2-pose windows, a window
    longer than a pose run, and output arrays that are not 16-byte aligned (two-kernel fallback)"""
    import ctypes as C
    import torch
    from localization_b200 import _ffi
    for W, N, A in ((1, 1, 3), (33, 2, 4), (95, 7, 8), (40, 50, 8), (31, 200, 16)):
        topo, batch, _ = synthetic.uwb_only(W, N, A, seed=W + N)
        cfg = Config()
        Hd, Ho, b, chi = solver.linearize(topo, batch, cfg)
        rHd, rHo, rb, rchi = oracle.linearize(topo, batch, cfg)
        assert np.array_equal(Hd, rHd) and np.array_equal(Ho, rHo) and np.array_equal(b, rb)
        assert np.array_equal(chi, rchi)
    # device API with output pointers 8 bytes off a 16-byte boundary
    W, N, A = 70, 12, 5
    topo, batch, _ = synthetic.uwb_only(W, N, A, seed=9)
    cfg = Config()
    dev = torch.device("cuda", 0)
    keep = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("pose_t", "anchors", "range_d", "range_info")}
    cb = _ffi.CBatch()
    cb.n_windows = W
    for k, v in keep.items():
        setattr(cb, k, C.cast(C.c_void_p(v.data_ptr()), C.POINTER(C.c_double)))
    outs = []
    for off in (0, 1):
        bufs = [torch.zeros(n + 1, dtype=torch.float64, device=dev) for n in (W * N * 36, W * (N - 1) * 36, W * N * 6, W * 2)]
        views = [t[off:off + t.numel() - 1] for t in bufs]
        assert all((v.data_ptr() % 16 == 0) == (off == 0) for v in views[:3])
        solver.linearize_device(topo, cb, cfg, *[v.data_ptr() for v in views], torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        outs.append([v.cpu().numpy() for v in views])
    rHd, rHo, rb, rchi = oracle.linearize(topo, batch, cfg)
    for o in outs:
        assert np.array_equal(o[0].reshape(rHd.shape), rHd)
        assert np.array_equal(o[1].reshape(rHo.shape), rHo)
        assert np.array_equal(o[2].reshape(rb.shape), rb)
        assert np.array_equal(o[3].reshape(rchi.shape), rchi)


def test_factor_solve_bit_exact(solver):
    topo, batch, _ = synthetic.uwb_imu_lidar(64, 20, 8)
    Hd, Ho, b, _ = oracle.linearize(topo, batch, Config())
    lam = np.full(64, 0.5)
    lam[::7] = 1e-3
    x, ok = solver.factor_solve(Hd, Ho, b, lam)
    rx, rok = oracle.factor_solve(Hd, Ho, b, lam)
    assert np.array_equal(ok, rok) and ok.all()
    assert np.array_equal(x, rx)
    # against a dense solve (not bit-exact: different elimination order)
    W, N = 64, 20
    for w in (0, 63):
        H = np.zeros((6 * N, 6 * N))
        for i in range(N):
            H[6 * i:6 * i + 6, 6 * i:6 * i + 6] = Hd[w, i]
            if i + 1 < N:
                H[6 * i:6 * i + 6, 6 * i + 6:6 * i + 12] = Ho[w, i]
                H[6 * i + 6:6 * i + 12, 6 * i:6 * i + 6] = Ho[w, i].T
        xd = np.linalg.solve(H + lam[w] * np.eye(6 * N), b[w].reshape(-1))
        assert np.allclose(x[w].reshape(-1), xd, rtol=1e-7, atol=1e-10)
    # a non-positive pivot is reported, x = 0
    Hd2 = Hd.copy()
    Hd2[3, 5] = -np.eye(6)
    x, ok = solver.factor_solve(Hd2, Ho, b, lam)
    rx, rok = oracle.factor_solve(Hd2, Ho, b, lam)
    assert ok[3] == 0 and np.array_equal(ok, rok)
    assert not x[3].any() and np.array_equal(x, rx)


def test_edge_cases(solver):
    cfg = Config(max_iterations=5)
    # single pose, single anchor edge
    topo = Topology.uwb_chain(1, 3)
    batch = Batch(pose_t=np.array([[[0.1, 0.2, 1.0]]]), anchors=np.array([[[3., 3, 0], [-3, 3, 1], [0, -3, 2]]]),
                  range_d=np.array([[4.0]]), range_info=np.array([[300.0]]))
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
    # two poses
    topo, batch, _ = synthetic.uwb_only(5, 2, 4, seed=9)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))
    # zero iterations: estimates untouched, chi2 of the initial estimate
    topo, batch, _ = synthetic.uwb_only(5, 6, 4, seed=9)
    got = solver.solve(topo, batch, Config(max_iterations=0))
    assert np.array_equal(got.pose_t, batch.pose_t)
    assert_parity(got, oracle.solve(topo, batch, Config(max_iterations=0)))
    # no edges at all: H = 0, lambda = 0, every trial fails, LM terminates after max_trials
    topo = Topology.from_edges(3, 0, 0, [])
    batch = Batch(pose_t=np.zeros((2, 3, 3)))
    got = solver.solve(topo, batch, cfg)
    ref = oracle.solve(topo, batch, cfg)
    assert_parity(got, ref)
    assert (got.status[:, 2] & FLAG_CHOL_FAIL).all() and (got.status[:, 2] & FLAG_TERMINATED).all()
    # empty batch
    topo, batch, _ = synthetic.uwb_only(4, 6, 4)
    got = solver.solve(topo, batch.slice(0, 0), cfg)
    assert got.pose_t.shape[0] == 0
    # exact measurements from the start: chi2 = 0
    topo = Topology.uwb_chain(4, 4)
    anchors = np.array([[[3., 3, 0], [-3, 3, 1], [0, -3, 2], [2, -2, 1.5]]])
    pts = np.array([[[0., 0, 1], [0.1, 0, 1], [0.2, 0, 1], [0.3, 0, 1]]])
    d = np.zeros((1, 7)); info = np.full((1, 7), 100.0)
    slot = 0
    for k in range(4):
        d[0, slot] = np.linalg.norm(pts[0, k] - anchors[0, k]); slot += 1
        if k > 0:
            d[0, slot] = np.linalg.norm(pts[0, k] - pts[0, k - 1]); slot += 1
    batch = Batch(pose_t=pts, anchors=anchors, range_d=d, range_info=info)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


def test_multiple_range_edges_per_pose_and_unordered_insertion(solver):
    """merged-covariance branch: extra anchor edges on a pose without a new vertex; two
    pose-pose range edges on one pair; edges inserted out of pose order"""
    rng = np.random.default_rng(4)
    N, A, W = 6, 4, 48
    edges = []
    for k in (3, 0, 5, 1, 4, 2):
        edges.append((EDGE_RANGE_ANCHOR, k, k % A, 0, 1))
        edges.append((EDGE_RANGE_ANCHOR, k, (k + 1) % A, 0, 0))
    for k in (4, 0, 2, 1, 3):
        edges.append((EDGE_RANGE_POSE, k, k + 1, 0, 1))
    edges.append((EDGE_RANGE_POSE, 2, 3, 0, 0))
    topo = Topology.from_edges(N, A, 0, edges)
    er = topo.counts()[0]
    batch = Batch(pose_t=rng.normal(0, 1, (W, N, 3)), anchors=rng.normal(0, 3, (W, A, 3)),
                  range_d=np.abs(rng.normal(3, 1, (W, er))), range_info=rng.uniform(10, 400, (W, er)))
    cfg = Config(max_iterations=6)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(1, 2)
    assert_parity(got, oracle.solve(topo, batch, cfg))
    # a third edge on the same pair exceeds the fast path's carry slots -> general path, same bits
    topo3 = Topology.from_edges(N, A, 0, edges + [(EDGE_RANGE_POSE, 2, 3, 0, 1)])
    b3 = Batch(pose_t=batch.pose_t, anchors=batch.anchors,
               range_d=np.concatenate([batch.range_d, batch.range_d[:, :1]], 1),
               range_info=np.concatenate([batch.range_info, batch.range_info[:, :1]], 1))
    got = solver.solve(topo3, b3, cfg)
    assert solver.path_ok(0)
    assert_parity(got, oracle.solve(topo3, b3, cfg))


def test_range_edges_between_non_adjacent_poses(solver):
    """pose-pose range edges to an older, non-adjacent pose (forest): general path"""
    rng = np.random.default_rng(12)
    N, A, W = 7, 3, 40
    edges = [(EDGE_RANGE_ANCHOR, k, k % A, 0, 1) for k in range(N)]
    edges += [(EDGE_RANGE_POSE, 0, 1, 0, 1), (EDGE_RANGE_POSE, 0, 2, 0, 1), (EDGE_RANGE_POSE, 2, 3, 0, 0),
              (EDGE_RANGE_POSE, 2, 4, 0, 1), (EDGE_RANGE_POSE, 2, 5, 0, 1), (EDGE_RANGE_POSE, 5, 6, 0, 1)]
    topo = Topology.from_edges(N, A, 0, edges)
    er = topo.counts()[0]
    batch = Batch(pose_t=rng.normal(0, 1, (W, N, 3)), anchors=rng.normal(0, 3, (W, A, 3)),
                  range_d=np.abs(rng.normal(2, 1, (W, er))), range_info=rng.uniform(10, 400, (W, er)))
    cfg = Config(max_iterations=6)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(0)
    assert_parity(got, oracle.solve(topo, batch, cfg))


def test_mixed_edge_kinds_general(solver):
    """every edge kind in one window, priors with dense information, robust and plain"""
    rng = np.random.default_rng(8)
    N, A, W, K = 5, 3, 40, 2
    edges = [(EDGE_RANGE_ANCHOR, 0, 0, 1, 1), (EDGE_PRIOR, 0, 0, 0, 1)]
    for k in range(1, N):
        edges += [(EDGE_SE3, k - 1, k, 0, k % 2), (EDGE_RANGE_ANCHOR, k, k % A, 1 + k % K, 1),
                  (EDGE_RANGE_POSE, k - 1, k, k % (K + 1), 0), (EDGE_PRIOR, k, 0, 0, 0)]
    topo = Topology.from_edges(N, A, K, edges)
    er, ep, es = topo.counts()

    def rand_R(shape):
        q = rng.normal(size=shape + (4,))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        R = np.empty(shape + (3, 3))
        R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
        R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
        R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
        return R

    def spd(shape):
        M = rng.normal(size=shape + (6, 6))
        return M @ np.swapaxes(M, -1, -2) + 0.5 * np.eye(6)
    pR = rand_R((W, N))
    pZ = np.concatenate([rand_R((W, ep)).reshape(W, ep, 9), rng.normal(0, 1, (W, ep, 3))], -1)
    sZ = np.concatenate([rand_R((W, es)).reshape(W, es, 9), rng.normal(0, .3, (W, es, 3))], -1)
    batch = Batch(pose_t=rng.normal(0, 1, (W, N, 3)), pose_R=pR, anchors=rng.normal(0, 3, (W, A, 3)),
                  range_d=np.abs(rng.normal(3, 1, (W, er))), range_info=rng.uniform(10, 400, (W, er)),
                  ant_offsets=rng.normal(0, 0.2, (K, 3)), prior_Z=pZ, prior_info=spd((W, ep)),
                  se3_Z=sZ, se3_info=spd((W, es)))
    cfg = Config(max_iterations=6)
    got = solver.solve(topo, batch, cfg)
    assert_parity(got, oracle.solve(topo, batch, cfg))
    Hd, Ho, b, chi = solver.linearize(topo, batch, cfg)
    rHd, rHo, rb, rchi = oracle.linearize(topo, batch, cfg)
    assert np.array_equal(Hd, rHd) and np.array_equal(Ho, rHo) and np.array_equal(b, rb)
    assert np.array_equal(chi, rchi)


def test_chunked_pipeline_is_shard_invariant(solver):
    """results do not depend on how windows are chunked over streams (nor, therefore, on how
    they are sharded over GPUs)"""
    topo, batch, _ = synthetic.uwb_only(1000, 12, 4, seed=21)
    cfg = Config(max_iterations=6)
    whole = solver.solve(topo, batch, cfg)
    solver.set_pipeline(96, 3)
    try:
        parts = solver.solve(topo, batch, cfg)
    finally:
        solver.set_pipeline(16384, 4)
    assert_parity(parts, whole)
    halves = [solver.solve(topo, batch.slice(0, 500), cfg), solver.solve(topo, batch.slice(500, 1000), cfg)]
    assert np.array_equal(np.concatenate([h.pose_t for h in halves]), whole.pose_t)
    assert np.array_equal(np.concatenate([h.chi2 for h in halves]), whole.chi2)


def test_properties_at_full_size(solver):
    """BASELINE headline size (65,536 x N=50): size-independent properties instead of the oracle
    on everything: accepted LM steps never increase the robust chi2, R stays identity, and a
    sample of windows matches the oracle bit for bit"""
    W = 65536
    topo, batch, truth = synthetic.uwb_only(W, 50, 8)
    cfg = Config(max_iterations=10)
    got = solver.solve(topo, batch, cfg)
    _, _, _, chi0 = solver.linearize(topo, batch.slice(0, 4096), cfg)
    assert (got.chi2[:4096, 1] <= chi0[:, 1]).all()
    assert np.array_equal(got.pose_R, np.broadcast_to(np.eye(3), got.pose_R.shape))
    assert (got.status[:, 0] == 10).all() and (got.status[:, 2] == 0).all()
    assert np.abs(got.pose_t - truth).mean() < np.abs(batch.pose_t - truth).mean()
    idx = np.arange(0, W, 257)
    sub = Batch(pose_t=batch.pose_t[idx], anchors=batch.anchors[idx], range_d=batch.range_d[idx],
                range_info=batch.range_info[idx])
    ref = oracle.solve(topo, sub, cfg)
    assert np.array_equal(got.pose_t[idx], ref.pose_t)
    assert np.array_equal(got.chi2[idx], ref.chi2)
    assert np.array_equal(got.status[idx], ref.status)


def test_c5_properties_at_full_size(_gpu_solver):
    """BASELINE C5 on one GPU (131,072 windows x N=200, A=16: the largest configuration): the same
    size-independent properties, and every 257th window against the oracle bit for bit"""
    solver = _gpu_solver
    W = 131072
    topo, batch, truth = synthetic.uwb_only(W, 200, 16, seed=synthetic.SEED_C3 + 5)
    cfg = Config(max_iterations=10)
    got = solver.solve(topo, batch, cfg)
    assert solver.last_path in (1, 2)
    _, _, _, chi0 = solver.linearize(topo, batch.slice(0, 2048), cfg)
    assert (got.chi2[:2048, 1] <= chi0[:, 1]).all()
    assert (got.status[:, 0] == 10).all() and (got.status[:, 2] == 0).all()
    assert np.abs(got.pose_t - truth).mean() < np.abs(batch.pose_t - truth).mean()
    idx = np.arange(0, W, 257)
    sub = Batch(pose_t=batch.pose_t[idx], anchors=batch.anchors[idx], range_d=batch.range_d[idx],
                range_info=batch.range_info[idx])
    ref = oracle.solve(topo, sub, cfg)
    assert np.array_equal(got.pose_t[idx], ref.pose_t)
    assert np.array_equal(got.chi2[idx], ref.chi2)
    assert np.array_equal(got.status[idx], ref.status)


@pytest.mark.parametrize("make,iters", [(synthetic.uwb_imu_lidar, 20), (synthetic.uwb_twist, 12)])
def test_general_properties_at_full_size(solver, make, iters):
    """BASELINE C4 size (two halves of 8,192 windows, 6x6 blocks, CTA-per-tile kernel, one chunk):
    accepted LM steps never increase the robust chi2, rotations stay orthonormal, every window runs
    its iterations, and a sample of windows matches the oracle bit for bit"""
    W = 8192
    topo, batch, _ = make(W)
    cfg = Config(max_iterations=iters)
    got = solver.solve(topo, batch, cfg)
    assert solver.path_ok(0)
    _, _, _, chi0 = solver.linearize(topo, batch.slice(0, 1024), cfg)
    assert (got.chi2[:1024, 1] <= chi0[:, 1]).all()
    R = got.pose_R.reshape(W, -1, 3, 3)
    assert np.abs(np.einsum("wnij,wnkj->wnik", R, R) - np.eye(3)).max() < 1e-9
    assert (got.status[:, 0] >= 1).all() and (got.status[:, 0] <= iters).all()
    assert ((got.status[:, 2] & ~2) == 0).all()   # at most UWBGO_FLAG_TERMINATED
    idx = np.arange(0, W, 61)
    sub = Batch(**{k: (getattr(batch, k)[idx] if k != "ant_offsets" else getattr(batch, k))
                   for k in ("pose_t", "pose_R", "anchors", "ant_offsets", "range_d", "range_info", "prior_Z",
                             "prior_info", "se3_Z", "se3_info") if getattr(batch, k) is not None})
    assert_parity(solver.solve(topo, sub, cfg), oracle.solve(topo, sub, cfg))
    ref = oracle.solve(topo, sub, cfg)
    assert np.array_equal(got.pose_t[idx], ref.pose_t) and np.array_equal(got.pose_R[idx], ref.pose_R)
    assert np.array_equal(got.chi2[idx], ref.chi2) and np.array_equal(got.status[idx], ref.status)


def test_non_finite_inputs_are_contained(solver):
    """a NaN / inf measurement poisons only its own window: flagged, no hang, neighbours bit-exact"""
    from localization_b200._ffi import FLAG_NONFINITE
    topo, batch, _ = synthetic.uwb_only(70, 10, 4, seed=41)
    cfg = Config(max_iterations=6)
    clean = solver.solve(topo, batch, cfg)
    batch.range_d[3, 4] = np.nan
    batch.range_d[40, 0] = np.inf
    batch.pose_t[65, 2, 1] = np.nan
    got = solver.solve(topo, batch, cfg)
    ref = oracle.solve(topo, batch, cfg)
    bad = np.array([3, 40, 65])
    good = np.setdiff1d(np.arange(70), bad)
    assert np.array_equal(got.pose_t[good], clean.pose_t[good]) and np.array_equal(got.chi2[good], clean.chi2[good])
    # NaN in H fails the first pivot test (a NaN is not > 0): every trial is rejected like a non-SPD system
    # (and rho = NaN ends the trial loop after one trial, as g2o's `while (rho < 0 ...)` would)
    assert (got.status[bad, 2] & (FLAG_NONFINITE | FLAG_CHOL_FAIL)).all()
    assert np.array_equal(got.status, ref.status)
    assert np.array_equal(got.pose_t[bad], ref.pose_t[bad], equal_nan=True)


def test_long_pose_window(solver):
    """cfg/uwb_pose.yaml keeps 500 poses; keyframe stars over a long window (general path, forest)"""
    topo, batch, _ = synthetic.uwb_pose(8, 500, 8, keyframe_len=20, seed=43)
    cfg = Config(max_iterations=4)
    assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


def test_random_forests(solver):
    """random window structures obeying the forest rule, every edge kind, random insertion order"""
    rng = np.random.default_rng(77)
    for trial in range(6):
        N, A, K, W = int(rng.integers(2, 12)), int(rng.integers(1, 5)), int(rng.integers(0, 3)), 33
        parent = [-1] + [int(rng.integers(0, j)) if rng.random() < 0.85 else -1 for j in range(1, N)]
        edges = []
        for j in range(N):
            for _ in range(int(rng.integers(0, 3))):
                edges.append((EDGE_RANGE_ANCHOR, j, int(rng.integers(0, A)), int(rng.integers(0, K + 1)), int(rng.integers(0, 2))))
            if rng.random() < 0.4:
                edges.append((EDGE_PRIOR, j, 0, 0, int(rng.integers(0, 2))))
            if parent[j] >= 0:
                for _ in range(int(rng.integers(1, 3))):
                    kind = EDGE_SE3 if rng.random() < 0.5 else EDGE_RANGE_POSE
                    edges.append((kind, parent[j], j, int(rng.integers(0, K + 1)) if kind == EDGE_RANGE_POSE else 0,
                                  int(rng.integers(0, 2))))
        order = rng.permutation(len(edges))
        topo = Topology.from_edges(N, A, K, [edges[k] for k in order])
        er, ep, es = topo.counts()

        def rand_R(shape):
            q = rng.normal(size=shape + (4,)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
            w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
            R = np.empty(shape + (3, 3))
            R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
            R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
            R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
            return R

        def spd(shape):
            M = rng.normal(size=shape + (6, 6))
            return M @ np.swapaxes(M, -1, -2) + 0.5 * np.eye(6)
        batch = Batch(pose_t=rng.normal(0, 1, (W, N, 3)), pose_R=rand_R((W, N)), anchors=rng.normal(0, 3, (W, A, 3)),
                      range_d=np.abs(rng.normal(3, 1, (W, er))), range_info=rng.uniform(10, 400, (W, er)),
                      ant_offsets=rng.normal(0, 0.2, (K, 3)) if K else None,
                      prior_Z=np.concatenate([rand_R((W, ep)).reshape(W, ep, 9), rng.normal(0, 1, (W, ep, 3))], -1),
                      prior_info=spd((W, ep)),
                      se3_Z=np.concatenate([rand_R((W, es)).reshape(W, es, 9), rng.normal(0, .3, (W, es, 3))], -1),
                      se3_info=spd((W, es)))
        cfg = Config(max_iterations=4)
        assert_parity(solver.solve(topo, batch, cfg), oracle.solve(topo, batch, cfg))


def test_abi_error_paths(solver):
    """the C ABI reports bad arguments with negative codes and a message, never crashes"""
    import ctypes as C
    from localization_b200 import UwbgoError, _ffi
    lib = _ffi.load_library()
    topo, batch, _ = synthetic.uwb_only(4, 6, 4)
    cfg = Config(max_iterations=2)
    t, b, c = topo.c_struct(), batch.c_struct(), cfg.c_struct()
    r = _ffi.CResult()   # pose_t NULL
    assert lib.uwbgo_solve_batch(solver._h, C.byref(t), C.byref(b), C.byref(c), C.byref(r)) == _ffi.E_INVALID
    assert b"pose_t" in lib.uwbgo_last_error()
    assert lib.uwbgo_solve_batch(None, C.byref(t), C.byref(b), C.byref(c), C.byref(r)) == _ffi.E_INVALID
    b.range_d = None
    res = solver.solve(topo, batch, cfg).c_struct()
    assert lib.uwbgo_solve_batch(solver._h, C.byref(t), C.byref(b), C.byref(c), C.byref(res)) == _ffi.E_INVALID
    bad = Topology.from_edges(3, 1, 0, [(EDGE_RANGE_ANCHOR, 0, 5, 0, 1)])          # anchor index out of range
    with pytest.raises(UwbgoError) as ei:
        solver.solve(bad, Batch(pose_t=np.zeros((1, 3, 3)), anchors=np.zeros((1, 1, 3)), range_d=np.ones((1, 1)),
                                range_info=np.ones((1, 1))), cfg)
    assert ei.value.code == _ffi.E_INVALID
    bad = Topology.from_edges(3, 1, 0, [(EDGE_RANGE_POSE, 2, 1, 0, 1)])            # vertex 1 older than vertex 0
    with pytest.raises(UwbgoError) as ei:
        solver.solve(bad, Batch(pose_t=np.zeros((1, 3, 3)), anchors=np.zeros((1, 1, 3)), range_d=np.ones((1, 1)),
                                range_info=np.ones((1, 1))), cfg)
    assert ei.value.code == _ffi.E_TOPOLOGY
    with pytest.raises(UwbgoError):
        solver.solve(topo, batch, Config(max_iterations=-1))
    with pytest.raises(UwbgoError):
        solver.set_pipeline(8, 1)
    got = solver.solve(topo, batch, cfg)                                            # context still healthy
    assert_parity(got, oracle.solve(topo, batch, cfg))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_branch_free_arithmetic_matches_ieee(solver, mode):
    """The branch-free sqrt / reciprocal / division / log / pivot sequences (NbMath) of the CHAIN and WINDOW
    kernels return the bits of the IEEE operations whenever they do not flag their operand; flagged
    operands make the solver re-run the trial (or the item) with the IEEE operations.  2^28 operands
    per mode; one operand in eight is forced to a significand ending in a run of ones, the one
    divisor class a Newton reciprocal can round wrongly (significand of ALL ones: flagged)."""
    compared, wrong, flagged = solver.selftest_math(1 << 28, mode=mode, seed=20260101 + mode)
    assert compared >= 1 << 28
    assert wrong == 0
    if mode == 1:
        # magnitudes the solver works with leave the fast path only through the all-ones divisors:
        # 1/8 forced x 1/4 with the full run x 1/2 not cleared, in 4 of the 5 operations
        assert 0 < flagged < compared // 20
    else:
        assert 0 < flagged < compared  # NaN / inf / denormal / extreme exponents are flagged


def test_coincident_points_and_extreme_magnitudes(solver):
    """sqrt(0) (poses initialised on top of each other, a pose on an anchor) stays on the branch-free
    arithmetic; denormal-scale and huge coordinates leave its safe range and the trial is re-run with
    the IEEE operations -- both bit-identical to the oracle, window by window inside one tile."""
    topo, batch, _ = synthetic.uwb_only(64, 12, 4, seed=43)
    cfg = Config(max_iterations=5)
    batch.pose_t[0, :, :] = batch.pose_t[0, 0, :]            # all poses of the window coincide
    batch.pose_t[1, 3, :] = batch.anchors[1, 0, :]                    # a pose sits on anchor 0
    batch.pose_t[2] *= 1e-160                                # squares underflow to denormals / zero
    batch.anchors[2] *= 1e-160
    batch.range_d[2] *= 1e-160
    batch.pose_t[3] *= 1e150                                 # squares overflow
    batch.anchors[3] *= 1e150
    batch.range_info[4] *= 1e-300                            # denormal-scale weights
    got = solver.solve(topo, batch, cfg)
    ref = oracle.solve(topo, batch, cfg)
    assert np.array_equal(got.status, ref.status)
    assert np.array_equal(got.pose_t, ref.pose_t, equal_nan=True)
    assert np.array_equal(got.chi2, ref.chi2, equal_nan=True)


@pytest.mark.parametrize("make,iters", [(lambda W: synthetic.uwb_imu_lidar(W, 12, 6, seed=41), 6),
                                        (lambda W: synthetic.uwb_twist(W, 9, 6, seed=42), 5)])
def test_diagonal_information_form(solver, make, iters):
    """UWBGO_DIAG_INFO: prior_info / se3_info as [W][E*][6] diagonals, the 6x6 matrices rebuilt on the device on the
    way into the tile layout (+0.0 off the diagonal): the bits of the full form, through the chunked host pipeline
    (ragged last tile, three lanes) and through the device-pointer entry point"""
    import ctypes as C
    import torch
    from localization_b200 import _ffi
    W = 333
    topo, b, _ = make(W)
    bd = b.with_info_diag()
    cfg = Config(max_iterations=iters)
    ref = oracle.solve(topo, b, cfg)
    solver.set_pipeline(96, 3)
    try:
        got = solver.solve(topo, bd, cfg)
    finally:
        solver.set_pipeline(8192, 8)
    assert solver.last_path == 0                                     # compact inputs take the tile kernels
    assert_parity(got, ref)
    dev = torch.device("cuda", 0)
    keep, cb = {}, _ffi.CBatch()
    cb.n_windows, cb.shared = W, _ffi.DIAG_INFO
    for k in ("pose_t", "pose_R", "anchors", "range_d", "range_info", "prior_Z", "prior_info", "se3_Z", "se3_info"):
        v = getattr(bd, k)
        if v is not None:
            keep[k] = torch.from_numpy(v).to(dev)
            setattr(cb, k, C.cast(C.c_void_p(keep[k].data_ptr()), C.POINTER(C.c_double)))
    if bd.ant_offsets is not None:
        cb.ant_offsets = bd.ant_offsets.ctypes.data_as(C.POINTER(C.c_double))
    N = topo.n_poses
    pose = torch.empty((W, N, 3), dtype=torch.float64, device=dev)
    rot = torch.empty((W, N, 9), dtype=torch.float64, device=dev)
    chi2 = torch.empty((W, 4), dtype=torch.float64, device=dev)
    cr = _ffi.CResult()
    cr.pose_t = C.cast(C.c_void_p(pose.data_ptr()), C.POINTER(C.c_double))
    cr.pose_R = C.cast(C.c_void_p(rot.data_ptr()), C.POINTER(C.c_double))
    cr.chi2 = C.cast(C.c_void_p(chi2.data_ptr()), C.POINTER(C.c_double))
    solver.solve_device(topo, cb, cfg, cr, torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    assert np.array_equal(pose.cpu().numpy(), ref.pose_t) and np.array_equal(chi2.cpu().numpy(), ref.chi2)
    assert np.array_equal(rot.cpu().numpy().reshape(W, N, 3, 3), ref.pose_R)


def test_compact_range_form_and_shared_anchors(solver):
    """uwbgo_range_msgs + UWBGO_SHARED_ANCHORS: the edge parameters of create_range_edge built on the device
    (localization.cpp:316-319,331,338,350) give the bits of the expanded form, through the chunked host pipeline
    and through the device-pointer entry point"""
    import ctypes as C
    import torch
    from localization_b200 import _ffi
    cfg = Config(max_iterations=10)
    W = 1000                                                         # ragged last tile
    tc, bc, _ = synthetic.uwb_only(W, 20, 8, seed=31, compact=True, shared_anchors=True)
    ref = oracle.solve(tc, bc.expanded(tc), cfg)
    solver.set_pipeline(128, 3)
    try:
        got = solver.solve(tc, bc, cfg)                              # compact form, host pointers, 8 chunks
    finally:
        solver.set_pipeline(8192, 8)
    assert solver.last_path == 2
    assert_parity(got, ref)
    assert_parity(solver.solve(tc, bc.expanded(tc), cfg), ref)       # expanded form
    # merged-covariance branch and a non-chain structure (table-driven kernel), per-window anchors
    from localization_b200.graph import RangeMsgs
    rng = np.random.default_rng(5)
    edges = []
    for k in range(9):
        edges += [(EDGE_RANGE_ANCHOR, k, k % 4, 0, 1), (EDGE_RANGE_ANCHOR, k, (k + 1) % 4, 0, 0)]
        if k > 0:
            edges.append((EDGE_RANGE_POSE, k - 1, k, 0, 1))
    t2 = Topology.from_edges(9, 4, 0, edges)
    W2 = 77
    anchors = np.array([[3., 3, 0.5], [-3, 3, 2], [-3, -3, 0.5], [3, -3, 2]])[None] + rng.uniform(-0.2, 0.2, (W2, 4, 3))
    p = rng.uniform(-1, 1, (W2, 9, 3))
    m = RangeMsgs(distance=rng.uniform(2, 5, (W2, 18)), distance_err=rng.choice([0.055, 0.024], (W2, 18)),
                  dt_pose=rng.uniform(0.02, 0.05, (W2, 8)), dt_anchor=rng.uniform(0.0, 0.05, (W2, 18)), v_max=2.5)
    b2 = Batch(pose_t=p, anchors=anchors, range_msgs=m)
    got2, ref2 = solver.solve(t2, b2, cfg), oracle.solve(t2, b2.expanded(t2), cfg)
    assert solver.last_path == 1
    assert_parity(got2, ref2)
    assert_parity(oracle.solve(t2, b2, cfg), ref2)
    # device-pointer entry point
    dev = torch.device("cuda", 0)
    pd = lambda x: C.cast(C.c_void_p(x.data_ptr()), C.POINTER(C.c_double))
    pf = lambda x: C.cast(C.c_void_p(x.data_ptr()), C.POINTER(C.c_float))
    d = {k: torch.from_numpy(v).to(dev) for k, v in dict(pose_t=bc.pose_t, anchors=bc.anchors, dist=bc.range_msgs.distance,
                                                          err=bc.range_msgs.distance_err, dtp=bc.range_msgs.dt_pose).items()}
    cm = _ffi.CRangeMsgs()
    cm.distance, cm.distance_err, cm.dt_pose, cm.v_max = pf(d["dist"]), pf(d["err"]), pd(d["dtp"]), bc.range_msgs.v_max
    cb = _ffi.CBatch()
    cb.n_windows, cb.pose_t, cb.anchors = W, pd(d["pose_t"]), pd(d["anchors"])
    cb.range_msgs, cb.shared = C.pointer(cm), _ffi.SHARED_ANCHORS
    o_t = torch.empty((W, 20, 3), dtype=torch.float64, device=dev)
    o_c = torch.empty((W, 4), dtype=torch.float64, device=dev)
    cr = _ffi.CResult()
    cr.pose_t, cr.chi2 = pd(o_t), pd(o_c)
    solver.solve_device(tc, cb, cfg, cr, torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize(dev)
    assert np.array_equal(o_t.cpu().numpy(), ref.pose_t) and np.array_equal(o_c.cpu().numpy(), ref.chi2)
    # one form only
    bad = Batch(pose_t=bc.pose_t, anchors=bc.anchors, range_msgs=bc.range_msgs, shared_anchors=True)
    cbad = bad.c_struct()
    e = bc.expanded(tc)
    cbad.range_d = e.range_d.ctypes.data_as(C.POINTER(C.c_double))
    res = __import__("localization_b200").Result.empty(W, 20)
    t_, c_, r_ = tc.c_struct(), cfg.c_struct(), res.c_struct()
    assert solver._lib.uwbgo_solve_batch(solver._h, C.byref(t_), C.byref(cbad), C.byref(c_), C.byref(r_)) == _ffi.E_INVALID


@pytest.mark.parametrize("make,iters,path", [
    (lambda: synthetic.uwb_only(300, 50, 8, seed=41), 10, 2),
    (lambda: synthetic.uwb_only(40, 12, 4, seed=42), 3, 2),
    (lambda: synthetic.uwb_imu_lidar(70, 20, 8, seed=43), 20, 0),
    (lambda: synthetic.uwb_imu_lidar(33, 12, 4, antennas=0, lidar=False, seed=44), 10, 0),
    (lambda: synthetic.uwb_twist(70, 15, 8, seed=45), 12, 0),
    (lambda: synthetic.uwb_pose(40, 24, 8, seed=46), 10, 0),
])
def test_edge_chi2_and_marginals(solver, make, iters, path):
    """the extras of Localization::solve()'s commented-out tail (localization.cpp:172-189): per-edge chi2 of the
    last trial and the covariance block of the newest pose, bit-identical to the oracle; the solve itself is
    unchanged by asking for them"""
    topo, batch, _ = make()
    cfg = Config(max_iterations=iters)
    ref = oracle.solve(topo, batch, cfg, edge_chi2=True, marginals=True)
    got = solver.solve(topo, batch, cfg, edge_chi2=True, marginals=True)
    assert solver.last_path == path                                   # extras always take the tile kernels
    assert_parity(got, ref)
    assert np.array_equal(got.edge_chi2, ref.edge_chi2)
    assert np.array_equal(got.marginal_ok, ref.marginal_ok)
    assert np.array_equal(got.marginal, ref.marginal, equal_nan=True)
    s = np.zeros(batch.n_windows)
    for e in range(topo.n_edges):
        s = s + got.edge_chi2[:, e]
    assert np.array_equal(s, got.chi2[:, 2])
    plain = solver.solve(topo, batch, cfg)
    assert np.array_equal(plain.pose_t, got.pose_t) and np.array_equal(plain.chi2, got.chi2)


def test_edge_chi2_after_a_rejected_last_trial_and_table_driven_path(solver):
    """max_trials = 1 ends many windows on a REJECTED trial: edge_chi2 must describe that trial's estimates, which
    the translation-only kernels keep in the other pose buffer; non-chain structure -> lm_fast_kernel"""
    rng = np.random.default_rng(9)
    edges = []
    for k in range(7):
        edges += [(EDGE_RANGE_ANCHOR, k, k % 4, 0, 1), (EDGE_RANGE_ANCHOR, k, (k + 2) % 4, 0, 1)]
        if k > 0:
            edges.append((EDGE_RANGE_POSE, k - 1, k, 0, 1))
    t = Topology.from_edges(7, 4, 0, edges)
    W = 130
    anchors = np.array([[3., 3, 0.5], [-3, 3, 2], [-3, -3, 0.5], [3, -3, 2]])[None] + rng.uniform(-0.2, 0.2, (W, 4, 3))
    b = Batch(pose_t=rng.uniform(-1.5, 1.5, (W, 7, 3)), anchors=anchors, range_d=rng.uniform(0.5, 6, (W, 20)),
              range_info=rng.uniform(50, 400, (W, 20)))
    tc, bc, _ = synthetic.uwb_only(130, 10, 4, seed=8)   # CHAIN windows with inconsistent ranges: first steps get rejected
    bc = Batch(pose_t=bc.pose_t, anchors=bc.anchors, range_info=bc.range_info,
               range_d=np.where(bc.range_d > 0, rng.uniform(0.5, 6, bc.range_d.shape), 0.0))
    rej = oracle.solve(tc, bc, Config(max_iterations=6, max_trials=2, tau=1e-10))
    assert (rej.chi2[:, 0] != rej.chi2[:, 2]).sum() > 60
    for topo, batch, path in ((t, b, 1), (tc, bc, 2)):
        for cfg in (Config(max_iterations=6, max_trials=1, tau=1e-9), Config(max_iterations=6, max_trials=2, tau=1e-10),
                    Config(max_iterations=0)):
            ref = oracle.solve(topo, batch, cfg, edge_chi2=True)
            got = solver.solve(topo, batch, cfg, edge_chi2=True)
            assert solver.last_path == path
            assert_parity(got, ref)
            assert np.array_equal(got.edge_chi2, ref.edge_chi2)
    rej = oracle.solve(t, b, Config(max_iterations=6, max_trials=1))
    assert (rej.chi2[:, 0] != rej.chi2[:, 2]).sum() > 10             # the case is really exercised
