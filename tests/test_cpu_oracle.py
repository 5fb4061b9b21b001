"""CPU tests (no GPU): the oracle against the independent numpy restatement, stage by stage,
against the committed golden vectors, and basic invariants.  The reference publishes no golden
vectors for this path: parity with g2o itself is unpinned."""
import math
import os

import numpy as np
import pytest

from localization_b200 import Batch, Config, Topology, synthetic
from oracle import oracle, oracle_np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dense_H(Hd, Ho, parents=None):
    """H_off[j] = block (parent(j+1), j+1); chains: parent(j) = j-1"""
    W, N = Hd.shape[:2]
    H = np.zeros((W, 6 * N, 6 * N))
    for i in range(N):
        H[:, 6 * i:6 * i + 6, 6 * i:6 * i + 6] = Hd[:, i]
    for j in range(1, N):
        a = j - 1 if parents is None else parents[j]
        if a >= 0:
            H[:, 6 * a:6 * a + 6, 6 * j:6 * j + 6] = Ho[:, j - 1]
            H[:, 6 * j:6 * j + 6, 6 * a:6 * a + 6] = np.swapaxes(Ho[:, j - 1], 1, 2)
    return H


def test_det_log_matches_libm():
    rng = np.random.default_rng(0)
    xs = np.concatenate([1.0 + rng.uniform(0, 1e-3, 200), rng.uniform(1, 50, 400), 10 ** rng.uniform(-300, 300, 200),
                         [1.0, 2.0, 0.5, math.sqrt(2), 1 / math.sqrt(2), 5e-324, 1.7e308]])
    for x in xs:
        a, b = oracle.det_log(x), math.log(x)
        assert abs(a - b) <= 2 ** -52 * max(abs(b), 2 ** -60) * 1.01, (x, a, b)
    assert oracle.det_log(0.0) == -math.inf and math.isnan(oracle.det_log(-1.0))
    assert oracle.det_log(math.inf) == math.inf and math.isnan(oracle.det_log(math.nan))


@pytest.mark.parametrize("make,kw", [
    (synthetic.uwb_only, dict(W=6, N=8, A=4, seed=1)),
    (synthetic.uwb_imu_lidar, dict(W=4, N=6, A=4, seed=2)),
    (synthetic.uwb_twist, dict(W=4, N=6, A=4, seed=3)),
    (synthetic.uwb_pose, dict(W=4, N=10, A=4, keyframe_len=4, seed=9)),
])
def test_linearize_against_numpy_restatement(make, kw):
    topo, batch, _ = make(**kw)
    cfg = Config()
    Hd, Ho, b, chi = oracle.linearize(topo, batch, cfg)
    H2, b2, chi2 = oracle_np.linearize(topo, batch, cfg)
    assert np.allclose(chi, chi2, rtol=1e-12, atol=1e-12)      # residuals, information, Cauchy rho
    H = dense_H(Hd, Ho, topo.parents())
    # numeric range Jacobians carry ~1e-7 relative round-off (delta = 1e-9), SURVEY Appendix B
    scale = np.abs(H).max(axis=(1, 2), keepdims=True)
    assert (np.abs(H - H2) / scale).max() < 5e-6
    assert (np.abs(b.reshape(b2.shape) - b2) / np.abs(b2).max(axis=1, keepdims=True)).max() < 5e-6


def test_vertex1_offsets_against_numpy_restatement():
    """offset[1] of EdgeSE3Range / pidTo of EdgeSE3RangeOffset (types_edge_se3range.cpp:99-114,
    types_edge_se3range_offset.cpp:126-149): antenna offsets on vertex 1, anchors included"""
    topo, batch, _ = synthetic.uwb_imu_lidar(4, 6, 4, seed=12)
    topo = synthetic.with_vertex1_offsets(topo, seed=3)
    assert topo.edge_ant_b is not None and (topo.edge_ant_b[topo.edge_kind == 1] > 0).any()
    assert (topo.edge_ant_b[topo.edge_kind == 0] > 0).any()
    cfg = Config(max_iterations=6)
    Hd, Ho, b, chi = oracle.linearize(topo, batch, cfg)
    H2, b2, chi2 = oracle_np.linearize(topo, batch, cfg)
    assert np.allclose(chi, chi2, rtol=1e-12, atol=1e-12)
    H = dense_H(Hd, Ho, topo.parents())
    scale = np.abs(H).max(axis=(1, 2), keepdims=True)
    assert (np.abs(H - H2) / scale).max() < 5e-6
    # the offsets are seen: the same window without them linearises differently
    base = synthetic.uwb_imu_lidar(4, 6, 4, seed=12)[0]
    assert not np.array_equal(oracle.linearize(base, batch, cfg)[0], Hd)
    ref = oracle.solve(topo, batch, cfg)
    pt, pR, c2, st = oracle_np.solve(topo, batch, cfg)
    assert np.array_equal(ref.status[:, :2], st)
    # two roundings of numeric Jacobians with random lever arms on both ends: ~1e-6 m floor
    assert np.abs(ref.pose_t - pt).max() < 5e-6
    assert np.allclose(ref.chi2[:, :2], c2, rtol=1e-4)


def test_factor_solve_against_dense():
    topo, batch, _ = synthetic.uwb_imu_lidar(8, 10, 4, seed=4)
    Hd, Ho, b, _ = oracle.linearize(topo, batch, Config())
    lam = np.linspace(1e-3, 2.0, 8)
    x, ok = oracle.factor_solve(Hd, Ho, b, lam)
    assert ok.all()
    H = dense_H(Hd, Ho)
    for w in range(8):
        xd = np.linalg.solve(H[w] + lam[w] * np.eye(60), b[w].reshape(-1))
        assert np.allclose(x[w].reshape(-1), xd, rtol=1e-8, atol=1e-11)
    Hd[2, 4] = -np.eye(6)
    x, ok = oracle.factor_solve(Hd, Ho, b, lam)
    assert ok[2] == 0 and not x[2].any() and ok.sum() == 7


@pytest.mark.parametrize("make,kw,iters", [
    (synthetic.uwb_only, dict(W=4, N=8, A=4, seed=5), 10),
    (synthetic.uwb_imu_lidar, dict(W=3, N=6, A=4, seed=6), 8),
    (synthetic.uwb_twist, dict(W=3, N=6, A=4, seed=7), 8),
    (synthetic.uwb_pose, dict(W=3, N=10, A=4, keyframe_len=4, seed=10), 8),     # stars: pose edges to a key vertex
])
def test_full_lm_against_numpy_restatement(make, kw, iters):
    """two independent roundings of the same algorithm: same accept/reject history, poses within
    1e-6 m, chi2 within 1e-4 relative (SURVEY Appendix B measured ~1e-6..1e-5 as the floor)"""
    topo, batch, _ = make(**kw)
    cfg = Config(max_iterations=iters)
    ref = oracle.solve(topo, batch, cfg)
    pt, pR, chi2, st = oracle_np.solve(topo, batch, cfg)
    assert np.array_equal(ref.status[:, :2], st)
    assert np.abs(ref.pose_t - pt).max() < 1e-6
    assert np.abs(ref.pose_R - pR).max() < 1e-5   # weakly observable rotations amplify round-off
    assert np.allclose(ref.chi2[:, :2], chi2, rtol=1e-4)


def test_invariants():
    topo, batch, truth = synthetic.uwb_only(64, 20, 8, seed=8)
    cfg = Config(max_iterations=10)
    r = oracle.solve(topo, batch, cfg, trace=True)
    # UWB-only keeps R = I bit-exactly (rotation columns of the numeric Jacobian are exactly 0)
    assert np.array_equal(r.pose_R, np.broadcast_to(np.eye(3), r.pose_R.shape))
    # accepted steps never increase the robust chi2
    chis = r.trace[:, :, 0]
    assert (np.diff(chis, axis=1) <= 0).all()
    _, _, _, chi0 = oracle.linearize(topo, batch, cfg)
    assert (r.chi2[:, 1] <= chi0[:, 1]).all()
    assert np.abs(r.pose_t - truth).mean() < np.abs(batch.pose_t - truth).mean()
    # the zero-length trajectory edge of the duplicated newest pose has J = 0 exactly:
    Hd, Ho, b, _ = oracle.linearize(topo, batch, cfg)
    assert not Ho[:, -1].any()
    # thread count does not change results
    r1 = oracle.solve(topo, batch, cfg, n_threads=1)
    assert np.array_equal(r1.pose_t, r.pose_t) and np.array_equal(r1.chi2, r.chi2)
    # oplus counter: 12 numeric calls per range-edge end per iteration + 1 per trial
    calls = np.array([12 + (12 if k > 0 else 0) + (12 if k < 19 else 0) for k in range(20)])
    assert np.array_equal(r.oplus_count, 10 * calls[None, :] + r.status[:, 1:2])


def test_golden_vectors():
    """committed outputs of the oracle (tests/golden/make_golden.py) — regression pin"""
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    cases = {"uwb_only": (synthetic.uwb_only, dict(W=8, N=10, A=4, seed=101), 10),
             "uwb_imu_lidar": (synthetic.uwb_imu_lidar, dict(W=6, N=8, A=4, seed=102), 20),
             "uwb_twist": (synthetic.uwb_twist, dict(W=6, N=7, A=4, seed=103), 12),
             "uwb_pose": (synthetic.uwb_pose, dict(W=6, N=11, A=4, keyframe_len=4, seed=104), 10)}
    for name, (make, kw, iters) in cases.items():
        topo, batch, _ = make(**kw)
        assert np.array_equal(batch.pose_t, g[f"{name}_in_pose_t"]), "generator changed"
        r = oracle.solve(topo, batch, Config(max_iterations=iters))
        assert np.array_equal(r.pose_t, g[f"{name}_pose_t"])
        assert np.array_equal(r.pose_R, g[f"{name}_pose_R"])
        assert np.array_equal(r.chi2, g[f"{name}_chi2"])
        assert np.array_equal(r.status, g[f"{name}_status"])
        Hd, Ho, b, chi = oracle.linearize(topo, batch, Config())
        assert np.array_equal(Hd, g[f"{name}_Hd"]) and np.array_equal(b, g[f"{name}_b"])


def test_forest_rule():
    """a pose with two different older neighbours is refused (would need fill-in handling)"""
    from localization_b200._ffi import EDGE_RANGE_ANCHOR, EDGE_SE3
    topo = Topology.from_edges(4, 1, 0, [(EDGE_RANGE_ANCHOR, 0, 0, 0, 1), (EDGE_SE3, 0, 1, 0, 1),
                                         (EDGE_SE3, 1, 2, 0, 1), (EDGE_SE3, 0, 3, 0, 1), (EDGE_SE3, 2, 3, 0, 1)])
    b = Batch(pose_t=np.zeros((1, 4, 3)), anchors=np.ones((1, 1, 3)), range_d=np.ones((1, 1)),
              range_info=np.ones((1, 1)), se3_Z=np.tile(np.r_[np.eye(3).ravel(), 0, 0, 0], (1, 4, 1)),
              se3_info=np.tile(np.eye(6), (1, 4, 1, 1)))
    with pytest.raises(RuntimeError):
        oracle.solve(topo, b, Config(max_iterations=1))


def test_diagonal_information_form_is_the_full_form():
    """UWBGO_DIAG_INFO: the information matrices of EdgeSE3Prior / EdgeSE3 passed as their diagonals ([W][E*][6];
    what Localization builds is Matrix<6,6>::Zero() plus diagonal entries, localization.cpp:478-479,515-518) are
    rebuilt with +0.0 elsewhere: the oracle gives the bits of the full form; a matrix with an entry (or a -0.0) off
    the diagonal cannot be passed this way"""
    for make, iters in ((lambda: synthetic.uwb_imu_lidar(40, 10, 6, seed=5), 5), (lambda: synthetic.uwb_twist(40, 8, 6, seed=6), 4)):
        topo, b, _ = make()
        bd = b.with_info_diag()
        assert bd.info_diag and (bd.prior_info is None or bd.prior_info.shape[-1] == 6 and bd.prior_info.ndim == 3)
        bd.check(topo)
        e = bd.expanded(topo)
        for name in ("prior_info", "se3_info"):
            assert (getattr(b, name) is None) == (getattr(e, name) is None)
            if getattr(b, name) is not None:
                assert np.array_equal(getattr(e, name).reshape(-1), getattr(b, name).reshape(-1))
        cfg = Config(max_iterations=iters)
        r0, r1 = oracle.solve(topo, b, cfg), oracle.solve(topo, bd, cfg)
        assert np.array_equal(r0.pose_t, r1.pose_t) and np.array_equal(r0.pose_R, r1.pose_R)
        assert np.array_equal(r0.chi2, r1.chi2) and np.array_equal(r0.status, r1.status)
        assert np.array_equal(bd.slice(3, 9).prior_info if bd.prior_info is not None else bd.slice(3, 9).se3_info,
                              (bd.prior_info if bd.prior_info is not None else bd.se3_info)[3:9])
    topo, b, _ = synthetic.uwb_imu_lidar(8, 6, 4, seed=1)
    b.prior_info[2, 1, 0, 3] = -0.0
    with pytest.raises(ValueError):
        b.with_info_diag()


def test_compact_range_form_is_the_expanded_form():
    """uwbgo_range_msgs (message fields: float32 distance / distance_err, stamp differences) expands to exactly
    the edge parameters Localization::addRangeEdge computes (localization.cpp:316-319,331,338,350); shared
    anchors are the per-window anchors of a fleet in one anchor field"""
    cfg = Config(max_iterations=10)
    t, b, _ = synthetic.uwb_only(48, 14, 4, seed=21)
    tc, bc, _ = synthetic.uwb_only(48, 14, 4, seed=21, compact=True)
    e = bc.expanded(tc)
    assert np.array_equal(e.range_d, b.range_d) and np.array_equal(e.range_info, b.range_info)
    r0, r1 = oracle.solve(t, b, cfg), oracle.solve(tc, bc, cfg)
    assert np.array_equal(r0.pose_t, r1.pose_t) and np.array_equal(r0.chi2, r1.chi2) and np.array_equal(r0.status, r1.status)
    ts, bs, _ = synthetic.uwb_only(48, 14, 4, seed=21, compact=True, shared_anchors=True)
    assert bs.anchors.shape == (4, 3)
    r2, r3 = oracle.solve(ts, bs, cfg), oracle.solve(ts, bs.expanded(ts), cfg)
    assert np.array_equal(r2.pose_t, r3.pose_t) and np.array_equal(r2.chi2, r3.chi2)
    assert not np.array_equal(r2.pose_t, r0.pose_t)
    # merged-covariance branch (localization.cpp:350): an anchor edge that carries the motion term as well
    from localization_b200.graph import RangeMsgs
    rng = np.random.default_rng(2)
    m = RangeMsgs(distance=bc.range_msgs.distance, distance_err=bc.range_msgs.distance_err, dt_pose=bc.range_msgs.dt_pose,
                  dt_anchor=rng.uniform(0.0, 0.04, bc.range_msgs.distance.shape), v_max=3.0)
    bm = Batch(pose_t=bc.pose_t, anchors=bc.anchors, range_msgs=m)
    rd, ri = m.expand(tc)
    e64 = m.distance_err.astype(np.float64)
    k = 0
    assert np.array_equal(ri[:, 0], 1.0 / (e64[:, 0] * e64[:, 0] + (3.0 * m.dt_anchor[:, 0] / 3.0) ** 2))
    r4 = oracle.solve(tc, bm, cfg)
    r5 = oracle.solve(tc, Batch(pose_t=bc.pose_t, anchors=bc.anchors, range_d=rd, range_info=ri), cfg)
    assert np.array_equal(r4.pose_t, r5.pose_t) and np.array_equal(r4.chi2, r5.chi2)
    with pytest.raises(ValueError):
        Batch(pose_t=bc.pose_t, anchors=bc.anchors, range_msgs=m, range_d=rd, range_info=ri).check(tc)


def _last_linearisation_point(t, b, K):
    """estimates (and oplus counters) at which iteration K of optimize(K) linearises: the result of optimize(K-1)"""
    rp = oracle.solve(t, b, Config(max_iterations=K - 1))
    kw = {k: getattr(b, k) for k in ("anchors", "range_d", "range_info", "ant_offsets", "prior_Z", "prior_info",
                                     "se3_Z", "se3_info")}
    return rp, Batch(pose_t=rp.pose_t, pose_R=rp.pose_R, oplus_count=rp.oplus_count, **kw)


@pytest.mark.parametrize("make,K", [(lambda: synthetic.uwb_twist(12, 15, 8, seed=6), 12),
                                    (lambda: synthetic.uwb_pose(12, 12, 8, seed=7), 10)])
def test_marginal_of_the_newest_pose_against_dense_inverse(make, K):
    """computeMarginals(spinv, last_vertex) (localization.cpp:185-189): the newest pose's block of H^-1, H from the
    last buildSystem; checked against numpy's dense inverse to 1e-9 (chains and key-vertex stars)"""
    t, b, _ = make()
    r = oracle.solve(t, b, Config(max_iterations=K), marginals=True)
    rp, at = _last_linearisation_point(t, b, K)
    Hd, Ho, _, _ = oracle.linearize(t, at, Config())
    checked = 0
    H = dense_H(Hd, Ho, t.parents())
    for w in range(b.n_windows):
        if rp.status[w, 0] != K - 1 or r.status[w, 0] != K:
            continue          # terminated early: the last buildSystem was somewhere else
        assert r.marginal_ok[w] == 1
        S = np.linalg.inv(H[w])[-6:, -6:]
        assert np.abs(S - r.marginal[w]).max() <= 1e-9 * np.abs(S).max()
        assert np.array_equal(r.marginal[w], r.marginal[w].T)
        checked += 1
    assert checked >= 8


def test_marginal_is_refused_when_H_is_singular():
    """UWB-only windows never observe the rotations, and in the uwb_imu stream the newest pose has no IMU prior yet
    when solve() runs: g2o would print "can't compute" (localization.cpp:187-188)"""
    for t, b in (synthetic.uwb_only(4, 10, 4, seed=3)[:2], synthetic.uwb_imu_lidar(4, 12, 4, antennas=0, lidar=False, seed=3)[:2]):
        r = oracle.solve(t, b, Config(max_iterations=5), marginals=True)
        assert (r.marginal_ok == 0).all() and np.isnan(r.marginal).all()
    t, b, _ = synthetic.uwb_twist(4, 15, 8, seed=6)
    r = oracle.solve(t, b, Config(max_iterations=0), marginals=True)      # no buildSystem ran at all
    assert (r.marginal_ok == 0).all()


@pytest.mark.parametrize("make,K", [(lambda: synthetic.uwb_only(64, 12, 4, seed=11), 10),
                                    (lambda: synthetic.uwb_imu_lidar(16, 20, 8, seed=4), 20),
                                    (lambda: synthetic.uwb_twist(16, 15, 8, seed=6), 12)])
def test_edge_chi2_is_the_last_trials(make, K):
    """edge->chi2() after optimize() (what the pruning of localization.cpp:172-181 would test): the terms of
    optimizer.chi2() -- the LAST trial's errors, accepted or not"""
    t, b, _ = make()
    r = oracle.solve(t, b, Config(max_iterations=K), edge_chi2=True)
    s = np.zeros(b.n_windows)
    for e in range(t.n_edges):
        s = s + r.edge_chi2[:, e]
    assert np.array_equal(s, r.chi2[:, 2])
    assert (r.edge_chi2 >= 0).all()
    # where the last trial was accepted the terms are those of the final estimate: re-evaluate them there
    kw = {k: getattr(b, k) for k in ("anchors", "range_d", "range_info", "ant_offsets", "prior_Z", "prior_info",
                                     "se3_Z", "se3_info")}
    at = Batch(pose_t=r.pose_t, pose_R=None if b.pose_R is None else r.pose_R, **kw)
    r0 = oracle.solve(t, at, Config(max_iterations=0), edge_chi2=True)
    accepted = r.chi2[:, 0] == r.chi2[:, 2]
    assert accepted.any()
    assert np.array_equal(r0.edge_chi2[accepted], r.edge_chi2[accepted])
    if (~accepted).any():
        assert not np.array_equal(r0.edge_chi2[~accepted], r.edge_chi2[~accepted])
