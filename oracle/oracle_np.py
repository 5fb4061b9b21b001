"""TEST INFRASTRUCTURE — second, independent restatement of the reference's hot path in numpy.

Exists only to validate oracle/uwbgo_oracle.c (the reference publishes no golden vectors and
g2o cannot be built here: "parity unpinned", see the header of uwbgo_oracle.c).  It is written
differently on purpose: 4x4 homogeneous matrices, dense 6N x 6N Hessian, numpy.linalg for the
linear solve, and EVERY Jacobian (range, EdgeSE3Prior, EdgeSE3) by central differences of the
error function, so the analytic Jacobians of the C oracle are checked against finite differences.
Agreement is therefore to round-off of two different operation orders (1e-7-ish), not bits.

Follows: EdgeSE3Range::computeError  reference src/types/types_edge_se3range.cpp:105-114;
g2o semantics as restated in SURVEY.md Appendix A (A.2 LM, A.3 numeric Jacobian, A.4 quadratic
form + Cauchy, A.5 EdgeSE3, A.6 EdgeSE3Prior, A.7 VertexSE3::oplus)."""
from __future__ import annotations

import numpy as np

from localization_b200._ffi import EDGE_PRIOR, EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE, EDGE_SE3


def quat_to_R(w, x, y, z):
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat_vec(R):
    """x,y,z of the unit quaternion with w >= 0 (g2o toCompactQuaternion)"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[3] = (R[k, j] - R[j, k]) / s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
    q = q / np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    return q[:3]


def T_from(R, t):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T


def oplus(T, v):
    q = v[3:]
    w2 = 1.0 - q @ q
    Rinc = np.eye(3) if w2 < 0 else quat_to_R(np.sqrt(w2), *q)
    return T @ T_from(Rinc, v[:3])


def mqt(T):
    return np.concatenate([T[:3, 3], R_to_quat_vec(T[:3, :3])])


class Window:
    def __init__(self, topo, batch, w, cfg):
        self.topo, self.cfg = topo, cfg
        N = topo.n_poses
        R = batch.pose_R[w].reshape(N, 3, 3) if batch.pose_R is not None else np.tile(np.eye(3), (N, 1, 1))
        self.X = [T_from(R[i], batch.pose_t[w].reshape(N, 3)[i]) for i in range(N)]
        self.anch = None if batch.anchors is None else batch.anchors[w].reshape(-1, 3)
        self.off = batch.ant_offsets
        g = lambda a, n: None if a is None else a[w].reshape(-1, n) if n > 1 else a[w].reshape(-1)
        self.rd, self.ri = g(batch.range_d, 1), g(batch.range_info, 1)
        self.pZ, self.pI = g(batch.prior_Z, 12), g(batch.prior_info, 36)
        self.sZ, self.sI = g(batch.se3_Z, 12), g(batch.se3_info, 36)
        self.slot = []
        c = [0, 0, 0]
        for k in topo.edge_kind:
            j = 0 if k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE) else (1 if k == EDGE_PRIOR else 2)
            self.slot.append(c[j])
            c[j] += 1

    def edge_error(self, e, X):
        tp = self.topo
        k, a, b, s = tp.edge_kind[e], tp.edge_a[e], tp.edge_b[e], self.slot[e]
        if k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE):
            o = self.off[tp.edge_ant[e] - 1] if tp.edge_ant[e] > 0 else np.zeros(3)
            P0 = X[a][:3, :3] @ o + X[a][:3, 3]
            ab = 0 if tp.edge_ant_b is None else tp.edge_ant_b[e]
            ob = self.off[ab - 1] if ab > 0 else np.zeros(3)
            Q = self.anch[b] + ob if k == EDGE_RANGE_ANCHOR else X[b][:3, :3] @ ob + X[b][:3, 3]
            return np.array([self.rd[s] - np.linalg.norm(P0 - Q)])
        if k == EDGE_PRIOR:
            Z = T_from(self.pZ[s][:9].reshape(3, 3), self.pZ[s][9:])
            return mqt(np.linalg.inv(Z) @ X[a])
        Z = T_from(self.sZ[s][:9].reshape(3, 3), self.sZ[s][9:])
        return mqt(np.linalg.inv(Z) @ np.linalg.inv(X[a]) @ X[b])

    def edge_info(self, e):
        k, s = self.topo.edge_kind[e], self.slot[e]
        if k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE):
            return np.array([[self.ri[s]]])
        return (self.pI if k == EDGE_PRIOR else self.sI)[s].reshape(6, 6)

    def chi2(self, X):
        plain = robust = 0.0
        d2 = self.cfg.kernel_delta ** 2
        for e in range(self.topo.n_edges):
            r = self.edge_error(e, X)
            c = float(r @ self.edge_info(e) @ r)
            plain += c
            robust += d2 * np.log(c / d2 + 1.0) if self.topo.edge_robust[e] else c
        return plain, robust

    def jac(self, e, X, v, delta):
        J = []
        for d in range(6):
            dv = np.zeros(6)
            dv[d] = delta
            Xp = list(X); Xp[v] = oplus(X[v], dv)
            Xm = list(X); Xm[v] = oplus(X[v], -dv)
            J.append((self.edge_error(e, Xp) - self.edge_error(e, Xm)) / (2 * delta))
        return np.stack(J, axis=1)

    def build(self, X):
        tp, N = self.topo, self.topo.n_poses
        H, b = np.zeros((6 * N, 6 * N)), np.zeros(6 * N)
        d2 = self.cfg.kernel_delta ** 2
        for e in range(tp.n_edges):
            k, a, bb = tp.edge_kind[e], tp.edge_a[e], tp.edge_b[e]
            rng = k in (EDGE_RANGE_ANCHOR, EDGE_RANGE_POSE)
            delta = self.cfg.jacobian_delta if rng else 1e-6
            r, O = self.edge_error(e, X), self.edge_info(e)
            rho1 = 1.0 / (float(r @ O @ r) / d2 + 1.0) if tp.edge_robust[e] else 1.0
            verts = [a] + ([bb] if k in (EDGE_RANGE_POSE, EDGE_SE3) else [])
            Js = [self.jac(e, X, v, delta) for v in verts]
            for vi, Ji in zip(verts, Js):
                b[6 * vi:6 * vi + 6] += -rho1 * Ji.T @ O @ r
                for vj, Jj in zip(verts, Js):
                    H[6 * vi:6 * vi + 6, 6 * vj:6 * vj + 6] += rho1 * Ji.T @ O @ Jj
        return H, b

    def solve(self):
        cfg, N = self.cfg, self.topo.n_poses
        X = self.X
        lam, ni = 0.0, 2.0
        plain, cur = self.chi2(X)
        iters = trials = 0
        for it in range(cfg.max_iterations):
            H, b = self.build(X)
            if it == 0:
                lam, ni = cfg.tau * np.max(np.abs(np.diag(H))), 2.0
            rho, q = 0.0, 0
            while True:
                try:
                    L = np.linalg.cholesky(H + lam * np.eye(6 * N))
                    x = np.linalg.solve(L.T, np.linalg.solve(L, b))
                    ok = True
                except np.linalg.LinAlgError:
                    x, ok = np.zeros(6 * N), False
                Xn = [oplus(X[i], x[6 * i:6 * i + 6]) for i in range(N)]
                tplain, tchi = self.chi2(Xn)
                if not ok:
                    tchi = np.finfo(float).max
                rho = (cur - tchi) / (float(x @ (lam * x + b)) + 1e-3)
                if rho > 0 and np.isfinite(tchi):
                    alpha = min(1.0 - (2 * rho - 1) ** 3, cfg.good_step_upper)
                    lam *= max(cfg.good_step_lower, alpha)
                    ni, cur, plain, X = 2.0, tchi, tplain, Xn
                else:
                    lam *= ni
                    ni *= 2
                q += 1
                trials += 1
                if not (rho < 0 and q < cfg.max_trials):
                    break
            iters += 1
            if q == cfg.max_trials or rho == 0:
                break
        self.X = X
        return X, plain, cur, iters, trials


def solve(topo, batch, cfg):
    W, N = batch.n_windows, topo.n_poses
    pose_t, pose_R = np.zeros((W, N, 3)), np.zeros((W, N, 3, 3))
    chi2, status = np.zeros((W, 2)), np.zeros((W, 2), np.int32)
    for w in range(W):
        X, p, r, it, tr = Window(topo, batch, w, cfg).solve()
        for i in range(N):
            pose_t[w, i], pose_R[w, i] = X[i][:3, 3], X[i][:3, :3]
        chi2[w], status[w] = (p, r), (it, tr)
    return pose_t, pose_R, chi2, status


def linearize(topo, batch, cfg):
    W, N = batch.n_windows, topo.n_poses
    H, b, chi = np.zeros((W, 6 * N, 6 * N)), np.zeros((W, 6 * N)), np.zeros((W, 2))
    for w in range(W):
        win = Window(topo, batch, w, cfg)
        H[w], b[w] = win.build(win.X)
        chi[w] = win.chi2(win.X)
    return H, b, chi
