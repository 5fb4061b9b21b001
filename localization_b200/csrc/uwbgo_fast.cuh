/*
 * uwbgo_fast.cuh — translation-only paths (UWB-only windows, R = I, no antenna offsets): the
 * table-driven FAST sweeps and their straight-line CHAIN specialisation.  EdgeSE3Range residual
 * (reference src/types/types_edge_se3range.cpp:105-114) with g2o's numeric Jacobian, Cauchy
 * weights, Hessian assembly, matrix-free factor sweep, fused substitution + chi2 sweep.
 */
#ifndef UWBGO_FAST_CUH
#define UWBGO_FAST_CUH

#include "uwbgo_block_solver.cuh"

namespace uwbgo {

/* ------------------------------------------------------------------------------------------ */
/* FAST path (translation-only): EdgeSE3Range with identity offsets on identity rotations       */
/* ------------------------------------------------------------------------------------------ */
constexpr int FAST_MAX_CARRY = 2; /* pose-pose range edges between one consecutive pair */

struct FastEnv {
    const DevTopo *tp;
    Ptrs p;
    Cauchy ck;
    double delta, scalar;
    double *stash;      /* shared memory: [2*FAST_MAX_CARRY][5][CTA_THREADS], this thread's column */
    const double *anch; /* anchors: shared-memory copy (stride blockDim.x) or the tile rows (TILE) */
    int anch_stride;
    int bs;             /* blockDim.x: stride of the shared-memory columns */
};
#define ANCH(E, k) ((E).anch[(size_t)(k) * (E).anch_stride])

/* computeActiveErrors + activeRobustChi2 / chi2, edges in insertion order */
UWBGO_DI void fast_chi_pass(const FastEnv &E, const double *__restrict__ T, double &plain,
                            double &robust)
{
    const DevTopo &tp = *E.tp;
    double p = 0.0, r = 0.0;
    for (int e = 0; e < tp.E; ++e) {
        EdgeRec er = load_edge(tp.edges + e);
        const double *ta = T + (size_t)er.a * 3 * TILE;
        double qx, qy, qz;
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
            qx = ANCH(E, er.b * 3); qy = ANCH(E, er.b * 3 + 1); qz = ANCH(E, er.b * 3 + 2);
        } else {
            const double *tb = T + (size_t)er.b * 3 * TILE;
            qx = ROW(tb, 0); qy = ROW(tb, 1); qz = ROW(tb, 2);
        }
        double n = dist3(ROW(ta, 0), ROW(ta, 1), ROW(ta, 2), qx, qy, qz);
        double err = ROW(E.p.rd, er.slot) - n;
        double Oe = ROW(E.p.ri, er.slot) * err;
        double chi = err * Oe;
        p = p + chi;
        r = r + (er.robust ? E.ck.rho0(chi) : chi);
    }
    plain = p;
    robust = r;
}

/* numeric Jacobian columns 0..2 of a range residual with respect to the translation of the
 * perturbed end point (px,py,pz); the other end point is (qx,qy,qz).  sign = +1: the perturbed
 * point is vertex 0 (dt = P - Q); sign = -1: vertex 1 (dt = Q - P, so pass P/Q swapped).
 * BaseBinaryEdge::linearizeOplus: J[d] = (e(+delta) - e(-delta)) / (2 delta). */
template <class M = IeeeMath>
UWBGO_DI void fast_jac_v0(double px, double py, double pz, double qx, double qy, double qz,
                          double d, double delta, double scalar, double *J, unsigned *badp = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
    double ep, em;
    ep = d - dist3m<M>(delta + px, py, pz, qx, qy, qz, bad);
    em = d - dist3m<M>(-delta + px, py, pz, qx, qy, qz, bad);
    J[0] = scalar * (ep - em);
    ep = d - dist3m<M>(px, delta + py, pz, qx, qy, qz, bad);
    em = d - dist3m<M>(px, -delta + py, pz, qx, qy, qz, bad);
    J[1] = scalar * (ep - em);
    ep = d - dist3m<M>(px, py, delta + pz, qx, qy, qz, bad);
    em = d - dist3m<M>(px, py, -delta + pz, qx, qy, qz, bad);
    J[2] = scalar * (ep - em);
}
template <class M = IeeeMath>
UWBGO_DI void fast_jac_v1(double px, double py, double pz, double qx, double qy, double qz,
                          double d, double delta, double scalar, double *J, unsigned *badp = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
    double ep, em;
    ep = d - dist3m<M>(px, py, pz, delta + qx, qy, qz, bad);
    em = d - dist3m<M>(px, py, pz, -delta + qx, qy, qz, bad);
    J[0] = scalar * (ep - em);
    ep = d - dist3m<M>(px, py, pz, qx, delta + qy, qz, bad);
    em = d - dist3m<M>(px, py, pz, qx, -delta + qy, qz, bad);
    J[1] = scalar * (ep - em);
    ep = d - dist3m<M>(px, py, pz, qx, qy, delta + qz, bad);
    em = d - dist3m<M>(px, py, pz, qx, qy, -delta + qz, bad);
    J[2] = scalar * (ep - em);
}

/* BlockSolver::buildSystem for one window: per pose, gather its edges in insertion order.
 * Writes the H records; returns max |H_kk| (computeLambdaInit). */
/* where fast_linearize puts the system: nowhere (lambda init only) or the tile-layout H records */
enum { SINK_NONE = 0, SINK_RECORDS = 1 };

template <int SINK>
UWBGO_DI double fast_linearize(const FastEnv &E, const double *__restrict__ T)
{
    const DevTopo &tp = *E.tp;
    const int N = tp.N;
    double maxdiag = 0.0;
    double cx = ROW(T, 0), cy = ROW(T, 1), cz = ROW(T, 2); /* pose i */
    double nx = 0.0, ny = 0.0, nz = 0.0;                   /* pose i+1 */
    double fx = 0.0, fy = 0.0, fz = 0.0;                   /* pose i+2, in flight */
    if (N > 1) {
        const double *tn = T + (size_t)3 * TILE;
        nx = ROW(tn, 0); ny = ROW(tn, 1); nz = ROW(tn, 2);
    }
    for (int i = 0; i < N; ++i) {
        if (i + 2 < N) {
            const double *tf = T + (size_t)(i + 2) * 3 * TILE;
            fx = ROW(tf, 0); fy = ROW(tf, 1); fz = ROW(tf, 2);
        }
        double hd[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double ho[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        double bb[3] = {0.0, 0.0, 0.0};
        const int ob = __ldg(tp.op_begin + i), oe = __ldg(tp.op_begin + i + 1);
        for (int o = ob; o < oe; ++o) {
            int2 op = __ldg(reinterpret_cast<const int2 *>(tp.ops + o));
            EdgeRec er = load_edge(tp.edges + op.x);
            double J[3], Ow, omega_r;
            if (op.y == 0) {
                double qx, qy, qz;
                if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                    qx = ANCH(E, er.b * 3); qy = ANCH(E, er.b * 3 + 1); qz = ANCH(E, er.b * 3 + 2);
                } else {
                    qx = nx; qy = ny; qz = nz;
                }
                double d = ROW(E.p.rd, er.slot), info = ROW(E.p.ri, er.slot);
                double err = d - dist3(cx, cy, cz, qx, qy, qz);
                fast_jac_v0(cx, cy, cz, qx, qy, qz, d, E.delta, E.scalar, J);
                double Oe = info * err;
                omega_r = -Oe;
                Ow = info;
                if (er.robust) {
                    double r1 = E.ck.rho1(err * Oe);
                    omega_r = omega_r * r1;
                    Ow = r1 * info;
                }
                if (er.kind == UWBGO_EDGE_RANGE_POSE) {
                    double B[3];
                    fast_jac_v1(cx, cy, cz, qx, qy, qz, d, E.delta, E.scalar, B);
                    double AtO[3] = {J[0] * Ow, J[1] * Ow, J[2] * Ow};
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) ho[3 * r + c] = fma(AtO[r], B[c], ho[3 * r + c]);
                    double *st = E.stash + (size_t)er.ant * 5 * E.bs; /* ant = carry slot */
                    st[0 * E.bs] = B[0];
                    st[1 * E.bs] = B[1];
                    st[2 * E.bs] = B[2];
                    st[3 * E.bs] = Ow;
                    st[4 * E.bs] = omega_r;
                }
            } else {
                const double *st = E.stash + (size_t)er.ant * 5 * E.bs;
                J[0] = st[0 * E.bs];
                J[1] = st[1 * E.bs];
                J[2] = st[2 * E.bs];
                Ow = st[3 * E.bs];
                omega_r = st[4 * E.bs];
            }
            /* constructQuadraticForm, 1-D error: b += J^T omega_r ; H += (J^T Ow) J */
#pragma unroll
            for (int r = 0; r < 3; ++r) bb[r] = fma(J[r], omega_r, bb[r]);
            double JtO[3] = {J[0] * Ow, J[1] * Ow, J[2] * Ow};
            hd[0] = fma(JtO[0], J[0], hd[0]);
            hd[1] = fma(JtO[0], J[1], hd[1]);
            hd[2] = fma(JtO[0], J[2], hd[2]);
            hd[3] = fma(JtO[1], J[1], hd[3]);
            hd[4] = fma(JtO[1], J[2], hd[4]);
            hd[5] = fma(JtO[2], J[2], hd[5]);
        }
        if (SINK == SINK_RECORDS) {
            double *h = E.p.HB + (size_t)i * HR_FAST * TILE;
#pragma unroll
            for (int k = 0; k < 6; ++k) ROW(h, k) = hd[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) ROW(h, 15 + k) = bb[k];
            if (i + 1 < N) {
                double *hn = h + (size_t)HR_FAST * TILE;
#pragma unroll
                for (int k = 0; k < 9; ++k) ROW(hn, 6 + k) = ho[k];
            }
        }
        double v;
        v = fabs(hd[0]); if (v > maxdiag) maxdiag = v;
        v = fabs(hd[3]); if (v > maxdiag) maxdiag = v;
        v = fabs(hd[5]); if (v > maxdiag) maxdiag = v;
        cx = nx; cy = ny; cz = nz;
        nx = fx; ny = fy; nz = fz;
    }
    return maxdiag;
}

/* Matrix-free factor sweep of one LM trial.  H + lambda I is never stored: walking the chain
 * from the newest pose down, the H record of pose i (Hd_i, H_{i-1,i}, b_i) is rebuilt in
 * registers from the current estimates and the measurements (7 rows of HBM traffic instead of
 * 18) and eliminated at once; only the substitution record (c_i, M_i, b_i) is written.  The
 * arithmetic is that of fast_linearize: the edges touching pose i are gathered in insertion
 * order; a pose-pose edge (i-1, i) is linearised when the sweep is at pose i (both Jacobians),
 * its vertex-0 terms travel to pose i-1 through the shared-memory stash. */
UWBGO_DI bool fast_factor_mf(const FastEnv &E, const double *__restrict__ T, double lambda)
{
    const DevTopo &tp = *E.tp;
    const int N = tp.N;
    double G[9], zn[3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 9; ++k) G[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) zn[k] = 0.0;
    const double *tl = T + (size_t)(N - 1) * 3 * TILE;
    double cx = ROW(tl, 0), cy = ROW(tl, 1), cz = ROW(tl, 2); /* pose i   */
    double px = 0.0, py = 0.0, pz = 0.0;                      /* pose i-1 */
    double fx = 0.0, fy = 0.0, fz = 0.0;                      /* pose i-2, in flight */
    if (N > 1) {
        const double *tq = tl - (size_t)3 * TILE;
        px = ROW(tq, 0); py = ROW(tq, 1); pz = ROW(tq, 2);
    }
    for (int i = N - 1; i >= 0; --i) {
        if (i >= 2) {
            const double *tf = T + (size_t)(i - 2) * 3 * TILE;
            fx = ROW(tf, 0); fy = ROW(tf, 1); fz = ROW(tf, 2);
        }
        double h[HR_FAST];
#pragma unroll
        for (int k = 0; k < HR_FAST; ++k) h[k] = 0.0;
        const int ob = __ldg(tp.op_begin + i), oe = __ldg(tp.op_begin + i + 1);
        for (int o = ob; o < oe; ++o) {
            int2 op = __ldg(reinterpret_cast<const int2 *>(tp.ops + o));
            EdgeRec er = load_edge(tp.edges + op.x);
            double J[3], Ow, omega_r;
            if (er.kind == UWBGO_EDGE_RANGE_POSE && op.y == 0) {
                /* edge (i, i+1): vertex-0 terms left by pose i+1 */
                const double *st = E.stash + (size_t)er.ant * 5 * E.bs;
                J[0] = st[0 * E.bs];
                J[1] = st[1 * E.bs];
                J[2] = st[2 * E.bs];
                Ow = st[3 * E.bs];
                omega_r = st[4 * E.bs];
            } else {
                if (UWBGO_L2PF_DIST > 0 && er.slot >= 3 * UWBGO_L2PF_DIST) {
                    prefetch_l2(E.p.rd + (size_t)(er.slot - 3 * UWBGO_L2PF_DIST) * TILE);
                    prefetch_l2(E.p.ri + (size_t)(er.slot - 3 * UWBGO_L2PF_DIST) * TILE);
                }
                const double d = ROW(E.p.rd, er.slot), info = ROW(E.p.ri, er.slot);
                double ax, ay, az, qx, qy, qz;
                if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                    ax = cx; ay = cy; az = cz;
                    qx = ANCH(E, er.b * 3); qy = ANCH(E, er.b * 3 + 1); qz = ANCH(E, er.b * 3 + 2);
                } else { /* edge (i-1, i): vertex 0 is pose i-1 */
                    ax = px; ay = py; az = pz;
                    qx = cx; qy = cy; qz = cz;
                }
                const double err = d - dist3(ax, ay, az, qx, qy, qz);
                const double Oe = info * err;
                omega_r = -Oe;
                Ow = info;
                if (er.robust) {
                    double r1 = E.ck.rho1(err * Oe);
                    omega_r = omega_r * r1;
                    Ow = r1 * info;
                }
                if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                    fast_jac_v0(ax, ay, az, qx, qy, qz, d, E.delta, E.scalar, J);
                } else {
                    double A[3];
                    fast_jac_v0(ax, ay, az, qx, qy, qz, d, E.delta, E.scalar, A);
                    fast_jac_v1(ax, ay, az, qx, qy, qz, d, E.delta, E.scalar, J);
                    double AtO[3] = {A[0] * Ow, A[1] * Ow, A[2] * Ow};
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) h[6 + 3 * r + c] = fma(AtO[r], J[c], h[6 + 3 * r + c]);
                    double *st = E.stash + (size_t)er.ant * 5 * E.bs;
                    st[0 * E.bs] = A[0];
                    st[1 * E.bs] = A[1];
                    st[2 * E.bs] = A[2];
                    st[3 * E.bs] = Ow;
                    st[4 * E.bs] = omega_r;
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) h[15 + r] = fma(J[r], omega_r, h[15 + r]);
            double JtO[3] = {J[0] * Ow, J[1] * Ow, J[2] * Ow};
            h[0] = fma(JtO[0], J[0], h[0]);
            h[1] = fma(JtO[0], J[1], h[1]);
            h[2] = fma(JtO[0], J[2], h[2]);
            h[3] = fma(JtO[1], J[1], h[3]);
            h[4] = fma(JtO[1], J[2], h[4]);
            h[5] = fma(JtO[2], J[2], h[5]);
        }
        double *l = E.p.LR + (size_t)i * LR_FAST * TILE;
        factor_step<3>(h, l, i + 1 < N, i > 0, lambda, G, zn, ok);
#pragma unroll
        for (int k = 0; k < 3; ++k) ROW(l, 12 + k) = h[15 + k];
        cx = px; cy = py; cz = pz;
        px = fx; py = fy; pz = fz;
    }
    return ok;
}

/* One pass after the factor sweep: substitution x_i = c_i - M_i x_{i-1} (ascending), computeScale(),
 * the estimate update (oplus with R = I: t + x) and computeActiveErrors + activeRobustChi2 at the
 * new estimates, following the schedule of DevTopo::sched: edges are summed in insertion order,
 * each as soon as both its poses exist.  The two newest poses stay in registers; an edge that
 * refers further back re-reads the pose it needs.  The L record / b / t of pose i+1 are in flight
 * while pose i and its edges are processed. */
UWBGO_DI void fast_solve_chi(const FastEnv &E, bool ok, double lambda, const double *__restrict__ Tc,
                             double *__restrict__ Tn, double &scale_out, double &plain,
                             double &robust)
{
    const DevTopo &tp = *E.tp;
    const int N = tp.N;
    double xp[3] = {0.0, 0.0, 0.0};
    double scale = 0.0, p = 0.0, r = 0.0;
    double c0 = 0.0, c1 = 0.0, c2 = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0; /* poses ic and ic-1 */
    int ic = -1;
    double nl[LR_FAST], nt[3]; /* prefetched inputs of the next pose */
    {
        const double *l = E.p.LR;
#pragma unroll
        for (int k = 0; k < LR_FAST; ++k) nl[k] = ROW(l, k);
#pragma unroll
        for (int k = 0; k < 3; ++k) nt[k] = ROW(Tc, k);
    }
    const int ns = tp.n_sched;
    for (int s = 0; s < ns; ++s) {
        int2 op = __ldg(reinterpret_cast<const int2 *>(tp.sched + s));
        if (op.x == 0) {
            const int i = op.y;
            double l[LR_FAST], b[3], t[3];
            if (UWBGO_SOLVE_REGPF) {
#pragma unroll
                for (int k = 0; k < LR_FAST; ++k) l[k] = nl[k];
            } else {
                const double *lc = E.p.LR + (size_t)i * LR_FAST * TILE;
#pragma unroll
                for (int k = 0; k < LR_FAST; ++k) l[k] = ROW(lc, k);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                b[k] = l[12 + k];
                t[k] = nt[k];
            }
            if (UWBGO_L2PF_DIST > 0 && i + UWBGO_L2PF_DIST < N) {
                prefetch_rows_l2<LR_FAST>(E.p.LR + (size_t)(i + UWBGO_L2PF_DIST) * LR_FAST * TILE);
                prefetch_rows_l2<3>(Tc + (size_t)(i + UWBGO_L2PF_DIST) * 3 * TILE);
            }
            if (i + 1 < N) {
                const double *ln = E.p.LR + (size_t)(i + 1) * LR_FAST * TILE;
                const double *tn = Tc + (size_t)(i + 1) * 3 * TILE;
                if (UWBGO_SOLVE_REGPF) {
#pragma unroll
                    for (int k = 0; k < LR_FAST; ++k) nl[k] = ROW(ln, k);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) nt[k] = ROW(tn, k);
            }
            subst_step<3>(l, i > 0, xp);
            if (!ok) xp[0] = xp[1] = xp[2] = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) scale = scale + xp[k] * (lambda * xp[k] + b[k]);
            v0 = c0; v1 = c1; v2 = c2;
            c0 = xp[0] + t[0]; c1 = xp[1] + t[1]; c2 = xp[2] + t[2];
            ic = i;
            double *to = Tn + (size_t)i * 3 * TILE;
            ROW(to, 0) = c0; ROW(to, 1) = c1; ROW(to, 2) = c2;
        } else {
            EdgeRec er = load_edge(tp.edges + op.y);
            double ax, ay, az, qx, qy, qz;
            if (er.a == ic) {
                ax = c0; ay = c1; az = c2;
            } else if (er.a == ic - 1) {
                ax = v0; ay = v1; az = v2;
            } else {
                const double *ta = Tn + (size_t)er.a * 3 * TILE;
                ax = ROW(ta, 0); ay = ROW(ta, 1); az = ROW(ta, 2);
            }
            if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                qx = ANCH(E, er.b * 3); qy = ANCH(E, er.b * 3 + 1); qz = ANCH(E, er.b * 3 + 2);
            } else if (er.b == ic) {
                qx = c0; qy = c1; qz = c2;
            } else if (er.b == ic - 1) {
                qx = v0; qy = v1; qz = v2;
            } else {
                const double *tb = Tn + (size_t)er.b * 3 * TILE;
                qx = ROW(tb, 0); qy = ROW(tb, 1); qz = ROW(tb, 2);
            }
            if (UWBGO_L2PF_DIST > 0 && er.slot + 2 * UWBGO_L2PF_DIST < tp.Er) {
                prefetch_l2(E.p.rd + (size_t)(er.slot + 2 * UWBGO_L2PF_DIST) * TILE);
                prefetch_l2(E.p.ri + (size_t)(er.slot + 2 * UWBGO_L2PF_DIST) * TILE);
            }
            double err = ROW(E.p.rd, er.slot) - dist3(ax, ay, az, qx, qy, qz);
            double Oe = ROW(E.p.ri, er.slot) * err;
            double chi = err * Oe;
            p = p + chi;
            r = r + (er.robust ? E.ck.rho0(chi) : chi);
        }
    }
    scale_out = scale;
    plain = p;
    robust = r;
}

/* ------------------------------------------------------------------------------------------ */
/* CHAIN path: the FAST path specialised for the window Localization::addRangeEdge builds         */
/* (localization.cpp:331-340): edges in insertion order are, for pose k = 0..N-1, the anchor     */
/* range edge of pose k followed (k > 0) by the trajectory edge (k-1, k).  With the structure     */
/* known, both sweeps are straight-line code per pose: no edge-table decode, everything the next  */
/* pose needs is loaded one pose ahead, and in the factor sweep the H record of pose i-1 is        */
/* rebuilt in the same basic block in which pose i is eliminated, so the 18 independent sqrt       */
/* chains of the numeric Jacobians fill the issue slots of the sqrt/div dependency chain of the   */
/* 3x3 potrf.  Same arithmetic, same order, same bits as the table-driven FAST path.              */
/* ------------------------------------------------------------------------------------------ */
template <class M = IeeeMath>
UWBGO_DI void chain_weights(const FastEnv &E, double err, double info, bool robust, double &Ow,
                            double &omega_r, unsigned *badp = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
    const double Oe = info * err;
    const double r1 = E.ck.template rho1m<M>(err * Oe, bad);
    omega_r = robust ? (-Oe) * r1 : -Oe;
    Ow = robust ? r1 * info : info;
}
UWBGO_DI void chain_acc(const double *J, double Ow, double omega_r, double *h)
{
#pragma unroll
    for (int r = 0; r < 3; ++r) h[15 + r] = fma(J[r], omega_r, h[15 + r]);
    const double JtO[3] = {J[0] * Ow, J[1] * Ow, J[2] * Ow};
    h[0] = fma(JtO[0], J[0], h[0]);
    h[1] = fma(JtO[0], J[1], h[1]);
    h[2] = fma(JtO[0], J[2], h[2]);
    h[3] = fma(JtO[1], J[1], h[3]);
    h[4] = fma(JtO[1], J[2], h[4]);
    h[5] = fma(JtO[2], J[2], h[5]);
}

/* inputs of one pose of the factor sweep, loaded one pose ahead */
struct ChainIn {
    double px, py, pz;      /* t_{i-1} */
    double da, ia, dt, it;  /* anchor edge of pose i: d, info; edge (i-1, i): d, info */
    int anchor, robust;
};
template <bool PREV>
UWBGO_DI void chain_load_t(const FastEnv &E, const double *__restrict__ T, int i, const int2 tb, ChainIn &in)
{
    in.anchor = tb.x;
    in.robust = tb.y;
    const int sa = i == 0 ? 0 : 2 * i - 1;
    in.da = ROW(E.p.rd, sa);
    in.ia = ROW(E.p.ri, sa);
    if (PREV) {
        const double *tq = T + (size_t)(i - 1) * 3 * TILE;
        in.px = ROW(tq, 0); in.py = ROW(tq, 1); in.pz = ROW(tq, 2);
        in.dt = ROW(E.p.rd, 2 * i);
        in.it = ROW(E.p.ri, 2 * i);
    } else {
        in.px = in.py = in.pz = in.dt = in.it = 0.0;
    }
}

template <bool PREV>
UWBGO_DI void chain_load(const FastEnv &E, const double *__restrict__ T, int i, ChainIn &in)
{
    chain_load_t<PREV>(E, T, i, __ldg(reinterpret_cast<const int2 *>(E.tp->chain + i)), in);
}

/* H record of pose i from (cx,cy,cz) = t_i and `in`; carry = vertex-0 terms of edge (i, i+1) on
 * entry (zeros at the newest pose: an exact no-op), of edge (i-1, i) on exit */
/* CHI: also return in chi4 the chi2 and robustified chi2 of the anchor edge of pose i, then of edge
 * (i-1, i) -- what computeActiveErrors + activeChi2 / activeRobustChi2 sum, term by term */
template <bool PREV, class M = IeeeMath, bool CHI = false>
UWBGO_DI void chain_build_q(const FastEnv &E, double cx, double cy, double cz, double qx, double qy,
                            double qz, const ChainIn &in, double *carry, double *h, unsigned *badp = nullptr,
                            double *chi4 = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
#pragma unroll
    for (int k = 0; k < HR_FAST; ++k) h[k] = 0.0;
    double J[3], Ow, omega_r;
    {
        const double err = in.da - dist3m<M>(cx, cy, cz, qx, qy, qz, bad);
        if (CHI) {
            const double chi = err * (in.ia * err);
            chi4[0] = chi;
            chi4[1] = (in.robust & 1) ? E.ck.template rho0m<M>(chi, bad) : chi;
        }
        fast_jac_v0<M>(cx, cy, cz, qx, qy, qz, in.da, E.delta, E.scalar, J, &bad);
        chain_weights<M>(E, err, in.ia, (in.robust & 1) != 0, Ow, omega_r, &bad);
        chain_acc(J, Ow, omega_r, h);
    }
    double nA[3] = {0.0, 0.0, 0.0}, nOw = 0.0, nOr = 0.0;
    if (PREV) {
        const double err = in.dt - dist3m<M>(in.px, in.py, in.pz, cx, cy, cz, bad);
        if (CHI) {
            const double chi = err * (in.it * err);
            chi4[2] = chi;
            chi4[3] = (in.robust & 2) ? E.ck.template rho0m<M>(chi, bad) : chi;
        }
        fast_jac_v0<M>(in.px, in.py, in.pz, cx, cy, cz, in.dt, E.delta, E.scalar, nA, &bad);
        fast_jac_v1<M>(in.px, in.py, in.pz, cx, cy, cz, in.dt, E.delta, E.scalar, J, &bad);
        chain_weights<M>(E, err, in.it, (in.robust & 2) != 0, nOw, nOr, &bad);
        const double AtO[3] = {nA[0] * nOw, nA[1] * nOw, nA[2] * nOw};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) h[6 + 3 * r + c] = fma(AtO[r], J[c], h[6 + 3 * r + c]);
        chain_acc(J, nOw, nOr, h);
    }
    chain_acc(carry, carry[3], carry[4], h);
    carry[0] = nA[0]; carry[1] = nA[1]; carry[2] = nA[2];
    carry[3] = nOw; carry[4] = nOr;
}

/* ... with the anchor of pose i read from the anchor rows */
template <bool PREV, class M = IeeeMath>
UWBGO_DI void chain_build(const FastEnv &E, double cx, double cy, double cz, const ChainIn &in,
                          double *carry, double *h, unsigned *badp = nullptr)
{
    chain_build_q<PREV, M>(E, cx, cy, cz, ANCH(E, in.anchor * 3), ANCH(E, in.anchor * 3 + 1),
                           ANCH(E, in.anchor * 3 + 2), in, carry, h, badp);
}

UWBGO_DI void chain_store_b(double *__restrict__ l, const double *h)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) ROW(l, 12 + k) = h[15 + k];
}

template <class M = IeeeMath>
UWBGO_DI bool chain_factor(const FastEnv &E, const double *__restrict__ T, double lambda, unsigned *badp = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
    const int N = E.tp->N;
    double G[9], zn[3], carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 9; ++k) G[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) zn[k] = 0.0;
    double *LR = E.p.LR;
    const double *tl = T + (size_t)(N - 1) * 3 * TILE;
    double cx = ROW(tl, 0), cy = ROW(tl, 1), cz = ROW(tl, 2);
    double hc[HR_FAST], hn[HR_FAST];
    ChainIn in, nx;
    if (N == 1) {
        chain_load<false>(E, T, 0, in);
        chain_build<false, M>(E, cx, cy, cz, in, carry, hc, &bad);
        factor_step<3, M>(hc, LR, true, false, lambda, G, zn, ok, &bad);
        chain_store_b(LR, hc);
        return ok;
    }
    chain_load<true>(E, T, N - 1, in);
    if (N > 2) chain_load<true>(E, T, N - 2, nx);
    else chain_load<false>(E, T, 0, nx);
    chain_build<true, M>(E, cx, cy, cz, in, carry, hc, &bad);
    cx = in.px; cy = in.py; cz = in.pz;
    in = nx;
    UWBGO_CHAIN_UNROLL_PRAGMA
    for (int i = N - 1; i >= 2; --i) {
        /* inputs of pose i-2 go in flight; pose i-1 is rebuilt while pose i is eliminated */
        if (i > 2) chain_load<true>(E, T, i - 2, nx);
        else chain_load<false>(E, T, 0, nx);
        if (UWBGO_L2PF_DIST > 0 && i >= 2 + 2 * UWBGO_L2PF_DIST) {
            const int j = i - 2 - 2 * UWBGO_L2PF_DIST;
            prefetch_rows_l2<3>(T + (size_t)j * 3 * TILE);
            prefetch_l2(E.p.rd + (size_t)(2 * j) * TILE);
            prefetch_l2(E.p.ri + (size_t)(2 * j) * TILE);
            prefetch_l2(E.p.rd + (size_t)(2 * j + 1) * TILE);
            prefetch_l2(E.p.ri + (size_t)(2 * j + 1) * TILE);
        }
        chain_build<true, M>(E, cx, cy, cz, in, carry, hn, &bad);
        double *l = LR + (size_t)i * LR_FAST * TILE;
        factor_step<3, M>(hc, l, true, true, lambda, G, zn, ok, &bad);
        chain_store_b(l, hc);
#pragma unroll
        for (int k = 0; k < HR_FAST; ++k) hc[k] = hn[k];
        cx = in.px; cy = in.py; cz = in.pz;
        in = nx;
    }
    /* i == 1: pose 0 has no predecessor */
    chain_build<false, M>(E, cx, cy, cz, in, carry, hn, &bad);
    {
        double *l = LR + (size_t)LR_FAST * TILE;
        factor_step<3, M>(hc, l, true, true, lambda, G, zn, ok, &bad);
        chain_store_b(l, hc);
    }
    factor_step<3, M>(hn, LR, true, false, lambda, G, zn, ok, &bad);
    chain_store_b(LR, hn);
    return ok;
}

/* inputs of one pose of the substitution sweep */
struct ChainSub {
    double l[LR_FAST];
    double tx, ty, tz;
    double da, ia, dt, it;
    int anchor, robust;
};
UWBGO_DI void chain_sub_load(const FastEnv &E, const double *__restrict__ Tc, int i, ChainSub &s)
{
    const double *l = E.p.LR + (size_t)i * LR_FAST * TILE;
#pragma unroll
    for (int k = 0; k < LR_FAST; ++k) s.l[k] = ROW(l, k);
    const double *t = Tc + (size_t)i * 3 * TILE;
    s.tx = ROW(t, 0); s.ty = ROW(t, 1); s.tz = ROW(t, 2);
    const int2 tb = __ldg(reinterpret_cast<const int2 *>(E.tp->chain + i));
    s.anchor = tb.x;
    s.robust = tb.y;
    const int sa = i == 0 ? 0 : 2 * i - 1;
    s.da = ROW(E.p.rd, sa);
    s.ia = ROW(E.p.ri, sa);
    if (i > 0) {
        s.dt = ROW(E.p.rd, 2 * i);
        s.it = ROW(E.p.ri, 2 * i);
    } else {
        s.dt = s.it = 0.0;
    }
}

template <class M = IeeeMath>
UWBGO_DI void chain_solve_chi(const FastEnv &E, bool ok, double lambda, const double *__restrict__ Tc,
                              double *__restrict__ Tn, double &scale_out, double &plain,
                              double &robust, unsigned *badp = nullptr)
{
    unsigned bl = 0;
    unsigned &bad = badp ? *badp : bl;
    const int N = E.tp->N;
    double xp[3] = {0.0, 0.0, 0.0};
    double scale = 0.0, p = 0.0, r = 0.0;
    double vx = 0.0, vy = 0.0, vz = 0.0; /* new estimate of pose i-1 */
    ChainSub cur, nxt;
    chain_sub_load(E, Tc, 0, cur);
    UWBGO_CHAIN_UNROLL_PRAGMA
    for (int i = 0; i < N; ++i) {
        if (i + 1 < N) chain_sub_load(E, Tc, i + 1, nxt);
        if (UWBGO_L2PF_DIST > 0 && i + 1 + UWBGO_L2PF_DIST < N) {
            const int j = i + 1 + UWBGO_L2PF_DIST;
            prefetch_rows_l2<LR_FAST>(E.p.LR + (size_t)j * LR_FAST * TILE);
            prefetch_rows_l2<3>(Tc + (size_t)j * 3 * TILE);
            prefetch_l2(E.p.rd + (size_t)(2 * j) * TILE);
            prefetch_l2(E.p.ri + (size_t)(2 * j) * TILE);
            prefetch_l2(E.p.rd + (size_t)(2 * j - 1) * TILE);
            prefetch_l2(E.p.ri + (size_t)(2 * j - 1) * TILE);
        }
        subst_step<3>(cur.l, i > 0, xp);
        if (!ok) xp[0] = xp[1] = xp[2] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) scale = scale + xp[k] * (lambda * xp[k] + cur.l[12 + k]);
        const double cx = xp[0] + cur.tx, cy = xp[1] + cur.ty, cz = xp[2] + cur.tz;
        double *to = Tn + (size_t)i * 3 * TILE;
        ROW(to, 0) = cx; ROW(to, 1) = cy; ROW(to, 2) = cz;
        {
            const double err = cur.da - dist3m<M>(cx, cy, cz, ANCH(E, cur.anchor * 3), ANCH(E, cur.anchor * 3 + 1),
                                                  ANCH(E, cur.anchor * 3 + 2), bad);
            const double chi = err * (cur.ia * err);
            p = p + chi;
            r = r + ((cur.robust & 1) ? E.ck.template rho0m<M>(chi, bad) : chi);
        }
        if (i > 0) {
            const double err = cur.dt - dist3m<M>(vx, vy, vz, cx, cy, cz, bad);
            const double chi = err * (cur.it * err);
            p = p + chi;
            r = r + ((cur.robust & 2) ? E.ck.template rho0m<M>(chi, bad) : chi);
        }
        vx = cx; vy = cy; vz = cz;
        cur = nxt;
    }
    scale_out = scale;
    plain = p;
    robust = r;
}

}  // namespace uwbgo
#endif
