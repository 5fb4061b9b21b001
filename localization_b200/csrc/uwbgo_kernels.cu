/*
 * uwbgo_kernels.cu — sm_100a kernels of the batched sliding-window LM solver.
 *
 * What is replaced (reference sair-lab/localization): everything below
 * Localization::solve() (src/localization/localization.cpp:164-192), i.e. g2o's
 * initializeOptimization() + optimize(iteration_max) with BlockSolver_6_3,
 * OptimizationAlgorithmLevenberg and LinearSolverCholmod (localization.h:82-85), over
 * EdgeSE3Range (src/types/types_edge_se3range.cpp:105-114, numeric Jacobian of
 * BaseBinaryEdge), EdgeSE3Prior (localization.cpp:462-535) and EdgeSE3 (localization.cpp:560-605).
 *
 * Mapping to the machine.  Windows are independent and every window of a batch has the same
 * graph structure, so ONE THREAD SOLVES ONE WINDOW: all 32 lanes of a warp execute the same
 * instruction on 32 different windows, there is no intra-window reduction, no idle lane in the
 * sequential block-Cholesky chain, and the FP64 pipe sees 32 independent dependency chains per
 * warp.  Per-window state (poses, H, L) does not fit in registers or shared memory for 512
 * windows per SM, so it is streamed through HBM/L2 in the tile layout of uwbgo_internal.h:
 * every access is a coalesced 256-byte row and every sweep walks a tile's rows monotonically.
 *
 * Two instantiations:
 *   FAST    range edges only, R = I, no antenna offsets.  Rotation rows/columns of H and b are
 *           exactly zero there (the residual never reads R), so only the translation 3x3 blocks
 *           are carried; this is bit-identical to carrying the full 6x6 blocks.
 *   GENERAL full 6x6 blocks: rotations, antenna offsets, EdgeSE3Prior, EdgeSE3.
 *
 * Compile with --fmad=false (see uwbgo_math.cuh).
 */
#include "uwbgo_internal.h"
#include "uwbgo_math.cuh"

namespace uwbgo {

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */
#define ROW(p, r) ((p)[(size_t)(r) * TILE])

/* developer knobs for A/B builds (see DESIGN.md "prefetch") */
#ifndef UWBGO_FACTOR_PF
#define UWBGO_FACTOR_PF 1 /* 0 none, 1 register double buffer, 2 L2 prefetch */
#endif
#ifndef UWBGO_SOLVE_REGPF
#define UWBGO_SOLVE_REGPF 1 /* L record of the next pose prefetched into registers */
#endif
#ifndef UWBGO_L2PF_DIST
#define UWBGO_L2PF_DIST 2 /* > 0: prefetch.global.L2 this many records ahead of the sweeps */
#endif

#ifndef UWBGO_CHAIN_UNROLL
#define UWBGO_CHAIN_UNROLL 1 /* unroll factor of the CHAIN sweeps' pose loops */
#endif
#define UWBGO_PRAGMA_(x) _Pragma(#x)
#define UWBGO_PRAGMA(x) UWBGO_PRAGMA_(x)
#define UWBGO_CHAIN_UNROLL_PRAGMA UWBGO_PRAGMA(unroll UWBGO_CHAIN_UNROLL)

UWBGO_DI void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int ROWS>
UWBGO_DI void prefetch_rows_l2(const double *p)
{
#pragma unroll
    for (int k = 0; k < ROWS; ++k) prefetch_l2(p + (size_t)k * TILE);
}

UWBGO_DI EdgeRec load_edge(const EdgeRec *e)
{
    const int4 *p = reinterpret_cast<const int4 *>(e);
    int4 u = __ldg(p), v = __ldg(p + 1);
    EdgeRec r;
    r.kind = u.x; r.a = u.y; r.b = u.z; r.slot = u.w;
    r.ant = v.x; r.robust = v.y; r.base_a = v.z; r.base_b = v.w;
    return r;
}

struct Cauchy {
    double dsqr, dsqrReci;
    UWBGO_DI void init(double delta)
    {
        dsqr = delta * delta;
        dsqrReci = 1.0 / dsqr;
    }
    UWBGO_DI double rho0(double e2) const { return dsqr * det_log(dsqrReci * e2 + 1.0); }
    UWBGO_DI double rho1(double e2) const { return 1.0 / (dsqrReci * e2 + 1.0); }
};

/* thread-private view of the tile-layout workspace */
struct Ptrs {
    double *T0, *T1;   /* translations [N*3] rows, two buffers (selected with ?: so the struct */
    double *Rm0, *Rm1; /* rotations    [N*9] rows (GENERAL)      never needs a local-memory copy) */
    UWBGO_DI double *T(int k) const { return k ? T1 : T0; }
    UWBGO_DI double *Rm(int k) const { return k ? Rm1 : Rm0; }
    int32_t *cnt;
    const double *anch, *rd, *ri, *pZ, *pI, *sZ, *sI;
    double *HB, *LR;
};

template <int HR, int LRR>
UWBGO_DI Ptrs thread_ptrs(const DevTopo &tp, const DevWs &ws, int64_t w)
{
    int64_t tile = w / TILE;
    int lane = (int)(w % TILE);
    Ptrs p;
    p.T0 = ws.T[0] + (tile * (size_t)tp.N * 3) * TILE + lane;
    p.T1 = ws.T[1] + (tile * (size_t)tp.N * 3) * TILE + lane;
    p.Rm0 = ws.Rm[0] ? ws.Rm[0] + (tile * (size_t)tp.N * 9) * TILE + lane : nullptr;
    p.Rm1 = ws.Rm[1] ? ws.Rm[1] + (tile * (size_t)tp.N * 9) * TILE + lane : nullptr;
    p.cnt = ws.cnt ? ws.cnt + (tile * (size_t)tp.N) * TILE + lane : nullptr;
    p.anch = ws.anch ? ws.anch + (tile * (size_t)tp.A * 3) * TILE + lane : nullptr;
    p.rd = ws.rd ? ws.rd + (tile * (size_t)tp.Er) * TILE + lane : nullptr;
    p.ri = ws.ri ? ws.ri + (tile * (size_t)tp.Er) * TILE + lane : nullptr;
    p.pZ = ws.pZ ? ws.pZ + (tile * (size_t)tp.Ep * 12) * TILE + lane : nullptr;
    p.pI = ws.pI ? ws.pI + (tile * (size_t)tp.Ep * 36) * TILE + lane : nullptr;
    p.sZ = ws.sZ ? ws.sZ + (tile * (size_t)tp.Es * 12) * TILE + lane : nullptr;
    p.sI = ws.sI ? ws.sI + (tile * (size_t)tp.Es * 36) * TILE + lane : nullptr;
    p.HB = ws.HB + (tile * (size_t)tp.N * HR) * TILE + lane;
    p.LR = ws.LR ? ws.LR + (tile * (size_t)tp.N * (tp.tree ? LR_TREE : LRR)) * TILE + lane : nullptr;
    return p;
}

/* ------------------------------------------------------------------------------------------ */
/* linear solver: block-tridiagonal Cholesky of H + lambda I, chain eliminated newest pose       */
/* first (replaces LinearSolverCholmod::solve).  D = 3 (FAST) or 6 (GENERAL).                   */
/*   H record of pose i:  Hd_i upper packed | H_{i-1,i} (rows i-1, cols i) | b_i                 */
/*   L record of pose i:  c_i | M_i      with the substitution  x_i = c_i - M_i x_{i-1}          */
/* The factor sweep walks the H records back to front and prefetches record i-1 into registers   */
/* while record i is being eliminated (D = 3), so the HBM latency of the stream hides behind     */
/* the sqrt/div dependency chain of the 3x3 potrf.                                               */
/* ------------------------------------------------------------------------------------------ */
template <int D>
struct Rec {
    static constexpr int TRI = D * (D + 1) / 2, SQ = D * D;
    static constexpr int H = TRI + SQ + D; /* rows of an H record */
    static constexpr int L = D + SQ;       /* rows of an L record */
};

template <int D>
UWBGO_DI void load_hrec(const double *__restrict__ h, double *r)
{
#pragma unroll
    for (int k = 0; k < Rec<D>::H; ++k) r[k] = ROW(h, k);
}

/* one elimination step on the H record held in `h`; G/zn carry G_i and z_{i+1} in, G_{i-1} and
 * z_i out */
template <int D>
UWBGO_DI void factor_step(const double *h, double *__restrict__ l, bool link, bool has_prev,
                          double lambda, double *G, double *zn, bool &ok)
{
    constexpr int TRI = Rec<D>::TRI, SQ = Rec<D>::SQ;
    double S[TRI], L[TRI], z[D], c[D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) {
            double s = h[up_idx(D, cc, r)];
            if (r == cc) s = s + lambda;
            if (link) {
#pragma unroll
                for (int k = 0; k < D; ++k) s = fma(-G[r * D + k], G[cc * D + k], s);
            }
            S[lo_idx(r, cc)] = s;
        }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = S[lo_idx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
        if (!(s > 0.0)) ok = false;
        double inv = 1.0 / sqrt(s);
        L[lo_idx(j, j)] = inv;
#pragma unroll
        for (int r = j + 1; r < D; ++r) {
            double t = S[lo_idx(r, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
            L[lo_idx(r, j)] = t * inv;
        }
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
        double s = h[TRI + SQ + r];
        if (link) {
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(-G[r * D + k], zn[k], s);
        }
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(-L[lo_idx(r, k)], z[k], s);
        z[r] = s * L[lo_idx(r, r)];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) zn[k] = z[k];
#pragma unroll
    for (int r = D - 1; r >= 0; --r) {
        double s = z[r];
#pragma unroll
        for (int k = r + 1; k < D; ++k) s = fma(-L[lo_idx(k, r)], c[k], s);
        c[r] = s * L[lo_idx(r, r)];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) ROW(l, k) = c[k];
    if (has_prev) {
        double M[SQ];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int cc = 0; cc < D; ++cc) {
                double s = h[TRI + r * D + cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) s = fma(-G[r * D + k], L[lo_idx(cc, k)], s);
                G[r * D + cc] = s * L[lo_idx(cc, cc)]; /* row r: entries k < cc are already new */
            }
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int r = D - 1; r >= 0; --r) {
                double s = G[j * D + r];
#pragma unroll
                for (int k = r + 1; k < D; ++k) s = fma(-L[lo_idx(k, r)], M[k * D + j], s);
                M[r * D + j] = s * L[lo_idx(r, r)];
            }
#pragma unroll
        for (int k = 0; k < SQ; ++k) ROW(l, D + k) = M[k];
    }
}

template <int D>
UWBGO_DI bool factor_sweep(const double *__restrict__ HB, double *__restrict__ LR, int N,
                           double lambda)
{
    constexpr int SQ = Rec<D>::SQ, RH = Rec<D>::H, RL = Rec<D>::L;
    double G[SQ], zn[D];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < SQ; ++k) G[k] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) zn[k] = 0.0;
    if (D == 3 && UWBGO_FACTOR_PF == 1) {
        double ra[RH], rb[RH];
        int i = N - 1;
        load_hrec<D>(HB + (size_t)i * RH * TILE, ra);
        while (i >= 0) {
            if (i > 0) load_hrec<D>(HB + (size_t)(i - 1) * RH * TILE, rb);
            factor_step<D>(ra, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok);
            --i;
            if (i < 0) break;
            if (i > 0) load_hrec<D>(HB + (size_t)(i - 1) * RH * TILE, ra);
            factor_step<D>(rb, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok);
            --i;
        }
    } else {
        for (int i = N - 1; i >= 0; --i) {
            double r[RH];
            if (D == 3 && UWBGO_L2PF_DIST > 0 && i - UWBGO_L2PF_DIST >= 0)
                prefetch_rows_l2<RH>(HB + (size_t)(i - UWBGO_L2PF_DIST) * RH * TILE);
            load_hrec<D>(HB + (size_t)i * RH * TILE, r);
            factor_step<D>(r, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok);
        }
    }
    return ok;
}

/* x_i = c_i - M_i x_{i-1}; l = L record values (registers); xp holds x_{i-1} in, x_i out */
template <int D>
UWBGO_DI void subst_step(const double *l, bool link, double *xp)
{
    double x[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        double s = l[r];
        if (link) {
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(-l[D + r * D + j], xp[j], s);
        }
        x[r] = s;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) xp[k] = x[k];
}

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
UWBGO_DI void fast_jac_v0(double px, double py, double pz, double qx, double qy, double qz,
                          double d, double delta, double scalar, double *J)
{
    double ep, em;
    ep = d - dist3(delta + px, py, pz, qx, qy, qz);
    em = d - dist3(-delta + px, py, pz, qx, qy, qz);
    J[0] = scalar * (ep - em);
    ep = d - dist3(px, delta + py, pz, qx, qy, qz);
    em = d - dist3(px, -delta + py, pz, qx, qy, qz);
    J[1] = scalar * (ep - em);
    ep = d - dist3(px, py, delta + pz, qx, qy, qz);
    em = d - dist3(px, py, -delta + pz, qx, qy, qz);
    J[2] = scalar * (ep - em);
}
UWBGO_DI void fast_jac_v1(double px, double py, double pz, double qx, double qy, double qz,
                          double d, double delta, double scalar, double *J)
{
    double ep, em;
    ep = d - dist3(px, py, pz, delta + qx, qy, qz);
    em = d - dist3(px, py, pz, -delta + qx, qy, qz);
    J[0] = scalar * (ep - em);
    ep = d - dist3(px, py, pz, qx, delta + qy, qz);
    em = d - dist3(px, py, pz, qx, -delta + qy, qz);
    J[1] = scalar * (ep - em);
    ep = d - dist3(px, py, pz, qx, qy, delta + qz);
    em = d - dist3(px, py, pz, qx, qy, -delta + qz);
    J[2] = scalar * (ep - em);
}

/* BlockSolver::buildSystem for one window: per pose, gather its edges in insertion order.
 * Writes the H records; returns max |H_kk| (computeLambdaInit). */
template <bool WRITE>
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
        if (WRITE) {
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
UWBGO_DI void chain_weights(const FastEnv &E, double err, double info, bool robust, double &Ow,
                            double &omega_r)
{
    const double Oe = info * err;
    const double r1 = E.ck.rho1(err * Oe);
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
UWBGO_DI void chain_load(const FastEnv &E, const double *__restrict__ T, int i, ChainIn &in)
{
    const int2 tb = __ldg(reinterpret_cast<const int2 *>(E.tp->chain + i));
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

/* H record of pose i from (cx,cy,cz) = t_i and `in`; carry = vertex-0 terms of edge (i, i+1) on
 * entry (zeros at the newest pose: an exact no-op), of edge (i-1, i) on exit */
template <bool PREV>
UWBGO_DI void chain_build(const FastEnv &E, double cx, double cy, double cz, const ChainIn &in,
                          double *carry, double *h)
{
#pragma unroll
    for (int k = 0; k < HR_FAST; ++k) h[k] = 0.0;
    double J[3], Ow, omega_r;
    {
        const double qx = ANCH(E, in.anchor * 3), qy = ANCH(E, in.anchor * 3 + 1), qz = ANCH(E, in.anchor * 3 + 2);
        const double err = in.da - dist3(cx, cy, cz, qx, qy, qz);
        fast_jac_v0(cx, cy, cz, qx, qy, qz, in.da, E.delta, E.scalar, J);
        chain_weights(E, err, in.ia, (in.robust & 1) != 0, Ow, omega_r);
        chain_acc(J, Ow, omega_r, h);
    }
    double nA[3] = {0.0, 0.0, 0.0}, nOw = 0.0, nOr = 0.0;
    if (PREV) {
        const double err = in.dt - dist3(in.px, in.py, in.pz, cx, cy, cz);
        fast_jac_v0(in.px, in.py, in.pz, cx, cy, cz, in.dt, E.delta, E.scalar, nA);
        fast_jac_v1(in.px, in.py, in.pz, cx, cy, cz, in.dt, E.delta, E.scalar, J);
        chain_weights(E, err, in.it, (in.robust & 2) != 0, nOw, nOr);
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

UWBGO_DI void chain_store_b(double *__restrict__ l, const double *h)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) ROW(l, 12 + k) = h[15 + k];
}

UWBGO_DI bool chain_factor(const FastEnv &E, const double *__restrict__ T, double lambda)
{
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
        chain_build<false>(E, cx, cy, cz, in, carry, hc);
        factor_step<3>(hc, LR, true, false, lambda, G, zn, ok);
        chain_store_b(LR, hc);
        return ok;
    }
    chain_load<true>(E, T, N - 1, in);
    if (N > 2) chain_load<true>(E, T, N - 2, nx);
    else chain_load<false>(E, T, 0, nx);
    chain_build<true>(E, cx, cy, cz, in, carry, hc);
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
        chain_build<true>(E, cx, cy, cz, in, carry, hn);
        double *l = LR + (size_t)i * LR_FAST * TILE;
        factor_step<3>(hc, l, true, true, lambda, G, zn, ok);
        chain_store_b(l, hc);
#pragma unroll
        for (int k = 0; k < HR_FAST; ++k) hc[k] = hn[k];
        cx = in.px; cy = in.py; cz = in.pz;
        in = nx;
    }
    /* i == 1: pose 0 has no predecessor */
    chain_build<false>(E, cx, cy, cz, in, carry, hn);
    {
        double *l = LR + (size_t)LR_FAST * TILE;
        factor_step<3>(hc, l, true, true, lambda, G, zn, ok);
        chain_store_b(l, hc);
    }
    factor_step<3>(hn, LR, true, false, lambda, G, zn, ok);
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

UWBGO_DI void chain_solve_chi(const FastEnv &E, bool ok, double lambda, const double *__restrict__ Tc,
                              double *__restrict__ Tn, double &scale_out, double &plain,
                              double &robust)
{
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
            const double err = cur.da - dist3(cx, cy, cz, ANCH(E, cur.anchor * 3), ANCH(E, cur.anchor * 3 + 1),
                                              ANCH(E, cur.anchor * 3 + 2));
            const double chi = err * (cur.ia * err);
            p = p + chi;
            r = r + ((cur.robust & 1) ? E.ck.rho0(chi) : chi);
        }
        if (i > 0) {
            const double err = cur.dt - dist3(vx, vy, vz, cx, cy, cz);
            const double chi = err * (cur.it * err);
            p = p + chi;
            r = r + ((cur.robust & 2) ? E.ck.rho0(chi) : chi);
        }
        vx = cx; vy = cy; vz = cz;
        cur = nxt;
    }
    scale_out = scale;
    plain = p;
    robust = r;
}

/* ------------------------------------------------------------------------------------------ */
/* GENERAL path: 6x6 blocks                                                                     */
/* ------------------------------------------------------------------------------------------ */
struct GenEnv {
    const DevTopo *tp;
    const DevCfg *cfg;
    Ptrs p;
    const double *ant;
    Cauchy ck;
    double delta, scalar;
};

/* a pose buffer of the GENERAL path: translations and rotations in separate tile arrays */
struct PoseBuf {
    double *t, *R;
};
UWBGO_DI void load_pose(const PoseBuf &T, int i, Pose &X)
{
    const double *q = T.t + (size_t)i * 3 * TILE;
    const double *m = T.R + (size_t)i * 9 * TILE;
#pragma unroll
    for (int k = 0; k < 3; ++k) X.t[k] = ROW(q, k);
#pragma unroll
    for (int k = 0; k < 9; ++k) X.R[k] = ROW(m, k);
}
UWBGO_DI void store_pose(const PoseBuf &T, int i, const Pose &X)
{
    double *q = T.t + (size_t)i * 3 * TILE;
    double *m = T.R + (size_t)i * 9 * TILE;
#pragma unroll
    for (int k = 0; k < 3; ++k) ROW(q, k) = X.t[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) ROW(m, k) = X.R[k];
}
UWBGO_DI void load_Zinv(const double *__restrict__ Zrows, int slot, Pose &Zinv)
{
    Pose Z;
    const double *q = Zrows + (size_t)slot * 12 * TILE;
#pragma unroll
    for (int k = 0; k < 9; ++k) Z.R[k] = ROW(q, k);
#pragma unroll
    for (int k = 0; k < 3; ++k) Z.t[k] = ROW(q, 9 + k);
    pose_inv(Z, Zinv);
}

/* (X * offset).translation() for a translation-only offset: R o + t */
UWBGO_DI void offset_point(const GenEnv &E, const Pose &X, int ant, double *P)
{
    if (ant > 0) {
        double o[3] = {__ldg(E.ant + 3 * (ant - 1)), __ldg(E.ant + 3 * (ant - 1) + 1),
                       __ldg(E.ant + 3 * (ant - 1) + 2)};
        mat3_vec_add(X.R, o, X.t, P);
    } else {
        P[0] = X.t[0]; P[1] = X.t[1]; P[2] = X.t[2];
    }
}

/* toVectorMQT(Zinv * Xi^-1 * Xj) */
UWBGO_DI void se3_error(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *e)
{
    Pose Xi_inv, T, Dl;
    pose_inv(Xi, Xi_inv);
    pose_mul(Zinv, Xi_inv, T);
    pose_mul(T, Xj, Dl);
    double q[4];
    R_to_quat(Dl.R, q);
    e[0] = Dl.t[0]; e[1] = Dl.t[1]; e[2] = Dl.t[2];
    e[3] = q[0]; e[4] = q[1]; e[5] = q[2];
}

/* chi2 = e . (Omega e) for a 6-D edge; Oe returned */
UWBGO_DI double chi2_6(const double *__restrict__ Irows, int slot, const double *e, double *Oe)
{
    const double *O = Irows + (size_t)slot * 36 * TILE;
    double chi = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = ROW(O, 6 * r) * e[0];
#pragma unroll
        for (int c = 1; c < 6; ++c) s = s + ROW(O, 6 * r + c) * e[c];
        Oe[r] = s;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) chi = chi + e[r] * Oe[r];
    return chi;
}

UWBGO_DI void gen_chi_pass(const GenEnv &E, const PoseBuf &T, double &plain, double &robust)
{
    const DevTopo &tp = *E.tp;
    double p = 0.0, r = 0.0;
    for (int e = 0; e < tp.E; ++e) {
        EdgeRec er = load_edge(tp.edges + e);
        double chi;
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
            Pose Xa;
            load_pose(T, er.a, Xa);
            double P0[3], Q[3];
            offset_point(E, Xa, er.ant, P0);
            if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                const double *an = E.p.anch + (size_t)er.b * 3 * TILE;
                Q[0] = ROW(an, 0); Q[1] = ROW(an, 1); Q[2] = ROW(an, 2);
            } else {
                const double *tb = T.t + (size_t)er.b * 3 * TILE;
                Q[0] = ROW(tb, 0); Q[1] = ROW(tb, 1); Q[2] = ROW(tb, 2);
            }
            double err = ROW(E.p.rd, er.slot) - dist3(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2]);
            double Oe = ROW(E.p.ri, er.slot) * err;
            chi = err * Oe;
        } else if (er.kind == UWBGO_EDGE_PRIOR) {
            Pose Zinv, X, Dl;
            load_Zinv(E.p.pZ, er.slot, Zinv);
            load_pose(T, er.a, X);
            pose_mul(Zinv, X, Dl);
            double q[4], e6[6], Oe[6];
            R_to_quat(Dl.R, q);
            e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
            e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
            chi = chi2_6(E.p.pI, er.slot, e6, Oe);
        } else {
            Pose Zinv, Xi, Xj;
            load_Zinv(E.p.sZ, er.slot, Zinv);
            load_pose(T, er.a, Xi);
            load_pose(T, er.b, Xj);
            double e6[6], Oe[6];
            se3_error(Zinv, Xi, Xj, e6);
            chi = chi2_6(E.p.sI, er.slot, e6, Oe);
        }
        p = p + chi;
        r = r + (er.robust ? E.ck.rho0(chi) : chi);
    }
    plain = p;
    robust = r;
}

/* numeric Jacobian of a range residual wrt vertex 0 (pose X with antenna offset `ant`); Q is the
 * other end point.  c0 = the pose's oplus counter when this linearisation started, base = oplus
 * calls made on it by earlier edges of this linearisation.  Call k trips the re-orthogonalisation
 * of the PERTURBED estimate when (c0 + k) % mod == 0 (VertexSE3::oplusImpl; push/pop restores the
 * estimate, not the counter). */
UWBGO_DI void gen_jac_v0(const GenEnv &E, const Pose &X, int ant, const double *Q, double d, int c0,
                         int base, double *J)
{
    const int mod = E.cfg->orth_mod;
    double o[3] = {0.0, 0.0, 0.0};
    if (ant > 0) {
        o[0] = __ldg(E.ant + 3 * (ant - 1));
        o[1] = __ldg(E.ant + 3 * (ant - 1) + 1);
        o[2] = __ldg(E.ant + 3 * (ant - 1) + 2);
    }
    int call = c0 + base;
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double epm[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            ++call;
            double v = sg == 0 ? E.delta : -E.delta;
            double tp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tp[r] = X.R[3 * r + dd] * v + X.t[r];
            double P[3];
            if (ant > 0) {
                if (call % mod == 0) {
                    double Rp[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rp[k] = X.R[k];
                    orthogonalize(Rp);
                    mat3_vec_add(Rp, o, tp, P);
                } else
                    mat3_vec_add(X.R, o, tp, P);
            } else {
                P[0] = tp[0]; P[1] = tp[1]; P[2] = tp[2];
            }
            epm[sg] = d - dist3(P[0], P[1], P[2], Q[0], Q[1], Q[2]);
        }
        J[dd] = E.scalar * (epm[0] - epm[1]);
    }
    if (ant > 0) {
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) {
            double epm[2];
#pragma unroll
            for (int sg = 0; sg < 2; ++sg) {
                ++call;
                double q[3] = {0.0, 0.0, 0.0};
                q[dd] = sg == 0 ? E.delta : -E.delta;
                double Rinc[9], Rp[9], P[3];
                increment_R(q, Rinc);
                mat3_mul(X.R, Rinc, Rp);
                if (call % mod == 0) orthogonalize(Rp);
                mat3_vec_add(Rp, o, X.t, P);
                epm[sg] = d - dist3(P[0], P[1], P[2], Q[0], Q[1], Q[2]);
            }
            J[3 + dd] = E.scalar * (epm[0] - epm[1]);
        }
    } else {
        J[3] = 0.0; J[4] = 0.0; J[5] = 0.0;
    }
}

/* numeric Jacobian wrt vertex 1 (pose X, identity offset); P0 is the unperturbed vertex-0 point.
 * Its point is X.t, which only translation increments move: rotation columns are exactly 0 and a
 * re-orthogonalisation of the perturbed R is unobservable. */
UWBGO_DI void gen_jac_v1(const GenEnv &E, const double *P0, const Pose &X, double d, double *J)
{
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double epm[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            double v = sg == 0 ? E.delta : -E.delta;
            double tp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tp[r] = X.R[3 * r + dd] * v + X.t[r];
            epm[sg] = d - dist3(P0[0], P0[1], P0[2], tp[0], tp[1], tp[2]);
        }
        J[dd] = E.scalar * (epm[0] - epm[1]);
    }
    J[3] = 0.0; J[4] = 0.0; J[5] = 0.0;
}

/* rows of the quaternion product matrices, 4-vectors ordered {w,x,y,z}; q = {x,y,z,w} */
UWBGO_DI void quat_left(const double *q, double *M)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    M[0] = w;  M[1] = -x; M[2] = -y;  M[3] = -z;
    M[4] = x;  M[5] = w;  M[6] = -z;  M[7] = y;
    M[8] = y;  M[9] = z;  M[10] = w;  M[11] = -x;
    M[12] = z; M[13] = -y; M[14] = x; M[15] = w;
}
UWBGO_DI void quat_right(const double *q, double *M)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    M[0] = w;  M[1] = -x; M[2] = -y;  M[3] = -z;
    M[4] = x;  M[5] = w;  M[6] = z;   M[7] = -y;
    M[8] = y;  M[9] = -z; M[10] = w;  M[11] = x;
    M[12] = z; M[13] = y; M[14] = -x; M[15] = w;
}

/* d(vector part of qE (x) dq)/d(dq) = w I + [q]x */
UWBGO_DI void set_jqq(const double *q, double *J /* 6x6, block (3,3) */)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    J[6 * 3 + 3] = w;  J[6 * 3 + 4] = -z; J[6 * 3 + 5] = y;
    J[6 * 4 + 3] = z;  J[6 * 4 + 4] = w;  J[6 * 4 + 5] = -x;
    J[6 * 5 + 3] = -y; J[6 * 5 + 4] = x;  J[6 * 5 + 5] = w;
}

/* analytic Jacobians of EdgeSE3 (computeEdgeSE3Gradient with identity offsets) */
UWBGO_DI void se3_jacobians(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *Ji, double *Jj,
                            bool want_i)
{
    Pose Xi_inv, Bm, AB;
    pose_inv(Xi, Xi_inv);
    pose_mul(Xi_inv, Xj, Bm);
    pose_mul(Zinv, Bm, AB);
#pragma unroll
    for (int k = 0; k < 36; ++k) Jj[k] = 0.0;
    double qE[4];
    R_to_quat(AB.R, qE);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Jj[6 * r + c] = AB.R[3 * r + c];
    set_jqq(qE, Jj);
    if (!want_i) return;
#pragma unroll
    for (int k = 0; k < 36; ++k) Ji[k] = 0.0;
    const double *Ra = Zinv.R, *tb = Bm.t;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ji[6 * r + c] = -Ra[3 * r + c];
    double S[9] = {0.0, -2.0 * tb[2], 2.0 * tb[1], 2.0 * tb[2], 0.0, -2.0 * tb[0],
                   -2.0 * tb[1], 2.0 * tb[0], 0.0};
    double RaS[9];
    mat3_mul(Ra, S, RaS);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = RaS[3 * r + c];
    double qA[4], qB[4], Lm[16], Rm[16];
    R_to_quat(Ra, qA);
    R_to_quat(Bm.R, qB);
    quat_left(qA, Lm);
    quat_right(qB, Rm);
    double wAB = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) wAB = wAB + Lm[k] * Rm[4 * k];
    double sgn = wAB < 0.0 ? 1.0 : -1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = s + Lm[4 * (r + 1) + k] * Rm[4 * k + (c + 1)];
            Ji[6 * (3 + r) + 3 + c] = sgn * s;
        }
}

/* constructQuadraticForm pieces.  hd = upper packed 6x6 (21), ho = 6x6, bb = 6 */
UWBGO_DI void acc1_diag(const double *J, double Ow, double omega_r, double *hd, double *bb)
{
#pragma unroll
    for (int r = 0; r < 6; ++r) bb[r] = fma(J[r], omega_r, bb[r]);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double JtO = J[r] * Ow;
#pragma unroll
        for (int c = r; c < 6; ++c) hd[up_idx(6, r, c)] = fma(JtO, J[c], hd[up_idx(6, r, c)]);
    }
}
UWBGO_DI void acc1_off(const double *A, const double *B, double Ow, double *ho)
{
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double AtO = A[r] * Ow;
#pragma unroll
        for (int c = 0; c < 6; ++c) ho[6 * r + c] = fma(AtO, B[c], ho[6 * r + c]);
    }
}
/* JtO = J^T Ow (6x6, Ow row-major rows in tile layout scaled by r1 when robust) */
UWBGO_DI void jt_omega(const double *J, const double *__restrict__ O, bool robust, double r1,
                       double *JtO)
{
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double ow[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            double v = ROW(O, 6 * k + c);
            ow[k] = robust ? r1 * v : v;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = J[r] * ow[0];
#pragma unroll
            for (int k = 1; k < 6; ++k) s = fma(J[6 * k + r], ow[k], s);
            JtO[6 * r + c] = s;
        }
    }
}
UWBGO_DI void acc6_b(const double *J, const double *omega_r, double *bb)
{
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = J[r] * omega_r[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) s = fma(J[6 * k + r], omega_r[k], s);
        bb[r] = bb[r] + s;
    }
}
UWBGO_DI void acc6_diag(const double *JtO, const double *J, double *hd)
{
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) {
            double s = JtO[6 * r] * J[c];
#pragma unroll
            for (int k = 1; k < 6; ++k) s = fma(JtO[6 * r + k], J[6 * k + c], s);
            hd[up_idx(6, r, c)] = hd[up_idx(6, r, c)] + s;
        }
}
UWBGO_DI void acc6_off(const double *AtO, const double *B, double *ho)
{
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = AtO[6 * r] * B[c];
#pragma unroll
            for (int k = 1; k < 6; ++k) s = fma(AtO[6 * r + k], B[6 * k + c], s);
            ho[6 * r + c] = ho[6 * r + c] + s;
        }
}

/* BlockSolver::buildSystem, general edges.  Advances the oplus counters by the numeric-Jacobian
 * calls.  Returns max |H_kk|. */
__device__ __noinline__ double gen_linearize(const GenEnv &E, const PoseBuf &T)
{
    const DevTopo &tp = *E.tp;
    const int N = tp.N, mod = E.cfg->orth_mod;
    double maxdiag = 0.0;
    for (int i = 0; i < N; ++i) {
        Pose Xi;
        load_pose(T, i, Xi);
        const int ci = E.p.cnt[(size_t)i * TILE];
        double hd[21], ho[36], bb[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) hd[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 36; ++k) ho[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) bb[k] = 0.0;
        const int ob = __ldg(tp.op_begin + i), oe = __ldg(tp.op_begin + i + 1);
        for (int o = ob; o < oe; ++o) {
            int2 op = __ldg(reinterpret_cast<const int2 *>(tp.ops + o));
            EdgeRec er = load_edge(tp.edges + op.x);
            if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
                double d = ROW(E.p.rd, er.slot), info = ROW(E.p.ri, er.slot);
                double P0[3], Q[3], J[6];
                Pose Xo; /* the other pose of a pose-pose edge */
                if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                    const double *an = E.p.anch + (size_t)er.b * 3 * TILE;
                    Q[0] = ROW(an, 0); Q[1] = ROW(an, 1); Q[2] = ROW(an, 2);
                    offset_point(E, Xi, er.ant, P0);
                } else if (op.y == 0) {
                    load_pose(T, er.b, Xo);
                    Q[0] = Xo.t[0]; Q[1] = Xo.t[1]; Q[2] = Xo.t[2];
                    offset_point(E, Xi, er.ant, P0);
                } else {
                    load_pose(T, er.a, Xo);
                    Q[0] = Xi.t[0]; Q[1] = Xi.t[1]; Q[2] = Xi.t[2];
                    offset_point(E, Xo, er.ant, P0);
                }
                double err = d - dist3(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2]);
                double Oe = info * err;
                double omega_r = -Oe, Ow = info;
                if (er.robust) {
                    double r1 = E.ck.rho1(err * Oe);
                    omega_r = omega_r * r1;
                    Ow = r1 * info;
                }
                if (op.y == 0) {
                    gen_jac_v0(E, Xi, er.ant, Q, d, ci, er.base_a, J);
                    acc1_diag(J, Ow, omega_r, hd, bb);
                } else {
                    /* vertex 1: its own terms, and the block H_{a,i} = A^T Ow B of the pair.  Pose a was
                     * swept earlier, so its counter already includes this linearisation's calls. */
                    gen_jac_v1(E, P0, Xi, d, J);
                    acc1_diag(J, Ow, omega_r, hd, bb);
                    double A[6];
                    const int ca_now = E.p.cnt[(size_t)er.a * TILE];
                    const int ca = ((ca_now - __ldg(tp.num_calls + er.a)) % mod + mod) % mod;
                    gen_jac_v0(E, Xo, er.ant, Q, d, ca, er.base_a, A);
                    acc1_off(A, J, Ow, ho);
                }
            } else if (er.kind == UWBGO_EDGE_PRIOR) {
                Pose Zinv, Dl;
                load_Zinv(E.p.pZ, er.slot, Zinv);
                pose_mul(Zinv, Xi, Dl);
                double q[4], e6[6], Oe[6], J[36], JtO[36];
                R_to_quat(Dl.R, q);
                e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
                e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
                double chi = chi2_6(E.p.pI, er.slot, e6, Oe);
                double r1 = er.robust ? E.ck.rho1(chi) : 1.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    Oe[k] = -Oe[k];
                    if (er.robust) Oe[k] = Oe[k] * r1;
                }
#pragma unroll
                for (int k = 0; k < 36; ++k) J[k] = 0.0;
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) J[6 * r + c] = Dl.R[3 * r + c];
                set_jqq(q, J);
                acc6_b(J, Oe, bb);
                jt_omega(J, E.p.pI + (size_t)er.slot * 36 * TILE, er.robust != 0, r1, JtO);
                acc6_diag(JtO, J, hd);
            } else { /* EdgeSE3 */
                Pose Zinv, Xo;
                load_Zinv(E.p.sZ, er.slot, Zinv);
                double e6[6], Oe[6], Ji[36], Jj[36], JtO[36];
                if (op.y == 0) {
                    load_pose(T, er.b, Xo);
                    se3_error(Zinv, Xi, Xo, e6);
                    se3_jacobians(Zinv, Xi, Xo, Ji, Jj, true);
                } else {
                    load_pose(T, er.a, Xo);
                    se3_error(Zinv, Xo, Xi, e6);
                    se3_jacobians(Zinv, Xo, Xi, Ji, Jj, true);
                }
                double chi = chi2_6(E.p.sI, er.slot, e6, Oe);
                double r1 = er.robust ? E.ck.rho1(chi) : 1.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    Oe[k] = -Oe[k];
                    if (er.robust) Oe[k] = Oe[k] * r1;
                }
                const double *O = E.p.sI + (size_t)er.slot * 36 * TILE;
                if (op.y == 0) {
                    acc6_b(Ji, Oe, bb);
                    jt_omega(Ji, O, er.robust != 0, r1, JtO);
                    acc6_diag(JtO, Ji, hd);
                } else {
                    acc6_b(Jj, Oe, bb);
                    jt_omega(Jj, O, er.robust != 0, r1, JtO);
                    acc6_diag(JtO, Jj, hd);
                    jt_omega(Ji, O, er.robust != 0, r1, JtO);
                    acc6_off(JtO, Jj, ho); /* H_{a,i}: rows of pose a, columns of pose i */
                }
            }
        }
        double *h = E.p.HB + (size_t)i * HR_GEN * TILE;
#pragma unroll
        for (int k = 0; k < 21; ++k) ROW(h, k) = hd[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) ROW(h, 57 + k) = bb[k];
#pragma unroll
        for (int k = 0; k < 36; ++k) ROW(h, 21 + k) = ho[k]; /* H_{parent(i), i} */
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double v = fabs(hd[up_idx(6, r, r)]);
            if (v > maxdiag) maxdiag = v;
        }
        E.p.cnt[(size_t)i * TILE] = (ci + __ldg(tp.num_calls + i)) % mod;
    }
    return maxdiag;
}

/* Forest windows (pose edges to a key vertex, localization.cpp:258-267): every pose has at most one
 * older neighbour parent(i) < i, not necessarily i-1.  Same elimination as factor_sweep<6>, newest
 * pose first and therefore without fill, but a pose may have several children, whose G and z are
 * read back from their L records:  L record (tree) = c 6 | M 36 | G 36 | z 6 | x 6. */
__device__ __noinline__ bool factor_sweep_tree(const DevTopo &tp, const double *__restrict__ HB,
                                               double *__restrict__ LR, double lambda)
{
    const int N = tp.N;
    bool ok = true;
    for (int i = N - 1; i >= 0; --i) {
        const double *h = HB + (size_t)i * HR_GEN * TILE;
        double *l = LR + (size_t)i * LR_TREE * TILE;
        double S[21], L[21], z[6], c[6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int cc = 0; cc <= r; ++cc) {
                double s = ROW(h, up_idx(6, cc, r));
                if (r == cc) s = s + lambda;
                S[lo_idx(r, cc)] = s;
            }
#pragma unroll
        for (int r = 0; r < 6; ++r) z[r] = ROW(h, 57 + r);
        const int cb = __ldg(tp.child_begin + i), ce = __ldg(tp.child_begin + i + 1);
        for (int q = cb; q < ce; ++q) { /* children in descending order */
            const double *lc = LR + (size_t)__ldg(tp.children + q) * LR_TREE * TILE;
            double G[36], zc[6];
#pragma unroll
            for (int k = 0; k < 36; ++k) G[k] = ROW(lc, 42 + k);
#pragma unroll
            for (int k = 0; k < 6; ++k) zc[k] = ROW(lc, 78 + k);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cc = 0; cc <= r; ++cc) {
                    double s = S[lo_idx(r, cc)];
#pragma unroll
                    for (int k = 0; k < 6; ++k) s = fma(-G[r * 6 + k], G[cc * 6 + k], s);
                    S[lo_idx(r, cc)] = s;
                }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double s = z[r];
#pragma unroll
                for (int k = 0; k < 6; ++k) s = fma(-G[r * 6 + k], zc[k], s);
                z[r] = s;
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double s = S[lo_idx(j, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
            if (!(s > 0.0)) ok = false;
            double inv = 1.0 / sqrt(s);
            L[lo_idx(j, j)] = inv;
#pragma unroll
            for (int r = j + 1; r < 6; ++r) {
                double t = S[lo_idx(r, j)];
#pragma unroll
                for (int k = 0; k < j; ++k) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
                L[lo_idx(r, j)] = t * inv;
            }
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = z[r];
#pragma unroll
            for (int k = 0; k < r; ++k) s = fma(-L[lo_idx(r, k)], z[k], s);
            z[r] = s * L[lo_idx(r, r)];
        }
#pragma unroll
        for (int r = 5; r >= 0; --r) {
            double s = z[r];
#pragma unroll
            for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], c[k], s);
            c[r] = s * L[lo_idx(r, r)];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ROW(l, k) = c[k];
            ROW(l, 78 + k) = z[k];
        }
        if (__ldg(tp.parent + i) >= 0) {
            double G[36], M[36];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) {
                    double s = ROW(h, 21 + r * 6 + cc);
#pragma unroll
                    for (int k = 0; k < cc; ++k) s = fma(-G[r * 6 + k], L[lo_idx(cc, k)], s);
                    G[r * 6 + cc] = s * L[lo_idx(cc, cc)];
                }
#pragma unroll
            for (int j = 0; j < 6; ++j)
#pragma unroll
                for (int r = 5; r >= 0; --r) {
                    double s = G[j * 6 + r];
#pragma unroll
                    for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], M[k * 6 + j], s);
                    M[r * 6 + j] = s * L[lo_idx(r, r)];
                }
#pragma unroll
            for (int k = 0; k < 36; ++k) {
                ROW(l, 6 + k) = M[k];
                ROW(l, 42 + k) = G[k];
            }
        }
    }
    return ok;
}

__device__ __noinline__ double gen_solve_update(const GenEnv &E, bool ok, double lambda,
                                                const PoseBuf &Tc, const PoseBuf &Tn)
{
    const int N = E.tp->N, mod = E.cfg->orth_mod;
    double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double scale = 0.0;
    const bool tree = E.tp->tree != 0;
    for (int i = 0; i < N; ++i) {
        if (tree) { /* x_i = c_i - M_i x_{parent(i)}; x kept in the L records */
            double *lp = E.p.LR + (size_t)i * LR_TREE * TILE;
            const int par = __ldg(E.tp->parent + i);
            double l[LR_GEN];
#pragma unroll
            for (int k = 0; k < LR_GEN; ++k) l[k] = ROW(lp, k);
            if (par >= 0) {
                const double *pp = E.p.LR + (size_t)par * LR_TREE * TILE;
#pragma unroll
                for (int k = 0; k < 6; ++k) xp[k] = ROW(pp, 84 + k);
            }
            subst_step<6>(l, par >= 0, xp);
#pragma unroll
            for (int k = 0; k < 6; ++k) ROW(lp, 84 + k) = ok ? xp[k] : 0.0;
        } else {
            const double *lp = E.p.LR + (size_t)i * LR_GEN * TILE;
            double l[LR_GEN];
#pragma unroll
            for (int k = 0; k < LR_GEN; ++k) l[k] = ROW(lp, k);
            subst_step<6>(l, i > 0, xp);
        }
        if (!ok) {
#pragma unroll
            for (int k = 0; k < 6; ++k) xp[k] = 0.0;
        }
        const double *h = E.p.HB + (size_t)i * HR_GEN * TILE;
#pragma unroll
        for (int k = 0; k < 6; ++k) scale = scale + xp[k] * (lambda * xp[k] + ROW(h, 57 + k));
        Pose X;
        load_pose(Tc, i, X);
        int c = E.p.cnt[(size_t)i * TILE];
        pose_oplus(X, xp, c, mod);
        E.p.cnt[(size_t)i * TILE] = c;
        store_pose(Tn, i, X);
    }
    return scale;
}

/* ------------------------------------------------------------------------------------------ */
/* optimize(iteration_max) with OptimizationAlgorithmLevenberg, one window per thread           */
/* ------------------------------------------------------------------------------------------ */
/* MODE 0: GENERAL, 1: FAST (table-driven), 2: CHAIN */
template <int MODE>
struct Path;

template <>
struct Path<2> {
    FastEnv E;
    UWBGO_DI void chi(int buf, double &p, double &r) const { fast_chi_pass(E, E.p.T(buf), p, r); }
    UWBGO_DI double linearize(int buf) const { return fast_linearize<false>(E, E.p.T(buf)); }
    UWBGO_DI bool trial(double lambda, int from, int to, double &scale, double &p, double &r) const
    {
        bool ok = chain_factor(E, E.p.T(from), lambda) && (lambda > 0.0);
        chain_solve_chi(E, ok, lambda, E.p.T(from), E.p.T(to), scale, p, r);
        return ok;
    }
};

template <>
struct Path<1> {
    FastEnv E;
    UWBGO_DI void chi(int buf, double &p, double &r) const { fast_chi_pass(E, E.p.T(buf), p, r); }
    UWBGO_DI double linearize(int buf) const { return fast_linearize<false>(E, E.p.T(buf)); }
    /* factor + substitution + update + residuals of one LM trial */
    UWBGO_DI bool trial(double lambda, int from, int to, double &scale, double &p, double &r) const
    {
        /* the rotation pivots of the 6x6 blocks are exactly lambda */
        bool ok = fast_factor_mf(E, E.p.T(from), lambda) && (lambda > 0.0);
        fast_solve_chi(E, ok, lambda, E.p.T(from), E.p.T(to), scale, p, r);
        return ok;
    }
};
template <>
struct Path<0> {
    GenEnv E;
    UWBGO_DI PoseBuf buf(int k) const { return PoseBuf{E.p.T(k), E.p.Rm(k)}; }
    UWBGO_DI void chi(int k, double &p, double &r) const { gen_chi_pass(E, buf(k), p, r); }
    UWBGO_DI double linearize(int k) const { return gen_linearize(E, buf(k)); }
    UWBGO_DI bool trial(double lambda, int from, int to, double &scale, double &p, double &r) const
    {
        bool ok = E.tp->tree ? factor_sweep_tree(*E.tp, E.p.HB, E.p.LR, lambda)
                             : factor_sweep<6>(E.p.HB, E.p.LR, E.tp->N, lambda);
        scale = gen_solve_update(E, ok, lambda, buf(from), buf(to));
        gen_chi_pass(E, buf(to), p, r);
        return ok;
    }
};

/* g2o's optimize() is a loop over iterations, each holding a loop over LM trials.  The 32
 * windows of a warp reject different numbers of trials, so that nesting would leave lanes idle
 * while their neighbours retry.  It is flattened into ONE loop whose body is "linearise if the
 * last trial ended an iteration, then run one trial": every pass of a warp does useful trial work
 * on every unfinished lane, and only the (cheaper) linearisation runs on a subset of lanes.  The
 * per-window arithmetic and its order are unchanged. */
template <int MODE>
UWBGO_DI void lm_window(const Path<MODE> &P, const DevCfg &cfg, double *chi2_out,
                        int32_t *status_out, int &cur_out)
{
    constexpr bool FAST = MODE != 0;
    double lambda = 0.0, ni = 2.0, stale, plainCur, currentChi, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0;
    P.chi(cur, plainCur, currentChi);
    stale = plainCur;
    bool need_lin = true, done = cfg.max_iterations <= 0;
    while (!done) {
        if (need_lin) {
            stale = plainCur; /* computeActiveErrors at unchanged estimates */
            /* GENERAL: buildSystem() into the H records.  FAST: H is rebuilt inside every trial's
             * factor sweep; only computeLambdaInit() needs a pass of its own. */
            if (!FAST || it == 0) {
                double maxdiag = P.linearize(cur);
                if (it == 0) {
                    lambda = cfg.tau * maxdiag;
                    ni = 2.0;
                }
            }
            rho = 0.0;
            q = 0;
            need_lin = false;
        }
        double scale, tplain, tempChi;
        const bool ok = P.trial(lambda, cur, cur ^ 1, scale, tplain, tempChi);
        if (!ok) flags |= UWBGO_FLAG_CHOL_FAIL;
        stale = tplain;
        if (!ok) tempChi = DBL_MAX;
        scale = scale + 1e-3;
        rho = (currentChi - tempChi) / scale;
        const bool fin = isfinite(tempChi);
        if (!fin) flags |= UWBGO_FLAG_NONFINITE;
        if (rho > 0.0 && fin) {
            double t = 2.0 * rho - 1.0;
            double alpha = 1.0 - (t * t) * t;
            alpha = (cfg.good_hi < alpha) ? cfg.good_hi : alpha;
            double sf = (cfg.good_lo < alpha) ? alpha : cfg.good_lo;
            lambda = lambda * sf;
            ni = 2.0;
            currentChi = tempChi;
            plainCur = tplain;
            cur ^= 1;
        } else {
            lambda = lambda * ni;
            ni = ni * 2.0;
        }
        ++q;
        ++trials_total;
        if (!(rho < 0.0 && q < cfg.max_trials)) { /* this iteration is over */
            ++iterations;
            qlast = q;
            if (q == cfg.max_trials || rho == 0.0) {
                flags |= UWBGO_FLAG_TERMINATED;
                done = true;
            } else if (++it >= cfg.max_iterations) {
                done = true;
            } else {
                need_lin = true;
            }
        }
    }
    ROW(chi2_out, 0) = plainCur;
    ROW(chi2_out, 1) = currentChi;
    ROW(chi2_out, 2) = stale;
    ROW(chi2_out, 3) = lambda;
    ROW(status_out, 0) = iterations;
    ROW(status_out, 1) = trials_total;
    ROW(status_out, 2) = flags;
    ROW(status_out, 3) = qlast;
    cur_out = cur;
}

/* dynamic shared memory of the FAST kernels: carry stash, then (when they fit) the anchors */
constexpr int STASH_PER_THREAD = 2 * FAST_MAX_CARRY * 5;

UWBGO_DI void fast_env_init(FastEnv &E, const DevTopo &tp, const DevCfg &cfg, const DevWs &ws,
                            int64_t w, double *smem, int anchors_in_smem)
{
    E.tp = &tp;
    E.p = thread_ptrs<HR_FAST, LR_FAST>(tp, ws, w);
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
    E.bs = (int)blockDim.x;
    E.stash = smem + threadIdx.x;
    if (anchors_in_smem) {
        double *a = smem + (size_t)STASH_PER_THREAD * blockDim.x + threadIdx.x;
        for (int k = 0; k < tp.A * 3; ++k) a[(size_t)k * blockDim.x] = ROW(E.p.anch, k);
        E.anch = a;
        E.anch_stride = (int)blockDim.x;
    } else {
        E.anch = E.p.anch;
        E.anch_stride = TILE;
    }
}

template <int MODE>
UWBGO_DI void lm_fast_body(const DevTopo &tp, const DevCfg &cfg, const DevWs &ws, int anchors_in_smem,
                           double *smem)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= ws.W) return;
    Path<MODE> P;
    fast_env_init(P.E, tp, cfg, ws, w, smem, anchors_in_smem);
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    int cur;
    lm_window<MODE>(P, cfg, ws.chi2 + tile * 4 * TILE + lane, ws.status + tile * 4 * TILE + lane, cur);
    if (cur) { /* result always leaves in buffer 0 */
        for (int r = 0; r < tp.N * 3; ++r) ROW(P.E.p.T0, r) = ROW(P.E.p.T1, r);
    }
    if (P.E.p.cnt) { /* VertexSE3::_numOplusCalls: 12 per range edge end per linearisation, 1 per trial */
        const int it = ROW(ws.status + tile * 4 * TILE + lane, 0), tr = ROW(ws.status + tile * 4 * TILE + lane, 1);
        for (int i = 0; i < tp.N; ++i) {
            long long c = (long long)P.E.p.cnt[(size_t)i * TILE] + (long long)it * __ldg(tp.num_calls + i) + tr;
            P.E.p.cnt[(size_t)i * TILE] = (int)(c % cfg.orth_mod);
        }
    }
}

#ifndef UWBGO_FAST_MINB
#define UWBGO_FAST_MINB 4
#endif
__global__ void __launch_bounds__(CTA_THREADS, UWBGO_FAST_MINB)
lm_fast_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
               const __grid_constant__ DevWs ws, int anchors_in_smem)
{
    extern __shared__ double smem[];
    lm_fast_body<1>(tp, cfg, ws, anchors_in_smem, smem);
}

#ifndef UWBGO_CHAIN_THREADS
#define UWBGO_CHAIN_THREADS 64
#define UWBGO_CHAIN_REGS 255
#endif
__global__ void __maxnreg__(UWBGO_CHAIN_REGS)
lm_chain_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                const __grid_constant__ DevWs ws, int anchors_in_smem)
{
    extern __shared__ double smem[];
    lm_fast_body<2>(tp, cfg, ws, anchors_in_smem, smem);
}

/* ------------------------------------------------------------------------------------------ */
/* CHAIN path, warp-specialised: one CTA = one tile of 32 windows = two warps.                   */
/*   warp 0 (P, "edge warp")   linearises the trajectory edge (i-1, i) of every pose -- both      */
/*                             numeric Jacobians, weights -- and, in the substitution phase,     */
/*                             evaluates all residuals / chi2 at the new estimates;              */
/*   warp 1 (C, "chain warp")  linearises the anchor edge, assembles the H record, runs the      */
/*                             elimination chain and the substitution, owns the LM state.        */
/* The two halves of a pose's work are about the same number of instructions, so the dependency  */
/* chains of twice as many warps are in flight per scheduler for the same register budget per    */
/* window.  The warps move in lockstep, one named barrier per pose and phase, handing 8 (edge     */
/* terms) resp. 3 (new estimate) doubles per window through shared memory.  Per-window arithmetic */
/* and its order are those of the single-warp CHAIN path: same bits.                             */
/* ------------------------------------------------------------------------------------------ */
struct WsShared {
    double traj[2][8][TILE]; /* P -> C: A(3), B(3), Ow, omega_r of edge (i-1, i), double buffered */
    double tnew[2][3][TILE]; /* C -> P: new estimate of pose i                                    */
    double chi[2][TILE];     /* P -> C: plain and robust chi2 of the trial                        */
    int cur[TILE];           /* C -> P: which pose buffer is current                              */
    int act[TILE];           /* C -> P: window still being optimised                              */
    double stash[STASH_PER_THREAD][TILE];
};

UWBGO_DI void ws_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

#ifndef UWBGO_WS_MINB
#define UWBGO_WS_MINB 8
#endif
__global__ void __launch_bounds__(64, UWBGO_WS_MINB)
lm_chain_ws_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                   const __grid_constant__ DevWs ws)
{
    __shared__ WsShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wreal = (int64_t)blockIdx.x * TILE + lane;
    const bool valid = wreal < ws.W;
    const int64_t w = wreal; /* tail lanes own zero-filled pad columns of the tile layout: harmless garbage */
    const int N = tp.N;
    FastEnv E;
    E.tp = &tp;
    E.p = thread_ptrs<HR_FAST, LR_FAST>(tp, ws, w);
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
    E.bs = TILE;
    E.stash = &sh.stash[0][lane];
    E.anch = E.p.anch;
    E.anch_stride = TILE;

    /* LM state, chain warp only */
    double lambda = 0.0, ni = 2.0, stale = 0.0, plainCur = 0.0, currentChi = 0.0, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0;
    bool done = !valid || cfg.max_iterations <= 0;
    if (warp == 1) {
        fast_chi_pass(E, E.p.T0, plainCur, currentChi);
        stale = plainCur;
        const double maxdiag = fast_linearize<false>(E, E.p.T0);
        if (cfg.max_iterations > 0) lambda = cfg.tau * maxdiag;
        sh.cur[lane] = 0;
        sh.act[lane] = done ? 0 : 1;
    }
    ws_barrier();
    for (;;) {
        const bool act = sh.act[lane] != 0;
        const int c = sh.cur[lane];
        if (__ballot_sync(0xffffffffu, act) == 0u) break;
        ws_barrier(); /* everybody has read act/cur before the chain warp may overwrite them */
        const double *Tc = E.p.T(c);
        double *Tn = E.p.T(c ^ 1);
        if (warp == 0) {
            /* ---------------- edge warp, factor phase: edges (i-1, i), i = N-1 .. 1 ---------------- */
            {
                double cx = 0.0, cy = 0.0, cz = 0.0, px = 0.0, py = 0.0, pz = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
                double dt = 0.0, it_ = 0.0, ndt = 0.0, nit = 0.0;
                int rob = 0, nrob = 0;
                {
                    const double *tl = Tc + (size_t)(N - 1) * 3 * TILE;
                    cx = ROW(tl, 0); cy = ROW(tl, 1); cz = ROW(tl, 2);
                    if (N > 1) {
                        const double *tq = tl - (size_t)3 * TILE;
                        px = ROW(tq, 0); py = ROW(tq, 1); pz = ROW(tq, 2);
                        dt = ROW(E.p.rd, 2 * (N - 1));
                        it_ = ROW(E.p.ri, 2 * (N - 1));
                        rob = __ldg(&tp.chain[N - 1].robust);
                    }
                }
                for (int k = 0; k <= N; ++k) {
                    const int i = N - 1 - k; /* edge (i-1, i) */
                    if (i >= 1) {
                        if (i >= 2) { /* inputs of the next edge go in flight */
                            const double *tf = Tc + (size_t)(i - 2) * 3 * TILE;
                            fx = ROW(tf, 0); fy = ROW(tf, 1); fz = ROW(tf, 2);
                            ndt = ROW(E.p.rd, 2 * (i - 1));
                            nit = ROW(E.p.ri, 2 * (i - 1));
                            nrob = __ldg(&tp.chain[i - 1].robust);
                        }
                        if (UWBGO_L2PF_DIST > 0 && i >= 2 + 2 * UWBGO_L2PF_DIST) {
                            const int j = i - 2 - 2 * UWBGO_L2PF_DIST;
                            prefetch_rows_l2<3>(Tc + (size_t)j * 3 * TILE);
                            prefetch_l2(E.p.rd + (size_t)(2 * j) * TILE);
                            prefetch_l2(E.p.ri + (size_t)(2 * j) * TILE);
                        }
                        double A[3], B[3], Ow, omega_r;
                        const double err = dt - dist3(px, py, pz, cx, cy, cz);
                        fast_jac_v0(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, A);
                        fast_jac_v1(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, B);
                        chain_weights(E, err, it_, (rob & 2) != 0, Ow, omega_r);
                        double(*o)[TILE] = sh.traj[k & 1];
                        o[0][lane] = A[0]; o[1][lane] = A[1]; o[2][lane] = A[2];
                        o[3][lane] = B[0]; o[4][lane] = B[1]; o[5][lane] = B[2];
                        o[6][lane] = Ow;   o[7][lane] = omega_r;
                        cx = px; cy = py; cz = pz;
                        px = fx; py = fy; pz = fz;
                        dt = ndt; it_ = nit; rob = nrob;
                    }
                    ws_barrier();
                }
            }
            /* ---------------- edge warp, substitution phase: chi2 at the new estimates ---------------- */
            {
                double p = 0.0, r = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
                double da = ROW(E.p.rd, 0), ia = ROW(E.p.ri, 0), dt = 0.0, it_ = 0.0;
                int2 tb = __ldg(reinterpret_cast<const int2 *>(tp.chain));
                double qx = ANCH(E, tb.x * 3), qy = ANCH(E, tb.x * 3 + 1), qz = ANCH(E, tb.x * 3 + 2);
                for (int k = 0; k <= N; ++k) {
                    if (k >= 1) {
                        const int i = k - 1; /* pose i was produced in the previous step */
                        double nda = 0.0, nia = 0.0, ndt = 0.0, nit = 0.0, nqx = 0.0, nqy = 0.0, nqz = 0.0;
                        int2 ntb = tb;
                        if (i + 1 < N) {
                            ntb = __ldg(reinterpret_cast<const int2 *>(tp.chain + i + 1));
                            nda = ROW(E.p.rd, 2 * i + 1);
                            nia = ROW(E.p.ri, 2 * i + 1);
                            ndt = ROW(E.p.rd, 2 * i + 2);
                            nit = ROW(E.p.ri, 2 * i + 2);
                            nqx = ANCH(E, ntb.x * 3); nqy = ANCH(E, ntb.x * 3 + 1); nqz = ANCH(E, ntb.x * 3 + 2);
                        }
                        if (UWBGO_L2PF_DIST > 0 && i + 1 + UWBGO_L2PF_DIST < N) {
                            const int j = i + 1 + UWBGO_L2PF_DIST;
                            prefetch_l2(E.p.rd + (size_t)(2 * j) * TILE);
                            prefetch_l2(E.p.ri + (size_t)(2 * j) * TILE);
                            prefetch_l2(E.p.rd + (size_t)(2 * j - 1) * TILE);
                            prefetch_l2(E.p.ri + (size_t)(2 * j - 1) * TILE);
                        }
                        const double(*tn)[TILE] = sh.tnew[i & 1];
                        const double cx = tn[0][lane], cy = tn[1][lane], cz = tn[2][lane];
                        {
                            const double err = da - dist3(cx, cy, cz, qx, qy, qz);
                            const double chi = err * (ia * err);
                            p = p + chi;
                            r = r + ((tb.y & 1) ? E.ck.rho0(chi) : chi);
                        }
                        if (i > 0) {
                            const double err = dt - dist3(vx, vy, vz, cx, cy, cz);
                            const double chi = err * (it_ * err);
                            p = p + chi;
                            r = r + ((tb.y & 2) ? E.ck.rho0(chi) : chi);
                        }
                        vx = cx; vy = cy; vz = cz;
                        da = nda; ia = nia; dt = ndt; it_ = nit; qx = nqx; qy = nqy; qz = nqz; tb = ntb;
                    }
                    ws_barrier();
                }
                sh.chi[0][lane] = p;
                sh.chi[1][lane] = r;
            }
            ws_barrier(); /* chi2 published */
            ws_barrier(); /* LM state published */
        } else {
            /* ---------------- chain warp, factor phase: poses i = N-1 .. 0 ---------------- */
            bool ok = true;
            {
                double G[9], zn[3], carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < 9; ++k) G[k] = 0.0;
#pragma unroll
                for (int k = 0; k < 3; ++k) zn[k] = 0.0;
                /* inputs of the anchor edge of pose i, one pose ahead */
                double cx, cy, cz, da, ia, qx, qy, qz;
                int rob;
                {
                    const int i = N - 1;
                    const double *tl = Tc + (size_t)i * 3 * TILE;
                    cx = ROW(tl, 0); cy = ROW(tl, 1); cz = ROW(tl, 2);
                    const int2 tb = __ldg(reinterpret_cast<const int2 *>(tp.chain + i));
                    const int sa = i == 0 ? 0 : 2 * i - 1;
                    da = ROW(E.p.rd, sa); ia = ROW(E.p.ri, sa);
                    qx = ANCH(E, tb.x * 3); qy = ANCH(E, tb.x * 3 + 1); qz = ANCH(E, tb.x * 3 + 2);
                    rob = tb.y;
                }
                ws_barrier(); /* step 0: the edge warp produces edge (N-2, N-1) */
                for (int k = 1; k <= N; ++k) {
                    const int i = N - k;
                    double ncx = 0.0, ncy = 0.0, ncz = 0.0, nda = 0.0, nia = 0.0, nqx = 0.0, nqy = 0.0, nqz = 0.0;
                    int nrob = 0;
                    if (i >= 1) {
                        const int j = i - 1;
                        const double *tl = Tc + (size_t)j * 3 * TILE;
                        ncx = ROW(tl, 0); ncy = ROW(tl, 1); ncz = ROW(tl, 2);
                        const int2 tb = __ldg(reinterpret_cast<const int2 *>(tp.chain + j));
                        const int sa = j == 0 ? 0 : 2 * j - 1;
                        nda = ROW(E.p.rd, sa); nia = ROW(E.p.ri, sa);
                        nqx = ANCH(E, tb.x * 3); nqy = ANCH(E, tb.x * 3 + 1); nqz = ANCH(E, tb.x * 3 + 2);
                        nrob = tb.y;
                    }
                    if (UWBGO_L2PF_DIST > 0 && i >= 1 + 2 * UWBGO_L2PF_DIST) {
                        const int j = i - 1 - 2 * UWBGO_L2PF_DIST;
                        prefetch_rows_l2<3>(Tc + (size_t)j * 3 * TILE);
                        prefetch_l2(E.p.rd + (size_t)(j == 0 ? 0 : 2 * j - 1) * TILE);
                        prefetch_l2(E.p.ri + (size_t)(j == 0 ? 0 : 2 * j - 1) * TILE);
                    }
                    double h[HR_FAST];
#pragma unroll
                    for (int m = 0; m < HR_FAST; ++m) h[m] = 0.0;
                    {
                        double J[3], Ow, omega_r;
                        const double err = da - dist3(cx, cy, cz, qx, qy, qz);
                        fast_jac_v0(cx, cy, cz, qx, qy, qz, da, E.delta, E.scalar, J);
                        chain_weights(E, err, ia, (rob & 1) != 0, Ow, omega_r);
                        chain_acc(J, Ow, omega_r, h);
                    }
                    double nA[3] = {0.0, 0.0, 0.0}, nOw = 0.0, nOr = 0.0;
                    if (i >= 1) { /* edge (i-1, i), linearised by the edge warp in step k-1 */
                        const double(*t)[TILE] = sh.traj[(k - 1) & 1];
                        nA[0] = t[0][lane]; nA[1] = t[1][lane]; nA[2] = t[2][lane];
                        double B[3] = {t[3][lane], t[4][lane], t[5][lane]};
                        nOw = t[6][lane];
                        nOr = t[7][lane];
                        const double AtO[3] = {nA[0] * nOw, nA[1] * nOw, nA[2] * nOw};
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) h[6 + 3 * r + cc] = fma(AtO[r], B[cc], h[6 + 3 * r + cc]);
                        chain_acc(B, nOw, nOr, h);
                    }
                    chain_acc(carry, carry[3], carry[4], h);
                    carry[0] = nA[0]; carry[1] = nA[1]; carry[2] = nA[2]; carry[3] = nOw; carry[4] = nOr;
                    /* tail / finished lanes write to a scratch record of their own window: harmless */
                    double *l = E.p.LR + (size_t)i * LR_FAST * TILE;
                    factor_step<3>(h, l, true, i > 0, lambda, G, zn, ok);
                    chain_store_b(l, h);
                    cx = ncx; cy = ncy; cz = ncz; da = nda; ia = nia; qx = nqx; qy = nqy; qz = nqz; rob = nrob;
                    ws_barrier();
                }
                ok = ok && (lambda > 0.0);
            }
            /* ---------------- chain warp, substitution phase ---------------- */
            double scale = 0.0;
            {
                double xp[3] = {0.0, 0.0, 0.0};
                double l[LR_FAST], nl[LR_FAST], t[3], nt[3];
#pragma unroll
                for (int m = 0; m < LR_FAST; ++m) nl[m] = ROW(E.p.LR, m);
#pragma unroll
                for (int m = 0; m < 3; ++m) nt[m] = ROW(Tc, m);
                for (int k = 0; k <= N; ++k) {
                    if (k < N) {
                        const int i = k;
#pragma unroll
                        for (int m = 0; m < LR_FAST; ++m) l[m] = nl[m];
#pragma unroll
                        for (int m = 0; m < 3; ++m) t[m] = nt[m];
                        if (i + 1 < N) {
                            const double *ln = E.p.LR + (size_t)(i + 1) * LR_FAST * TILE;
                            const double *tn = Tc + (size_t)(i + 1) * 3 * TILE;
#pragma unroll
                            for (int m = 0; m < LR_FAST; ++m) nl[m] = ROW(ln, m);
#pragma unroll
                            for (int m = 0; m < 3; ++m) nt[m] = ROW(tn, m);
                        }
                        if (UWBGO_L2PF_DIST > 0 && i + 1 + UWBGO_L2PF_DIST < N) {
                            prefetch_rows_l2<LR_FAST>(E.p.LR + (size_t)(i + 1 + UWBGO_L2PF_DIST) * LR_FAST * TILE);
                            prefetch_rows_l2<3>(Tc + (size_t)(i + 1 + UWBGO_L2PF_DIST) * 3 * TILE);
                        }
                        subst_step<3>(l, i > 0, xp);
                        if (!ok) xp[0] = xp[1] = xp[2] = 0.0;
#pragma unroll
                        for (int m = 0; m < 3; ++m) scale = scale + xp[m] * (lambda * xp[m] + l[12 + m]);
                        const double nx = xp[0] + t[0], ny = xp[1] + t[1], nz = xp[2] + t[2];
                        double(*o)[TILE] = sh.tnew[i & 1];
                        o[0][lane] = nx; o[1][lane] = ny; o[2][lane] = nz;
                        if (act) {
                            double *to = Tn + (size_t)i * 3 * TILE;
                            ROW(to, 0) = nx; ROW(to, 1) = ny; ROW(to, 2) = nz;
                        }
                    }
                    ws_barrier();
                }
            }
            ws_barrier(); /* chi2 published by the edge warp */
            if (act) {
                const double tplain = sh.chi[0][lane];
                double tempChi = sh.chi[1][lane];
                if (!ok) flags |= UWBGO_FLAG_CHOL_FAIL;
                stale = tplain;
                if (!ok) tempChi = DBL_MAX;
                scale = scale + 1e-3;
                rho = (currentChi - tempChi) / scale;
                const bool fin = isfinite(tempChi);
                if (!fin) flags |= UWBGO_FLAG_NONFINITE;
                if (rho > 0.0 && fin) {
                    double tt = 2.0 * rho - 1.0;
                    double alpha = 1.0 - (tt * tt) * tt;
                    alpha = (cfg.good_hi < alpha) ? cfg.good_hi : alpha;
                    double sf = (cfg.good_lo < alpha) ? alpha : cfg.good_lo;
                    lambda = lambda * sf;
                    ni = 2.0;
                    currentChi = tempChi;
                    plainCur = tplain;
                    cur ^= 1;
                } else {
                    lambda = lambda * ni;
                    ni = ni * 2.0;
                }
                ++q;
                ++trials_total;
                if (!(rho < 0.0 && q < cfg.max_trials)) {
                    ++iterations;
                    qlast = q;
                    if (q == cfg.max_trials || rho == 0.0) {
                        flags |= UWBGO_FLAG_TERMINATED;
                        done = true;
                    } else if (++it >= cfg.max_iterations) {
                        done = true;
                    }
                    rho = 0.0;
                    q = 0;
                }
                sh.cur[lane] = cur;
                sh.act[lane] = done ? 0 : 1;
            }
            ws_barrier(); /* LM state published */
        }
    }
    if (warp == 1 && valid) {
        const int64_t tile = wreal / TILE;
        double *chi2_out = ws.chi2 + tile * 4 * TILE + lane;
        int32_t *status_out = ws.status + tile * 4 * TILE + lane;
        ROW(chi2_out, 0) = plainCur;
        ROW(chi2_out, 1) = currentChi;
        ROW(chi2_out, 2) = stale;
        ROW(chi2_out, 3) = lambda;
        ROW(status_out, 0) = iterations;
        ROW(status_out, 1) = trials_total;
        ROW(status_out, 2) = flags;
        ROW(status_out, 3) = qlast;
        if (cur) {
            for (int r = 0; r < N * 3; ++r) ROW(E.p.T0, r) = ROW(E.p.T1, r);
        }
        if (E.p.cnt) {
            for (int i = 0; i < N; ++i) {
                long long cc = (long long)E.p.cnt[(size_t)i * TILE] + (long long)iterations * __ldg(tp.num_calls + i) + trials_total;
                E.p.cnt[(size_t)i * TILE] = (int)(cc % cfg.orth_mod);
            }
        }
    }
}

UWBGO_DI void gen_env_init(GenEnv &E, const DevTopo &tp, const DevCfg &cfg, const DevWs &ws, int64_t w)
{
    E.tp = &tp;
    E.cfg = &cfg;
    E.p = thread_ptrs<HR_GEN, LR_GEN>(tp, ws, w);
    E.ant = ws.ant;
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
}

__global__ void __launch_bounds__(CTA_THREADS)
lm_general_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                  const __grid_constant__ DevWs ws)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    Path<0> P;
    gen_env_init(P.E, tp, cfg, ws, w);
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    int cur;
    lm_window<0>(P, cfg, ws.chi2 + tile * 4 * TILE + lane, ws.status + tile * 4 * TILE + lane, cur);
    if (cur) {
        for (int r = 0; r < tp.N * 3; ++r) ROW(P.E.p.T0, r) = ROW(P.E.p.T1, r);
        for (int r = 0; r < tp.N * 9; ++r) ROW(P.E.p.Rm0, r) = ROW(P.E.p.Rm1, r);
    }
}

/* one linearisation: computeActiveErrors + buildSystem; chi2 = {plain, robust} */
__global__ void __launch_bounds__(CTA_THREADS, 4)
linearize_fast_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                      const __grid_constant__ DevWs ws, int anchors_in_smem)
{
    extern __shared__ double smem[];
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    FastEnv E;
    fast_env_init(E, tp, cfg, ws, w, smem, anchors_in_smem);
    double p, r;
    fast_chi_pass(E, E.p.T0, p, r);
    fast_linearize<true>(E, E.p.T0);
    double *c = ws.chi2 + (w / TILE) * 2 * TILE + (w % TILE);
    ROW(c, 0) = p;
    ROW(c, 1) = r;
}

__global__ void __launch_bounds__(CTA_THREADS)
linearize_general_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                         const __grid_constant__ DevWs ws)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    GenEnv E;
    gen_env_init(E, tp, cfg, ws, w);
    double p, r;
    PoseBuf T0{E.p.T0, E.p.Rm0};
    gen_chi_pass(E, T0, p, r);
    gen_linearize(E, T0);
    double *c = ws.chi2 + (w / TILE) * 2 * TILE + (w % TILE);
    ROW(c, 0) = p;
    ROW(c, 1) = r;
}

/* ------------------------------------------------------------------------------------------ */
/* layout kernels                                                                               */
/* ------------------------------------------------------------------------------------------ */
/* window-major [W][C] -> tile layout [tile][C][32] (pack) and back (unpack), through a padded
 * shared-memory tile so both sides are coalesced.  grid = (tiles, column chunks of 32, jobs),
 * block = (32, 8). */
template <typename T, bool PACK>
UWBGO_DI void xpose_body(const XposeJob &j, int64_t W, T (*sm)[33])
{
    const int64_t tile = blockIdx.x;
    const int c0 = blockIdx.y * 32;
    if (c0 >= j.C) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T *src = static_cast<const T *>(j.src);
    T *dst = static_cast<T *>(j.dst);
    if (PACK) {
        /* read: x runs along columns of one window */
        for (int y = ty; y < 32; y += 8) {
            int64_t w = tile * TILE + y;
            int c = c0 + tx;
            T v = T(0);
            if (w < W && c < j.C) v = src[w * j.C + c];
            sm[y][tx] = v;
        }
        __syncthreads();
        for (int y = ty; y < 32; y += 8) {
            int c = c0 + y;
            if (c < j.C) dst[(tile * j.C + c) * TILE + tx] = sm[tx][y];
        }
    } else {
        for (int y = ty; y < 32; y += 8) {
            int c = c0 + y;
            T v = T(0);
            if (c < j.C) v = src[(tile * j.C + c) * TILE + tx];
            sm[y][tx] = v;
        }
        __syncthreads();
        for (int y = ty; y < 32; y += 8) {
            int64_t w = tile * TILE + y;
            int c = c0 + tx;
            if (w < W && c < j.C) dst[w * j.C + c] = sm[tx][y];
        }
    }
}

template <bool PACK>
__global__ void __launch_bounds__(256) xpose_kernel(XposeJobs jobs)
{
    __shared__ double sm[32][33];
    const XposeJob &j = jobs.job[blockIdx.z];
    if (j.mode == 1) { /* identity rotations into a tile-layout [tile][C = N*9][32] array */
        const int64_t tile = blockIdx.x;
        const int c0 = blockIdx.y * 32;
        for (int y = threadIdx.y; y < 32; y += 8) {
            int c = c0 + y;
            if (c >= j.C) continue;
            int k = c % 9;
            static_cast<double *>(j.dst)[(tile * j.C + c) * TILE + threadIdx.x] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
        }
        return;
    }
    if (j.mode == 2) { /* identity rotations into a window-major [W][C = N*9] array */
        const int64_t tile = blockIdx.x;
        const int c0 = blockIdx.y * 32;
        for (int y = threadIdx.y; y < 32; y += 8) {
            int64_t w = tile * TILE + y;
            int c = c0 + threadIdx.x;
            if (w >= jobs.W || c >= j.C) continue;
            int k = c % 9;
            static_cast<double *>(j.dst)[w * j.C + c] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;
        }
        return;
    }
    if (j.mode == 3) { /* zero fill of a tile-layout int32/double array, C rows */
        const int64_t tile = blockIdx.x;
        const int c0 = blockIdx.y * 32;
        for (int y = threadIdx.y; y < 32; y += 8) {
            int c = c0 + y;
            if (c >= j.C) continue;
            if (j.elem == 8)
                static_cast<double *>(j.dst)[(tile * j.C + c) * TILE + threadIdx.x] = 0.0;
            else
                static_cast<int32_t *>(j.dst)[(tile * j.C + c) * TILE + threadIdx.x] = 0;
        }
        return;
    }
    if (j.elem == 8)
        xpose_body<double, PACK>(j, jobs.W, sm);
    else
        xpose_body<int32_t, PACK>(j, jobs.W, reinterpret_cast<int32_t(*)[33]>(sm));
}

static cudaError_t launch_xpose(const XposeJobs &jobs, bool pack, cudaStream_t st)
{
    if (jobs.n <= 0 || jobs.W <= 0) return cudaSuccess;
    int maxC = 0;
    for (int k = 0; k < jobs.n; ++k) maxC = jobs.job[k].C > maxC ? jobs.job[k].C : maxC;
    dim3 grid((unsigned)n_tiles(jobs.W), (unsigned)((maxC + 31) / 32), (unsigned)jobs.n);
    dim3 block(32, 8);
    if (pack)
        xpose_kernel<true><<<grid, block, 0, st>>>(jobs);
    else
        xpose_kernel<false><<<grid, block, 0, st>>>(jobs);
    return cudaGetLastError();
}
cudaError_t launch_pack(const XposeJobs &jobs, cudaStream_t st) { return launch_xpose(jobs, true, st); }
cudaError_t launch_unpack(const XposeJobs &jobs, cudaStream_t st) { return launch_xpose(jobs, false, st); }

static unsigned window_blocks(int64_t W) { return (unsigned)((W + CTA_THREADS - 1) / CTA_THREADS); }

/* dynamic shared memory of a FAST launch; anchors go to shared memory when 4 CTAs/SM still fit */
static size_t fast_smem_bytes(const DevTopo &topo, int threads, int *anchors_in_smem)
{
    size_t stash = sizeof(double) * STASH_PER_THREAD * threads;
    size_t anch = sizeof(double) * (size_t)topo.A * 3 * threads;
    /* budget: the CTAs that fill an SM (512 threads) must fit in its 227 KB */
    *anchors_in_smem = (topo.A > 0 && (stash + anch) * (512 / threads) <= 220 * 1024) ? 1 : 0;
    return stash + (*anchors_in_smem ? anch : 0);
}

cudaError_t launch_solve(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws, cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    if (topo.fast) {
#ifndef UWBGO_CHAIN_WS
#define UWBGO_CHAIN_WS 1 /* 1: warp-specialised CHAIN kernel, 0: single-warp CHAIN kernel */
#endif
        if (topo.fast == 2 && UWBGO_CHAIN_WS) {
            lm_chain_ws_kernel<<<(unsigned)n_tiles(ws.W), 64, 0, st>>>(topo, cfg, ws);
            return cudaGetLastError();
        }
        int ais = 0;
        const int threads = topo.fast == 2 ? UWBGO_CHAIN_THREADS : CTA_THREADS;
        size_t sm = fast_smem_bytes(topo, threads, &ais);
        auto kern = topo.fast == 2 ? lm_chain_kernel : lm_fast_kernel;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)((ws.W + threads - 1) / threads), threads, sm, st>>>(topo, cfg, ws, ais);
    } else
        lm_general_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, cfg, ws);
    return cudaGetLastError();
}

cudaError_t launch_linearize(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws,
                             cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    if (topo.fast) {
        int ais = 0;
        size_t sm = fast_smem_bytes(topo, CTA_THREADS, &ais);
        cudaError_t e = cudaFuncSetAttribute(linearize_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        linearize_fast_kernel<<<window_blocks(ws.W), CTA_THREADS, sm, st>>>(topo, cfg, ws, ais);
    } else
        linearize_general_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, cfg, ws);
    return cudaGetLastError();
}

/* H records (tile layout) -> public H_diag [W][N][36] (both triangles), H_off [W][N-1][36],
 * b [W][N][6].  One thread per (window, pose); reads coalesced, writes are 36-double runs. */
__global__ void __launch_bounds__(CTA_THREADS)
expand_H_kernel(DevTopo tp, DevWs ws, double *__restrict__ H_diag, double *__restrict__ H_off,
                double *__restrict__ b)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    const int i = blockIdx.y;
    if (w >= ws.W) return;
    const int N = tp.N;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    double *hd = H_diag + ((size_t)w * N + i) * 36;
    double *bo = b + ((size_t)w * N + i) * 6;
    if (tp.fast) {
        const double *h = ws.HB + ((tile * N + i) * (size_t)HR_FAST) * TILE + lane;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double v = 0.0;
                if (r < 3 && c < 3) v = r <= c ? ROW(h, up_idx(3, r, c)) : ROW(h, up_idx(3, c, r));
                hd[6 * r + c] = v;
            }
        for (int r = 0; r < 6; ++r) bo[r] = r < 3 ? ROW(h, 15 + r) : 0.0;
        if (i > 0) {
            double *ho = H_off + ((size_t)w * (N - 1) + (i - 1)) * 36;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) ho[6 * r + c] = (r < 3 && c < 3) ? ROW(h, 6 + 3 * r + c) : 0.0;
        }
    } else {
        const double *h = ws.HB + ((tile * N + i) * (size_t)HR_GEN) * TILE + lane;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c)
                hd[6 * r + c] = r <= c ? ROW(h, up_idx(6, r, c)) : ROW(h, up_idx(6, c, r));
        for (int r = 0; r < 6; ++r) bo[r] = ROW(h, 57 + r);
        if (i > 0) {
            double *ho = H_off + ((size_t)w * (N - 1) + (i - 1)) * 36;
            for (int k = 0; k < 36; ++k) ho[k] = ROW(h, 21 + k);
        }
    }
}

cudaError_t launch_expand_H(const DevTopo &topo, const DevWs &ws, double *H_diag, double *H_off,
                            double *b, cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    dim3 grid(window_blocks(ws.W), (unsigned)topo.N);
    expand_H_kernel<<<grid, CTA_THREADS, 0, st>>>(topo, ws, H_diag, H_off, b);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* stand-alone linear solve on the public window-major arrays (LinearSolverCholmod::solve)      */
/* ------------------------------------------------------------------------------------------ */
/* gather one window's H_diag/H_off/b into H records (tile layout), one thread per (window,pose) */
__global__ void __launch_bounds__(CTA_THREADS)
compress_H_kernel(int N, int64_t W, const double *__restrict__ H_diag,
                  const double *__restrict__ H_off, const double *__restrict__ b,
                  double *__restrict__ HB)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    const int i = blockIdx.y;
    if (w >= W) return;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    double *h = HB + ((tile * N + i) * (size_t)HR_GEN) * TILE + lane;
    const double *hd = H_diag + ((size_t)w * N + i) * 36;
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) ROW(h, up_idx(6, r, c)) = hd[6 * c + r]; /* lower triangle is read */
    const double *bo = b + ((size_t)w * N + i) * 6;
    for (int r = 0; r < 6; ++r) ROW(h, 57 + r) = bo[r];
    if (i > 0) {
        const double *ho = H_off + ((size_t)w * (N - 1) + (i - 1)) * 36;
        for (int k = 0; k < 36; ++k) ROW(h, 21 + k) = ho[k];
    }
}

__global__ void __launch_bounds__(CTA_THREADS)
factor_solve_kernel(int N, int64_t W, double *__restrict__ HB, double *__restrict__ LR,
                    const double *__restrict__ lambda, double *__restrict__ x,
                    int32_t *__restrict__ okv)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= W) return;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    double *hb = HB + (tile * (size_t)N * HR_GEN) * TILE + lane;
    double *lr = LR + (tile * (size_t)N * LR_GEN) * TILE + lane;
    bool ok = factor_sweep<6>(hb, lr, N, lambda[w]);
    double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = 0; i < N; ++i) {
        const double *lp = lr + (size_t)i * LR_GEN * TILE;
        double l[LR_GEN];
        for (int k = 0; k < LR_GEN; ++k) l[k] = ROW(lp, k);
        subst_step<6>(l, i > 0, xp);
        double *xo = x + ((size_t)w * N + i) * 6;
        for (int k = 0; k < 6; ++k) xo[k] = ok ? xp[k] : 0.0;
    }
    if (okv) okv[w] = ok ? 1 : 0;
}

size_t factor_solve_scratch_bytes(int32_t N, int64_t W)
{
    return 2 * (size_t)n_tiles(W) * TILE * (size_t)N * HR_GEN * sizeof(double);
}

cudaError_t launch_factor_solve(int32_t N, int64_t W, const double *H_diag, const double *H_off,
                                const double *b, const double *lambda, double *x, int32_t *ok,
                                double *scratch, cudaStream_t st)
{
    if (W <= 0) return cudaSuccess;
    double *HB = scratch;
    double *LR = scratch + (size_t)n_tiles(W) * TILE * (size_t)N * HR_GEN;
    dim3 grid(window_blocks(W), (unsigned)N);
    compress_H_kernel<<<grid, CTA_THREADS, 0, st>>>(N, W, H_diag, H_off, b, HB);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    factor_solve_kernel<<<window_blocks(W), CTA_THREADS, 0, st>>>(N, W, HB, LR, lambda, x, ok);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* FP64 FMA micro-benchmark: 8 independent chains per thread                                    */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0,
           a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-6;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t launch_fp64_peak(double *out, int iters, cudaStream_t st, int *blocks, int *threads)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *blocks = sms * 8;
    *threads = 256;
    fp64_peak_kernel<<<*blocks, *threads, 0, st>>>(out, iters);
    return cudaGetLastError();
}

}  // namespace uwbgo
