/*
 * uwbgo_window.cu — the WINDOW path: one CTA per window, every quantity of the window resident in
 * shared memory for the whole optimize(iteration_max) call.
 *
 * This is the shape of the reference's own call pattern: Localization::addRangeEdge ends in
 * solve() once per range message (reference src/localization/localization.cpp:164-192,371-375),
 * i.e. ONE window of 10-20 poses at a time, and small fleets of a few hundred windows.  The tile
 * kernels (lane = window, state in HBM) leave 31 of 32 lanes idle there and need three launches
 * plus fifteen copies; this kernel reads the window-major arrays of the public ABI directly
 * (device memory or mapped pinned host memory), so a host call is one launch.
 *
 * The parallelism is inside the window.  Per LM iteration, buildSystem is spread over the CTA:
 *   J   numeric Jacobians of the range edges: one thread per (edge, vertex, column) = two
 *       perturbed residuals each; analytic Jacobians of EdgeSE3Prior / EdgeSE3: one thread per edge
 *   O   J^T Omega of the 6-D edges: one thread per entry
 *   H   every entry of H_ii (upper), H_{parent(i),i} and b_i is OWNED by one thread, which walks the
 *       pose's edges in g2o insertion order and adds their terms: every entry sees exactly the
 *       accumulation order of the oracle, no atomics, no reductions
 * and per trial
 *   F   block elimination newest pose first on warp 0: the 21 + 6 entries of S_i and of the
 *       right-hand side are assembled by one lane each, every lane then runs the 6x6 potrf in
 *       registers (redundantly: the sqrt -> reciprocal chain is latency, not throughput), and lanes
 *       0..6 run the SAME triangular substitutions on seven right-hand sides: rows of
 *       H_{parent(i),i} give the rows of G_i and the columns of M_i, b_i gives z_i and c_i.
 *       Then the substitution x_i = c_i - M_i x_parent(i), one lane per row
 *   U   estimate (+) x_i, one thread per pose; computeScale() on the last warp meanwhile
 *   C   computeActiveErrors, one thread per edge; chi2 terms summed in insertion order by thread 0
 *   D   accept / reject, lambda update (thread 0)
 * Per-entry operation sequences are those of the tile kernels (uwbgo_general.cuh) and of the CPU checker:
 * the results are the same bits.
 */
#include "uwbgo_device.cuh"
#ifdef UWBGO_WIN_TIMING
#include <cstdio>
#endif

namespace uwbgo {

namespace {

#ifndef UWBGO_WIN_THREADS
#define UWBGO_WIN_THREADS 256
#endif
constexpr int WT = UWBGO_WIN_THREADS;
constexpr int WNW = WT / 32;

/* per-edge linearisation records (doubles) */
constexpr int RJ = 14;  /* range: A 6 | B 6 | Ow | omega_r                      */
constexpr int PJ = 79;  /* prior: J 36 | J^T Ow 36 | omega_r 6 | rho1           */
constexpr int SJ = 151; /* se3:   A 36 | B 36 | A^T Ow 36 | B^T Ow 36 | omega_r 6 | rho1 */

struct WinCarve {
    int X0, X1, anch, ant, rd, ri, pZi, pI, sZi, sI, rJ, pJ, sJ, Hd, Ho, b, G, M, c, z, x, S, echi, cnt, edges, total;
};

__host__ __device__ inline WinCarve win_carve(const DevTopo &t)
{
    WinCarve c;
    int o = 0;
    auto take = [&](int n) {
        int at = o;
        o += (n + 1) & ~1; /* keep 16-byte alignment */
        return at;
    };
    const int N = t.N;
    c.X0 = take(N * 12);
    c.X1 = take(N * 12);
    c.anch = take(t.A * 3);
    c.ant = take(t.K * 3);
    c.rd = take(t.Er);
    c.ri = take(t.Er);
    c.pZi = take(t.Ep * 12);
    c.pI = take(t.Ep * 36);
    c.sZi = take(t.Es * 12);
    c.sI = take(t.Es * 36);
    c.rJ = take(t.Er * RJ);
    c.pJ = take(t.Ep * PJ);
    c.sJ = take(t.Es * SJ);
    c.Hd = take(N * 21);
    c.Ho = take(N * 36);
    c.b = take(N * 6);
    c.G = take(N * 36);
    c.M = take(N * 36);
    c.c = take(N * 6);
    c.z = take(N * 6);
    c.x = take(N * 6);
    c.S = take(28);
    c.echi = take(t.E * 2);
    c.cnt = take((N + 1) / 2);                                  /* int32 [N]        */
    c.edges = take((int)(sizeof(EdgeRec) / 8) * (t.E > 0 ? t.E : 1)); /* EdgeRec [E]      */
    c.total = o;
    return c;
}

struct WinSm {
    double *X[2], *anch, *ant, *rd, *ri, *pZi, *pI, *sZi, *sI, *rJ, *pJ, *sJ, *Hd, *Ho, *b, *G, *M, *c, *z, *x, *S, *echi;
    int *cnt;
    EdgeRec *edges;
};

struct WinCtl {
    int go, lin, cur, ok;
    double lambda, scale;
};

UWBGO_DI void ld_pose(const double *p, Pose &X)
{
#pragma unroll
    for (int k = 0; k < 9; ++k) X.R[k] = p[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) X.t[k] = p[9 + k];
}
UWBGO_DI void st_pose(double *p, const Pose &X)
{
#pragma unroll
    for (int k = 0; k < 9; ++k) p[k] = X.R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[9 + k] = X.t[k];
}

/* (X * offset).translation() for a translation-only offset: R o + t; identity offset: t */
UWBGO_DI void offset_point(const WinSm &sm, const double *Xp, int ant, double *P)
{
    if (ant > 0) {
        const double *o = sm.ant + 3 * (ant - 1);
        mat3_vec_add(Xp, o, Xp + 9, P);
    } else {
        P[0] = Xp[9]; P[1] = Xp[10]; P[2] = Xp[11];
    }
}
/* vertex 1 of an anchor range edge: fixed identity-rotation vertex at the anchor, times its offset */
UWBGO_DI void anchor_point(const WinSm &sm, int b, int ant_b, double *Q)
{
    const double *an = sm.anch + 3 * b;
    if (ant_b > 0) {
        const double I3[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
        mat3_vec_add(I3, sm.ant + 3 * (ant_b - 1), an, Q);
    } else {
        Q[0] = an[0]; Q[1] = an[1]; Q[2] = an[2];
    }
}

/* both end points of range edge er at the estimates X */
UWBGO_DI void range_points(const WinSm &sm, const double *X, const EdgeRec &er, double *P0, double *Q)
{
    offset_point(sm, X + 12 * er.a, er.ant, P0);
    if (er.kind == UWBGO_EDGE_RANGE_ANCHOR)
        anchor_point(sm, er.b, er.ant_b, Q);
    else
        offset_point(sm, X + 12 * er.b, er.ant_b, Q);
}

UWBGO_DI double chi2_6s(const double *O, const double *e, double *Oe)
{
    double chi = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = O[6 * r] * e[0];
#pragma unroll
        for (int c = 1; c < 6; ++c) s = s + O[6 * r + c] * e[c];
        Oe[r] = s;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) chi = chi + e[r] * Oe[r];
    return chi;
}

/* toVectorMQT(Zinv * X) and the product itself (EdgeSE3Prior::computeError, identity offset) */
UWBGO_DI void prior_error(const double *Zi, const double *Xp, Pose &Dl, double *q, double *e6)
{
    Pose Zinv, X;
    ld_pose(Zi, Zinv);
    ld_pose(Xp, X);
    pose_mul(Zinv, X, Dl);
    R_to_quat(Dl.R, q);
    e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
    e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
}

/* toVectorMQT(Zinv * Xi^-1 * Xj), evaluated left to right (EdgeSE3::computeError) */
UWBGO_DI void se3_error_s(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *e)
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

UWBGO_DI void quat_left_s(const double *q, double *M)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    M[0] = w;  M[1] = -x; M[2] = -y;  M[3] = -z;
    M[4] = x;  M[5] = w;  M[6] = -z;  M[7] = y;
    M[8] = y;  M[9] = z;  M[10] = w;  M[11] = -x;
    M[12] = z; M[13] = -y; M[14] = x; M[15] = w;
}
UWBGO_DI void quat_right_s(const double *q, double *M)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    M[0] = w;  M[1] = -x; M[2] = -y;  M[3] = -z;
    M[4] = x;  M[5] = w;  M[6] = z;   M[7] = -y;
    M[8] = y;  M[9] = -z; M[10] = w;  M[11] = x;
    M[12] = z; M[13] = y; M[14] = -x; M[15] = w;
}
/* d(vector part of qE (x) dq)/d(dq) = w I + [q]x into block (3,3) of the 6x6 J */
UWBGO_DI void set_jqq_s(const double *q, double *J)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    J[6 * 3 + 3] = w;  J[6 * 3 + 4] = -z; J[6 * 3 + 5] = y;
    J[6 * 4 + 3] = z;  J[6 * 4 + 4] = w;  J[6 * 4 + 5] = -x;
    J[6 * 5 + 3] = -y; J[6 * 5 + 4] = x;  J[6 * 5 + 5] = w;
}

/* analytic Jacobians of EdgeSE3 (g2o computeEdgeSE3Gradient with identity offsets), written to
 * shared memory */
UWBGO_DI void se3_jacobians_s(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *Ji, double *Jj)
{
    Pose Xi_inv, Bm, AB;
    pose_inv(Xi, Xi_inv);
    pose_mul(Xi_inv, Xj, Bm);
    pose_mul(Zinv, Bm, AB);
#pragma unroll
    for (int k = 0; k < 36; ++k) {
        Ji[k] = 0.0;
        Jj[k] = 0.0;
    }
    double qE[4];
    R_to_quat(AB.R, qE);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Jj[6 * r + c] = AB.R[3 * r + c];
    set_jqq_s(qE, Jj);
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
    quat_left_s(qA, Lm);
    quat_right_s(qB, Rm);
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

/* ONE column of the numeric central-difference Jacobian of a range residual (g2o
 * BaseBinaryEdge::linearizeOplus): the pose Xp (shared memory, R|t) with antenna offset `ant` is
 * perturbed along dd, `other` is the other end point.  callbase = (counter of the pose at the start of
 * this linearisation + oplus calls made on it by earlier edges) mod `mod`; the k-th call of this edge
 * trips the re-orthogonalisation of the PERTURBED estimate iff (callbase + k) mod `mod` == 0
 * (VertexSE3::oplusImpl; push / pop restores the estimate, not the counter).  Shortcuts, all exact:
 * a translation increment multiplies by the identity rotation, so the perturbed translation is
 * R[:,dd] * (+-delta) + t in one rounded product and one rounded sum; with the identity offset the
 * point is the translation, rotation columns are exactly 0 and a re-orthogonalisation is invisible. */
UWBGO_DI double range_jac_col(const WinSm &sm, const double *Xp, int ant, const double *other, double d, int dd,
                              int callbase, int mod, double delta, double scalar)
{
    if (dd >= 3 && ant <= 0) return 0.0;
    double epm[2];
    int call = callbase + 2 * dd;
    if (call >= mod) call -= mod;
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
        if (++call == mod) call = 0;
        const double v = sg == 0 ? delta : -delta;
        double P[3];
        if (dd < 3) {
            double tp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tp[r] = Xp[3 * r + dd] * v + Xp[9 + r];
            if (ant > 0) {
                const double *o = sm.ant + 3 * (ant - 1);
                if (call == 0) {
                    double Rp[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rp[k] = Xp[k];
                    orthogonalize(Rp);
                    mat3_vec_add(Rp, o, tp, P);
                } else
                    mat3_vec_add(Xp, o, tp, P);
            } else {
                P[0] = tp[0]; P[1] = tp[1]; P[2] = tp[2];
            }
        } else {
            const double *o = sm.ant + 3 * (ant - 1);
            double q[3] = {0.0, 0.0, 0.0};
            q[dd - 3] = v;
            double Rinc[9], Rp[9], R[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Xp[k];
            increment_R(q, Rinc);
            mat3_mul(R, Rinc, Rp);
            if (call == 0) orthogonalize(Rp);
            mat3_vec_add(Rp, o, Xp + 9, P);
        }
        epm[sg] = d - dist3(P[0], P[1], P[2], other[0], other[1], other[2]);
    }
    return scalar * (epm[0] - epm[1]);
}

/* computeError + chi2 of one edge at the estimates X: plain chi2 and its robustified value */
UWBGO_DI void edge_chi(const WinSm &sm, const double *X, const EdgeRec &er, const Cauchy &ck, double &chi_out,
                       double &rob_out)
{
    double chi;
    if (er.kind <= UWBGO_EDGE_RANGE_POSE) {
        double P0[3], Q[3];
        range_points(sm, X, er, P0, Q);
        const double err = sm.rd[er.slot] - dist3(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2]);
        const double Oe = sm.ri[er.slot] * err;
        chi = err * Oe;
    } else if (er.kind == UWBGO_EDGE_PRIOR) {
        Pose Dl;
        double q[4], e6[6], Oe[6];
        prior_error(sm.pZi + 12 * er.slot, X + 12 * er.a, Dl, q, e6);
        chi = chi2_6s(sm.pI + 36 * er.slot, e6, Oe);
    } else {
        Pose Zinv, Xi, Xj;
        ld_pose(sm.sZi + 12 * er.slot, Zinv);
        ld_pose(X + 12 * er.a, Xi);
        ld_pose(X + 12 * er.b, Xj);
        double e6[6], Oe[6];
        se3_error_s(Zinv, Xi, Xj, e6);
        chi = chi2_6s(sm.sI + 36 * er.slot, e6, Oe);
    }
    chi_out = chi;
    rob_out = er.robust ? ck.rho0(chi) : chi;
}

/* named barrier over the first `count` threads of the CTA */
UWBGO_DI void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

}  // namespace

__global__ void __launch_bounds__(WT, 1)
lm_window_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                 const __grid_constant__ WinIo io)
{
    extern __shared__ __align__(16) double wsm[];
    __shared__ WinCtl ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w = blockIdx.x;
    const int N = tp.N, NE = tp.E, mod = cfg.orth_mod;
    const WinCarve cv = win_carve(tp);
    WinSm sm;
    sm.X[0] = wsm + cv.X0; sm.X[1] = wsm + cv.X1;
    sm.anch = wsm + cv.anch; sm.ant = wsm + cv.ant; sm.rd = wsm + cv.rd; sm.ri = wsm + cv.ri;
    sm.pZi = wsm + cv.pZi; sm.pI = wsm + cv.pI; sm.sZi = wsm + cv.sZi; sm.sI = wsm + cv.sI;
    sm.rJ = wsm + cv.rJ; sm.pJ = wsm + cv.pJ; sm.sJ = wsm + cv.sJ;
    sm.Hd = wsm + cv.Hd; sm.Ho = wsm + cv.Ho; sm.b = wsm + cv.b;
    sm.G = wsm + cv.G; sm.M = wsm + cv.M; sm.c = wsm + cv.c; sm.z = wsm + cv.z; sm.x = wsm + cv.x;
    sm.S = wsm + cv.S; sm.echi = wsm + cv.echi;
    sm.cnt = reinterpret_cast<int *>(wsm + cv.cnt);
    sm.edges = reinterpret_cast<EdgeRec *>(wsm + cv.edges);
    Cauchy ck;
    ck.init(cfg.kdelta);
    const double delta = cfg.jdelta, scalar = 1.0 / (2.0 * delta);

    /* ---- the window into shared memory -------------------------------------------------------- */
    for (int k = tid; k < N * 12; k += WT) {
        const int i = k / 12, j = k - 12 * i;
        double v;
        if (j >= 9)
            v = io.pose_t[(w * N + i) * 3 + (j - 9)];
        else if (io.pose_R)
            v = io.pose_R[(w * N + i) * 9 + j];
        else
            v = (j == 0 || j == 4 || j == 8) ? 1.0 : 0.0;
        sm.X[0][k] = v;
    }
    for (int k = tid; k < N; k += WT) sm.cnt[k] = io.cnt_in ? io.cnt_in[w * N + k] : 0;
    for (int k = tid; k < tp.A * 3; k += WT) sm.anch[k] = io.anchors[w * tp.A * 3 + k];
    for (int k = tid; k < tp.K * 3; k += WT) sm.ant[k] = io.ant[k];
    for (int k = tid; k < tp.Er; k += WT) {
        sm.rd[k] = io.rd[w * tp.Er + k];
        sm.ri[k] = io.ri[w * tp.Er + k];
    }
    for (int k = tid; k < tp.Ep * 36; k += WT) sm.pI[k] = io.pI[w * tp.Ep * 36 + k];
    for (int k = tid; k < tp.Es * 36; k += WT) sm.sI[k] = io.sI[w * tp.Es * 36 + k];
    for (int k = tid; k < tp.Ep + tp.Es; k += WT) { /* measurement inverses, once */
        const bool pr = k < tp.Ep;
        const int s = pr ? k : k - tp.Ep;
        const double *zp = pr ? io.pZ + (w * tp.Ep + s) * 12 : io.sZ + (w * tp.Es + s) * 12;
        Pose Z, Zinv;
        ld_pose(zp, Z);
        pose_inv(Z, Zinv);
        st_pose((pr ? sm.pZi : sm.sZi) + 12 * s, Zinv);
    }
    {
        const int words = (int)(sizeof(EdgeRec) / 4) * NE;
        const int *src = reinterpret_cast<const int *>(tp.edges);
        int *dst = reinterpret_cast<int *>(sm.edges);
        for (int k = tid; k < words; k += WT) dst[k] = __ldg(src + k);
    }
    __syncthreads();

    /* phase C: chi2 terms of every edge at buffer `sel`; kinds are kept in separate warps */
    const int nC = (WNW > 1 ? WNW - 1 : 1) * 32; /* threads of the U / C phases (the last warp sums computeScale) */
    auto chi_phase = [&](int sel) {
        const double *X = sm.X[sel];
        const int *se = tp.slot_edge;
        const int n6 = tp.Es + tp.Ep;
        const int base = ((n6 + 31) & ~31) % nC;
        for (int k = tid; k < n6; k += nC) { /* se3 slots, then prior slots */
            const int e = __ldg(se + tp.Er + (k < tp.Es ? tp.Ep + k : k - tp.Es));
            double chi, rob;
            edge_chi(sm, X, sm.edges[e], ck, chi, rob);
            sm.echi[2 * e] = chi;
            sm.echi[2 * e + 1] = rob;
        }
        for (int k = (tid - base + nC) % nC; k < tp.Er; k += nC) {
            const int e = __ldg(se + k);
            double chi, rob;
            edge_chi(sm, X, sm.edges[e], ck, chi, rob);
            sm.echi[2 * e] = chi;
            sm.echi[2 * e + 1] = rob;
        }
    };
    auto chi_sum = [&](double &p, double &r) {
        double pp = 0.0, rr = 0.0;
        for (int e = 0; e < NE; ++e) {
            pp = pp + sm.echi[2 * e];
            rr = rr + sm.echi[2 * e + 1];
        }
        p = pp;
        r = rr;
    };

    /* LM state (thread 0) */
    double lambda = 0.0, ni = 2.0, stale = 0.0, plainCur = 0.0, currentChi = 0.0, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0;
    bool need_lin = true;

    if (tid < nC) chi_phase(0);
    __syncthreads();
    if (tid == 0) {
        chi_sum(plainCur, currentChi);
        stale = plainCur;
        ctl.go = cfg.max_iterations > 0;
        ctl.lin = 1;
        ctl.cur = 0;
    }

#ifdef UWBGO_WIN_TIMING
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = clock64();
#define WIN_TICK(k) do { long long tn_ = clock64(); tph[k] += tn_ - tq; tq = tn_; } while (0)
#else
#define WIN_TICK(k)
#endif
    while (true) {
        __syncthreads(); /* control published */
        WIN_TICK(7);
        if (!ctl.go) break;
        const int cb = ctl.cur;
        if (ctl.lin) {
            const double *X = sm.X[cb];
            /* ---- J: Jacobians, errors, weights ---------------------------------------------- */
            {
                const int *se = tp.slot_edge;
                const int n6 = tp.Es + tp.Ep;
                for (int k = tid; k < n6; k += WT) {
                    if (k < tp.Es) {
                        const EdgeRec er = sm.edges[__ldg(se + tp.Er + tp.Ep + k)];
                        double *rec = sm.sJ + SJ * er.slot;
                        Pose Zinv, Xi, Xj;
                        ld_pose(sm.sZi + 12 * er.slot, Zinv);
                        ld_pose(X + 12 * er.a, Xi);
                        ld_pose(X + 12 * er.b, Xj);
                        double e6[6], Oe[6];
                        se3_error_s(Zinv, Xi, Xj, e6);
                        se3_jacobians_s(Zinv, Xi, Xj, rec, rec + 36);
                        const double chi = chi2_6s(sm.sI + 36 * er.slot, e6, Oe);
                        const double r1 = er.robust ? ck.rho1(chi) : 1.0;
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            double v = -Oe[j];
                            if (er.robust) v = v * r1;
                            rec[144 + j] = v;
                        }
                        rec[150] = r1;
                    } else {
                        const EdgeRec er = sm.edges[__ldg(se + tp.Er + (k - tp.Es))];
                        double *rec = sm.pJ + PJ * er.slot;
                        Pose Dl;
                        double qq[4], e6[6], Oe[6];
                        prior_error(sm.pZi + 12 * er.slot, X + 12 * er.a, Dl, qq, e6);
                        const double chi = chi2_6s(sm.pI + 36 * er.slot, e6, Oe);
                        const double r1 = er.robust ? ck.rho1(chi) : 1.0;
#pragma unroll
                        for (int j = 0; j < 36; ++j) rec[j] = 0.0;
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) rec[6 * r + c] = Dl.R[3 * r + c];
                        set_jqq_s(qq, rec);
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            double v = -Oe[j];
                            if (er.robust) v = v * r1;
                            rec[72 + j] = v;
                        }
                        rec[78] = r1;
                    }
                }
                const int base1 = ((n6 + 31) & ~31) % WT;
                for (int k = (tid - base1 + WT) % WT; k < tp.Er; k += WT) { /* error and weights of a range edge */
                    const EdgeRec er = sm.edges[__ldg(se + k)];
                    double P0[3], Q[3];
                    range_points(sm, X, er, P0, Q);
                    const double info = sm.ri[k];
                    const double err = sm.rd[k] - dist3(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2]);
                    const double Oe = info * err;
                    double omega_r = -Oe, Ow = info;
                    if (er.robust) {
                        const double r1 = ck.rho1(err * Oe);
                        omega_r = omega_r * r1;
                        Ow = r1 * info;
                    }
                    sm.rJ[RJ * k + 12] = Ow;
                    sm.rJ[RJ * k + 13] = omega_r;
                }
                const int base2 = (base1 + ((tp.Er + 31) & ~31)) % WT;
                for (int u = (tid - base2 + WT) % WT; u < 12 * tp.Er; u += WT) { /* one Jacobian column */
                    const int k = u / 12, j = u - 12 * k, which = j / 6, dd = j - 6 * which;
                    const EdgeRec er = sm.edges[__ldg(se + k)];
                    double v = 0.0;
                    if (which == 0) {
                        double Q[3];
                        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR)
                            anchor_point(sm, er.b, er.ant_b, Q);
                        else
                            offset_point(sm, X + 12 * er.b, er.ant_b, Q);
                        v = range_jac_col(sm, X + 12 * er.a, er.ant, Q, sm.rd[k], dd, (sm.cnt[er.a] + er.base_a) % mod, mod,
                                          delta, scalar);
                    } else if (er.kind == UWBGO_EDGE_RANGE_POSE) {
                        double P0[3];
                        offset_point(sm, X + 12 * er.a, er.ant, P0);
                        v = range_jac_col(sm, X + 12 * er.b, er.ant_b, P0, sm.rd[k], dd, (sm.cnt[er.b] + er.base_b) % mod,
                                          mod, delta, scalar);
                    }
                    sm.rJ[RJ * k + j] = v;
                }
            }
            __syncthreads();
            WIN_TICK(0);
            /* ---- O: J^T Ow of the 6-D edges, one entry per thread; oplus counters advance ----- */
            {
                const int nP = 36 * tp.Ep, nS = 72 * tp.Es;
                for (int u = tid; u < nP + nS; u += WT) {
                    const double *J, *O;
                    double *out;
                    double r1;
                    bool robust;
                    int rc;
                    if (u < nS) {
                        const int s = u / 72, j = u - 72 * s, which = j / 36;
                        rc = j - 36 * which;
                        const double *rec = sm.sJ + SJ * s;
                        J = rec + 36 * which;
                        out = sm.sJ + SJ * s + 72 + 36 * which;
                        O = sm.sI + 36 * s;
                        r1 = rec[150];
                        robust = sm.edges[__ldg(tp.slot_edge + tp.Er + tp.Ep + s)].robust != 0;
                    } else {
                        const int v = u - nS, s = v / 36;
                        rc = v - 36 * s;
                        const double *rec = sm.pJ + PJ * s;
                        J = rec;
                        out = sm.pJ + PJ * s + 36;
                        O = sm.pI + 36 * s;
                        r1 = rec[78];
                        robust = sm.edges[__ldg(tp.slot_edge + tp.Er + s)].robust != 0;
                    }
                    const int r = rc / 6, c = rc - 6 * r;
                    double ow[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const double v = O[6 * k + c];
                        ow[k] = robust ? r1 * v : v;
                    }
                    double s = J[r] * ow[0];
#pragma unroll
                    for (int k = 1; k < 6; ++k) s = fma(J[6 * k + r], ow[k], s);
                    out[rc] = s;
                }
                /* the numeric Jacobians made num_calls[i] oplus calls on pose i (all columns were read
                 * before the barrier above) */
                for (int i = tid; i < N; i += WT) sm.cnt[i] = (sm.cnt[i] + __ldg(tp.num_calls + i)) % mod;
            }
            __syncthreads();
            WIN_TICK(1);
            /* ---- H: every entry of the H record of every pose, owned by one thread -------------- */
            for (int u = tid; u < 63 * N; u += WT) {
                const int i = u / 63, k = u - 63 * i;
                /* k < 21: H_ii upper (r, c); 21..56: H_{parent(i), i} (r, c); 57..62: b_i[r] */
                int r, c, what;
                if (k < 21) {
                    what = 0;
                    r = 0;
                    int kk = k;
                    while (kk >= 6 - r) {
                        kk -= 6 - r;
                        ++r;
                    }
                    c = r + kk;
                } else if (k < 57) {
                    what = 1;
                    r = (k - 21) / 6;
                    c = (k - 21) - 6 * r;
                } else {
                    what = 2;
                    r = k - 57;
                    c = 0;
                }
                double acc = 0.0;
                const int ob = __ldg(tp.op_begin + i), oe = __ldg(tp.op_begin + i + 1);
                for (int o = ob; o < oe; ++o) {
                    const int2 op = __ldg(reinterpret_cast<const int2 *>(tp.ops + o));
                    const EdgeRec &er = sm.edges[op.x];
                    const int role = op.y;
                    if (er.kind <= UWBGO_EDGE_RANGE_POSE) {
                        const double *rec = sm.rJ + RJ * er.slot;
                        const double *J = rec + 6 * role;
                        const double Ow = rec[12];
                        if (what == 0)
                            acc = fma(J[r] * Ow, J[c], acc);
                        else if (what == 2)
                            acc = fma(J[r], rec[13], acc);
                        else if (role == 1)
                            acc = fma(rec[r] * Ow, rec[6 + c], acc);
                    } else {
                        const bool se3 = er.kind == UWBGO_EDGE_SE3;
                        const double *rec = se3 ? sm.sJ + SJ * er.slot : sm.pJ + PJ * er.slot;
                        const double *J = rec + 36 * role;                          /* A or B           */
                        const double *JtO = rec + (se3 ? 72 : 36) + 36 * role;      /* A^T Ow or B^T Ow */
                        const double *om = rec + (se3 ? 144 : 72);
                        if (what == 0) {
                            double s = JtO[6 * r] * J[c];
#pragma unroll
                            for (int j = 1; j < 6; ++j) s = fma(JtO[6 * r + j], J[6 * j + c], s);
                            acc = acc + s;
                        } else if (what == 2) {
                            double s = J[r] * om[0];
#pragma unroll
                            for (int j = 1; j < 6; ++j) s = fma(J[6 * j + r], om[j], s);
                            acc = acc + s;
                        } else if (role == 1) { /* rows of pose a, columns of pose i: (A^T Ow) B */
                            const double *AtO = rec + 72, *B = rec + 36;
                            double s = AtO[6 * r] * B[c];
#pragma unroll
                            for (int j = 1; j < 6; ++j) s = fma(AtO[6 * r + j], B[6 * j + c], s);
                            acc = acc + s;
                        }
                    }
                }
                if (what == 0)
                    sm.Hd[21 * i + k] = acc;
                else if (what == 1)
                    sm.Ho[36 * i + (k - 21)] = acc;
                else
                    sm.b[6 * i + r] = acc;
            }
            __syncthreads();
            WIN_TICK(2);
        }

        /* ---- F: the trial's linear solve, warp 0 ------------------------------------------------ */
        if (warp == 0) {
            if (tid == 0) {
                if (need_lin) {
                    stale = plainCur;
                    if (it == 0) {
                        double maxdiag = 0.0;
                        for (int i = 0; i < N; ++i)
#pragma unroll
                            for (int r = 0; r < 6; ++r) {
                                const double v = fabs(sm.Hd[21 * i + up_idx(6, r, r)]);
                                if (v > maxdiag) maxdiag = v;
                            }
                        lambda = cfg.tau * maxdiag;
                        ni = 2.0;
                    }
                    rho = 0.0;
                    q = 0;
                    need_lin = false;
                }
                ctl.lambda = lambda;
            }
            __syncwarp();
            const double lam = ctl.lambda;
            /* lane roles of the assembly: 0..20 entry (r, c) of the lower triangle of S, 21..26 row of the
             * right-hand side */
            int er_ = 0, ec_ = 0;
            if (lane < 21) {
                int kk = lane;
                while (kk > er_) {
                    kk -= er_ + 1;
                    ++er_;
                }
                ec_ = kk;
            } else if (lane < 27) {
                er_ = lane - 21;
            }
            bool ok = true;
            for (int i = N - 1; i >= 0; --i) {
                double v = 0.0;
                if (lane < 21) {
                    v = sm.Hd[21 * i + up_idx(6, ec_, er_)];
                    if (er_ == ec_) v = v + lam;
                } else if (lane < 27) {
                    v = sm.b[6 * i + er_];
                }
                const int qb = __ldg(tp.child_begin + i), qe = __ldg(tp.child_begin + i + 1);
                for (int qq = qb; qq < qe; ++qq) { /* children in descending order */
                    const int ch = __ldg(tp.children + qq);
                    const double *Gc = sm.G + 36 * ch, *zc = sm.z + 6 * ch;
                    if (lane < 21) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) v = fma(-Gc[6 * er_ + k], Gc[6 * ec_ + k], v);
                    } else if (lane < 27) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) v = fma(-Gc[6 * er_ + k], zc[k], v);
                    }
                }
                if (lane < 27) sm.S[lane] = v;
                __syncwarp();
                double S[21], L[21], rhs[6], fw[6], bw[6];
#pragma unroll
                for (int k = 0; k < 21; ++k) S[k] = sm.S[k];
                const bool has_parent = __ldg(tp.parent + i) >= 0;
                {
                    const double *src = lane < 6 ? sm.Ho + 36 * i + 6 * lane : sm.S + 21;
#pragma unroll
                    for (int k = 0; k < 6; ++k) rhs[k] = src[k];
                }
                __syncwarp();
                /* potrf, every lane the whole block; the diagonal slot keeps 1 / L_jj */
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double s = S[lo_idx(j, j)];
#pragma unroll
                    for (int k = 0; k < j; ++k) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
                    if (!(s > 0.0)) ok = false;
                    const double inv = 1.0 / sqrt(s);
                    L[lo_idx(j, j)] = inv;
#pragma unroll
                    for (int r = j + 1; r < 6; ++r) {
                        double t = S[lo_idx(r, j)];
#pragma unroll
                        for (int k = 0; k < j; ++k) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
                        L[lo_idx(r, j)] = t * inv;
                    }
                }
                /* forward substitution: row of G_i (lanes 0..5) or z_i (lane 6) */
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) {
                    double s = rhs[cc];
#pragma unroll
                    for (int k = 0; k < cc; ++k) s = fma(-fw[k], L[lo_idx(cc, k)], s);
                    fw[cc] = s * L[lo_idx(cc, cc)];
                }
                /* backward substitution: column of M_i (lanes 0..5) or c_i (lane 6) */
#pragma unroll
                for (int r = 5; r >= 0; --r) {
                    double s = fw[r];
#pragma unroll
                    for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], bw[k], s);
                    bw[r] = s * L[lo_idx(r, r)];
                }
                if (lane < 6 && has_parent) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        sm.G[36 * i + 6 * lane + k] = fw[k];
                        sm.M[36 * i + 6 * k + lane] = bw[k];
                    }
                } else if (lane == 6) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        sm.z[6 * i + k] = fw[k];
                        sm.c[6 * i + k] = bw[k];
                    }
                }
                __syncwarp();
            }
            ok = __all_sync(0xffffffffu, ok);
            WIN_TICK(3);
            /* substitution x_i = c_i - M_i x_parent(i), ascending */
            for (int i = 0; i < N; ++i) {
                const int par = __ldg(tp.parent + i);
                if (lane < 6) {
                    double s = sm.c[6 * i + lane];
                    if (par >= 0) {
                        const double *Mi = sm.M + 36 * i + 6 * lane, *xp = sm.x + 6 * par;
#pragma unroll
                        for (int j = 0; j < 6; ++j) s = fma(-Mi[j], xp[j], s);
                    }
                    sm.x[6 * i + lane] = ok ? s : 0.0;
                }
                __syncwarp();
            }
            if (tid == 0) {
                ctl.ok = ok ? 1 : 0;
                if (!ok) flags |= UWBGO_FLAG_CHOL_FAIL;
            }
        }
        __syncthreads();
        WIN_TICK(4);

        /* ---- U: estimate (+) x into the trial buffer; computeScale() on the last warp ----------- */
        if (tid < nC) {
            const double *Xc = sm.X[cb];
            double *Xn = sm.X[cb ^ 1];
            for (int i = tid; i < N; i += nC) {
                Pose X;
                ld_pose(Xc + 12 * i, X);
                double xv[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) xv[k] = sm.x[6 * i + k];
                int c = sm.cnt[i];
                pose_oplus(X, xv, c, mod);
                sm.cnt[i] = c;
                st_pose(Xn + 12 * i, X);
            }
            if (WNW > 1) bar_named(1, nC);
            else __syncthreads();
            chi_phase(cb ^ 1);
        }
        if (tid == WT - 1 || (WNW == 1 && tid == 0)) {
            const double lam = ctl.lambda;
            double scale = 0.0;
            for (int j = 0; j < 6 * N; ++j) scale = scale + sm.x[j] * (lam * sm.x[j] + sm.b[j]);
            ctl.scale = scale;
        }
        __syncthreads();
        WIN_TICK(5);

        /* ---- D: accept / reject ------------------------------------------------------------------ */
        if (tid == 0) {
            const bool ok = ctl.ok != 0;
            double scale = ctl.scale, tplain, tempChi;
            chi_sum(tplain, tempChi);
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
            bool done = false;
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
            ctl.go = done ? 0 : 1;
            ctl.lin = need_lin ? 1 : 0;
            ctl.cur = cur;
            WIN_TICK(6);
        }
    }
#ifdef UWBGO_WIN_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("window 0 cycles: J %lld  O %lld  H %lld  factor %lld  subst %lld  U+C %lld  D %lld  sync %lld  (trials %d, iterations %d)\n",
               tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7], trials_total, iterations);
#endif

    /* ---- results ---------------------------------------------------------------------------------- */
    const double *Xf = sm.X[ctl.cur];
    for (int k = tid; k < N * 3; k += WT) {
        const int i = k / 3, j = k - 3 * i;
        io.o_pose_t[(w * N + i) * 3 + j] = Xf[12 * i + 9 + j];
    }
    if (io.o_pose_R)
        for (int k = tid; k < N * 9; k += WT) {
            const int i = k / 9, j = k - 9 * i;
            io.o_pose_R[(w * N + i) * 9 + j] = Xf[12 * i + j];
        }
    if (io.o_cnt)
        for (int k = tid; k < N; k += WT) io.o_cnt[w * N + k] = sm.cnt[k];
    if (tid == 0) {
        if (io.o_chi2) {
            double *o = io.o_chi2 + w * UWBGO_CHI2_STRIDE;
            o[0] = plainCur;
            o[1] = currentChi;
            o[2] = stale;
            o[3] = lambda;
        }
        if (io.o_status) {
            int32_t *o = io.o_status + w * UWBGO_STATUS_STRIDE;
            o[0] = iterations;
            o[1] = trials_total;
            o[2] = flags;
            o[3] = qlast;
        }
    }
}

size_t window_path_smem_bytes(const DevTopo &topo) { return sizeof(double) * (size_t)win_carve(topo).total; }

cudaError_t launch_solve_window(const DevTopo &topo, const DevCfg &cfg, const WinIo &io, int device, cudaStream_t st)
{
    if (io.W <= 0) return cudaSuccess;
    const size_t sm = window_path_smem_bytes(topo);
    static size_t configured[64] = {0}; /* the attribute is per device */
    size_t &conf = configured[device & 63];
    if (sm > conf) {
        cudaError_t e = cudaFuncSetAttribute(lm_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        conf = sm;
    }
    lm_window_kernel<<<(unsigned)io.W, WT, sm, st>>>(topo, cfg, io);
    return cudaGetLastError();
}

}  // namespace uwbgo
