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
 *       accumulation order of the CPU checker, no atomics, no reductions
 * and per trial
 *   F   block elimination newest pose first on ONE warp: the 21 + 6 entries of S_i and of the
 *       right-hand side are assembled by one lane each, every lane then runs the 6x6 potrf in
 *       registers (redundantly: the sqrt -> reciprocal chain is latency, not throughput), and lanes
 *       0..6 run the SAME triangular substitutions on seven right-hand sides: rows of
 *       H_{parent(i),i} give the rows of G_i and the columns of M_i, b_i gives z_i and c_i.
 *       Then the substitution x_i = c_i - M_i x_parent(i), one lane per row, while a lane of the
 *       helper warp follows it with the strictly ordered sum of computeScale()
 *   U   estimate (+) x_i, one lane per pose
 *   C   computeActiveErrors: 6-D edges on the trial's main warp, range edges on its helper warp
 *   D   accept / reject, lambda update (thread 0), chi2 terms summed in insertion order
 * SPECULATION.  What LM does after a REJECTED trial is known before the trial is evaluated: the
 * same H and b with lambda * nu.  So a round evaluates up to KS candidates lambda, lambda nu,
 * lambda nu 2nu, ... at once, each on its own pair of warps (F, U, C are per candidate), and D
 * consumes them in order exactly as the serial loop would: the first accepted candidate ends the
 * iteration, the others are discarded.  A rejection costs no extra round.  The oplus counter of a
 * pose advances by one per CONSUMED trial, so candidate k updates with counter + k.
 * Per-entry operation sequences are those of the tile kernels (uwbgo_general.cuh) and of the CPU
 * checker: the results are the same bits.
 */
#include <type_traits>

#include "uwbgo_device.cuh"
#ifdef UWBGO_WIN_TIMING
#include <cstdio>
#endif

namespace uwbgo {

namespace {

#ifdef UWBGO_WIN_TIMING
__device__ unsigned long long g_bad_chi6 = 0, g_bad_lin6 = 0, g_n_chi6 = 0, g_n_lin6 = 0; /* IEEE re-evaluations */
__device__ __forceinline__ long long clk()
{
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
#endif

/* per-edge linearisation records (doubles) */
constexpr int RJ = 14;  /* range: A 6 | B 6 | Ow | omega_r                      */
constexpr int PJ = 79;  /* prior: J 36 | J^T Ow 36 | omega_r 6 | rho1           */
constexpr int SJ = 151; /* se3:   A 36 | B 36 | A^T Ow 36 | B^T Ow 36 | omega_r 6 | rho1 */

constexpr int WIN_MAX_KS = 4;
#ifndef UWBGO_WIN_IEEE
#define UWBGO_WIN_IEEE 0 /* experiment: bit 0 C phase, bit 1 J phase, bit 2 U phase run the IEEE sequences */
#endif
/* The branch-free policy of the WINDOW kernel.  NbMath::div flags a divisor with a significand of all ones (its
 * Newton sequence is not guaranteed to round correctly there) and the caller then repeats the WHOLE item with the
 * IEEE sequences.  That divisor is systematic here: the quaternion of a near-identity rotation divides by
 * sqrt(4 - 2 ulp) = 2 - ulp, and the error rotation of an IMU prior on the pose whose estimate the same message
 * set (localization.cpp:499-535) is exactly that -- one prior per window sent every J and C phase of its warp
 * through both code paths.  Here such a quotient alone is formed by the IEEE division (same bits, by definition)
 * and the item is not flagged. */
using WMC = std::conditional<(UWBGO_WIN_IEEE & 1) != 0, IeeeMath, NbMathW>::type;
using WMJ = std::conditional<(UWBGO_WIN_IEEE & 2) != 0, IeeeMath, NbMathW>::type;
using WMU = std::conditional<(UWBGO_WIN_IEEE & 4) != 0, IeeeMath, NbMathW>::type;
#ifdef UWBGO_WIN_NO6D
constexpr bool NO6D = true; /* experiment: range-only build, to see what the 6-D code costs in instruction fetch */
#else
constexpr bool NO6D = false;
#endif

/* shared-memory map, offsets in doubles */
struct WinCarve {
    int X, anch, ant, rd, ri, pZi, pI, sZi, sI, rJ, pJ, sJ, Hd, Ho, b, cand, cand_stride;
    int cnt, parent, cbeg, chl, opb, ncalls, slot_edge, ops, edges, total;
};

__host__ __device__ inline WinCarve win_carve(const DevTopo &t, int ks)
{
    WinCarve c;
    int o = 0;
    auto take = [&](int n) {
        int at = o;
        o += (n + 1) & ~1; /* keep 16-byte alignment */
        return at;
    };
    auto ints = [&](int n) { return take((n + 1) / 2); };
    const int N = t.N;
    c.X = take((ks + 1) * N * 12); /* pool of ks + 1 pose buffers: the estimate and one trial per candidate */
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
    /* per candidate: G 36N | M 36N | c 6N | z 6N | x 6N | S 28 | L staging 2 x 22 | echi 2E | 42 zeros (the
     * G and z of the child the newest pose does not have) */
    c.cand_stride = ((90 * N + 28 + 44 + 2 * t.E + 1) & ~1) + 42;
    c.cand = take(ks * c.cand_stride);
    c.cnt = ints(N);
    c.parent = ints(N);
    c.cbeg = ints(N + 1);
    c.chl = ints(N);
    c.opb = ints(N + 1);
    c.ncalls = ints(N);
    c.slot_edge = ints(t.E > 0 ? t.E : 1);
    c.ops = ints(2 * (t.n_ops > 0 ? t.n_ops : 1));
    c.edges = take((int)(sizeof(EdgeRec) / 8) * (t.E > 0 ? t.E : 1));
    c.total = o;
    return c;
}

struct WinSm {
    double *X, *anch, *ant, *rd, *ri, *pZi, *pI, *sZi, *sI, *rJ, *pJ, *sJ, *Hd, *Ho, *b, *cand;
    int cand_stride, xstride;
    int *cnt, *parent, *cbeg, *chl, *opb, *ncalls, *slot_edge;
    int2 *ops;
    EdgeRec *edges;
};

struct WinCtl {
    int go, lin, cur, nk, adv, first;
    int ok[WIN_MAX_KS];
    int redo[WIN_MAX_KS]; /* the branch-free arithmetic flagged an operand: the chain runs again with IEEE sequences */
    double lam[WIN_MAX_KS], scale[WIN_MAX_KS];
    /* per candidate, formed by lane 0 of its main warp before D: the two chi2 sums and the gain ratio */
    double tplain[WIN_MAX_KS], tchi[WIN_MAX_KS], rho[WIN_MAX_KS];
    double chi_cur; /* robust chi2 at the current estimate (thread 0's currentChi) */
    int bd;         /* block-diagonal window (factor_main_bd) */
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
template <class M>
UWBGO_DI void prior_error(const double *Zi, const double *Xp, Pose &Dl, double *q, double *e6, unsigned &bad)
{
    Pose Zinv, X;
    ld_pose(Zi, Zinv);
    ld_pose(Xp, X);
    pose_mul(Zinv, X, Dl);
    R_to_quat_m<M>(Dl.R, q, bad);
    e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
    e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
}

/* toVectorMQT(Zinv * Xi^-1 * Xj), evaluated left to right (EdgeSE3::computeError) */
template <class M>
UWBGO_DI void se3_error_s(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *e, unsigned &bad)
{
    Pose Xi_inv, T, Dl;
    pose_inv(Xi, Xi_inv);
    pose_mul(Zinv, Xi_inv, T);
    pose_mul(T, Xj, Dl);
    double q[4];
    R_to_quat_m<M>(Dl.R, q, bad);
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
template <class M>
UWBGO_DI void se3_jacobians_s(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *Ji, double *Jj, unsigned &bad)
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
    R_to_quat_m<M>(AB.R, qE, bad);
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
    R_to_quat_m<M>(Ra, qA, bad);
    R_to_quat_m<M>(Bm.R, qB, bad);
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
template <class M>
UWBGO_DI double range_jac_col(const WinSm &sm, const double *Xp, int ant, const double *other, double d, int dd,
                              int callbase, int mod, double delta, double scalar, unsigned &bad)
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
            const double q[3] = {dd == 3 ? v : 0.0, dd == 4 ? v : 0.0, dd == 5 ? v : 0.0};
            double Rinc[9], Rp[9], R[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Xp[k];
            increment_R_m<M>(q, Rinc, bad);
            mat3_mul(R, Rinc, Rp);
            if (call == 0) orthogonalize(Rp);
            mat3_vec_add(Rp, o, Xp + 9, P);
        }
        epm[sg] = d - dist3m<M>(P[0], P[1], P[2], other[0], other[1], other[2], bad);
    }
    return scalar * (epm[0] - epm[1]);
}

/* computeError + chi2 of one edge at the estimates X: plain chi2 and its robustified value */
template <class M>
UWBGO_DI void range_chi(const WinSm &sm, const double *X, const EdgeRec &er, const Cauchy &ck, double &chi, double &rob,
                        unsigned &bad)
{
    double P0[3], Q[3];
    range_points(sm, X, er, P0, Q);
    const double err = sm.rd[er.slot] - dist3m<M>(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2], bad);
    const double Oe = sm.ri[er.slot] * err;
    chi = err * Oe;
    rob = er.robust ? ck.rho0m<M>(chi, bad) : chi;
}
template <class M>
UWBGO_DI void six_chi(const WinSm &sm, const double *X, const EdgeRec &er, const Cauchy &ck, double &chi, double &rob,
                      unsigned &bad)
{
    double e6[6], Oe[6];
    if (er.kind == UWBGO_EDGE_PRIOR) {
        Pose Dl;
        double q[4];
        prior_error<M>(sm.pZi + 12 * er.slot, X + 12 * er.a, Dl, q, e6, bad);
        chi = chi2_6s(sm.pI + 36 * er.slot, e6, Oe);
    } else {
        Pose Zinv, Xi, Xj;
        ld_pose(sm.sZi + 12 * er.slot, Zinv);
        ld_pose(X + 12 * er.a, Xi);
        ld_pose(X + 12 * er.b, Xj);
        se3_error_s<M>(Zinv, Xi, Xj, e6, bad);
        chi = chi2_6s(sm.sI + 36 * er.slot, e6, Oe);
    }
    rob = er.robust ? ck.rho0m<M>(chi, bad) : chi;
}

/* ---- J-phase items: what one thread computes for one edge (6-D), one range edge's weights, or one
 * column of a range Jacobian.  Records: see RJ / PJ / SJ. ---------------------------------------------- */
template <class M>
UWBGO_DI void lin_six(const WinSm &sm, const double *X, const EdgeRec &er, const Cauchy &ck, unsigned &bad)
{
    double e6[6], Oe[6];
    if (er.kind == UWBGO_EDGE_SE3) {
        double *rec = sm.sJ + SJ * er.slot;
        Pose Zinv, Xi, Xj;
        ld_pose(sm.sZi + 12 * er.slot, Zinv);
        ld_pose(X + 12 * er.a, Xi);
        ld_pose(X + 12 * er.b, Xj);
        se3_error_s<M>(Zinv, Xi, Xj, e6, bad);
        se3_jacobians_s<M>(Zinv, Xi, Xj, rec, rec + 36, bad);
        const double chi = chi2_6s(sm.sI + 36 * er.slot, e6, Oe);
        const double r1 = er.robust ? ck.rho1m<M>(chi, bad) : 1.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double v = -Oe[j];
            if (er.robust) v = v * r1;
            rec[144 + j] = v;
        }
        rec[150] = r1;
    } else {
        double *rec = sm.pJ + PJ * er.slot;
        Pose Dl;
        double qq[4];
        prior_error<M>(sm.pZi + 12 * er.slot, X + 12 * er.a, Dl, qq, e6, bad);
        const double chi = chi2_6s(sm.pI + 36 * er.slot, e6, Oe);
        const double r1 = er.robust ? ck.rho1m<M>(chi, bad) : 1.0;
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
template <class M>
UWBGO_DI void lin_range_weights(const WinSm &sm, const double *X, const EdgeRec &er, const Cauchy &ck, unsigned &bad)
{
    const int k = er.slot;
    double P0[3], Q[3];
    range_points(sm, X, er, P0, Q);
    const double info = sm.ri[k];
    const double err = sm.rd[k] - dist3m<M>(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2], bad);
    const double Oe = info * err;
    double omega_r = -Oe, Ow = info;
    if (er.robust) {
        const double r1 = ck.rho1m<M>(err * Oe, bad);
        omega_r = omega_r * r1;
        Ow = r1 * info;
    }
    sm.rJ[RJ * k + 12] = Ow;
    sm.rJ[RJ * k + 13] = omega_r;
}
template <class M>
UWBGO_DI void lin_range_col(const WinSm &sm, const double *X, const EdgeRec &er, const int j, const int adv, const int mod,
                            const double delta, const double scalar, unsigned &bad)
{
    const int k = er.slot, which = j / 6, dd = j - 6 * which;
    double v = 0.0;
    if (which == 0) {
        double Q[3];
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR)
            anchor_point(sm, er.b, er.ant_b, Q);
        else
            offset_point(sm, X + 12 * er.b, er.ant_b, Q);
        v = range_jac_col<M>(sm, X + 12 * er.a, er.ant, Q, sm.rd[k], dd, (sm.cnt[er.a] + adv + er.base_a) % mod, mod,
                             delta, scalar, bad);
    } else if (er.kind == UWBGO_EDGE_RANGE_POSE) {
        double P0[3];
        offset_point(sm, X + 12 * er.a, er.ant, P0);
        v = range_jac_col<M>(sm, X + 12 * er.b, er.ant_b, P0, sm.rd[k], dd, (sm.cnt[er.b] + adv + er.base_b) % mod, mod,
                             delta, scalar, bad);
    }
    sm.rJ[RJ * k + j] = v;
}

/* named barrier over `count` threads */
UWBGO_DI void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__constant__ unsigned char UP_R[21] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5};
__constant__ unsigned char UP_C[21] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5};
/* the entries of a pose's H record that a block-diagonal window forms, as indices of the 63-entry enumeration
 * (0..20 H_ii upper, 21..56 H_{parent,i} row-major, 57..62 b_i): both diagonal blocks, the translation block of
 * H_{parent,i}, b_i */
__constant__ unsigned char BD_ENTRY[27] = {0,  1,  2,  6,  7,  11, 15, 16, 17, 18, 19, 20, 21, 22,
                                           23, 27, 28, 29, 33, 34, 35, 57, 58, 59, 60, 61, 62};

/* F, main warp of a candidate: block elimination of H + lam I, newest pose first.  cd = the candidate's
 * block (G | M | c | z | x | S | echi), Lst = its two staging slots for L_i.  Per pose: the 21 + 6
 * entries of S_i and of the right-hand side are assembled by one lane each, every lane runs the potrf,
 * lanes 0..6 the forward substitutions; G_i, z_i and L_i go to shared memory and the helper warp is
 * released (bar.sync with it: it is always waiting, its step is a tenth of this one).  Returns false
 * when a pivot was not positive. */
/* D = 6: the full blocks.  D = 3 (windows of range edges only, identity rotations, no lever arms, see
 * WinIo::t3): the rotation rows and columns of every block are exact zeros -- S_i is diag(T_i, lam I), the
 * rotation parts of L_i off its diagonal, of G_i, M_i, z_i, c_i and x_i are zeros -- and a zero operand leaves
 * an fma chain's value as it is, so only the translation entries are formed: the same operations in the same
 * order on them, half the dependent sqrt -> reciprocal pivots.  The rotation pivots are lam itself, which
 * leaves their positivity test.  Same storage strides as D = 6. */
template <class MATH, bool CHAIN, int D>
UWBGO_DI bool factor_main(const WinSm &sm, double *cd, double *Lst, const int bar_id, const int N, const double lam,
                          const int lane, unsigned &bad
#ifdef UWBGO_WIN_TIMING
                          , long long *tf
#endif
)
{
#ifdef UWBGO_WIN_TIMING
    long long tq = clock64();
#define F_TICK(k) do { long long tn_ = clock64(); tf[k] += tn_ - tq; tq = tn_; } while (0)
#else
#define F_TICK(k)
#endif
    constexpr int NS = D * (D + 1) / 2;
    double *G = cd, *zv = cd + 78 * N, *Sst = cd + 90 * N;
    /* lane roles of the assembly: 0..NS-1 entry (r, c) of the lower triangle of S, NS..NS+D-1 row of the rhs */
    int er_ = 0, ec_ = 0;
    if (lane < NS) {
        int kk = lane;
        while (kk > er_) {
            kk -= er_ + 1;
            ++er_;
        }
        ec_ = kk;
    } else if (lane < NS + D) {
        er_ = lane - NS;
    }
    const int hd_idx = up_idx(6, ec_, er_);
    /* operand offsets of the child update inside the candidate block: row er_ of G_c times row ec_ of G_c
     * (S lanes) or times z_c (rhs lanes) */
    const int a_off = 6 * er_, y_mul = lane < NS ? 36 : 6, y_off = lane < NS ? 6 * ec_ : 78 * N;
    const int pad_off = sm.cand_stride - 42; /* 42 zeros at the end of the candidate block */
    bool ok = D == 6 || lam > 0.0; /* D = 3: the rotation pivots are (0 + lam) - 0 */
    /* the assembly inputs of pose i are fetched during the potrf of pose i + 1 */
    auto fetch = [&](int i, double &v, int &qb, int &qe) {
        v = 0.0;
        if (lane < NS)
            v = sm.Hd[21 * i + hd_idx];
        else if (lane < NS + D)
            v = sm.b[6 * i + er_];
        if (!CHAIN) {
            qb = sm.cbeg[i];
            qe = sm.cbeg[i + 1];
        }
    };
    double vn;
    int qbn = 0, qen = 0;
    fetch(N - 1, vn, qbn, qen);
    for (int i = N - 1; i >= 0; --i) {
        double v = vn;
        const int qb = qbn, qe = qen;
        if (lane < NS && er_ == ec_) v = v + lam;
        if (CHAIN) {
            /* the one child is pose i + 1; the newest pose reads the zero block instead, so the body has no
             * branch: S(r, c) -= G[r,:] . G[c,:] and rhs(r) -= G[r,:] . z in one stream for all lanes */
            const int gch = i < N - 1 ? 36 * (i + 1) : pad_off, zch = i < N - 1 ? 78 * N + 6 * (i + 1) : pad_off + 36;
            const double2 *Ga = reinterpret_cast<const double2 *>(cd + gch + a_off);
            const double2 *Yb = reinterpret_cast<const double2 *>(cd + (lane < NS ? gch + 6 * ec_ : zch));
            const double2 a0 = Ga[0], a1 = Ga[1], y0 = Yb[0], y1 = Yb[1];
            v = fma(-a0.x, y0.x, v);
            v = fma(-a0.y, y0.y, v);
            v = fma(-a1.x, y1.x, v);
            if (D == 6) {
                const double2 a2 = Ga[2], y2 = Yb[2];
                v = fma(-a1.y, y1.y, v);
                v = fma(-a2.x, y2.x, v);
                v = fma(-a2.y, y2.y, v);
            }
        } else {
#pragma unroll 1
            for (int qq = qb; qq < qe; ++qq) { /* children in descending order */
                const int ch = sm.chl[qq];
                const double2 *Ga = reinterpret_cast<const double2 *>(cd + 36 * ch + a_off);
                const double2 *Yb = reinterpret_cast<const double2 *>(cd + y_mul * ch + y_off);
                const double2 a0 = Ga[0], a1 = Ga[1], y0 = Yb[0], y1 = Yb[1];
                v = fma(-a0.x, y0.x, v);
                v = fma(-a0.y, y0.y, v);
                v = fma(-a1.x, y1.x, v);
                if (D == 6) {
                    const double2 a2 = Ga[2], y2 = Yb[2];
                    v = fma(-a1.y, y1.y, v);
                    v = fma(-a2.x, y2.x, v);
                    v = fma(-a2.y, y2.y, v);
                }
            }
        }
        if (lane < NS + D) Sst[lane] = v;
        __syncwarp();
        F_TICK(0);
        const bool has_parent = CHAIN ? true : sm.parent[i] >= 0; /* chains: G_0 of the root is zero and unread */
        double L[NS], rhs[D], fw[D];
        {
            const double *src = lane < D ? sm.Ho + 36 * i + 6 * lane : Sst + NS;
#pragma unroll
            for (int k = 0; k < D; ++k) rhs[k] = src[k];
        }
        fetch(i > 0 ? i - 1 : 0, vn, qbn, qen);
        /* potrf, every lane the whole block, S read where it is used; the diagonal slot keeps 1 / L_jj.
         * (loops over the full 0..5 range with the triangle as a predicate: constant trip counts, so
         * that they are all unrolled and L stays in registers) */
#pragma unroll
        for (int j = 0; j < D; ++j) {
            double s = Sst[lo_idx(j, j)];
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < j) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
            if (!(s > 0.0)) ok = false;
            const double inv = MATH::rsqrt_pivot(s, bad);
            L[lo_idx(j, j)] = inv;
#pragma unroll
            for (int r = 0; r < D; ++r)
                if (r > j) {
                    double t = Sst[lo_idx(r, j)];
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        if (k < j) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
                    L[lo_idx(r, j)] = t * inv;
                }
        }
        F_TICK(1);
        /* forward substitution: row of G_i (lanes 0..D-1) or z_i (lane D) */
#pragma unroll
        for (int cc = 0; cc < D; ++cc) {
            double s = rhs[cc];
#pragma unroll
            for (int k = 0; k < D; ++k)
                if (k < cc) s = fma(-fw[k], L[lo_idx(cc, k)], s);
            fw[cc] = s * L[lo_idx(cc, cc)];
        }
        { /* G_i rows (lanes 0..D-1) and z_i (lane D): one predicated stream of 16-byte stores */
            double *dst = lane < D ? G + 36 * i + 6 * lane : zv + 6 * i;
            if ((lane < D && has_parent) || lane == D) {
                *reinterpret_cast<double2 *>(dst) = make_double2(fw[0], fw[1]);
                if (D == 6) {
                    *reinterpret_cast<double2 *>(dst + 2) = make_double2(fw[2], fw[3]);
                    *reinterpret_cast<double2 *>(dst + 4) = make_double2(fw[D - 2], fw[D - 1]);
                } else {
                    dst[2] = fw[2];
                }
            }
            if (lane == D + 1) { /* L_i for the helper warp */
                double *Lo = Lst + 22 * (i & 1);
#pragma unroll
                for (int k = 0; k + 1 < NS; k += 2) *reinterpret_cast<double2 *>(Lo + k) = make_double2(L[k], L[k + 1]);
                if (NS & 1) Lo[NS - 1] = L[NS - 1];
            }
        }
        F_TICK(2);
        bar_named(bar_id, 64); /* G_i, z_i (the parent's step reads them) and L_i (the helper's) are visible */
        F_TICK(3);
    }
    return ok;
}

/* F, helper warp of a candidate: behind every elimination step the backward substitutions
 * M_i = L_i^-T G_i^T (lanes 0..5, a column each) and c_i = L_i^-T z_i (lane 6), which nothing in the
 * elimination waits for */
template <bool CHAIN, int D>
UWBGO_DI void factor_helper(const WinSm &sm, double *cd, const double *Lst, const int bar_id, const int N, const int lane)
{
    double *G = cd, *Mm = cd + 36 * N, *cv = cd + 72 * N, *zv = cd + 78 * N;
    for (int i = N - 1; i >= 0; --i) {
        bar_named(bar_id, 64);
        const bool has_parent = CHAIN ? true : sm.parent[i] >= 0;
        if (lane <= D && (lane == D || has_parent)) {
            constexpr int NS = D * (D + 1) / 2;
            const double *Li = Lst + 22 * (i & 1);
            const double *src = lane < D ? G + 36 * i + 6 * lane : zv + 6 * i;
            double L[NS], fw[D], bw[D];
#pragma unroll
            for (int k = 0; k < NS; ++k) L[k] = Li[k];
#pragma unroll
            for (int k = 0; k < D; ++k) fw[k] = src[k];
#pragma unroll
            for (int rr = 0; rr < D; ++rr) {
                const int r = D - 1 - rr;
                double s = fw[r];
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (k > r) s = fma(-L[lo_idx(k, r)], bw[k], s);
                bw[r] = s * L[lo_idx(r, r)];
            }
            if (lane < D) {
#pragma unroll
                for (int k = 0; k < D; ++k) Mm[36 * i + 6 * k + lane] = bw[k];
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) cv[6 * i + k] = bw[k];
            }
        }
    }
    __syncwarp();
}

/* BLOCK-DIAGONAL windows (WinCtl::bd): range edges without lever arms plus EdgeSE3Prior edges whose information
 * has zero translation / rotation cross blocks (what Localization builds for IMU and lidar, localization.cpp:478-479,
 * 515-518), on a simple chain.  Every H_ii is then diag(T_i, R_i) up to exact zeros and H_{i-1,i} has translation
 * entries only: the translation blocks form the chain, the rotation blocks are not coupled at all.  The two halves
 * of the main warp run the SAME 3x3 instruction stream side by side -- lanes 0..15 on T_i (with the child update),
 * lanes 16..31 on R_i (whose child update reads the zero pad) -- so a pose costs three dependent pivots instead of
 * six.  Entry by entry the operations are those of the 6x6 elimination minus terms with an exact-zero operand:
 * the same bits.  M_i keeps zeros outside its translation block (zeroed once), so the substitution and
 * computeScale() run unchanged on all six rows. */
template <class MATH>
UWBGO_DI bool factor_main_bd(const WinSm &sm, double *cd, double *Lst, const int bar_id, const int N, const double lam,
                             const int lane, unsigned &bad)
{
    double *G = cd, *zv = cd + 78 * N, *Sst = cd + 90 * N;
    const int hb = lane >> 4, l = lane & 15; /* half: 0 = translation block, 1 = rotation block */
    int er_ = 0, ec_ = 0;
    if (l < 6) {
        int kk = l;
        while (kk > er_) {
            kk -= er_ + 1;
            ++er_;
        }
        ec_ = kk;
    } else if (l < 9) {
        er_ = l - 6;
    }
    const int hd_idx = up_idx(6, ec_ + 3 * hb, er_ + 3 * hb);
    const int pad_off = sm.cand_stride - 42;
    double *Sh = Sst + 12 * hb;
    bool ok = true;
    auto fetch = [&](int i, double &v) {
        v = 0.0;
        if (l < 6)
            v = sm.Hd[21 * i + hd_idx];
        else if (l < 9)
            v = sm.b[6 * i + er_ + 3 * hb];
    };
    double vn;
    fetch(N - 1, vn);
    for (int i = N - 1; i >= 0; --i) {
        double v = vn;
        if (l < 6 && er_ == ec_) v = v + lam;
        {
            /* child i + 1 (translation half); the rotation half and the newest pose read zeros */
            const bool upd = hb == 0 && i < N - 1;
            const int gch = upd ? 36 * (i + 1) : pad_off, zch = upd ? 78 * N + 6 * (i + 1) : pad_off + 36;
            const double2 *Ga = reinterpret_cast<const double2 *>(cd + gch + 6 * er_);
            const double2 *Yb = reinterpret_cast<const double2 *>(cd + (l < 6 ? gch + 6 * ec_ : zch));
            const double2 a0 = Ga[0], a1 = Ga[1], y0 = Yb[0], y1 = Yb[1];
            v = fma(-a0.x, y0.x, v);
            v = fma(-a0.y, y0.y, v);
            v = fma(-a1.x, y1.x, v);
        }
        if (l < 9) Sh[l] = v;
        __syncwarp();
        double L[6], rhs[3], fw[3];
        {
            const double *src = (hb == 0 && l < 3) ? sm.Ho + 36 * i + 6 * l : Sh + 6;
#pragma unroll
            for (int k = 0; k < 3; ++k) rhs[k] = src[k];
        }
        fetch(i > 0 ? i - 1 : 0, vn);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double s = Sh[lo_idx(j, j)];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < j) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
            if (!(s > 0.0)) ok = false;
            const double inv = MATH::rsqrt_pivot(s, bad);
            L[lo_idx(j, j)] = inv;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                if (r > j) {
                    double t = Sh[lo_idx(r, j)];
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (k < j) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
                    L[lo_idx(r, j)] = t * inv;
                }
        }
        /* forward substitution: rows of G_i (translation half, lanes 0..2), z_i (lane 3 of either half) */
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
            double s = rhs[cc];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (k < cc) s = fma(-fw[k], L[lo_idx(cc, k)], s);
            fw[cc] = s * L[lo_idx(cc, cc)];
        }
        {
            double *dst = l < 3 ? G + 36 * i + 6 * l : zv + 6 * i + 3 * hb;
            if ((hb == 0 && l < 3) || l == 3) {
                dst[0] = fw[0];
                dst[1] = fw[1];
                dst[2] = fw[2];
            }
            if (l == 4) { /* L_i of this half for the helper warp */
                double *Lo = Lst + 22 * (i & 1) + 6 * hb;
#pragma unroll
                for (int k = 0; k < 6; ++k) Lo[k] = L[k];
            }
        }
        bar_named(bar_id, 64);
    }
    return ok;
}

/* helper warp: M_i = L_T^-T G_i^T (lanes 0..2, a column each), c_i = L^-T z_i of each half (lanes 3 and 19) */
UWBGO_DI void factor_helper_bd(double *cd, const double *Lst, const int bar_id, const int N, const int lane)
{
    double *G = cd, *Mm = cd + 36 * N, *cv = cd + 72 * N, *zv = cd + 78 * N;
    const int hb = lane >> 4, l = lane & 15;
    for (int i = N - 1; i >= 0; --i) {
        bar_named(bar_id, 64);
        if ((hb == 0 && l < 3) || l == 3) {
            const double *Li = Lst + 22 * (i & 1) + 6 * hb;
            const double *src = l < 3 ? G + 36 * i + 6 * l : zv + 6 * i + 3 * hb;
            double L[6], fw[3], bw[3];
#pragma unroll
            for (int k = 0; k < 6; ++k) L[k] = Li[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) fw[k] = src[k];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const int r = 2 - rr;
                double s = fw[r];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (k > r) s = fma(-L[lo_idx(k, r)], bw[k], s);
                bw[r] = s * L[lo_idx(r, r)];
            }
            if (l < 3) {
#pragma unroll
                for (int k = 0; k < 3; ++k) Mm[36 * i + 6 * k + l] = bw[k];
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) cv[6 * i + 3 * hb + k] = bw[k];
            }
        }
    }
    __syncwarp();
}

/* T3: windows whose 6x6 blocks are zero outside the translation entries (WinIo::t3 says when): the H phase, the
 * elimination and the substitution run on the 3x3 blocks (factor_main), and the 6-D edge code is not compiled in */
template <int KS, int NT, bool T3>
__global__ void __launch_bounds__(NT, 1)
lm_window_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                 const __grid_constant__ WinIo io)
{
    static_assert(KS >= 1 && KS <= WIN_MAX_KS && 2 * KS * 32 <= NT, "one main and one helper warp per candidate");
    extern __shared__ __align__(16) double wsm[];
    __shared__ WinCtl ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef UWBGO_WIN_PAD
    const int64_t w = blockIdx.x % io.W; /* experiment: redundant CTAs keep the other SMs busy */
    const bool writer = blockIdx.x < io.W;
#else
    const int64_t w = blockIdx.x;
    const bool writer = true;
#endif
    constexpr bool NO6 = NO6D || T3; /* no EdgeSE3Prior / EdgeSE3 in a T3 window */
    constexpr int DB = T3 ? 3 : 6;    /* block dimension of the linear solve */
    const int N = tp.N, NE = tp.E, mod = cfg.orth_mod;
    const WinCarve cv = win_carve(tp, KS);
    WinSm sm;
    sm.X = wsm + cv.X; sm.xstride = N * 12;
    sm.anch = wsm + cv.anch; sm.ant = wsm + cv.ant; sm.rd = wsm + cv.rd; sm.ri = wsm + cv.ri;
    sm.pZi = wsm + cv.pZi; sm.pI = wsm + cv.pI; sm.sZi = wsm + cv.sZi; sm.sI = wsm + cv.sI;
    sm.rJ = wsm + cv.rJ; sm.pJ = wsm + cv.pJ; sm.sJ = wsm + cv.sJ;
    sm.Hd = wsm + cv.Hd; sm.Ho = wsm + cv.Ho; sm.b = wsm + cv.b;
    sm.cand = wsm + cv.cand; sm.cand_stride = cv.cand_stride;
    sm.cnt = reinterpret_cast<int *>(wsm + cv.cnt);
    sm.parent = reinterpret_cast<int *>(wsm + cv.parent);
    sm.cbeg = reinterpret_cast<int *>(wsm + cv.cbeg);
    sm.chl = reinterpret_cast<int *>(wsm + cv.chl);
    sm.opb = reinterpret_cast<int *>(wsm + cv.opb);
    sm.ncalls = reinterpret_cast<int *>(wsm + cv.ncalls);
    sm.slot_edge = reinterpret_cast<int *>(wsm + cv.slot_edge);
    sm.ops = reinterpret_cast<int2 *>(wsm + cv.ops);
    sm.edges = reinterpret_cast<EdgeRec *>(wsm + cv.edges);
    Cauchy ck;
    ck.init(cfg.kdelta);
    const double delta = cfg.jdelta, scalar = 1.0 / (2.0 * delta);
    auto xbuf = [&](int k) { return sm.X + k * sm.xstride; };
    auto cand = [&](int k) { return sm.cand + k * sm.cand_stride; };
    auto cand_x = [&](int k) { return cand(k) + 84 * N; };
    auto cand_L = [&](int k) { return cand(k) + 90 * N + 28; };
    auto cand_echi = [&](int k) { return cand(k) + 90 * N + 72; };

    /* ---- the window and its topology into shared memory -------------------------------------- */
    for (int k = tid; k < N * 12; k += NT) {
        const int i = k / 12, j = k - 12 * i;
        double v;
        if (j >= 9)
            v = io.pose_t[(w * N + i) * 3 + (j - 9)];
        else if (io.pose_R)
            v = io.pose_R[(w * N + i) * 9 + j];
        else
            v = (j == 0 || j == 4 || j == 8) ? 1.0 : 0.0;
        sm.X[k] = v;
    }
    for (int k = tid; k < N; k += NT) {
        sm.cnt[k] = io.cnt_in ? io.cnt_in[w * N + k] : 0;
        sm.parent[k] = __ldg(tp.parent + k);
        sm.ncalls[k] = __ldg(tp.num_calls + k);
    }
    for (int k = tid; k <= N; k += NT) {
        sm.cbeg[k] = __ldg(tp.child_begin + k);
        sm.opb[k] = __ldg(tp.op_begin + k);
    }
    {
        const int nch = __ldg(tp.child_begin + N);
        for (int k = tid; k < nch; k += NT) sm.chl[k] = __ldg(tp.children + k);
        for (int k = tid; k < tp.n_ops; k += NT) sm.ops[k] = __ldg(reinterpret_cast<const int2 *>(tp.ops + k));
        for (int k = tid; k < NE; k += NT) sm.slot_edge[k] = __ldg(tp.slot_edge + k);
        const int words = (int)(sizeof(EdgeRec) / 4) * NE;
        const int *src = reinterpret_cast<const int *>(tp.edges);
        int *dst = reinterpret_cast<int *>(sm.edges);
        for (int k = tid; k < words; k += NT) dst[k] = __ldg(src + k);
    }
    for (int k = tid; k < tp.A * 3; k += NT) sm.anch[k] = io.anchors[w * tp.A * 3 + k];
    for (int k = tid; k < tp.K * 3; k += NT) sm.ant[k] = io.ant[k];
    for (int k = tid; k < tp.Er; k += NT) {
        sm.rd[k] = io.rd[w * tp.Er + k];
        sm.ri[k] = io.ri[w * tp.Er + k];
    }
    for (int k = tid; k < tp.Ep * 36; k += NT) sm.pI[k] = io.pI[w * tp.Ep * 36 + k];
    for (int k = tid; k < tp.Es * 36; k += NT) sm.sI[k] = io.sI[w * tp.Es * 36 + k];
    for (int k = tid; k < KS * 42; k += NT) sm.cand[(k / 42) * sm.cand_stride + sm.cand_stride - 42 + k % 42] = 0.0;
    for (int k = tid; k < tp.Ep + tp.Es; k += NT) { /* measurement inverses, once */
        const bool pr = k < tp.Ep;
        const int s = pr ? k : k - tp.Ep;
        const double *zp = pr ? io.pZ + (w * tp.Ep + s) * 12 : io.sZ + (w * tp.Es + s) * 12;
        Pose Z, Zinv;
        ld_pose(zp, Z);
        pose_inv(Z, Zinv);
        st_pose((pr ? sm.pZi : sm.sZi) + 12 * s, Zinv);
    }
    __syncthreads();
    /* block-diagonal window?  The structure (io.bd_ok: range edges without lever arms and priors only, a simple
     * chain) is the host's; the data condition is that no prior's information couples translation and rotation */
    bool bd = false;
    if (!T3 && io.bd_ok) { /* (uniform over the CTA) */
        int clean = 1;
        for (int k = tid; k < tp.Ep * 18; k += NT) {
            const int s = k / 18, j = k - 18 * s, r = (j / 3) % 3, c = j % 3;
            const double v = sm.pI[36 * s + (j < 9 ? 6 * r + 3 + c : 6 * (3 + r) + c)];
            if (!(v == 0.0)) clean = 0;
        }
        bd = __syncthreads_and(clean) != 0;
        if (bd) { /* M keeps zeros outside its translation block: only that block is ever written */
            for (int k = tid; k < KS * 36 * N; k += NT) sm.cand[(k / (36 * N)) * sm.cand_stride + 36 * N + k % (36 * N)] = 0.0;
            /* rotation columns of the range Jacobians: zeros, never recomputed (J phase) */
            for (int k = tid; k < 6 * tp.Er; k += NT) sm.rJ[RJ * (k / 6) + (k % 6 < 3 ? 3 + k % 6 : 6 + k % 6)] = 0.0;
            __syncthreads();
        }
    }

    const int n6 = NO6 ? 0 : tp.Es + tp.Ep;
    /* chi2 terms of the 6-D edges (se3 slots, then prior slots) resp. the range edges at the estimates X,
     * over the lanes of one warp */
    /* (branch-free arithmetic first; an item whose operands it flags is evaluated again with the IEEE
     * sequences: same bits, and the cold copy stays out of the instruction stream) */
    auto chi_six = [&](const double *X, double *echi) {
        if (NO6) return;
        for (int k = lane; k < n6; k += 32) {
            const int e = sm.slot_edge[tp.Er + (k < tp.Es ? tp.Ep + k : k - tp.Es)];
            double chi, rob;
            unsigned bad = 0;
            six_chi<WMC>(sm, X, sm.edges[e], ck, chi, rob, bad);
#ifdef UWBGO_WIN_TIMING
            atomicAdd(&g_n_chi6, 1ull);
            if (bad) atomicAdd(&g_bad_chi6, 1ull);
#endif
            if (bad) six_chi<IeeeMath>(sm, X, sm.edges[e], ck, chi, rob, bad);
            echi[2 * e] = chi;
            echi[2 * e + 1] = rob;
        }
    };
    auto chi_range = [&](const double *X, double *echi) {
        for (int k = lane; k < tp.Er; k += 32) {
            const int e = sm.slot_edge[k];
            double chi, rob;
            unsigned bad = 0;
            range_chi<WMC>(sm, X, sm.edges[e], ck, chi, rob, bad);
            if (bad) range_chi<IeeeMath>(sm, X, sm.edges[e], ck, chi, rob, bad);
            echi[2 * e] = chi;
            echi[2 * e + 1] = rob;
        }
    };
    /* the two ordered sums over the edges: loads in blocks of 8 edges, then the additions */
    auto chi_sum = [&](const double *echi, double &p, double &r) {
        double pp = 0.0, rr = 0.0;
        int e = 0;
        for (; e + 8 <= NE; e += 8) {
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = echi[2 * e + k];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                pp = pp + v[2 * k];
                rr = rr + v[2 * k + 1];
            }
        }
        for (; e < NE; ++e) {
            pp = pp + echi[2 * e];
            rr = rr + echi[2 * e + 1];
        }
        p = pp;
        r = rr;
    };

    /* LM state (thread 0) */
    double lambda = 0.0, ni = 2.0, stale = 0.0, plainCur = 0.0, currentChi = 0.0, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0;
    bool need_lin = true;
    /* candidates of the next round: lambda, lambda nu, lambda nu 2nu, ... as many as trials are left */
    auto publish_candidates = [&]() {
        double l = lambda, n = ni;
        int nk = cfg.max_trials - q;
        if (nk > KS) nk = KS;
        for (int k = 0; k < nk; ++k) {
            ctl.lam[k] = l;
            l = l * n;
            n = n * 2.0;
        }
        ctl.nk = nk;
    };

    /* initial computeActiveErrors at buffer 0 */
    if (warp == 0) chi_six(xbuf(0), cand_echi(0));
    else if (warp == 1) chi_range(xbuf(0), cand_echi(0));
    __syncthreads();
    if (tid == 0) {
        chi_sum(cand_echi(0), plainCur, currentChi);
        stale = plainCur;
        ctl.chi_cur = currentChi;
        ctl.go = cfg.max_iterations > 0;
        ctl.lin = 1;
        ctl.cur = 0;
        ctl.adv = 0;
        ctl.first = 1;
    }

#ifdef UWBGO_WIN_TIMING
    __shared__ long long dbg[8];
    if (tid < 8) dbg[tid] = 0;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = clock64(), tfac[4] = {0, 0, 0, 0};
    int rounds = 0;
#define WIN_TICK(k) do { long long tn_ = clock64(); tph[k] += tn_ - tq; tq = tn_; } while (0)
#else
#define WIN_TICK(k)
#endif
    while (true) {
        __syncthreads(); /* control published */
        WIN_TICK(7);
        if (!ctl.go) break;
#ifdef UWBGO_WIN_TIMING
        ++rounds;
#endif
        const int cb = ctl.cur;
        const double *X = xbuf(cb);
        if (ctl.lin) {
            const int adv = ctl.adv; /* oplus calls of the trials consumed since the counters were last touched */
            /* ---- J: Jacobians, errors, weights (branch-free arithmetic, IEEE on a flagged item) ---- */
            {
                for (int k = tid; k < n6; k += NT) {
                    const EdgeRec er = sm.edges[sm.slot_edge[tp.Er + (k < tp.Es ? tp.Ep + k : k - tp.Es)]];
                    unsigned bad = 0;
                    lin_six<WMJ>(sm, X, er, ck, bad);
#ifdef UWBGO_WIN_TIMING
                    atomicAdd(&g_n_lin6, 1ull);
                    if (bad) atomicAdd(&g_bad_lin6, 1ull);
#endif
                    if (bad) lin_six<IeeeMath>(sm, X, er, ck, bad);
                }
                const int base1 = ((n6 + 31) & ~31) % NT;
                for (int k = (tid - base1 + NT) % NT; k < tp.Er; k += NT) {
                    const EdgeRec er = sm.edges[sm.slot_edge[k]];
                    unsigned bad = 0;
                    lin_range_weights<WMJ>(sm, X, er, ck, bad);
                    if (bad) lin_range_weights<IeeeMath>(sm, X, er, ck, bad);
                }
                const int base2 = (base1 + ((tp.Er + 31) & ~31)) % NT;
                /* translation-only and block-diagonal windows: the rotation columns of a range edge (no lever arms)
                 * are (e - e) / 2 delta = +0 -- the perturbed point IS the point -- and they are the longest items
                 * of the phase.  T3 never reads them; block-diagonal windows hold zeros there, written once */
                const bool tcols = T3 || bd;
                const int per = tcols ? 6 : 12;
                for (int u = (tid - base2 + NT) % NT; u < per * tp.Er; u += NT) {
                    const int k = u / per, jj = u - per * k, j = (tcols && jj >= 3) ? jj + 3 : jj;
                    const EdgeRec er = sm.edges[sm.slot_edge[k]];
                    unsigned bad = 0;
                    lin_range_col<WMJ>(sm, X, er, j, adv, mod, delta, scalar, bad);
                    if (bad) lin_range_col<IeeeMath>(sm, X, er, j, adv, mod, delta, scalar, bad);
                }
            }
            __syncthreads();
            WIN_TICK(0);
            /* ---- O: J^T Ow of the 6-D edges, one entry per thread; oplus counters advance ----- */
            {
                const int nP = NO6 ? 0 : 36 * tp.Ep, nS = NO6 ? 0 : 72 * tp.Es;
                for (int u = tid; u < nP + nS; u += NT) {
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
                        robust = sm.edges[sm.slot_edge[tp.Er + tp.Ep + s]].robust != 0;
                    } else {
                        const int v = u - nS, s = v / 36;
                        rc = v - 36 * s;
                        const double *rec = sm.pJ + PJ * s;
                        J = rec;
                        out = sm.pJ + PJ * s + 36;
                        O = sm.pI + 36 * s;
                        r1 = rec[78];
                        robust = sm.edges[sm.slot_edge[tp.Er + s]].robust != 0;
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
                /* the trials consumed so far made one oplus call each, the numeric Jacobians ncalls[i]
                 * (all columns were read before the barrier above) */
                for (int i = tid; i < N; i += NT) sm.cnt[i] = (sm.cnt[i] + adv + sm.ncalls[i]) % mod;
            }
            __syncthreads();
            WIN_TICK(1);
            /* ---- H: every entry of the H record of every pose, owned by one thread -------------- */
            constexpr int HN = T3 ? 18 : 63, HU = T3 ? 6 : 21, HO = DB * DB;
            /* block-diagonal windows: 12 + 9 + 6 of the 63 entries per pose (the others are exact zeros nobody reads) */
            const int hn = bd ? 27 : HN;
            /* entry-major: the threads of a warp own the SAME entry of consecutive poses, so that they walk the
             * same kind of term through the same branches (pose-major, a warp mixed diagonal, off-diagonal and
             * right-hand-side entries and ran each path in turn) */
            for (int u = tid; u < hn * N; u += NT) {
                int k = u / N;
                const int i = u - N * k;
                if (bd) k = BD_ENTRY[k];
                /* k < HU: H_ii upper (r, c); then H_{parent(i), i} (r, c); the last DB: b_i[r].  (T3: the entries of
                 * the translation blocks; the others are exact zeros nobody reads) */
                int r, c, what;
                if (k < HU) {
                    what = 0;
                    r = T3 ? (k < 3 ? 0 : (k < 5 ? 1 : 2)) : UP_R[k];
                    c = T3 ? (k < 3 ? k : (k < 5 ? k - 2 : 2)) : UP_C[k];
                } else if (k < HU + HO) {
                    what = 1;
                    r = (k - HU) / DB;
                    c = (k - HU) - DB * r;
                } else {
                    what = 2;
                    r = k - HU - HO;
                    c = 0;
                }
                double acc = 0.0;
                const int ob = sm.opb[i], oe = sm.opb[i + 1];
                for (int o = ob; o < oe; ++o) {
                    const int2 op = sm.ops[o];
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
                    } else if (!NO6) {
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
                    sm.Hd[21 * i + up_idx(6, r, c)] = acc;
                else if (what == 1)
                    sm.Ho[36 * i + 6 * r + c] = acc;
                else
                    sm.b[6 * i + r] = acc;
            }
            __syncthreads();
            WIN_TICK(2);
            if (tid == 0) { /* a new iteration starts */
                stale = plainCur;
                if (it == 0) {
                    double maxdiag = 0.0;
                    for (int i = 0; i < N; ++i)
#pragma unroll
                        for (int r = 0; r < DB; ++r) { /* (T3: the rotation diagonal is +0) */
                            const double v = fabs(sm.Hd[21 * i + up_idx(6, r, r)]);
                            if (v > maxdiag) maxdiag = v;
                        }
                    lambda = cfg.tau * maxdiag;
                    ni = 2.0;
                }
                rho = 0.0;
                q = 0;
                need_lin = false;
                if (ctl.first) publish_candidates();
            }
            if (ctl.first) __syncthreads(); /* lambda_0 needed H; later rounds publish their candidates in D */
        }
        const int adv0 = ctl.lin ? 0 : ctl.adv; /* trials consumed since the counters were last advanced */
        const int nk = ctl.nk;

        /* ---- F: the linear solves of the candidates, a pair of warps each ----------------------- */
        if (warp < nk) { /* main: the elimination chain */
            double *cd = cand(warp);
            const double lam = ctl.lam[warp];
            unsigned bad = 0;
#ifdef UWBGO_WIN_TIMING
#define F_TARG , tfac
#else
#define F_TARG
#endif
            const bool chain = tp.simple_chain != 0;
            bool ok;
            if (!T3 && bd)
                ok = factor_main_bd<NbMath>(sm, cd, cand_L(warp), 1 + warp, N, lam, lane, bad);
            else
                ok = (T3 || chain) ? factor_main<NbMath, true, DB>(sm, cd, cand_L(warp), 1 + warp, N, lam, lane, bad F_TARG)
                                   : factor_main<NbMath, false, DB>(sm, cd, cand_L(warp), 1 + warp, N, lam, lane, bad F_TARG);
            const bool redo = __any_sync(0xffffffffu, bad != 0);
            if (lane == 0) ctl.redo[warp] = redo ? 1 : 0;
            bar_named(1 + warp, 64);
            if (redo) { /* operands outside the branch-free range: the IEEE sequences */
                bad = 0;
                if (!T3 && bd)
                    ok = factor_main_bd<IeeeMath>(sm, cd, cand_L(warp), 1 + warp, N, lam, lane, bad);
                else
                    ok = factor_main<IeeeMath, false, DB>(sm, cd, cand_L(warp), 1 + warp, N, lam, lane, bad F_TARG);
            }
            ok = __all_sync(0xffffffffu, ok);
            if (lane == 0) ctl.ok[warp] = ok ? 1 : 0;
            bar_named(1 + warp, 64);
            WIN_TICK(3);
        } else if (warp >= KS && warp - KS < nk) { /* helper: backward substitutions, then x and computeScale() */
            const int k = warp - KS;
            double *cd = cand(k);
            if (!T3 && bd) factor_helper_bd(cd, cand_L(k), 1 + k, N, lane);
            else if (T3 || tp.simple_chain) factor_helper<true, DB>(sm, cd, cand_L(k), 1 + k, N, lane);
            else factor_helper<false, DB>(sm, cd, cand_L(k), 1 + k, N, lane);
            bar_named(1 + k, 64);
            if (ctl.redo[k]) {
                if (!T3 && bd) factor_helper_bd(cd, cand_L(k), 1 + k, N, lane);
                else factor_helper<false, DB>(sm, cd, cand_L(k), 1 + k, N, lane);
            }
            bar_named(1 + k, 64);
            const bool ok = ctl.ok[k] != 0;
            const double lam = ctl.lam[k];
#ifdef UWBGO_WIN_TIMING
            const long long ts0 = clk();
#endif
            /* x_i = c_i - M_i x_parent(i), ascending, and computeScale() = sum over j of x_j (lambda x_j + b_j)
             * strictly in order.  Lanes 0..5 own a row each; x_(i-1) travels between them by shuffle (every lane
             * keeps all six), the row of M_(i+1) is fetched a pose ahead, and lane 0 adds the six terms of pose
             * i - 1 in the shadow of pose i's FMA chain. */
            {
                const double *Mm = cd + 36 * N, *cvv = cd + 72 * N;
                double *xv = cd + 84 * N;
                const int r = lane < DB ? lane : 0;
                double scale = 0.0, tprev = 0.0;
                double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                double mrow[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, ci, bi;
                int par;
                auto fetch = [&](int i) {
                    const double2 *mp = reinterpret_cast<const double2 *>(Mm + 36 * i + 6 * r);
                    const double2 m0 = mp[0], m1 = mp[1];
                    mrow[0] = m0.x; mrow[1] = m0.y; mrow[2] = m1.x;
                    if (DB == 6) {
                        const double2 m2 = mp[2];
                        mrow[3] = m1.y; mrow[4] = m2.x; mrow[5] = m2.y;
                    }
                    ci = cvv[6 * i + r];
                    bi = sm.b[6 * i + r];
                    par = sm.parent[i];
                };
                fetch(0);
                if (tp.simple_chain) {
                    /* chains: parent(i) = i - 1.  No branch in the body (a lone warp pays for every one):
                     * the fetch index is clamped, pose 0 multiplies zeros */
#pragma unroll 2
                    for (int i = 0; i < N; ++i) {
                        double x = ci;
                        const double b_i = bi;
                        if (i > 0) {
#pragma unroll
                            for (int j = 0; j < DB; ++j) x = fma(-mrow[j], xp[j], x);
                        }
                        fetch(i + 1 < N ? i + 1 : N - 1);
#pragma unroll
                        for (int j = 0; j < DB; ++j) scale = scale + __shfl_sync(0xffffffffu, tprev, j); /* pose i - 1 */
                        x = ok ? x : 0.0;
                        if (lane < DB) xv[6 * i + lane] = x;
                        else if (T3 && lane < 6) xv[6 * i + lane] = 0.0; /* the rotation increment: zero */
                        tprev = x * (lam * x + b_i);
#pragma unroll
                        for (int j = 0; j < DB; ++j) xp[j] = __shfl_sync(0xffffffffu, x, j);
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < N; ++i) {
                        double x = ci;
                        const double b_i = bi;
                        const int p_i = par;
                        if (p_i >= 0 && p_i != i - 1) { /* forests: the parent is not the previous pose */
                            __syncwarp();
#pragma unroll
                            for (int j = 0; j < DB; ++j) xp[j] = xv[6 * p_i + j];
                        }
                        if (p_i >= 0) {
#pragma unroll
                            for (int j = 0; j < DB; ++j) x = fma(-mrow[j], xp[j], x);
                        }
                        fetch(i + 1 < N ? i + 1 : N - 1);
#pragma unroll
                        for (int j = 0; j < DB; ++j) scale = scale + __shfl_sync(0xffffffffu, tprev, j); /* pose i - 1 */
                        x = ok ? x : 0.0;
                        if (lane < DB) xv[6 * i + lane] = x;
                        else if (T3 && lane < 6) xv[6 * i + lane] = 0.0; /* the rotation increment: zero */
                        tprev = x * (lam * x + b_i);
#pragma unroll
                        for (int j = 0; j < DB; ++j) xp[j] = __shfl_sync(0xffffffffu, x, j);
                    }
                }
#pragma unroll
                for (int j = 0; j < DB; ++j) scale = scale + __shfl_sync(0xffffffffu, tprev, j);
                if (lane == 0) ctl.scale[k] = scale;
            }
            __syncwarp();
#ifdef UWBGO_WIN_TIMING
            if (warp == KS && lane == 0) dbg[7] += clk() - ts0;
#endif
        }
#ifdef UWBGO_WIN_TIMING
        const long long ta0 = clk();
#endif
        __syncthreads();
#ifdef UWBGO_WIN_TIMING
        if (warp == KS && lane == 0) dbg[3] += clk() - ta0;
#endif
        WIN_TICK(6);

        /* ---- U and C per candidate: main warp updates the poses and evaluates the 6-D edges, its helper
         * the range edges ------------------------------------------------------------------------------ */
#ifdef UWBGO_WIN_TIMING
        long long t0_ = clk(), t1_ = t0_, t2_ = t0_, t3_ = t0_;
#endif
        if (warp < nk) {
            const int k = warp;
            double *Xn = xbuf((cb + 1 + k) % (KS + 1));
            const double *xv = cand_x(k);
#ifdef UWBGO_WIN_TIMING
            long long u0 = clk(), u1 = u0, u2 = u0, u3 = u0;
#endif
            for (int i = lane; i < N; i += 32) {
                Pose P;
                ld_pose(X + 12 * i, P);
                double dx[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) dx[j] = xv[6 * i + j];
#ifdef UWBGO_WIN_TIMING
                u1 = clk();
#endif
                int c0 = sm.cnt[i] + adv0 + k; /* cnt < mod, a handful of trials since: no division */
                while (c0 >= mod) c0 -= mod;
                int c = c0;
#ifdef UWBGO_WIN_TIMING
                if (c < 0) dx[0] = 0.0;
                u2 = clk();
#endif
                unsigned bad = 0;
                {
                    Pose Q = P;
                    pose_oplus_m<WMU>(Q, dx, c, mod, bad);
                    if (bad) {
                        c = c0;
                        pose_oplus_m<IeeeMath>(P, dx, c, mod, bad);
                    } else
                        P = Q;
                }
#ifdef UWBGO_WIN_TIMING
                if (P.R[0] == 12345.0) dx[0] = 0.0;
                u3 = clk();
#endif
                st_pose(Xn + 12 * i, P);
            }
#ifdef UWBGO_WIN_TIMING
            t1_ = clk();
            if (tid == 0) { dbg[4] += u0 - t0_; }
#endif
            bar_named(1 + k, 64);
#ifdef UWBGO_WIN_TIMING
            t2_ = clk();
#endif
            chi_six(Xn, cand_echi(k));
#ifdef UWBGO_WIN_TIMING
            t3_ = clk();
#endif
        } else if (warp >= KS && warp - KS < nk) {
            const int k = warp - KS;
#ifdef UWBGO_WIN_TIMING
            t1_ = clk();
#endif
            bar_named(1 + k, 64);
#ifdef UWBGO_WIN_TIMING
            t2_ = clk();
#endif
            chi_range(xbuf((cb + 1 + k) % (KS + 1)), cand_echi(k));
#ifdef UWBGO_WIN_TIMING
            t3_ = clk();
#endif
        }
#ifdef UWBGO_WIN_TIMING
        if (lane == 0 && (warp == 0 || warp == KS)) {
            if (warp == 0) { dbg[0] += t1_ - t0_; dbg[1] += t2_ - t1_; dbg[2] += t3_ - t2_; }
            else { dbg[5] += t2_ - t1_; dbg[6] += t3_ - t2_; }
        }
#endif
        __syncthreads();
        WIN_TICK(4);
        /* the ordered chi2 sums and the gain ratio of every candidate, side by side (the current chi2 is the same
         * for all of them: an accepted candidate ends the round) */
        if (lane == 0 && warp < nk) {
            double tp_, tc_;
            chi_sum(cand_echi(warp), tp_, tc_);
            if (ctl.ok[warp] == 0) tc_ = DBL_MAX;
            ctl.tplain[warp] = tp_;
            ctl.tchi[warp] = tc_;
            ctl.rho[warp] = (ctl.chi_cur - tc_) / (ctl.scale[warp] + 1e-3);
        }
        __syncthreads();

        /* ---- D: the trials of this round, consumed in order --------------------------------------------- */
        if (tid == 0) {
            int used = 0;
            bool done = false, more = true;
            while (more) {
                const int k = used;
                const bool ok = ctl.ok[k] != 0;
                if (!ok) flags |= UWBGO_FLAG_CHOL_FAIL;
                const double tplain = ctl.tplain[k], tempChi = ctl.tchi[k];
                stale = tplain;
                rho = ctl.rho[k];
                const bool fin = isfinite(tempChi);
                if (!fin) flags |= UWBGO_FLAG_NONFINITE;
                ++used;
                if (rho > 0.0 && fin) {
                    double t = 2.0 * rho - 1.0;
                    double alpha = 1.0 - (t * t) * t;
                    alpha = (cfg.good_hi < alpha) ? cfg.good_hi : alpha;
                    double sf = (cfg.good_lo < alpha) ? alpha : cfg.good_lo;
                    lambda = lambda * sf;
                    ni = 2.0;
                    currentChi = tempChi;
                    plainCur = tplain;
                    cur = (cb + 1 + k) % (KS + 1);
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
                        q = 0;
                    }
                    more = false;
                } else if (used == nk) {
                    more = false; /* every candidate rejected: another round on the same H */
                }
            }
            ctl.go = done ? 0 : 1;
            ctl.lin = need_lin ? 1 : 0;
            ctl.cur = cur;
            ctl.adv = adv0 + used;
            ctl.first = 0;
            ctl.chi_cur = currentChi;
            if (!done) publish_candidates();
            WIN_TICK(5);
        }
    }
#ifdef UWBGO_WIN_TIMING
    if (tid == 0 && blockIdx.x == 0)
        printf("window 0 cycles: J %lld  O %lld  H %lld  factor %lld (assemble %lld potrf %lld fwd+store %lld bar %lld)  wait for x %lld  U+C %lld  D %lld  sync %lld  (trials %d, iterations %d, rounds %d)\n",
               tph[0], tph[1], tph[2], tph[3], tfac[0], tfac[1], tfac[2], tfac[3], tph[6], tph[4], tph[5], tph[7], trials_total, iterations, rounds);
    if (tid == 0 && blockIdx.x == 0)
        printf("   IEEE re-evaluations so far: chi6 %llu of %llu, lin6 %llu of %llu\n", g_bad_chi6, g_n_chi6, g_bad_lin6, g_n_lin6);
    if (tid == 0 && blockIdx.x == 0)
        printf("   main 0: U %lld (pre %lld)  bar %lld  chi6 %lld   helper 0: subst %lld  wait at A %lld  bar %lld  chi-range %lld\n", dbg[0], dbg[4], dbg[1], dbg[2], dbg[7], dbg[3], dbg[5], dbg[6]);
#endif

    /* ---- results ---------------------------------------------------------------------------------- */
    if (!writer) return;
    const double *Xf = xbuf(ctl.cur);
    for (int k = tid; k < N * 3; k += NT) {
        const int i = k / 3, j = k - 3 * i;
        io.o_pose_t[(w * N + i) * 3 + j] = Xf[12 * i + 9 + j];
    }
    if (io.o_pose_R)
        for (int k = tid; k < N * 9; k += NT) {
            const int i = k / 9, j = k - 9 * i;
            io.o_pose_R[(w * N + i) * 9 + j] = Xf[12 * i + j];
        }
    if (io.o_cnt) {
        const int adv = ctl.adv;
        for (int k = tid; k < N; k += NT) io.o_cnt[w * N + k] = (sm.cnt[k] + adv) % mod;
    }
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

}  // namespace

size_t window_path_smem_bytes(const DevTopo &topo, int ks) { return sizeof(double) * (size_t)win_carve(topo, ks).total; }

/* candidates evaluated per round: 4 while the batch leaves SMs to spare, 1 (no speculation, 64 threads)
 * when every SM has several windows to work on anyway */
int window_path_candidates(int64_t W) { return W <= 148 ? 4 : (W <= 296 ? 2 : 1); }

template <int KS, int NT, bool T3>
static cudaError_t launch_win(const DevTopo &topo, const DevCfg &cfg, const WinIo &io, int device, cudaStream_t st)
{
    const size_t sm = window_path_smem_bytes(topo, KS);
    static size_t configured[64] = {0}; /* the attribute is per device */
    size_t &conf = configured[device & 63];
    if (sm > conf) {
        cudaError_t e = cudaFuncSetAttribute(lm_window_kernel<KS, NT, T3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        conf = sm;
    }
#ifdef UWBGO_WIN_PAD
    lm_window_kernel<KS, NT, T3><<<(unsigned)(io.W < UWBGO_WIN_PAD ? UWBGO_WIN_PAD : io.W), NT, sm, st>>>(topo, cfg, io);
#else
    lm_window_kernel<KS, NT, T3><<<(unsigned)io.W, NT, sm, st>>>(topo, cfg, io);
#endif
    return cudaGetLastError();
}

cudaError_t launch_solve_window(const DevTopo &topo, const DevCfg &cfg, const WinIo &io, int ks, int device,
                                cudaStream_t st)
{
    if (io.W <= 0) return cudaSuccess;
    if (io.t3) {
        if (ks >= 4) return launch_win<4, 256, true>(topo, cfg, io, device, st);
        if (ks >= 2) return launch_win<2, 256, true>(topo, cfg, io, device, st);
        return launch_win<1, 128, true>(topo, cfg, io, device, st);
    }
    if (ks >= 4) return launch_win<4, 256, false>(topo, cfg, io, device, st);
    if (ks >= 2) return launch_win<2, 256, false>(topo, cfg, io, device, st);
    return launch_win<1, 128, false>(topo, cfg, io, device, st);
}

}  // namespace uwbgo
