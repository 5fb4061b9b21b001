/*
 * uwbgo_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the hot path of sair-lab/localization:
 *   Localization::solve()  (src/localization/localization.cpp:164-192)
 *     = g2o SparseOptimizer::initializeOptimization() + optimize(iteration_max)
 *       with BlockSolver_6_3 + OptimizationAlgorithmLevenberg + LinearSolverCholmod
 *       (src/localization/localization.h:82-85, localization.cpp:44-52)
 *   over the edges the reference creates:
 *     EdgeSE3Range::computeError           src/types/types_edge_se3range.cpp:105-114
 *     (no linearizeOplus override => g2o BaseBinaryEdge numeric central differences)
 *     EdgeSE3Prior (IMU / lidar)           localization.cpp:462-535
 *     EdgeSE3 (twist)                      localization.cpp:438-459,560-605
 *     RobustKernelCauchy on range/twist    localization.cpp:602,624
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in third-party g2o, pinned by the
 * reference at commit deafc01ee8315b9405351fb145238c5d62f82dc7 (README.md:26-33), plus
 * SuiteSparse CHOLMOD and Eigen; none of them is vendored under /root/reference and none can
 * be built offline.  The reference has no tests, golden vectors or known answers for this
 * path.  This file restates g2o's published algorithm and is checked
 * against an independent numpy restatement (oracle/oracle_np.py) and against physical truth
 * (Vicon) on bag/data_example.bag — not against output of the reference itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product (localization_b200/) never does.
 *
 * Arithmetic contract (so that the CUDA kernels can be compared bit-for-bit):
 *   IEEE-754 binary64, round-to-nearest, compiled with -ffp-contract=off; fma() is used ONLY
 *   where written explicitly (quadratic-form accumulation and the linear solver); sums run
 *   left to right in the order written; sqrt and / are correctly rounded; the natural log of
 *   the Cauchy kernel is det_log() below, not libm's.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/uwbgo.h"

/* ------------------------------------------------------------------------------------------ */
/* deterministic natural logarithm (argument-reduction + atanh series, classic fdlibm scheme)  */
/* ------------------------------------------------------------------------------------------ */
static double det_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    if (!(x > 0.0) || (bits >> 52) == 0x7ff) {      /* <=0, NaN, +inf */
        if (x == 0.0) return -HUGE_VAL;
        if (x > 0.0) return x;                       /* +inf */
        return NAN;
    }
    int32_t hx = (int32_t)(bits >> 32);
    int32_t k = 0;
    if (hx < 0x00100000) {                           /* subnormal: scale up by 2^54 */
        x *= 18014398509481984.0;
        memcpy(&bits, &x, 8);
        hx = (int32_t)(bits >> 32);
        k = -54;
    }
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;           /* mantissa >= sqrt(2): halve it */
    bits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (bits & 0xffffffffu);
    k += i >> 20;
    double m;
    memcpy(&m, &bits, 8);
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double w = z * z;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    double R = t2 + t1;
    double hfsq = 0.5 * f * f;
    double dk = (double)k;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

double uwbgo_oracle_log(double x) { return det_log(x); }

/* ------------------------------------------------------------------------------------------ */
/* SE(3) helpers: pose = R[9] row-major, t[3]  (g2o VertexSE3 estimate, an Eigen Isometry3d)    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double R[9];
    double t[3];
} pose_t;

/* Eigen Quaternion::toRotationMatrix operation order */
static void quat_to_R(double w, double x, double y, double z, double *R)
{
    double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w;
    double txx = tx * x, txy = ty * x, txz = tz * x;
    double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0 - (txx + tyy);
}

/* C = A*B for 3x3 row-major; each entry (a0*b0 + a1*b1) + a2*b2 */
static void mat3_mul(const double *A, const double *B, double *C)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            C[3 * r + c] = (A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
}

/* y = A*v + t, each entry ((a0*v0 + a1*v1) + a2*v2) + t */
static void mat3_vec_add(const double *A, const double *v, const double *t, double *y)
{
    for (int r = 0; r < 3; ++r)
        y[r] = ((A[3 * r] * v[0] + A[3 * r + 1] * v[1]) + A[3 * r + 2] * v[2]) + t[r];
}

/* Isometry product X = P*Q: R = Rp*Rq, t = Rp*tq + tp */
static void pose_mul(const pose_t *P, const pose_t *Q, pose_t *X)
{
    pose_t out;
    mat3_mul(P->R, Q->R, out.R);
    mat3_vec_add(P->R, Q->t, P->t, out.t);
    *X = out;
}

/* Isometry inverse: R' = R^T, t' = (-R^T) * t */
static void pose_inv(const pose_t *P, pose_t *X)
{
    pose_t out;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) out.R[3 * r + c] = P->R[3 * c + r];
    for (int r = 0; r < 3; ++r)
        out.t[r] = ((-out.R[3 * r]) * P->t[0] + (-out.R[3 * r + 1]) * P->t[1]) +
                   (-out.R[3 * r + 2]) * P->t[2];
    *X = out;
}

/* g2o internal::approximateNearestOrthogonalMatrix: E = R^T R - I; R -= 0.5 * R * E */
static void orthogonalize(double *R)
{
    double Rt[9], E[9], RE[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Rt[3 * r + c] = R[3 * c + r];
    mat3_mul(Rt, R, E);
    E[0] -= 1.0;
    E[4] -= 1.0;
    E[8] -= 1.0;
    mat3_mul(R, E, RE);
    for (int k = 0; k < 9; ++k) R[k] = R[k] - 0.5 * RE[k];
}

/* VertexSE3::oplusImpl: estimate = estimate * fromVectorMQT(v); SURVEY.md A.7 */
static void pose_oplus(pose_t *X, const double *v, int32_t *count, int32_t orth_after)
{
    pose_t inc;
    double n2 = (v[3] * v[3] + v[4] * v[4]) + v[5] * v[5];
    double w = 1.0 - n2;
    if (w < 0.0) {
        memset(inc.R, 0, sizeof inc.R);
        inc.R[0] = inc.R[4] = inc.R[8] = 1.0;
    } else {
        w = sqrt(w);
        quat_to_R(w, v[3], v[4], v[5], inc.R);
    }
    inc.t[0] = v[0];
    inc.t[1] = v[1];
    inc.t[2] = v[2];
    pose_mul(X, &inc, X);
    if (++(*count) > orth_after) {
        *count = 0;
        orthogonalize(X->R);
    }
}

/* Eigen Quaternion(Matrix3) (Shoemake) followed by g2o normalize(): q /= |q|; w >= 0.
 * q = {x, y, z, w} */
static void R_to_quat(const double *R, double *q)
{
    double t = (R[0] + R[4]) + R[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(((R[4 * i] - R[4 * j]) - R[4 * k]) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
    double n = sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    q[0] = q[0] / n;
    q[1] = q[1] / n;
    q[2] = q[2] / n;
    q[3] = q[3] / n;
    if (q[3] < 0.0) {
        q[0] = -q[0];
        q[1] = -q[1];
        q[2] = -q[2];
        q[3] = -q[3];
    }
}

void uwbgo_oracle_quat_to_R(const double *q_xyzw, double *R)
{
    quat_to_R(q_xyzw[3], q_xyzw[0], q_xyzw[1], q_xyzw[2], R);
}
void uwbgo_oracle_R_to_quat(const double *R, double *q_xyzw) { R_to_quat(R, q_xyzw); }

/* ------------------------------------------------------------------------------------------ */
/* edge errors                                                                                  */
/* ------------------------------------------------------------------------------------------ */

/* EdgeSE3Range::computeError (types_edge_se3range.cpp:105-114):
 *   dt = (X0 * offset0).translation() - (X1 * offset1).translation();  e = d - |dt|
 * Only the translation o of an offset isometry O reaches the residual: (X * O).translation() =
 * R_X o + t_X.  Same for EdgeSE3RangeOffset (types_edge_se3range_offset.cpp:126-131, n2w of the
 * CacheSE3Offset = estimate * offset parameter). */
static double range_error(const pose_t *X0, const double *o0, const double *P1, double d)
{
    double P0[3];
    mat3_vec_add(X0->R, o0, X0->t, P0);
    double dx = P0[0] - P1[0], dy = P0[1] - P1[1], dz = P0[2] - P1[2];
    double n = sqrt((dx * dx + dy * dy) + dz * dz);
    return d - n;
}

/* translation of (X * offset): R*o + t (exactly t for o = 0 and finite R) */
static void pose_point(const pose_t *X, const double *o, double *P)
{
    mat3_vec_add(X->R, o, X->t, P);
}

/* vertex 1 of an anchor range edge: a fixed vertex with identity rotation at the anchor position */
static void anchor_point(const double *anchor, const double *o, double *P)
{
    static const double I3[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
    mat3_vec_add(I3, o, anchor, P);
}

/* g2o internal::toVectorMQT(Isometry): [t; x,y,z of the normalised quaternion with w >= 0] */
static void to_vector_mqt(const pose_t *D, double *e, double *q_out)
{
    double q[4];
    R_to_quat(D->R, q);
    e[0] = D->t[0];
    e[1] = D->t[1];
    e[2] = D->t[2];
    e[3] = q[0];
    e[4] = q[1];
    e[5] = q[2];
    if (q_out) memcpy(q_out, q, sizeof q);
}

/* ------------------------------------------------------------------------------------------ */
/* Cauchy kernel (g2o RobustKernelCauchy::robustify), delta = cfg->kernel_delta                */
/* ------------------------------------------------------------------------------------------ */
static void cauchy(double e2, double delta, double *rho0, double *rho1)
{
    double dsqr = delta * delta;
    double dsqrReci = 1.0 / dsqr;
    double aux = dsqrReci * e2 + 1.0;
    *rho0 = dsqr * det_log(aux);
    *rho1 = 1.0 / aux;
}

/* ------------------------------------------------------------------------------------------ */
/* per-window problem                                                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const uwbgo_topology *topo;
    const uwbgo_config *cfg;
    int N, A, E, Er, Ep, Es;
    const int32_t *slot; /* [E] per-kind slot */
    /* window data */
    const double *anchors, *ant, *range_d, *range_info, *prior_Z, *prior_info, *se3_Z, *se3_info;
    /* state */
    pose_t *X, *Xbak;
    int32_t *cnt;
    double *err;  /* [E][6] last computed errors */
    double *Hd;   /* [N][36] diagonal blocks (both triangles) */
    double *Ho;   /* [N][36] Ho[j] = block (parent(j), j): rows of the parent, columns of pose j */
    const int32_t *parent; /* [N] the one older neighbour of pose j (-1: none); chains: j-1 */
    double *b;    /* [N][6] */
    double *x;    /* [N][6] */
    double *Ld;   /* [N][36] M_i of the substitution x_i = c_i - M_i x_{i-1} */
    double *Lo;   /* [N-1][36] scratch G_i */
    double *y;    /* [N][6] c_i */
    double *zs;   /* [N][6] z_i */
    double *rd_own, *ri_own; /* [Er] edge parameters built from range messages (compact input form) */
    double *pi_own, *si_own; /* [Ep][36], [Es][36] information matrices rebuilt from their diagonals (UWBGO_DIAG_INFO) */
    double *estale;          /* [E] chi2() of every edge at the last computeActiveErrors of optimize() */
    double *mg;              /* [N][36 + 36 + 36] marginal scratch: G_i | M_i | S_i^-1 */
    int built;               /* build_system() ran at least once */
} window_t;

static const double ZERO3[3] = {0.0, 0.0, 0.0};

static const double *edge_offset(const window_t *W, int e)
{
    int a = W->topo->edge_ant ? W->topo->edge_ant[e] : 0;
    return a > 0 ? W->ant + 3 * (a - 1) : ZERO3;
}
static const double *edge_offset_b(const window_t *W, int e)
{
    int a = W->topo->edge_ant_b ? W->topo->edge_ant_b[e] : 0;
    return a > 0 ? W->ant + 3 * (a - 1) : ZERO3;
}

/* error of range edge e at the current estimates */
static double range_edge_error(const window_t *W, int e)
{
    const uwbgo_topology *T = W->topo;
    double P1[3];
    if (T->edge_kind[e] == UWBGO_EDGE_RANGE_ANCHOR) {
        const double *an = W->anchors + 3 * T->edge_b[e];
        if (T->edge_ant_b && T->edge_ant_b[e] > 0)
            anchor_point(an, edge_offset_b(W, e), P1);
        else
            memcpy(P1, an, sizeof P1);
    } else
        pose_point(&W->X[T->edge_b[e]], edge_offset_b(W, e), P1);
    return range_error(&W->X[T->edge_a[e]], edge_offset(W, e), P1, W->range_d[W->slot[e]]);
}

/* EdgeSE3::computeError: delta = Zinv * Xi^-1 * Xj, evaluated left to right */
static void se3_error(const pose_t *Zinv, const pose_t *Xi, const pose_t *Xj, double *e)
{
    pose_t Xi_inv, T, D;
    pose_inv(Xi, &Xi_inv);
    pose_mul(Zinv, &Xi_inv, &T);
    pose_mul(&T, Xj, &D);
    to_vector_mqt(&D, e, NULL);
}

/* EdgeSE3Prior::computeError: delta = Zinv * (X * P), P = identity offset parameter id 0
 * (localization.cpp:54-56,485,524) */
static void prior_error(const pose_t *Zinv, const pose_t *X, double *e, pose_t *D_out, double *q_out)
{
    pose_t D;
    pose_mul(Zinv, X, &D);
    to_vector_mqt(&D, e, q_out);
    if (D_out) *D_out = D;
}

static void load_Z(const double *z12, pose_t *Z)
{
    memcpy(Z->R, z12, 9 * sizeof(double));
    memcpy(Z->t, z12 + 9, 3 * sizeof(double));
}

/* computeActiveErrors: one error per edge, in edge order */
static void compute_errors(window_t *W)
{
    const uwbgo_topology *T = W->topo;
    for (int e = 0; e < W->E; ++e) {
        double *er = W->err + 6 * e;
        int a = T->edge_a[e], s = W->slot[e];
        switch (T->edge_kind[e]) {
        case UWBGO_EDGE_RANGE_ANCHOR:
        case UWBGO_EDGE_RANGE_POSE:
            er[0] = range_edge_error(W, e);
            break;
        case UWBGO_EDGE_PRIOR: {
            pose_t Z, Zinv;
            load_Z(W->prior_Z + 12 * s, &Z);
            pose_inv(&Z, &Zinv);
            prior_error(&Zinv, &W->X[a], er, NULL, NULL);
            break;
        }
        case UWBGO_EDGE_SE3: {
            pose_t Z, Zinv;
            load_Z(W->se3_Z + 12 * s, &Z);
            pose_inv(&Z, &Zinv);
            se3_error(&Zinv, &W->X[a], &W->X[T->edge_b[e]], er);
            break;
        }
        }
    }
}

static const double *edge_info6(const window_t *W, int e)
{
    return W->topo->edge_kind[e] == UWBGO_EDGE_PRIOR ? W->prior_info + 36 * W->slot[e]
                                                      : W->se3_info + 36 * W->slot[e];
}

/* edge chi2 = e . (Omega e) */
static double edge_chi2(const window_t *W, int e, double *Oe_out)
{
    const double *er = W->err + 6 * e;
    int kind = W->topo->edge_kind[e];
    if (kind == UWBGO_EDGE_RANGE_ANCHOR || kind == UWBGO_EDGE_RANGE_POSE) {
        double Oe = W->range_info[W->slot[e]] * er[0];
        if (Oe_out) Oe_out[0] = Oe;
        return er[0] * Oe;
    }
    const double *O = edge_info6(W, e);
    double Oe[6], chi = 0.0;
    for (int r = 0; r < 6; ++r) {
        double s = O[6 * r] * er[0];
        for (int c = 1; c < 6; ++c) s = s + O[6 * r + c] * er[c];
        Oe[r] = s;
    }
    for (int r = 0; r < 6; ++r) chi = chi + er[r] * Oe[r];
    if (Oe_out) memcpy(Oe_out, Oe, sizeof Oe);
    return chi;
}

/* activeRobustChi2 and (plain) chi2 over edges in order */
static void chi2_sums(const window_t *W, double *plain, double *robust)
{
    double p = 0.0, r = 0.0;
    for (int e = 0; e < W->E; ++e) {
        double c = edge_chi2(W, e, NULL);
        p = p + c;
        if (W->topo->edge_robust[e]) {
            double r0, r1;
            cauchy(c, W->cfg->kernel_delta, &r0, &r1);
            r = r + r0;
        } else
            r = r + c;
    }
    *plain = p;
    *robust = r;
}

/* ------------------------------------------------------------------------------------------ */
/* linearisation: BaseBinaryEdge::linearizeOplus (numeric) + constructQuadraticForm             */
/* ------------------------------------------------------------------------------------------ */

/* numeric Jacobian of a range edge wrt vertex `which` (0: edge_a pose; 1: edge_b pose) */
static void range_numeric_jacobian(window_t *W, int e, int which, double *J)
{
    const uwbgo_topology *T = W->topo;
    const double delta = W->cfg->jacobian_delta;
    const double scalar = 1.0 / (2.0 * delta);
    int a = T->edge_a[e];
    int vidx = which == 0 ? a : T->edge_b[e];
    for (int d = 0; d < 6; ++d) {
        double ep[2];
        for (int sgn = 0; sgn < 2; ++sgn) {
            double add[6] = {0, 0, 0, 0, 0, 0};
            add[d] = sgn == 0 ? delta : -delta;
            pose_t keep = W->X[vidx];                                  /* push */
            pose_oplus(&W->X[vidx], add, &W->cnt[vidx], W->cfg->orthogonalize_after);
            ep[sgn] = range_edge_error(W, e);
            W->X[vidx] = keep;                                         /* pop */
        }
        J[d] = scalar * (ep[0] - ep[1]);
    }
}

/* vector part of the quaternion-product matrices used by the analytic SE3 Jacobians */
static void quat_left(const double *q, double M[16])
{ /* L(q) p = q (x) p, q = {x,y,z,w}, ordering of the 4-vectors {w,x,y,z} */
    double w = q[3], x = q[0], y = q[1], z = q[2];
    double L[16] = {w, -x, -y, -z, x, w, -z, y, y, z, w, -x, z, -y, x, w};
    memcpy(M, L, sizeof L);
}
static void quat_right(const double *q, double M[16])
{ /* R(q) p = p (x) q */
    double w = q[3], x = q[0], y = q[1], z = q[2];
    double Rm[16] = {w, -x, -y, -z, x, w, z, -y, y, -z, w, x, z, y, -x, w};
    memcpy(M, Rm, sizeof Rm);
}

/* J_qq of an error quaternion right-multiplied by the increment: w I + [q]x   */
static void jqq_right(const double *q, double *J /*3x3*/)
{
    double w = q[3], x = q[0], y = q[1], z = q[2];
    J[0] = w;  J[1] = -z; J[2] = y;
    J[3] = z;  J[4] = w;  J[5] = -x;
    J[6] = -y; J[7] = x;  J[8] = w;
}

/* accumulate a D-dimensional edge with Jacobians A (vertex i) and B (vertex j, may be NULL)
 * into H and b.  A, B are D x 6 row-major.  Ow = weighted information (D x D), omega_r = -rho1*Omega*e */
static void accumulate(window_t *W, int D, int i, int j, const double *A, const double *B,
                       const double *Ow, const double *omega_r)
{
    double *Hii = W->Hd + 36 * i, *bi = W->b + 6 * i;
    if (D == 1) {
        double AtO[6];
        for (int r = 0; r < 6; ++r) bi[r] = fma(A[r], omega_r[0], bi[r]);
        for (int r = 0; r < 6; ++r) AtO[r] = A[r] * Ow[0];
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) Hii[6 * r + c] = fma(AtO[r], A[c], Hii[6 * r + c]);
        if (B) {
            double *Hij = W->Ho + 36 * j, *Hjj = W->Hd + 36 * j, *bj = W->b + 6 * j;
            double BtO[6];
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c) Hij[6 * r + c] = fma(AtO[r], B[c], Hij[6 * r + c]);
            for (int r = 0; r < 6; ++r) bj[r] = fma(B[r], omega_r[0], bj[r]);
            for (int r = 0; r < 6; ++r) BtO[r] = B[r] * Ow[0];
            for (int r = 0; r < 6; ++r)
                for (int c = r; c < 6; ++c) Hjj[6 * r + c] = fma(BtO[r], B[c], Hjj[6 * r + c]);
        }
        return;
    }
    /* D == 6: b_i += A^T omega_r; H_ii += (A^T Ow) A; H_ij += (A^T Ow) B; H_jj += (B^T Ow) B */
    double AtO[36], BtO[36];
    for (int r = 0; r < 6; ++r) {
        double s = A[r] * omega_r[0];
        for (int k = 1; k < 6; ++k) s = fma(A[6 * k + r], omega_r[k], s);
        bi[r] = bi[r] + s;
    }
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
            double s = A[r] * Ow[c];
            for (int k = 1; k < 6; ++k) s = fma(A[6 * k + r], Ow[6 * k + c], s);
            AtO[6 * r + c] = s;
        }
    for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) {
            double s = AtO[6 * r] * A[c];
            for (int k = 1; k < 6; ++k) s = fma(AtO[6 * r + k], A[6 * k + c], s);
            Hii[6 * r + c] = Hii[6 * r + c] + s;
        }
    if (B) {
        double *Hij = W->Ho + 36 * j, *Hjj = W->Hd + 36 * j, *bj = W->b + 6 * j;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double s = AtO[6 * r] * B[c];
                for (int k = 1; k < 6; ++k) s = fma(AtO[6 * r + k], B[6 * k + c], s);
                Hij[6 * r + c] = Hij[6 * r + c] + s;
            }
        for (int r = 0; r < 6; ++r) {
            double s = B[r] * omega_r[0];
            for (int k = 1; k < 6; ++k) s = fma(B[6 * k + r], omega_r[k], s);
            bj[r] = bj[r] + s;
        }
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double s = B[r] * Ow[c];
                for (int k = 1; k < 6; ++k) s = fma(B[6 * k + r], Ow[6 * k + c], s);
                BtO[6 * r + c] = s;
            }
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) {
                double s = BtO[6 * r] * B[c];
                for (int k = 1; k < 6; ++k) s = fma(BtO[6 * r + k], B[6 * k + c], s);
                Hjj[6 * r + c] = Hjj[6 * r + c] + s;
            }
    }
}

/* analytic Jacobian of EdgeSE3Prior: D = Zinv*X */
static void prior_jacobian(const pose_t *D, const double *qE, double *J /*6x6*/)
{
    memset(J, 0, 36 * sizeof(double));
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[6 * r + c] = D->R[3 * r + c];
    double Jq[9];
    jqq_right(qE, Jq);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[6 * (3 + r) + 3 + c] = Jq[3 * r + c];
}

/* analytic Jacobians of EdgeSE3: A = Zinv, B = Xi^-1 Xj, E = A*B */
static void se3_jacobians(const pose_t *Zinv, const pose_t *Xi, const pose_t *Xj, double *Ji,
                          double *Jj)
{
    pose_t Xi_inv, Bm, AB;
    pose_inv(Xi, &Xi_inv);
    pose_mul(&Xi_inv, Xj, &Bm);
    pose_mul(Zinv, &Bm, &AB);
    memset(Ji, 0, 36 * sizeof(double));
    memset(Jj, 0, 36 * sizeof(double));
    const double *Ra = Zinv->R, *tb = Bm.t;
    /* dte/dti = -Ra ; dte/dtj = R_AB */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            Ji[6 * r + c] = -Ra[3 * r + c];
            Jj[6 * r + c] = AB.R[3 * r + c];
        }
    /* dte/dqi = Ra * 2[tb]x */
    double S[9] = {0.0, -2.0 * tb[2], 2.0 * tb[1], 2.0 * tb[2], 0.0, -2.0 * tb[0],
                   -2.0 * tb[1], 2.0 * tb[0], 0.0};
    double RaS[9];
    mat3_mul(Ra, S, RaS);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ji[6 * r + 3 + c] = RaS[3 * r + c];
    /* rotation blocks: q_E = q_A (x) dq^-1 (x) q_B for vertex i, q_E (x) dq for vertex j */
    double qA[4], qB[4], qE[4];
    R_to_quat(Ra, qA);
    R_to_quat(Bm.R, qB);
    R_to_quat(AB.R, qE);
    double Jq[9];
    jqq_right(qE, Jq);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Jj[6 * (3 + r) + 3 + c] = Jq[3 * r + c];
    double L[16], Rm[16];
    quat_left(qA, L);
    quat_right(qB, Rm);
    /* M = L(qA) * R(qB); the sign of q_E was normalised to w >= 0: if w(qA (x) qB) < 0 flip */
    double wAB = 0.0;
    for (int k = 0; k < 4; ++k) wAB = wAB + L[k] * Rm[4 * k];
    double sgn = wAB < 0.0 ? 1.0 : -1.0; /* Ji_qq = -M_vv (times -1 again if flipped) */
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s = s + L[4 * (r + 1) + k] * Rm[4 * k + (c + 1)];
            Ji[6 * (3 + r) + 3 + c] = sgn * s;
        }
}

/* BlockSolver::buildSystem: errors must be current (compute_errors) */
static void build_system(window_t *W)
{
    const uwbgo_topology *T = W->topo;
    int N = W->N;
    memset(W->Hd, 0, (size_t)N * 36 * sizeof(double));
    memset(W->Ho, 0, (size_t)N * 36 * sizeof(double));
    memset(W->b, 0, (size_t)N * 6 * sizeof(double));
    for (int e = 0; e < W->E; ++e) {
        int kind = T->edge_kind[e], a = T->edge_a[e], s = W->slot[e];
        const double *er = W->err + 6 * e;
        if (kind == UWBGO_EDGE_RANGE_ANCHOR || kind == UWBGO_EDGE_RANGE_POSE) {
            double A[6], B[6];
            range_numeric_jacobian(W, e, 0, A);
            if (kind == UWBGO_EDGE_RANGE_POSE) range_numeric_jacobian(W, e, 1, B);
            double info = W->range_info[s];
            double Oe = info * er[0];
            double chi = er[0] * Oe;
            double omega_r = -Oe, Ow = info;
            if (T->edge_robust[e]) {
                double r0, r1;
                cauchy(chi, W->cfg->kernel_delta, &r0, &r1);
                omega_r = omega_r * r1;
                Ow = r1 * info;
            }
            accumulate(W, 1, a, kind == UWBGO_EDGE_RANGE_POSE ? T->edge_b[e] : -1, A,
                       kind == UWBGO_EDGE_RANGE_POSE ? B : NULL, &Ow, &omega_r);
        } else {
            double Oe[6], omega_r[6], Ow[36], Ji[36], Jj[36];
            const double *O = edge_info6(W, e);
            double chi = edge_chi2(W, e, Oe);
            double r1 = 1.0;
            if (T->edge_robust[e]) {
                double r0;
                cauchy(chi, W->cfg->kernel_delta, &r0, &r1);
            }
            for (int r = 0; r < 6; ++r) omega_r[r] = -Oe[r];
            if (T->edge_robust[e]) {
                for (int r = 0; r < 6; ++r) omega_r[r] = omega_r[r] * r1;
                for (int k = 0; k < 36; ++k) Ow[k] = r1 * O[k];
            } else
                memcpy(Ow, O, sizeof Ow);
            pose_t Z, Zinv;
            if (kind == UWBGO_EDGE_PRIOR) {
                pose_t D;
                double e6[6], qE[4];
                load_Z(W->prior_Z + 12 * s, &Z);
                pose_inv(&Z, &Zinv);
                prior_error(&Zinv, &W->X[a], e6, &D, qE);
                prior_jacobian(&D, qE, Ji);
                accumulate(W, 6, a, -1, Ji, NULL, Ow, omega_r);
            } else {
                load_Z(W->se3_Z + 12 * s, &Z);
                pose_inv(&Z, &Zinv);
                se3_jacobians(&Zinv, &W->X[a], &W->X[T->edge_b[e]], Ji, Jj);
                accumulate(W, 6, a, T->edge_b[e], Ji, Jj, Ow, omega_r);
            }
        }
    }
    /* mirror the upper triangles of the diagonal blocks */
    for (int i = 0; i < N; ++i)
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < r; ++c) W->Hd[36 * i + 6 * r + c] = W->Hd[36 * i + 6 * c + r];
}

/* ------------------------------------------------------------------------------------------ */
/* linear solver: sparse block Cholesky of H + lambda I for windows whose pose graph is a        */
/* forest in which every pose has at most ONE older neighbour, parent(j) < j.  Chains            */
/* (parent(j) = j-1: range / twist windows) give block-tridiagonal H; pose edges to a key vertex */
/* (localization.cpp:258-267) give stars.  Eliminating the NEWEST pose first creates no fill in  */
/* such a graph (a pose's only not-yet-eliminated neighbour is its parent).                      */
/* Stands in for LinearSolverCholmod (localization.h:84): CHOLMOD factorises P H P^T for its     */
/* own AMD permutation P, so ANY exact FP64 Cholesky agrees with it up to round-off (SURVEY      */
/* A.8).  The factor sweep runs j = N-1..0, the substitution sweep j = 0..N-1, so x comes out in  */
/* ascending order, the order g2o's computeScale() and update() consume it:                      */
/*   S_j   = H_jj + lambda I - sum over children c of j (descending c) of G_c G_c^T               */
/*   L_j   = lower Cholesky factor of S_j, diagonal kept inverted                                 */
/*   z_j   = L_j^-1 (b_j - sum over children c (descending) of G_c z_c)                           */
/*   G_j   = H_{parent(j),j} L_j^-T      (rows of the parent, columns of pose j)   -> Lo[j]       */
/*   c_j   = L_j^-T z_j                  -> y[j]                                                  */
/*   M_j   = L_j^-T G_j^T                -> Ld[j]                                                 */
/*   x_j   = c_j - M_j x_{parent(j)}                                                             */
/* zs [N][6] is scratch for the z_j.                                                             */
/* ------------------------------------------------------------------------------------------ */
static int factor_solve(int N, const int32_t *parent, const double *Hd, const double *Ho,
                        const double *b, double lambda, double *Ld, double *Lo, double *y,
                        double *zs, double *x)
{
    for (int i = N - 1; i >= 0; --i) {
        double S[36], L[36];
        double *z = zs + 6 * i;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = Hd[36 * i + 6 * r + c];
                if (r == c) s = s + lambda;
                S[6 * r + c] = s;
            }
        for (int ch = N - 1; ch > i; --ch) {
            if ((parent ? parent[ch] : ch - 1) != i) continue;
            const double *G = Lo + 36 * ch;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c <= r; ++c) {
                    double s = S[6 * r + c];
                    for (int k = 0; k < 6; ++k) s = fma(-G[6 * r + k], G[6 * c + k], s);
                    S[6 * r + c] = s;
                }
        }
        memset(L, 0, sizeof L);
        for (int j = 0; j < 6; ++j) {
            double s = S[6 * j + j];
            for (int k = 0; k < j; ++k) s = fma(-L[6 * j + k], L[6 * j + k], s);
            if (!(s > 0.0)) return 0;
            double inv = 1.0 / sqrt(s);
            L[6 * j + j] = inv; /* the diagonal slot stores 1/L_jj */
            for (int r = j + 1; r < 6; ++r) {
                double t = S[6 * r + j];
                for (int k = 0; k < j; ++k) t = fma(-L[6 * r + k], L[6 * j + k], t);
                L[6 * r + j] = t * inv;
            }
        }
        for (int r = 0; r < 6; ++r) {
            double s = b[6 * i + r];
            for (int ch = N - 1; ch > i; --ch) {
                if ((parent ? parent[ch] : ch - 1) != i) continue;
                const double *G = Lo + 36 * ch, *zc = zs + 6 * ch;
                for (int k = 0; k < 6; ++k) s = fma(-G[6 * r + k], zc[k], s);
            }
            for (int k = 0; k < r; ++k) s = fma(-L[6 * r + k], z[k], s);
            z[r] = s * L[6 * r + r];
        }
        double *c = y + 6 * i;
        for (int r = 5; r >= 0; --r) {
            double s = z[r];
            for (int k = r + 1; k < 6; ++k) s = fma(-L[6 * k + r], c[k], s);
            c[r] = s * L[6 * r + r];
        }
        if ((parent ? parent[i] : i - 1) >= 0) {
            double *X = Lo + 36 * i;
            for (int r = 0; r < 6; ++r)
                for (int cc = 0; cc < 6; ++cc) {
                    double s = Ho[36 * i + 6 * r + cc];
                    for (int k = 0; k < cc; ++k) s = fma(-X[6 * r + k], L[6 * cc + k], s);
                    X[6 * r + cc] = s * L[6 * cc + cc];
                }
            double *M = Ld + 36 * i;
            for (int j = 0; j < 6; ++j)
                for (int r = 5; r >= 0; --r) {
                    double s = X[6 * j + r];
                    for (int k = r + 1; k < 6; ++k) s = fma(-L[6 * k + r], M[6 * k + j], s);
                    M[6 * r + j] = s * L[6 * r + r];
                }
        }
    }
    for (int i = 0; i < N; ++i) {
        const double *M = Ld + 36 * i;
        const int p = parent ? parent[i] : i - 1;
        for (int r = 0; r < 6; ++r) {
            double s = y[6 * i + r];
            if (p >= 0)
                for (int j = 0; j < 6; ++j) s = fma(-M[6 * r + j], x[6 * p + j], s);
            x[6 * i + r] = s;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* computeMarginals(spinv, newest vertex) -- the commented-out tail of Localization::solve(),     */
/* localization.cpp:185-189: the diagonal block of H^-1 that belongs to the newest pose, H as the  */
/* last buildSystem() left it (no damping).  Same elimination as factor_solve (newest pose first,  */
/* lambda = 0): S_i = H_ii - sum over children G_c G_c^T = L_i L_i^T, G_i = H_{p(i),i} L_i^-T,      */
/* M_i = L_i^-T G_i^T.  x_i given x_p(i) is Gaussian with mean c_i - M_i x_p(i) and covariance      */
/* S_i^-1, so along the path root -> ... -> newest                                                  */
/*     Sigma_root = S_root^-1,      Sigma_i = S_i^-1 + M_i Sigma_p(i) M_i^T.                        */
/* mg: [N][108] scratch (G_i | M_i | S_i^-1).  Returns 0 when a pivot is not positive.              */
/* ------------------------------------------------------------------------------------------ */
static int marginal_newest(int N, const int32_t *parent, const double *Hd, const double *Ho, double *mg,
                           double *Sigma /* 36 */)
{
    for (int i = N - 1; i >= 0; --i) {
        double S[36], L[36], Li[36];
        double *G = mg + 108 * (size_t)i, *M = G + 36, *Sinv = G + 72;
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c <= r; ++c) S[6 * r + c] = Hd[36 * i + 6 * r + c];
        for (int ch = N - 1; ch > i; --ch) {
            if (parent[ch] != i) continue;
            const double *Gc = mg + 108 * (size_t)ch;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c <= r; ++c) {
                    double s = S[6 * r + c];
                    for (int k = 0; k < 6; ++k) s = fma(-Gc[6 * r + k], Gc[6 * c + k], s);
                    S[6 * r + c] = s;
                }
        }
        memset(L, 0, sizeof L);
        for (int j = 0; j < 6; ++j) {
            double s = S[6 * j + j];
            for (int k = 0; k < j; ++k) s = fma(-L[6 * j + k], L[6 * j + k], s);
            if (!(s > 0.0)) return 0;
            double inv = 1.0 / sqrt(s);
            L[6 * j + j] = inv; /* the diagonal slot stores 1/L_jj */
            for (int r = j + 1; r < 6; ++r) {
                double t = S[6 * r + j];
                for (int k = 0; k < j; ++k) t = fma(-L[6 * r + k], L[6 * j + k], t);
                L[6 * r + j] = t * inv;
            }
        }
        /* Li = L^-1 (lower), then S^-1 = Li^T Li */
        memset(Li, 0, sizeof Li);
        for (int j = 0; j < 6; ++j) {
            Li[6 * j + j] = L[6 * j + j];
            for (int r = j + 1; r < 6; ++r) {
                double s = 0.0;
                for (int k = j; k < r; ++k) s = fma(-L[6 * r + k], Li[6 * k + j], s);
                Li[6 * r + j] = s * L[6 * r + r];
            }
        }
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = 0.0;
                for (int k = r; k < 6; ++k) s = fma(Li[6 * k + r], Li[6 * k + c], s);
                Sinv[6 * r + c] = s;
                Sinv[6 * c + r] = s;
            }
        if (parent[i] >= 0) {
            for (int r = 0; r < 6; ++r)
                for (int cc = 0; cc < 6; ++cc) {
                    double s = Ho[36 * i + 6 * r + cc];
                    for (int k = 0; k < cc; ++k) s = fma(-G[6 * r + k], L[6 * cc + k], s);
                    G[6 * r + cc] = s * L[6 * cc + cc];
                }
            for (int j = 0; j < 6; ++j)
                for (int r = 5; r >= 0; --r) {
                    double s = G[6 * j + r];
                    for (int k = r + 1; k < 6; ++k) s = fma(-L[6 * k + r], M[6 * k + j], s);
                    M[6 * r + j] = s * L[6 * r + r];
                }
        }
    }
    /* the path from the newest pose up to its root, walked back down */
    int path[4096], len = 0;
    for (int i = N - 1; i >= 0 && len < 4096; i = parent[i]) path[len++] = i;
    memcpy(Sigma, mg + 108 * (size_t)path[len - 1] + 72, 36 * sizeof(double));
    for (int k = len - 2; k >= 0; --k) {
        const double *M = mg + 108 * (size_t)path[k] + 36, *Sinv = M + 36;
        double T[36], Nw[36];
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
                for (int j = 0; j < 6; ++j) s = fma(M[6 * r + j], Sigma[6 * j + c], s);
                T[6 * r + c] = s;
            }
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c <= r; ++c) {
                double s = Sinv[6 * r + c];
                for (int j = 0; j < 6; ++j) s = fma(T[6 * r + j], M[6 * c + j], s);
                Nw[6 * r + c] = s;
                Nw[6 * c + r] = s;
            }
        memcpy(Sigma, Nw, sizeof Nw);
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* window set-up / tear-down                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* per-kind data slots, argument checks, and the parent of every pose: the one older neighbour it
 * shares a pose-pose edge with (-1: none).  Two different older neighbours = not a forest. */
static int count_slots(const uwbgo_topology *T, int32_t *slot, int32_t *parent, int *Er, int *Ep,
                       int *Es)
{
    int er = 0, ep = 0, es = 0;
    for (int i = 0; i < T->n_poses; ++i) parent[i] = -1;
    for (int e = 0; e < T->n_edges; ++e) {
        int a = T->edge_a[e], b = T->edge_b[e], kind = T->edge_kind[e];
        if (a < 0 || a >= T->n_poses) return UWBGO_E_INVALID;
        switch (kind) {
        case UWBGO_EDGE_RANGE_ANCHOR:
            if (b < 0 || b >= T->n_anchors) return UWBGO_E_INVALID;
            slot[e] = er++;
            break;
        case UWBGO_EDGE_PRIOR:
            slot[e] = ep++;
            break;
        case UWBGO_EDGE_RANGE_POSE:
        case UWBGO_EDGE_SE3:
            if (b <= a || b >= T->n_poses) return UWBGO_E_TOPOLOGY;
            if (parent[b] >= 0 && parent[b] != a) return UWBGO_E_TOPOLOGY;
            parent[b] = a;
            slot[e] = kind == UWBGO_EDGE_SE3 ? es++ : er++;
            break;
        default:
            return UWBGO_E_INVALID;
        }
        if (kind <= UWBGO_EDGE_RANGE_POSE && T->edge_ant &&
            (T->edge_ant[e] < 0 || T->edge_ant[e] > T->n_antennas))
            return UWBGO_E_INVALID;
        if (kind <= UWBGO_EDGE_RANGE_POSE && T->edge_ant_b &&
            (T->edge_ant_b[e] < 0 || T->edge_ant_b[e] > T->n_antennas))
            return UWBGO_E_INVALID;
    }
    *Er = er;
    *Ep = ep;
    *Es = es;
    return 0;
}

static int window_alloc(window_t *W, const uwbgo_topology *T, const uwbgo_config *cfg,
                        const int32_t *slot, const int32_t *parent, int Er, int Ep, int Es)
{
    memset(W, 0, sizeof *W);
    W->topo = T;
    W->cfg = cfg;
    W->N = T->n_poses;
    W->A = T->n_anchors;
    W->E = T->n_edges;
    W->Er = Er;
    W->Ep = Ep;
    W->Es = Es;
    W->slot = slot;
    W->parent = parent;
    size_t N = (size_t)W->N, E = (size_t)W->E;
    W->X = (pose_t *)malloc(N * sizeof(pose_t));
    W->Xbak = (pose_t *)malloc(N * sizeof(pose_t));
    W->cnt = (int32_t *)malloc(N * sizeof(int32_t));
    W->err = (double *)calloc(E * 6 + 1, sizeof(double));
    W->Hd = (double *)malloc(N * 36 * sizeof(double));
    W->Ho = (double *)malloc(N * 36 * sizeof(double));
    W->b = (double *)malloc(N * 6 * sizeof(double));
    W->x = (double *)malloc(N * 6 * sizeof(double));
    W->Ld = (double *)malloc(N * 36 * sizeof(double));
    W->Lo = (double *)malloc(N * 36 * sizeof(double));
    W->y = (double *)malloc(N * 6 * sizeof(double));
    W->zs = (double *)malloc(N * 6 * sizeof(double));
    W->rd_own = (double *)malloc(((size_t)Er + 1) * sizeof(double));
    W->ri_own = (double *)malloc(((size_t)Er + 1) * sizeof(double));
    W->pi_own = (double *)calloc((size_t)W->Ep * 36 + 1, sizeof(double));
    W->si_own = (double *)calloc((size_t)W->Es * 36 + 1, sizeof(double));
    W->estale = (double *)calloc(E + 1, sizeof(double));
    W->mg = (double *)malloc(N * 108 * sizeof(double));
    return (W->X && W->Xbak && W->cnt && W->err && W->Hd && W->Ho && W->b && W->x && W->Ld &&
            W->Lo && W->y && W->zs && W->rd_own && W->ri_own && W->pi_own && W->si_own && W->estale && W->mg)
               ? 0
               : UWBGO_E_NOMEM;
}

static void window_free(window_t *W)
{
    free(W->X); free(W->Xbak); free(W->cnt); free(W->err); free(W->Hd); free(W->Ho);
    free(W->b); free(W->x); free(W->Ld); free(W->Lo); free(W->y); free(W->zs);
    free(W->rd_own); free(W->ri_own); free(W->pi_own); free(W->si_own); free(W->estale); free(W->mg);
}

static void window_load(window_t *W, const uwbgo_batch *in, int64_t w)
{
    int N = W->N;
    for (int i = 0; i < N; ++i) {
        memcpy(W->X[i].t, in->pose_t + ((size_t)w * N + i) * 3, 3 * sizeof(double));
        if (in->pose_R)
            memcpy(W->X[i].R, in->pose_R + ((size_t)w * N + i) * 9, 9 * sizeof(double));
        else {
            memset(W->X[i].R, 0, sizeof W->X[i].R);
            W->X[i].R[0] = W->X[i].R[4] = W->X[i].R[8] = 1.0;
        }
        W->cnt[i] = in->oplus_count ? in->oplus_count[(size_t)w * N + i] : 0;
    }
    W->built = 0;
    /* UWBGO_SHARED_ANCHORS: one constellation for every window */
    W->anchors = in->anchors ? in->anchors + ((in->shared & UWBGO_SHARED_ANCHORS) ? 0 : (size_t)w * W->A * 3) : NULL;
    W->ant = in->ant_offsets;
    W->range_d = in->range_d ? in->range_d + (size_t)w * W->Er : NULL;
    W->range_info = in->range_info ? in->range_info + (size_t)w * W->Er : NULL;
    if (in->range_msgs) {
        /* compact form: the edge parameters as Localization::addRangeEdge / create_range_edge compute them
         * (localization.cpp:316-319: distance_cov = pow(distance_err, 2), cov_requester =
         * pow(robot_max_velocity * dt_requester / 3, 2); :331 (distance, distance_cov); :338 (0, cov_requester);
         * :350 (distance, distance_cov + cov_requester); :613 information = covariance_matrix.inverse()) */
        const uwbgo_range_msgs *m = in->range_msgs;
        const uwbgo_topology *T = W->topo;
        int era = 0, erp = 0, ka = 0, kp = 0;
        for (int e = 0; e < T->n_edges; ++e) {
            era += T->edge_kind[e] == UWBGO_EDGE_RANGE_ANCHOR;
            erp += T->edge_kind[e] == UWBGO_EDGE_RANGE_POSE;
        }
        for (int e = 0; e < T->n_edges; ++e) {
            const int kind = T->edge_kind[e];
            if (kind == UWBGO_EDGE_RANGE_ANCHOR) {
                const double derr = (double)m->distance_err[(size_t)w * era + ka];
                double cov = derr * derr;
                if (m->dt_anchor) {
                    const double mv = m->v_max * m->dt_anchor[(size_t)w * era + ka] / 3;
                    cov = cov + mv * mv;
                }
                W->rd_own[W->slot[e]] = (double)m->distance[(size_t)w * era + ka];
                W->ri_own[W->slot[e]] = 1.0 / cov;
                ++ka;
            } else if (kind == UWBGO_EDGE_RANGE_POSE) {
                const double mv = m->v_max * m->dt_pose[(size_t)w * erp + kp] / 3;
                W->rd_own[W->slot[e]] = 0.0;
                W->ri_own[W->slot[e]] = 1.0 / (mv * mv);
                ++kp;
            }
        }
        W->range_d = W->rd_own;
        W->range_info = W->ri_own;
    }
    W->prior_Z = in->prior_Z ? in->prior_Z + (size_t)w * W->Ep * 12 : NULL;
    W->prior_info = in->prior_info ? in->prior_info + (size_t)w * W->Ep * 36 : NULL;
    W->se3_Z = in->se3_Z ? in->se3_Z + (size_t)w * W->Es * 12 : NULL;
    W->se3_info = in->se3_info ? in->se3_info + (size_t)w * W->Es * 36 : NULL;
    if (in->shared & UWBGO_DIAG_INFO) {
        /* the information matrices as Localization fills them: Matrix<double, 6, 6>::Zero() plus the diagonal
         * entries (localization.cpp:478-479, 515-518); pi_own / si_own were calloc'ed, only the diagonals change */
        for (int k = 0; k < W->Ep; ++k)
            for (int d = 0; d < 6; ++d) W->pi_own[(size_t)k * 36 + 7 * d] = in->prior_info[((size_t)w * W->Ep + k) * 6 + d];
        for (int k = 0; k < W->Es; ++k)
            for (int d = 0; d < 6; ++d) W->si_own[(size_t)k * 36 + 7 * d] = in->se3_info[((size_t)w * W->Es + k) * 6 + d];
        W->prior_info = W->Ep ? W->pi_own : NULL;
        W->se3_info = W->Es ? W->si_own : NULL;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* optimize(iteration_max) with OptimizationAlgorithmLevenberg — SURVEY.md A.2                  */
/* ------------------------------------------------------------------------------------------ */
/* edge->chi2() of every edge as the last computeActiveErrors left it */
static void snapshot_edge_chi2(window_t *W, int want)
{
    if (!want) return;
    for (int e = 0; e < W->E; ++e) W->estale[e] = edge_chi2(W, e, NULL);
}

static void solve_window(window_t *W, double *chi2_out, int32_t *status_out, double *trace,
                         int trace_stride, int want_edge_chi2)
{
    const uwbgo_config *cfg = W->cfg;
    int N = W->N;
    double lambda = 0.0, ni = 2.0, stale = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0;
    {
        double p, r;
        compute_errors(W);
        chi2_sums(W, &p, &r);
        stale = p;
        snapshot_edge_chi2(W, want_edge_chi2);
    }
    for (int it = 0; it < cfg->max_iterations; ++it) {
        double plain, currentChi;
        compute_errors(W);
        chi2_sums(W, &plain, &currentChi);
        stale = plain;
        build_system(W);
        W->built = 1;
        if (it == 0) {
            double maxdiag = 0.0;
            for (int i = 0; i < N; ++i)
                for (int r = 0; r < 6; ++r) {
                    double v = fabs(W->Hd[36 * i + 7 * r]);
                    if (v > maxdiag) maxdiag = v;
                }
            lambda = cfg->tau * maxdiag;
            ni = 2.0;
        }
        double rho = 0.0;
        int q = 0;
        do {
            memcpy(W->Xbak, W->X, (size_t)N * sizeof(pose_t)); /* push */
            int ok = factor_solve(N, W->parent, W->Hd, W->Ho, W->b, lambda, W->Ld, W->Lo, W->y, W->zs, W->x);
            if (!ok) {
                memset(W->x, 0, (size_t)N * 6 * sizeof(double));
                flags |= UWBGO_FLAG_CHOL_FAIL;
            }
            for (int i = 0; i < N; ++i)
                pose_oplus(&W->X[i], W->x + 6 * i, &W->cnt[i], cfg->orthogonalize_after);
            double tplain, tempChi;
            compute_errors(W);
            chi2_sums(W, &tplain, &tempChi);
            stale = tplain;
            snapshot_edge_chi2(W, want_edge_chi2);
            if (!ok) tempChi = DBL_MAX;
            double scale = 0.0;
            for (int j = 0; j < 6 * N; ++j) scale = scale + W->x[j] * (lambda * W->x[j] + W->b[j]);
            scale = scale + 1e-3;
            rho = (currentChi - tempChi) / scale;
            if (!isfinite(tempChi)) flags |= UWBGO_FLAG_NONFINITE;
            if (rho > 0.0 && isfinite(tempChi)) {
                double t = 2.0 * rho - 1.0;
                double alpha = 1.0 - (t * t) * t;
                alpha = (cfg->good_step_upper < alpha) ? cfg->good_step_upper : alpha; /* std::min */
                double sf = (cfg->good_step_lower < alpha) ? alpha : cfg->good_step_lower; /* std::max */
                lambda = lambda * sf;
                ni = 2.0;
                currentChi = tempChi;
            } else {
                lambda = lambda * ni;
                ni = ni * 2.0;
                memcpy(W->X, W->Xbak, (size_t)N * sizeof(pose_t)); /* pop */
            }
            ++q;
            ++trials_total;
        } while (rho < 0.0 && q < cfg->max_trials);
        ++iterations;
        qlast = q;
        if (trace) {
            double *t = trace + (size_t)it * trace_stride;
            t[0] = currentChi;
            t[1] = lambda;
            t[2] = (double)q;
            t[3] = rho;
        }
        if (q == cfg->max_trials || rho == 0.0) {
            flags |= UWBGO_FLAG_TERMINATED;
            break;
        }
    }
    {
        double p, r;
        compute_errors(W);
        chi2_sums(W, &p, &r);
        chi2_out[0] = p;
        chi2_out[1] = r;
        chi2_out[2] = stale;
        chi2_out[3] = lambda;
    }
    status_out[0] = iterations;
    status_out[1] = trials_total;
    status_out[2] = flags;
    status_out[3] = qlast;
}

/* ------------------------------------------------------------------------------------------ */
/* batch drivers (one window per task over n_threads pthreads)                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int mode; /* 0 solve, 1 linearize */
    const uwbgo_topology *topo;
    const uwbgo_batch *in;
    const uwbgo_config *cfg;
    uwbgo_result *out;
    double *H_diag, *H_off, *b, *chi2, *trace;
    const int32_t *slot, *parent;
    int Er, Ep, Es, tid, nthreads, rc;
} job_t;

static void *worker(void *arg)
{
    job_t *J = (job_t *)arg;
    window_t W;
    J->rc = window_alloc(&W, J->topo, J->cfg, J->slot, J->parent, J->Er, J->Ep, J->Es);
    if (J->rc) {
        window_free(&W);
        return NULL;
    }
    int N = W.N;
    for (int64_t w = J->tid; w < J->in->n_windows; w += J->nthreads) {
        window_load(&W, J->in, w);
        if (J->mode == 0) {
            double chi2[UWBGO_CHI2_STRIDE];
            int32_t status[UWBGO_STATUS_STRIDE];
            solve_window(&W, chi2, status,
                         J->trace ? J->trace + (size_t)w * J->cfg->max_iterations * 4 : NULL, 4,
                         J->out->edge_chi2 != NULL);
            if (J->out->edge_chi2)
                memcpy(J->out->edge_chi2 + (size_t)w * W.E, W.estale, (size_t)W.E * sizeof(double));
            if (J->out->marginal) {
                double *Sg = J->out->marginal + (size_t)w * 36;
                int good = W.built && N <= 4096 && marginal_newest(N, W.parent, W.Hd, W.Ho, W.mg, Sg);
                if (!good)
                    for (int k = 0; k < 36; ++k) Sg[k] = NAN;
                if (J->out->marginal_ok) J->out->marginal_ok[w] = good;
            }
            for (int i = 0; i < N; ++i) {
                memcpy(J->out->pose_t + ((size_t)w * N + i) * 3, W.X[i].t, 3 * sizeof(double));
                if (J->out->pose_R)
                    memcpy(J->out->pose_R + ((size_t)w * N + i) * 9, W.X[i].R, 9 * sizeof(double));
                if (J->out->oplus_count) J->out->oplus_count[(size_t)w * N + i] = W.cnt[i];
            }
            if (J->out->chi2) memcpy(J->out->chi2 + (size_t)w * UWBGO_CHI2_STRIDE, chi2, sizeof chi2);
            if (J->out->status)
                memcpy(J->out->status + (size_t)w * UWBGO_STATUS_STRIDE, status, sizeof status);
        } else {
            double p, r;
            compute_errors(&W);
            chi2_sums(&W, &p, &r);
            build_system(&W);
            memcpy(J->H_diag + (size_t)w * N * 36, W.Hd, (size_t)N * 36 * sizeof(double));
            if (N > 1)
                memcpy(J->H_off + (size_t)w * (N - 1) * 36, W.Ho + 36, (size_t)(N - 1) * 36 * sizeof(double));
            memcpy(J->b + (size_t)w * N * 6, W.b, (size_t)N * 6 * sizeof(double));
            if (J->chi2) {
                J->chi2[2 * w] = p;
                J->chi2[2 * w + 1] = r;
            }
        }
    }
    window_free(&W);
    return NULL;
}

static int run_jobs(job_t *proto, int n_threads)
{
    const uwbgo_topology *T = proto->topo;
    if (!T || !proto->in || !proto->cfg || T->n_poses < 1 || T->n_edges < 0) return UWBGO_E_INVALID;
    int32_t *slot = (int32_t *)malloc(((size_t)T->n_edges + 1 + (size_t)T->n_poses) * sizeof(int32_t));
    if (!slot) return UWBGO_E_NOMEM;
    int32_t *parent = slot + T->n_edges + 1;
    int rc = count_slots(T, slot, parent, &proto->Er, &proto->Ep, &proto->Es);
    if (rc) {
        free(slot);
        return rc;
    }
    proto->slot = slot;
    proto->parent = parent;
    if (n_threads < 1) n_threads = 1;
    if ((int64_t)n_threads > proto->in->n_windows) n_threads = (int)proto->in->n_windows;
    if (n_threads < 1) n_threads = 1;
    job_t *jobs = (job_t *)malloc((size_t)n_threads * sizeof(job_t));
    pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = *proto;
        jobs[t].tid = t;
        jobs[t].nthreads = n_threads;
        jobs[t].rc = 0;
    }
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], NULL, worker, &jobs[t]);
    worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], NULL);
    for (int t = 0; t < n_threads; ++t)
        if (jobs[t].rc) rc = jobs[t].rc;
    free(jobs);
    free(th);
    free(slot);
    return rc;
}

/* optimize(iteration_max) for every window.  trace (optional): [W][max_iterations][4] =
 * {currentChi, lambda, trials, rho} after each LM iteration. */
int uwbgo_oracle_solve_batch(const uwbgo_topology *topo, const uwbgo_batch *in,
                             const uwbgo_config *cfg, uwbgo_result *out, double *trace,
                             int n_threads)
{
    job_t J;
    memset(&J, 0, sizeof J);
    J.mode = 0;
    J.topo = topo;
    J.in = in;
    J.cfg = cfg;
    J.out = out;
    J.trace = trace;
    if (!out || !out->pose_t) return UWBGO_E_INVALID;
    return run_jobs(&J, n_threads);
}

int uwbgo_oracle_linearize_batch(const uwbgo_topology *topo, const uwbgo_batch *in,
                                 const uwbgo_config *cfg, double *H_diag, double *H_off, double *b,
                                 double *chi2, int n_threads)
{
    job_t J;
    memset(&J, 0, sizeof J);
    J.mode = 1;
    J.topo = topo;
    J.in = in;
    J.cfg = cfg;
    J.H_diag = H_diag;
    J.H_off = H_off;
    J.b = b;
    J.chi2 = chi2;
    if (!H_diag || !b || (topo && topo->n_poses > 1 && !H_off)) return UWBGO_E_INVALID;
    return run_jobs(&J, n_threads);
}

int uwbgo_oracle_factor_solve_batch(int32_t n_poses, int64_t n_windows, const double *H_diag,
                                    const double *H_off, const double *b, const double *lambda,
                                    double *x, int32_t *ok)
{
    /* chains only: H_off[w][j] = block (j, j+1) = Ho[j+1] of the window */
    size_t N = (size_t)n_poses;
    double *Ld = (double *)malloc(N * 36 * sizeof(double));
    double *Lo = (double *)malloc(N * 36 * sizeof(double));
    double *Ho = (double *)calloc(N * 36, sizeof(double));
    double *y = (double *)malloc(N * 6 * sizeof(double));
    double *zs = (double *)malloc(N * 6 * sizeof(double));
    if (!Ld || !Lo || !Ho || !y || !zs) {
        free(Ld); free(Lo); free(Ho); free(y); free(zs);
        return UWBGO_E_NOMEM;
    }
    for (int64_t w = 0; w < n_windows; ++w) {
        if (N > 1) memcpy(Ho + 36, H_off + (size_t)w * (N - 1) * 36, (N - 1) * 36 * sizeof(double));
        int good = factor_solve(n_poses, NULL, H_diag + (size_t)w * N * 36, Ho, b + (size_t)w * N * 6,
                                lambda[w], Ld, Lo, y, zs, x + (size_t)w * N * 6);
        if (!good) memset(x + (size_t)w * N * 6, 0, N * 6 * sizeof(double));
        if (ok) ok[w] = good;
    }
    free(Ld); free(Lo); free(Ho); free(y); free(zs);
    return 0;
}

void uwbgo_oracle_config_default(uwbgo_config *cfg)
{
    cfg->max_iterations = 20;       /* localization.cpp:65 */
    cfg->max_trials = 10;
    cfg->orthogonalize_after = 1000;
    cfg->reserved = 0;
    cfg->tau = 1e-5;
    cfg->good_step_lower = 1.0 / 3.0;
    cfg->good_step_upper = 2.0 / 3.0;
    cfg->kernel_delta = 1.0;
    cfg->jacobian_delta = 1e-9;
}
