/*
 * uwbgo_general.cuh — the GENERAL path: full 6x6 blocks with rotations, antenna offsets,
 * EdgeSE3Prior (IMU / lidar, reference localization.cpp:462-535), EdgeSE3 (twist / pose,
 * localization.cpp:254-290,560-605), VertexSE3 oplus counters, chains and forests.
 */
#ifndef UWBGO_GENERAL_CUH
#define UWBGO_GENERAL_CUH

#include "uwbgo_block_solver.cuh"

namespace uwbgo {

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
        double o[3] = {E.ant[3 * (ant - 1)], E.ant[3 * (ant - 1) + 1], E.ant[3 * (ant - 1) + 2]};
        mat3_vec_add(X.R, o, X.t, P);
    } else {
        P[0] = X.t[0]; P[1] = X.t[1]; P[2] = X.t[2];
    }
}

/* vertex 1 of an anchor range edge: a fixed identity-rotation vertex at the anchor, times its offset */
UWBGO_DI void anchor_point(const GenEnv &E, int b, int ant_b, double *Q)
{
    const double *an = E.p.anch + (size_t)b * 3 * TILE;
    if (ant_b > 0) {
        const double I3[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
        double o[3] = {E.ant[3 * (ant_b - 1)], E.ant[3 * (ant_b - 1) + 1], E.ant[3 * (ant_b - 1) + 2]};
        double a[3] = {ROW(an, 0), ROW(an, 1), ROW(an, 2)};
        mat3_vec_add(I3, o, a, Q);
    } else {
        Q[0] = ROW(an, 0); Q[1] = ROW(an, 1); Q[2] = ROW(an, 2);
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

/* the 6x6 information matrix of a 6-D edge into registers, ahead of the arithmetic that produces
 * the error (its IEEE sqrt / division sequences end in branches, which the loads cannot cross) */
UWBGO_DI void load_info6(const double *__restrict__ Irows, int slot, double *O)
{
    const double *p = Irows + (size_t)slot * 36 * TILE;
#pragma unroll
    for (int k = 0; k < 36; ++k) O[k] = ROW(p, k);
}
UWBGO_DI double chi2_6_reg(const double *O, const double *e, double *Oe)
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

#ifndef UWBGO_GCTA_PF
#define UWBGO_GCTA_PF 3 /* bit 0: next op's rows in the linearise phase, bit 1: next edge's rows in the residual phase */
#endif
/* L2 prefetch of everything the evaluation / linearisation of edge e will read */
UWBGO_DI void gen_edge_prefetch(const GenEnv &E, const PoseBuf &T, int e)
{
    EdgeRec er = load_edge(E.tp->edges + e);
    prefetch_rows_l2<3>(T.t + (size_t)er.a * 3 * TILE);
    prefetch_rows_l2<9>(T.R + (size_t)er.a * 9 * TILE);
    if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
        prefetch_l2(E.p.rd + (size_t)er.slot * TILE);
        prefetch_l2(E.p.ri + (size_t)er.slot * TILE);
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR)
            prefetch_rows_l2<3>(E.p.anch + (size_t)er.b * 3 * TILE);
        else
            prefetch_rows_l2<3>(T.t + (size_t)er.b * 3 * TILE);
    } else if (er.kind == UWBGO_EDGE_PRIOR) {
        prefetch_rows_l2<12>(E.p.pZ + (size_t)er.slot * 12 * TILE);
        prefetch_rows_l2<36>(E.p.pI + (size_t)er.slot * 36 * TILE);
    } else {
        prefetch_rows_l2<3>(T.t + (size_t)er.b * 3 * TILE);
        prefetch_rows_l2<9>(T.R + (size_t)er.b * 9 * TILE);
        prefetch_rows_l2<12>(E.p.sZ + (size_t)er.slot * 12 * TILE);
        prefetch_rows_l2<36>(E.p.sI + (size_t)er.slot * 36 * TILE);
    }
}

/* computeError + chi2 of edge e at the estimates T: plain chi2 and its robustified value */
UWBGO_DI void gen_edge_chi(const GenEnv &E, const PoseBuf &T, int e, double &chi_out, double &rob_out)
{
    const DevTopo &tp = *E.tp;
    EdgeRec er = load_edge_full(tp.edges + e);
    double chi;
    if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
        Pose Xa;
        load_pose(T, er.a, Xa);
        double P0[3], Q[3];
        offset_point(E, Xa, er.ant, P0);
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
            anchor_point(E, er.b, er.ant_b, Q);
        } else if (er.ant_b > 0) {
            Pose Xb;
            load_pose(T, er.b, Xb);
            offset_point(E, Xb, er.ant_b, Q);
        } else {
            const double *tb = T.t + (size_t)er.b * 3 * TILE;
            Q[0] = ROW(tb, 0); Q[1] = ROW(tb, 1); Q[2] = ROW(tb, 2);
        }
        double err = ROW(E.p.rd, er.slot) - dist3(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2]);
        double Oe = ROW(E.p.ri, er.slot) * err;
        chi = err * Oe;
    } else if (er.kind == UWBGO_EDGE_PRIOR) {
        Pose Zinv, X, Dl;
        double O[36];
        load_info6(E.p.pI, er.slot, O);
        load_pose(T, er.a, X);
        load_Zinv(E.p.pZ, er.slot, Zinv);
        pose_mul(Zinv, X, Dl);
        double q[4], e6[6], Oe[6];
        R_to_quat(Dl.R, q);
        e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
        e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
        chi = chi2_6_reg(O, e6, Oe);
    } else {
        Pose Zinv, Xi, Xj;
        double O[36];
        load_info6(E.p.sI, er.slot, O);
        load_pose(T, er.a, Xi);
        load_pose(T, er.b, Xj);
        load_Zinv(E.p.sZ, er.slot, Zinv);
        double e6[6], Oe[6];
        se3_error(Zinv, Xi, Xj, e6);
        chi = chi2_6_reg(O, e6, Oe);
    }
    chi_out = chi;
    rob_out = er.robust ? E.ck.rho0(chi) : chi;
}

/* computeActiveErrors + activeChi2 / activeRobustChi2: edges summed in insertion order */
UWBGO_DI void gen_chi_pass(const GenEnv &E, const PoseBuf &T, double &plain, double &robust)
{
    double p = 0.0, r = 0.0;
    for (int e = 0; e < E.tp->E; ++e) {
        double chi, rob;
        gen_edge_chi(E, T, e, chi, rob);
        p = p + chi;
        r = r + rob;
    }
    plain = p;
    robust = r;
}

/* numeric Jacobian of a range residual wrt vertex 0 (pose X with antenna offset `ant`); Q is the
 * other end point.  c0 = the pose's oplus counter when this linearisation started, base = oplus
 * calls made on it by earlier edges of this linearisation.  Call k trips the re-orthogonalisation
 * of the PERTURBED estimate when (c0 + k) % mod == 0 (VertexSE3::oplusImpl; push/pop restores the
 * estimate, not the counter). */
template <bool SECOND = false> /* SECOND: the perturbed pose is vertex 1, its point the subtrahend */
UWBGO_DI void gen_jac_v0(const GenEnv &E, const Pose &X, int ant, const double *Q, double d, int c0,
                         int base, double *J)
{
    const int mod = E.cfg->orth_mod;
    double o[3] = {0.0, 0.0, 0.0};
    if (ant > 0) {
        o[0] = E.ant[3 * (ant - 1)];
        o[1] = E.ant[3 * (ant - 1) + 1];
        o[2] = E.ant[3 * (ant - 1) + 2];
    }
    /* call = (c0 + base + k) mod `mod` of the k-th oplus, kept reduced: one division per edge */
    int call = (c0 + base) % mod;
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double epm[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            if (++call == mod) call = 0;
            double v = sg == 0 ? E.delta : -E.delta;
            double tp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tp[r] = X.R[3 * r + dd] * v + X.t[r];
            double P[3];
            if (ant > 0) {
                if (call == 0) {
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
            epm[sg] = SECOND ? d - dist3(Q[0], Q[1], Q[2], P[0], P[1], P[2])
                             : d - dist3(P[0], P[1], P[2], Q[0], Q[1], Q[2]);
        }
        J[dd] = E.scalar * (epm[0] - epm[1]);
    }
    if (ant > 0) {
#pragma unroll
        for (int dd = 0; dd < 3; ++dd) {
            double epm[2];
#pragma unroll
            for (int sg = 0; sg < 2; ++sg) {
                if (++call == mod) call = 0;
                double q[3] = {0.0, 0.0, 0.0};
                q[dd] = sg == 0 ? E.delta : -E.delta;
                double Rinc[9], Rp[9], P[3];
                increment_R(q, Rinc);
                mat3_mul(X.R, Rinc, Rp);
                if (call == 0) orthogonalize(Rp);
                mat3_vec_add(Rp, o, X.t, P);
                epm[sg] = SECOND ? d - dist3(Q[0], Q[1], Q[2], P[0], P[1], P[2])
                                 : d - dist3(P[0], P[1], P[2], Q[0], Q[1], Q[2]);
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
/* accumulator views: a plain array (registers) or a shared-memory column, element k at p[k * STRIDE] */
template <int STRIDE>
struct SmemAcc {
    double *p;
    UWBGO_DI double &operator[](int k) const { return p[(size_t)k * STRIDE]; }
};

template <class AH, class AB>
UWBGO_DI void acc1_diag(const double *J, double Ow, double omega_r, AH hd, AB bb)
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
template <class AO>
UWBGO_DI void acc1_off(const double *A, const double *B, double Ow, AO ho)
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
template <class AB>
UWBGO_DI void acc6_b(const double *J, const double *omega_r, AB bb)
{
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = J[r] * omega_r[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) s = fma(J[6 * k + r], omega_r[k], s);
        bb[r] = bb[r] + s;
    }
}
template <class AH>
UWBGO_DI void acc6_diag(const double *JtO, const double *J, AH hd)
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
template <class AO>
UWBGO_DI void acc6_off(const double *AtO, const double *B, AO ho)
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

/* BlockSolver::buildSystem for ONE pose: the H record of pose i (H_ii, H_{parent(i),i}, b_i) gathered
 * from the edges touching it, in insertion order.  Returns max |H_kk| over the pose's diagonal.
 * Reads the oplus counters, does not advance them.  SWEPT: poses are linearised in ascending order
 * by one thread that advances each pose's counter right after its record (gen_linearize), so the
 * counter of an older pose already includes this linearisation's calls; !SWEPT: every counter
 * still holds its value from before the linearisation (poses linearised concurrently). */
template <bool SWEPT, class AH, class AO, class AB>
UWBGO_DI double gen_linearize_pose_acc(const GenEnv &E, const PoseBuf &T, const int i, AH hd, AO ho, AB bb)
{
    const DevTopo &tp = *E.tp;
    const int mod = E.cfg->orth_mod;
    double maxdiag = 0.0;
    {
        Pose Xi;
        load_pose(T, i, Xi);
        const int ci = E.p.cnt[(size_t)i * TILE];
#pragma unroll
        for (int k = 0; k < 21; ++k) hd[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 36; ++k) ho[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) bb[k] = 0.0;
        const int ob = __ldg(tp.op_begin + i), oe = __ldg(tp.op_begin + i + 1);
        for (int o = ob; o < oe; ++o) {
            int2 op = __ldg(reinterpret_cast<const int2 *>(tp.ops + o));
            EdgeRec er = load_edge_full(tp.edges + op.x);
            if (!SWEPT && (UWBGO_GCTA_PF & 1) && o + 1 < oe) gen_edge_prefetch(E, T, __ldg(&tp.ops[o + 1].edge));
            if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
                double d = ROW(E.p.rd, er.slot), info = ROW(E.p.ri, er.slot);
                double P0[3], Q[3], J[6];
                Pose Xo; /* the other pose of a pose-pose edge */
                if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
                    anchor_point(E, er.b, er.ant_b, Q);
                    offset_point(E, Xi, er.ant, P0);
                } else if (op.y == 0) {
                    load_pose(T, er.b, Xo);
                    offset_point(E, Xo, er.ant_b, Q);
                    offset_point(E, Xi, er.ant, P0);
                } else {
                    load_pose(T, er.a, Xo);
                    offset_point(E, Xi, er.ant_b, Q);
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
                    if (er.ant_b > 0) /* lever arm on vertex 1: rotation columns, counter-driven re-orthogonalisation */
                        gen_jac_v0<true>(E, Xi, er.ant_b, P0, d, ci, er.base_b, J);
                    else
                        gen_jac_v1(E, P0, Xi, d, J);
                    acc1_diag(J, Ow, omega_r, hd, bb);
                    double A[6];
                    const int ca_now = E.p.cnt[(size_t)er.a * TILE];
                    const int ca = SWEPT ? ((ca_now - __ldg(tp.num_calls + er.a)) % mod + mod) % mod : ca_now;
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
    }
    return maxdiag;
}

/* ... with the accumulators of the pose in registers */
template <bool SWEPT>
UWBGO_DI double gen_linearize_pose(const GenEnv &E, const PoseBuf &T, const int i)
{
    double hd[21], ho[36], bb[6];
    return gen_linearize_pose_acc<SWEPT, double *, double *, double *>(E, T, i, hd, ho, bb);
}

/* BlockSolver::buildSystem, general edges, one thread per window.  Advances the oplus counters by
 * the numeric-Jacobian calls.  Returns max |H_kk|. */
static __device__ __noinline__ double gen_linearize(const GenEnv &E, const PoseBuf &T)
{
    const DevTopo &tp = *E.tp;
    const int N = tp.N, mod = E.cfg->orth_mod;
    double maxdiag = 0.0;
    for (int i = 0; i < N; ++i) {
        double m = gen_linearize_pose<true>(E, T, i);
        if (m > maxdiag) maxdiag = m;
        E.p.cnt[(size_t)i * TILE] = (E.p.cnt[(size_t)i * TILE] + __ldg(tp.num_calls + i)) % mod;
    }
    return maxdiag;
}

/* Forest windows (pose edges to a key vertex, localization.cpp:258-267): every pose has at most one
 * older neighbour parent(i) < i, not necessarily i-1.  Same elimination as factor_sweep<6>, newest
 * pose first and therefore without fill, but a pose may have several children, whose G and z are
 * read back from their L records:  L record (tree) = c 6 | M 36 | G 36 | z 6 | x 6. */
static __device__ __noinline__ bool factor_sweep_tree(const DevTopo &tp, const double *__restrict__ HB,
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

static __device__ __noinline__ double gen_solve_update(const GenEnv &E, bool ok, double lambda,
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

/* gen_solve_update cut in two for the CTA kernel.  First the serial part, one thread per window:
 * the substitution x_i = c_i - M_i x_parent(i) and computeScale(); x_i is left in the L record
 * (chain: over c_i, which is dead once x_i exists; forest: its x slot). */
static __device__ __noinline__ double gen_subst_scale(const GenEnv &E, bool ok, double lambda)
{
    const int N = E.tp->N;
    double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double scale = 0.0;
    const bool tree = E.tp->tree != 0;
    const size_t lstride = (size_t)(tree ? LR_TREE : LR_GEN) * TILE;
    /* the record and b of pose i + 1 are fetched while pose i is substituted */
    double l[LR_GEN], ln[LR_GEN], b[6], bn[6];
#pragma unroll
    for (int k = 0; k < LR_GEN; ++k) ln[k] = ROW(E.p.LR, k);
#pragma unroll
    for (int k = 0; k < 6; ++k) bn[k] = ROW(E.p.HB, 57 + k);
    for (int i = 0; i < N; ++i) {
        double *lp = E.p.LR + (size_t)i * lstride;
#pragma unroll
        for (int k = 0; k < LR_GEN; ++k) l[k] = ln[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) b[k] = bn[k];
        if (i + 1 < N) {
            const double *h = E.p.HB + (size_t)(i + 1) * HR_GEN * TILE;
#pragma unroll
            for (int k = 0; k < LR_GEN; ++k) ln[k] = ROW(lp + lstride, k);
#pragma unroll
            for (int k = 0; k < 6; ++k) bn[k] = ROW(h, 57 + k);
        }
        bool link = i > 0;
        if (tree) {
            const int par = __ldg(E.tp->parent + i);
            link = par >= 0;
            if (link) {
                const double *pp = E.p.LR + (size_t)par * LR_TREE * TILE;
#pragma unroll
                for (int k = 0; k < 6; ++k) xp[k] = ROW(pp, 84 + k);
            }
        }
        subst_step<6>(l, link, xp);
        if (!ok) {
#pragma unroll
            for (int k = 0; k < 6; ++k) xp[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) ROW(lp, (tree ? 84 : 0) + k) = xp[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) scale = scale + xp[k] * (lambda * xp[k] + b[k]);
    }
    return scale;
}

/* ... then the update of pose i, independent of every other pose: estimate (+) x_i into the trial
 * buffer, oplus counter advanced */
UWBGO_DI void gen_update_pose(const GenEnv &E, int i, const PoseBuf &Tc, const PoseBuf &Tn, bool linearised)
{
    const bool tree = E.tp->tree != 0;
    const double *lp = E.p.LR + (size_t)i * (tree ? LR_TREE : LR_GEN) * TILE;
    double x[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = ROW(lp, (tree ? 84 : 0) + k);
    Pose X;
    load_pose(Tc, i, X);
    int c = E.p.cnt[(size_t)i * TILE];
    /* the numeric Jacobians of this iteration's buildSystem advanced the counter first */
    if (linearised) c = (c + __ldg(E.tp->num_calls + i)) % E.cfg->orth_mod;
    pose_oplus(X, x, c, E.cfg->orth_mod);
    E.p.cnt[(size_t)i * TILE] = c;
    store_pose(Tn, i, X);
}

}  // namespace uwbgo
#endif
