/*
 * uwbgo_general_items.cu — GENERAL path (6x6 blocks), ITEM kernel: one CTA per tile of 32 windows, lane = window,
 * and every phase of the LM trial cut into ITEMS small enough to run at 128 registers per thread, so that a tile
 * has 8 warps (two tiles per SM) instead of the 4 warps at 255 registers of lm_general_cta_kernel
 * (uwbgo_general_cta.cuh), whose linearisation of a whole pose holds ~170 doubles per thread and spills.
 *
 * The arithmetic is Localization::solve()'s optimize(iteration_max) (reference src/localization/localization.cpp:164-192)
 * over the graph of localization.cpp:254-376,462-605, as restated by the CPU checker; every H / b entry, every
 * chi2 term and every elimination step sees the operation sequence of lm_general_cta_kernel, so the results are
 * the same bits.  What changes is who computes what:
 *
 *   J   buildSystem, part 1: one item per 6-D edge (error, Jacobians, weights), per range edge (error, weights,
 *       numeric Jacobian wrt vertex 0) and per pose-pose range edge (numeric Jacobian wrt vertex 1); warp k takes
 *       items k, k + NW, ...; the records go to a per-tile scratch (tile layout, L2)
 *   H   buildSystem, part 2: one item per (pose, row of H_ii + b_i) and per (pose, row of H_{i-1,i}); the item
 *       walks the pose's edges in g2o insertion order and adds their terms (J^T Omega rows of the 6-D edges are
 *       formed on the fly), so every entry is accumulated by one thread in the order of the checker
 *   F   the serial chain of the window on TWO warps: warp 0 eliminates (S_i, potrf, z_i, G_i; G of the child pose
 *       is read from shared memory), warp 1 follows one pose behind with what nothing in the elimination waits for
 *       (c_i, M_i and the L record store) and then runs the substitution and computeScale()
 *   U   estimate (+) x, one item per pose;  C  computeActiveErrors, one item per edge, chi2 terms to shared memory
 *   D   warp 0: ordered sums, accept / reject (the flat LM loop of lm_general_cta_kernel)
 *
 * Chains only (parent(i) = i - 1 or none); forest windows (pose edges to a key vertex) keep lm_general_cta_kernel.
 */
#include <cstdlib>
#include <type_traits>
#include "uwbgo_general.cuh"
#ifdef UWBGO_GIT_TIMING
#include <cstdio>
/* (clock64() may be moved across barriers by the compiler; the volatile asm with a memory clobber may not) */
__device__ __forceinline__ long long git_clk()
{
    /* the clock is read behind a shared-memory load whose value it formally depends on: a barrier releases the
     * warp's next memory access, not its next instruction, and a bare clock read after bar.sync returns early */
    extern __shared__ double git_clk_dyn[];
    long long t;
    unsigned z;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(z) : "r"((uint32_t)__cvta_generic_to_shared(git_clk_dyn)) : "memory");
    asm volatile("{\n\t.reg .u64 c;\n\tmov.u64 c, %%clock64;\n\tand.b32 %1, %1, 0;\n\tcvt.u64.u32 %0, %1;\n\tadd.u64 %0, %0, c;\n\t}" : "=l"(t), "+r"(z)::"memory");
    return t;
}
#endif

namespace uwbgo {

namespace {

constexpr int GIT_MAX_SMEM_ANTENNAS = 16;

/* per-edge linearisation records, rows of the per-tile scratch (DevWs::jrec) */
constexpr int GR_RANGE = 14; /* A 6 | B 6 | Ow | omega_r                                   */
constexpr int GR_PRIOR = 27; /* the edge's terms of H_ii (upper, 21) and of b_i (6): J^T Ow J and J^T omega_r */
constexpr int GR_SE3 = 151;  /* A 36 | B 36 | omega_r 6 | rho1 | A^T Ow 36 | B^T Ow 36     */

/* hand-off buffer of the elimination, one per step parity: G_i 36 | L_i 21 | z_i 6, [row][lane] */
constexpr int GIT_HAND = 63;

struct GitShared {
    double ant[3 * GIT_MAX_SMEM_ANTENNAS];
    double lam[TILE];                 /* F: lambda of the trial (warp 0 -> warp 1)                        */
    double tscale[TILE];              /* F -> D: computeScale() of the trial (warp 1 -> warp 0)           */
    int tok[TILE];                    /* F: the factorisation succeeded (warp 0 -> warp 1)                */
    int lin[TILE];                    /* D -> all: window starts an iteration (buildSystem)               */
    int cur[TILE];                    /* D -> all: which pose buffer holds the estimate                   */
    int act[TILE];                    /* D -> all: window still being optimised                           */
    int go, anylin, redo;
    double rinc[6][9];                /* the six rotation increments of the numeric Jacobians: fromVectorMQT(+-delta e_k) */
    int rinc_sparse;                  /* each of them is I plus one antisymmetric pair (bit patterns), see jac_v0_m */
    int diag;                         /* every prior information matrix of the tile is diagonal (bit patterns) */
    int ctr[4];                       /* next item of the J, H, U, C phase: the warps of the tile draw their items from a queue */
    int zero;                         /* 0, read through a volatile pointer: keeps per-phase address arithmetic per phase */
    /* LM state of every window, parked here between the phases of warp 0 */
    double lm_d[6][TILE]; /* lambda, ni, stale, plainCur, currentChi, rho */
    int lm_i[8][TILE];    /* iterations, trials_total, flags, qlast, cur, q, it, bit 0 need_lin | bit 1 done */
};

struct LmState {
    double lambda, ni, stale, plainCur, currentChi, rho;
    int iterations, trials_total, flags, qlast, cur, q, it;
    bool need_lin, done;
    UWBGO_DI void load(const GitShared &sh, int lane)
    {
        lambda = sh.lm_d[0][lane]; ni = sh.lm_d[1][lane]; stale = sh.lm_d[2][lane];
        plainCur = sh.lm_d[3][lane]; currentChi = sh.lm_d[4][lane]; rho = sh.lm_d[5][lane];
        iterations = sh.lm_i[0][lane]; trials_total = sh.lm_i[1][lane]; flags = sh.lm_i[2][lane]; qlast = sh.lm_i[3][lane];
        cur = sh.lm_i[4][lane]; q = sh.lm_i[5][lane]; it = sh.lm_i[6][lane];
        need_lin = (sh.lm_i[7][lane] & 1) != 0;
        done = (sh.lm_i[7][lane] & 2) != 0;
    }
    UWBGO_DI void store(GitShared &sh, int lane) const
    {
        sh.lm_d[0][lane] = lambda; sh.lm_d[1][lane] = ni; sh.lm_d[2][lane] = stale;
        sh.lm_d[3][lane] = plainCur; sh.lm_d[4][lane] = currentChi; sh.lm_d[5][lane] = rho;
        sh.lm_i[0][lane] = iterations; sh.lm_i[1][lane] = trials_total; sh.lm_i[2][lane] = flags; sh.lm_i[3][lane] = qlast;
        sh.lm_i[4][lane] = cur; sh.lm_i[5][lane] = q; sh.lm_i[6][lane] = it;
        sh.lm_i[7][lane] = (need_lin ? 1 : 0) | (done ? 2 : 0);
    }
};

UWBGO_DI void bar_pair() { asm volatile("bar.sync 1, 64;" ::: "memory"); }
/* barrier 0 over the whole CTA from warp-specialised branches (what __syncthreads() compiles to) */
UWBGO_DI void cta_bar() { asm volatile("bar.sync 0;" ::: "memory"); }

/* asynchronous copies global -> shared (LDGSTS), 16 bytes per lane and instruction, past L1 */
UWBGO_DI void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
UWBGO_DI void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
UWBGO_DI void cp_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
/* one warp stages `rows` consecutive rows of a tile array ([row][32] doubles, contiguous) */
UWBGO_DI void stage_rows(double *dst, const double *src, int rows, int lane)
{
    const int chunks = rows * (TILE * (int)sizeof(double) / 16);
    for (int c = lane; c < chunks; c += 32)
        cp_async16(reinterpret_cast<char *>(dst) + 16 * c, reinterpret_cast<const char *>(src) + 16 * c);
}

struct GitEnv {
    GenEnv E;
    double *jrec; /* this lane's column of the tile's linearisation records */
    const double *rinc; /* [6][9] rotation increments of the numeric Jacobians (shared memory) */
    bool rinc_sparse;
    int Er, Ep, Es;
};

/* the topology tables every item walks, copied to shared memory once per CTA (they are the same for all tiles) */
struct GitTopo {
    const EdgeRec *edges; /* [E]     */
    const int2 *ops;      /* [n_ops] */
    const int *op_begin;  /* [N + 1] */
    const int *slot_edge; /* [E]     */
};
UWBGO_DI EdgeRec smem_edge(const EdgeRec *e)
{
    const int4 *p = reinterpret_cast<const int4 *>(e);
    const int4 u = p[0], v = p[1];
    EdgeRec r;
    r.kind = u.x; r.a = u.y; r.b = u.z; r.slot = u.w;
    r.ant = v.x; r.robust = v.y; r.base_a = v.z; r.base_b = v.w;
    r.ant_b = e->ant_b;
    return r;
}

/* The information matrix of an EdgeSE3Prior.  The priors Localization builds are diagonal (lidar: one entry,
 * localization.cpp:478-479; IMU: three, localization.cpp:515-518), and a tile whose prior information matrices all have
 * exactly +0.0 off the diagonal (checked once per launch, bit patterns) reads the six diagonal rows only and forms
 * the sums of the terms that are not structurally zero (prior_terms_diag: same chains, same bits in H, b and chi2).
 * Anything else takes the dense reads and loops. */
template <bool DIAG>
UWBGO_DI double info_at(const double *__restrict__ O, int k, int c)
{
    if (DIAG) return k == c ? ROW(O, 7 * k) : 0.0;
    return ROW(O, 6 * k + c);
}
template <bool DIAG>
UWBGO_DI double chi2_6_t(const double *__restrict__ O, const double *e, double *Oe)
{
    double chi = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double s = info_at<DIAG>(O, r, 0) * e[0];
#pragma unroll
        for (int c = 1; c < 6; ++c) s = s + info_at<DIAG>(O, r, c) * e[c];
        Oe[r] = s;
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) chi = chi + e[r] * Oe[r];
    return chi;
}
template <bool DIAG>
UWBGO_DI void jt_omega_t(const double *J, const double *__restrict__ O, bool robust, double r1, double *JtO)
{
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double ow[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double v = info_at<DIAG>(O, k, c);
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

/* gen_jac_v0 / gen_jac_v1 of uwbgo_general.cuh with the roots of a math policy M (NbMath: branch-free, so that the
 * twelve perturbed residuals of an edge interleave; an operand outside its range raises `bad` and the item is
 * evaluated again with IeeeMath: same bits) */
template <class M, bool SPARSE, bool SECOND = false>
UWBGO_DI void jac_v0_m(const GenEnv &E, const double *rinc, const Pose &X, int ant, const double *Q, double d, int c0,
                       int base, double *J, unsigned &bad)
{
    /* Sparse increments (rinc_sparse: B = I + B[j][k] + B[k][j], every other entry 1 or (+-)0): mat3_mul's
     * (A[r][0] B[0][c] + A[r][1] B[1][c]) + A[r][2] B[2][c] is then A[r][c] plus at most one rounded product -- the terms
     * dropped are products of a FINITE number with a zero, which leave a non-zero sum unchanged, and where the sum is
     * zero they can only change its sign, which the squares of dist3 absorb (the perturbed rotation is used for this
     * residual alone).  12 instead of 45 operations per perturbed rotation, same bits in J.  Only with the branch-free
     * policy: a non-finite rotation entry raises `bad`, and the repeat with IeeeMath takes the dense product. */
    static_assert(!SPARSE || std::is_same<M, NbMath>::value, "the sparse product is the branch-free policy's");
    constexpr bool sparse = SPARSE;
    if (sparse && ant > 0) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) t = fma(X.R[k], 0.0, t);
        if (t != t) bad |= 1u;
    }
    const int mod = E.cfg->orth_mod;
    double o[3] = {0.0, 0.0, 0.0};
    if (ant > 0) {
        o[0] = E.ant[3 * (ant - 1)];
        o[1] = E.ant[3 * (ant - 1) + 1];
        o[2] = E.ant[3 * (ant - 1) + 2];
    }
    int call = (c0 + base) % mod;
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double epm[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            if (++call == mod) call = 0;
            const double v = sg == 0 ? E.delta : -E.delta;
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
            epm[sg] = SECOND ? d - dist3m<M>(Q[0], Q[1], Q[2], P[0], P[1], P[2], bad)
                             : d - dist3m<M>(P[0], P[1], P[2], Q[0], Q[1], Q[2], bad);
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
                /* the increment depends on delta alone: the six matrices are formed once per CTA (same operations) */
                double Rp[9], P[3];
                if (sparse) {
                    const int jj = (dd + 1) % 3, kk = (dd + 2) % 3;
                    const double bjk = rinc[9 * (2 * dd + sg) + 3 * jj + kk], bkj = rinc[9 * (2 * dd + sg) + 3 * kk + jj];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        Rp[3 * r + dd] = X.R[3 * r + dd];
                        Rp[3 * r + kk] = X.R[3 * r + kk] + X.R[3 * r + jj] * bjk;
                        Rp[3 * r + jj] = X.R[3 * r + jj] + X.R[3 * r + kk] * bkj;
                    }
                } else {
                    double Rinc[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) Rinc[k] = rinc[9 * (2 * dd + sg) + k];
                    mat3_mul(X.R, Rinc, Rp);
                }
                if (call == 0) orthogonalize(Rp);
                mat3_vec_add(Rp, o, X.t, P);
                epm[sg] = SECOND ? d - dist3m<M>(Q[0], Q[1], Q[2], P[0], P[1], P[2], bad)
                                 : d - dist3m<M>(P[0], P[1], P[2], Q[0], Q[1], Q[2], bad);
            }
            J[3 + dd] = E.scalar * (epm[0] - epm[1]);
        }
    } else {
        J[3] = 0.0; J[4] = 0.0; J[5] = 0.0;
    }
}
template <class M>
UWBGO_DI void jac_v1_m(const GenEnv &E, const double *P0, const Pose &X, double d, double *J, unsigned &bad)
{
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
        double epm[2];
#pragma unroll
        for (int sg = 0; sg < 2; ++sg) {
            const double v = sg == 0 ? E.delta : -E.delta;
            double tp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) tp[r] = X.R[3 * r + dd] * v + X.t[r];
            epm[sg] = d - dist3m<M>(P0[0], P0[1], P0[2], tp[0], tp[1], tp[2], bad);
        }
        J[dd] = E.scalar * (epm[0] - epm[1]);
    }
    J[3] = 0.0; J[4] = 0.0; J[5] = 0.0;
}
/* toVectorMQT(Zinv * Xi^-1 * Xj) (se3_error of uwbgo_general.cuh) */
template <class M>
UWBGO_DI void se3_error_m(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *e, unsigned &bad)
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

/* the next item of a phase for this warp (items are queued heaviest kind first; a static round robin leaves warps
 * idle: the edges of a pose repeat with a period that divides the warp count, so every warp would always get the
 * same kind -- measured on C4a: the residual phase took 6.1 M cycles with 5.2 M of them one warp's wait) */
UWBGO_DI int next_item(int *ctr, int lane)
{
    int u = 0;
    if (lane == 0) u = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, u, 0);
}

/* ---- J items ------------------------------------------------------------------------------------------------ */

/* range edge, slot k: error, weights and the numeric Jacobian wrt vertex 0 (gen_linearize_pose_acc, role 0) */
template <class M, bool SPARSE = false>
UWBGO_DI unsigned item_range_v0(const GitEnv &G, const GitTopo &tt, const PoseBuf &T, int k)
{
    unsigned bad = 0;
    const GenEnv &E = G.E;
    const EdgeRec er = smem_edge(tt.edges + tt.slot_edge[k]);
    double *rec = G.jrec + (size_t)k * GR_RANGE * TILE;
    const double d = ROW(E.p.rd, er.slot), info = ROW(E.p.ri, er.slot);
    Pose Xa;
    load_pose(T, er.a, Xa);
    double P0[3], Q[3], A[6];
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
    const double err = d - dist3m<M>(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2], bad);
    const double Oe = info * err;
    double omega_r = -Oe, Ow = info;
    if (er.robust) {
        const double r1 = E.ck.rho1m<M>(err * Oe, bad);
        omega_r = omega_r * r1;
        Ow = r1 * info;
    }
    jac_v0_m<M, SPARSE>(E, G.rinc, Xa, er.ant, Q, d, E.p.cnt[(size_t)er.a * TILE], er.base_a, A, bad);
#pragma unroll
    for (int j = 0; j < 6; ++j) ROW(rec, j) = A[j];
    ROW(rec, 12) = Ow;
    ROW(rec, 13) = omega_r;
    return bad;
}

/* pose-pose range edge, slot k: numeric Jacobian wrt vertex 1 (gen_linearize_pose_acc, role 1) */
template <class M, bool SPARSE = false>
UWBGO_DI unsigned item_range_v1(const GitEnv &G, const GitTopo &tt, const PoseBuf &T, int k)
{
    unsigned bad = 0;
    const GenEnv &E = G.E;
    const EdgeRec er = smem_edge(tt.edges + tt.slot_edge[k]);
    if (er.kind != UWBGO_EDGE_RANGE_POSE) return 0;
    double *rec = G.jrec + (size_t)k * GR_RANGE * TILE;
    const double d = ROW(E.p.rd, er.slot);
    Pose Xa, Xb;
    load_pose(T, er.a, Xa);
    load_pose(T, er.b, Xb);
    double P0[3], B[6];
    offset_point(E, Xa, er.ant, P0);
    if (er.ant_b > 0)
        jac_v0_m<M, SPARSE, true>(E, G.rinc, Xb, er.ant_b, P0, d, E.p.cnt[(size_t)er.b * TILE], er.base_b, B, bad);
    else
        jac_v1_m<M>(E, P0, Xb, d, B, bad);
#pragma unroll
    for (int j = 0; j < 6; ++j) ROW(rec, 6 + j) = B[j];
    return bad;
}

/* J^T Ow of a 6-D edge (jt_omega of uwbgo_general.cuh), column by column, straight into the record: J is in
 * registers with its structural zeros, Ow = rho1 * Omega when the edge is robust */
UWBGO_DI void store_jto(const double *J, const double *__restrict__ O, bool robust, double r1, double *out)
{
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        double ow[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double v = ROW(O, 6 * k + c);
            ow[k] = robust ? r1 * v : v;
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = J[r] * ow[0];
#pragma unroll
            for (int k = 1; k < 6; ++k) s = fma(J[6 * k + r], ow[k], s);
            ROW(out, 6 * r + c) = s;
        }
    }
}

/* The sums of acc6_b / jt_omega / acc6_diag for a prior whose information matrix is diagonal, restricted to the terms
 * that are not structurally zero: J is block diagonal (two 3x3 blocks), so J^T Ow and J^T Ow J are, and every sum is
 * the three-term chain of its own block in the order of the dense loops.  What the dense loops add on top are
 * products of a FINITE number with a literal zero: they leave a non-zero partial sum unchanged, and where the sum
 * is zero they can only change its sign, which the accumulators of the H item (started at +0.0) absorb.  The caller
 * guards finiteness (any NaN / inf operand -> the dense arithmetic). */
UWBGO_DI void prior_terms_diag(const double *J, const double *Oe, const double *od, bool robust, double r1, double *rec)
{
#pragma unroll
    for (int r = 0; r < 6; ++r) { /* acc6_b */
        const int k0 = r < 3 ? 0 : 3;
        double v = J[6 * k0 + r] * Oe[k0];
        v = fma(J[6 * (k0 + 1) + r], Oe[k0 + 1], v);
        v = fma(J[6 * (k0 + 2) + r], Oe[k0 + 2], v);
        ROW(rec, 21 + r) = v;
    }
    double ow[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ow[c] = robust ? r1 * od[c] : od[c];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const int k0 = r < 3 ? 0 : 3;
        double jto[3]; /* J^T Ow (r, k0 .. k0 + 2) */
#pragma unroll
        for (int k = 0; k < 3; ++k) jto[k] = J[6 * (k0 + k) + r] * ow[k0 + k];
#pragma unroll
        for (int c = r; c < 6; ++c) {
            double v = 0.0;
            if ((c < 3) == (r < 3)) {
                v = jto[0] * J[6 * k0 + c];
                v = fma(jto[1], J[6 * (k0 + 1) + c], v);
                v = fma(jto[2], J[6 * (k0 + 2) + c], v);
            }
            ROW(rec, up_idx(6, r, c)) = v;
        }
    }
}
/* NaN if any of the n values is not finite, else 0 */
UWBGO_DI double nonfinite_probe(const double *x, int n, double t)
{
#pragma unroll
    for (int k = 0; k < 36; ++k)
        if (k < n) t = fma(x[k], 0.0, t);
    return t;
}

/* EdgeSE3Prior, slot s: the edge touches one pose, so its whole contribution to H_ii and b_i is formed here
 * (error, rho1, omega_r, J, J^T Ow, then the sums of acc6_b / acc6_diag) and the H item only adds it */
template <class M, bool DIAG>
UWBGO_DI unsigned item_prior(const GitEnv &G, const GitTopo &tt, const PoseBuf &T, int s)
{
    unsigned bad = 0;
    const GenEnv &E = G.E;
    const EdgeRec er = smem_edge(tt.edges + tt.slot_edge[G.Er + s]);
    double *rec = G.jrec + ((size_t)G.Er * GR_RANGE + (size_t)s * GR_PRIOR) * TILE;
    Pose Zinv, Xi, Dl;
    load_pose(T, er.a, Xi);
    load_Zinv(E.p.pZ, er.slot, Zinv);
    pose_mul(Zinv, Xi, Dl);
    double q[4], e6[6], Oe[6];
    R_to_quat_m<M>(Dl.R, q, bad);
    e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
    e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
    const double *O = E.p.pI + (size_t)er.slot * 36 * TILE;
    double od[6];
    double chi;
    if (DIAG) { /* Oe = Omega e and chi2 = e . Oe, the non-zero terms (see prior_terms_diag) */
#pragma unroll
        for (int k = 0; k < 6; ++k) od[k] = ROW(O, 7 * k);
#pragma unroll
        for (int k = 0; k < 6; ++k) Oe[k] = od[k] * e6[k];
        chi = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) chi = chi + e6[k] * Oe[k];
    } else
        chi = chi2_6_t<false>(O, e6, Oe);
    const double r1 = er.robust ? E.ck.rho1m<M>(chi, bad) : 1.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        Oe[k] = -Oe[k];
        if (er.robust) Oe[k] = Oe[k] * r1;
    }
    double J[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) J[k] = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[6 * r + c] = Dl.R[3 * r + c];
    set_jqq(q, J);
    if (DIAG) {
        double t = nonfinite_probe(Dl.R, 9, 0.0);
        t = nonfinite_probe(e6, 6, t);
        t = nonfinite_probe(od, 6, t);
        t = fma(q[3], 0.0, fma(r1, 0.0, t));
        if (t != t) bad |= 1u; /* a NaN / inf operand: the dense IEEE arithmetic decides what it propagates to */
        prior_terms_diag(J, Oe, od, er.robust != 0, r1, rec);
        return bad;
    }
    double JtO[36];
#pragma unroll
    for (int r = 0; r < 6; ++r) { /* acc6_b */
        double v = J[r] * Oe[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) v = fma(J[6 * k + r], Oe[k], v);
        ROW(rec, 21 + r) = v;
    }
    jt_omega_t<false>(J, O, er.robust != 0, r1, JtO);
#pragma unroll
    for (int r = 0; r < 6; ++r) /* acc6_diag */
#pragma unroll
        for (int c = r; c < 6; ++c) {
            double v = JtO[6 * r] * J[c];
#pragma unroll
            for (int k = 1; k < 6; ++k) v = fma(JtO[6 * r + k], J[6 * k + c], v);
            ROW(rec, up_idx(6, r, c)) = v;
        }
    return bad;
}

/* one of the two analytic Jacobians of EdgeSE3 (se3_jacobians of uwbgo_general.cuh, g2o computeEdgeSE3Gradient with
 * identity offsets): WHICH = 0 wrt vertex 0 (Ji), 1 wrt vertex 1 (Jj) */
template <class M, int WHICH>
UWBGO_DI void se3_jacobian_m(const Pose &Zinv, const Pose &Xi, const Pose &Xj, double *J, unsigned &bad)
{
    Pose Xi_inv, Bm;
    pose_inv(Xi, Xi_inv);
    pose_mul(Xi_inv, Xj, Bm);
#pragma unroll
    for (int k = 0; k < 36; ++k) J[k] = 0.0;
    if (WHICH == 1) {
        Pose AB;
        pose_mul(Zinv, Bm, AB);
        double qE[4];
        R_to_quat_m<M>(AB.R, qE, bad);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[6 * r + c] = AB.R[3 * r + c];
        set_jqq(qE, J);
        return;
    }
    const double *Ra = Zinv.R, *tb = Bm.t;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[6 * r + c] = -Ra[3 * r + c];
    double S[9] = {0.0, -2.0 * tb[2], 2.0 * tb[1], 2.0 * tb[2], 0.0, -2.0 * tb[0], -2.0 * tb[1], 2.0 * tb[0], 0.0};
    double RaS[9];
    mat3_mul(Ra, S, RaS);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) J[6 * r + 3 + c] = RaS[3 * r + c];
    double qA[4], qB[4], Lm[16], Rm[16];
    R_to_quat_m<M>(Ra, qA, bad);
    R_to_quat_m<M>(Bm.R, qB, bad);
    quat_left(qA, Lm);
    quat_right(qB, Rm);
    double wAB = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) wAB = wAB + Lm[k] * Rm[4 * k];
    const double sgn = wAB < 0.0 ? 1.0 : -1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) v = v + Lm[4 * (r + 1) + k] * Rm[4 * k + (c + 1)];
            J[6 * (3 + r) + 3 + c] = sgn * v;
        }
}

/* EdgeSE3, slot s, one item per vertex: the vertex's Jacobian and its J^T Ow (item 0 also leaves omega_r); each
 * item evaluates the error it needs for rho1 itself, so that neither holds both Jacobians */
template <class M, int WHICH>
UWBGO_DI unsigned item_se3(const GitEnv &G, const GitTopo &tt, const PoseBuf &T, int s)
{
    unsigned bad = 0;
    const GenEnv &E = G.E;
    const EdgeRec er = smem_edge(tt.edges + tt.slot_edge[G.Er + G.Ep + s]);
    double *rec = G.jrec + ((size_t)G.Er * GR_RANGE + (size_t)G.Ep * GR_PRIOR + (size_t)s * GR_SE3) * TILE;
    const double *O = E.p.sI + (size_t)er.slot * 36 * TILE;
    Pose Zinv, Xi, Xj;
    load_Zinv(E.p.sZ, er.slot, Zinv);
    load_pose(T, er.a, Xi);
    load_pose(T, er.b, Xj);
    double r1 = 1.0;
    if (WHICH == 0 || er.robust) {
        double e6[6], Oe[6];
        se3_error_m<M>(Zinv, Xi, Xj, e6, bad);
        const double chi = chi2_6_t<false>(O, e6, Oe);
        if (er.robust) r1 = E.ck.rho1m<M>(chi, bad);
        if (WHICH == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double v = -Oe[k];
                if (er.robust) v = v * r1;
                ROW(rec, 72 + k) = v;
            }
            ROW(rec, 78) = r1;
        }
    }
    double J[36];
    se3_jacobian_m<M, WHICH>(Zinv, Xi, Xj, J, bad);
#pragma unroll
    for (int k = 0; k < 36; ++k) ROW(rec, 36 * WHICH + k) = J[k];
    store_jto(J, O, er.robust != 0, r1, rec + (size_t)(79 + 36 * WHICH) * TILE);
    return bad;
}

/* ---- H items ------------------------------------------------------------------------------------------------ */

/* the Jacobian of a 6-D edge wrt the pose playing `role` into registers, structural zeros included (they take
 * part in the sums exactly as in gen_linearize_pose_acc); prior records keep the two 3x3 diagonal blocks only */
UWBGO_DI void load_six_J(const double *rec, bool prior, int role, double *J)
{
    if (prior) {
#pragma unroll
        for (int k = 0; k < 36; ++k) J[k] = 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                J[6 * r + c] = ROW(rec, 3 * r + c);
                J[6 * (3 + r) + 3 + c] = ROW(rec, 9 + 3 * r + c);
            }
    } else {
        const double *p = rec + (size_t)(36 * role) * TILE;
#pragma unroll
        for (int k = 0; k < 36; ++k) J[k] = ROW(p, k);
    }
}

/* H_ii (upper) and b_i of pose i, gathered from the records of the edges touching it in insertion order (the
 * accumulation part of gen_linearize_pose_acc); returns max |H_ii(k, k)| */
UWBGO_DI double item_h_diag(const GitEnv &G, const GitTopo &tt, int i)
{
    const GenEnv &E = G.E;
    double hd[21], bb[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) hd[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) bb[k] = 0.0;
    const int ob = tt.op_begin[i], oe = tt.op_begin[i + 1];
    for (int o = ob; o < oe; ++o) {
        const int2 op = tt.ops[o];
        const EdgeRec er = smem_edge(tt.edges + op.x);
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR || er.kind == UWBGO_EDGE_RANGE_POSE) {
            const double *rec = G.jrec + (size_t)er.slot * GR_RANGE * TILE;
            const double *Jp = rec + (size_t)(6 * op.y) * TILE;
            double J[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) J[k] = ROW(Jp, k);
            acc1_diag(J, ROW(rec, 12), ROW(rec, 13), hd, bb);
        } else if (er.kind == UWBGO_EDGE_PRIOR) {
            const double *rec = G.jrec + ((size_t)G.Er * GR_RANGE + (size_t)er.slot * GR_PRIOR) * TILE;
            double v[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) v[k] = ROW(rec, k);
#pragma unroll
            for (int k = 0; k < 6; ++k) bb[k] = bb[k] + v[21 + k];
#pragma unroll
            for (int k = 0; k < 21; ++k) hd[k] = hd[k] + v[k];
        } else {
            const double *rec = G.jrec + ((size_t)G.Er * GR_RANGE + (size_t)G.Ep * GR_PRIOR + (size_t)er.slot * GR_SE3) * TILE;
            const double *om = rec + (size_t)72 * TILE;
            const double *jto = rec + (size_t)(79 + 36 * op.y) * TILE;
            double J[36], omv[6];
            load_six_J(rec, false, op.y, J);
#pragma unroll
            for (int k = 0; k < 6; ++k) omv[k] = ROW(om, k);
            acc6_b(J, omv, bb);
#pragma unroll
            for (int r = 0; r < 6; ++r) { /* acc6_diag, J^T Ow read row by row */
                double t[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) t[k] = ROW(jto, 6 * r + k);
#pragma unroll
                for (int c = r; c < 6; ++c) {
                    double s = t[0] * J[c];
#pragma unroll
                    for (int k = 1; k < 6; ++k) s = fma(t[k], J[6 * k + c], s);
                    hd[up_idx(6, r, c)] = hd[up_idx(6, r, c)] + s;
                }
            }
        }
    }
    double *h = E.p.HB + (size_t)i * HR_GEN * TILE;
#pragma unroll
    for (int k = 0; k < 21; ++k) ROW(h, k) = hd[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) ROW(h, 57 + k) = bb[k];
    double maxdiag = 0.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const double v = fabs(hd[up_idx(6, r, r)]);
        if (v > maxdiag) maxdiag = v;
    }
    return maxdiag;
}

/* H_{a, i} (rows of the older pose a, columns of pose i): the pose-pose edges in which pose i is vertex 1 */
UWBGO_DI void item_h_off(const GitEnv &G, const GitTopo &tt, int i)
{
    const GenEnv &E = G.E;
    double ho[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) ho[k] = 0.0;
    const int ob = tt.op_begin[i], oe = tt.op_begin[i + 1];
    for (int o = ob; o < oe; ++o) {
        const int2 op = tt.ops[o];
        if (op.y != 1) continue;
        const EdgeRec er = smem_edge(tt.edges + op.x);
        if (er.kind == UWBGO_EDGE_RANGE_POSE) {
            const double *rec = G.jrec + (size_t)er.slot * GR_RANGE * TILE;
            double A[6], B[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                A[k] = ROW(rec, k);
                B[k] = ROW(rec, 6 + k);
            }
            acc1_off(A, B, ROW(rec, 12), ho);
        } else if (er.kind == UWBGO_EDGE_SE3) {
            const double *rec = G.jrec + ((size_t)G.Er * GR_RANGE + (size_t)G.Ep * GR_PRIOR + (size_t)er.slot * GR_SE3) * TILE;
            const double *ato = rec + (size_t)79 * TILE;
            double B[36];
            load_six_J(rec, false, 1, B);
#pragma unroll
            for (int r = 0; r < 6; ++r) { /* acc6_off, A^T Ow read row by row */
                double t[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) t[k] = ROW(ato, 6 * r + k);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    double s = t[0] * B[c];
#pragma unroll
                    for (int k = 1; k < 6; ++k) s = fma(t[k], B[6 * k + c], s);
                    ho[6 * r + c] = ho[6 * r + c] + s;
                }
            }
        }
    }
    double *h = E.p.HB + (size_t)i * HR_GEN * TILE;
#pragma unroll
    for (int k = 0; k < 36; ++k) ROW(h, 21 + k) = ho[k];
}

/* ---- U item: estimate (+) x_i into the trial buffer, oplus counter advanced (gen_update_pose with the branch-free
 * root first) ---- */
UWBGO_DI void item_update(const GenEnv &E, int i, const PoseBuf &Tc, const PoseBuf &Tn, bool linearised)
{
    const double *lp = E.p.LR + (size_t)i * LR_GEN * TILE;
    double x[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = ROW(lp, k);
    Pose X;
    load_pose(Tc, i, X);
    int c = E.p.cnt[(size_t)i * TILE];
    /* the numeric Jacobians of this iteration's buildSystem advanced the counter first */
    if (linearised) c = (c + __ldg(E.tp->num_calls + i)) % E.cfg->orth_mod;
    Pose Q = X;
    int cq = c;
    unsigned bad = 0;
    pose_oplus_m<NbMath>(Q, x, cq, E.cfg->orth_mod, bad);
    if (bad) {
        Q = X;
        cq = c;
        pose_oplus_m<IeeeMath>(Q, x, cq, E.cfg->orth_mod, bad);
    }
    E.p.cnt[(size_t)i * TILE] = cq;
    store_pose(Tn, i, Q);
}

/* ---- C item: computeError + chi2 of edge e (gen_edge_chi with the information matrix streamed) ---------------- */
template <class M, bool DIAG>
UWBGO_DI unsigned item_chi(const GenEnv &E, const GitTopo &tt, const PoseBuf &T, int e, double &chi_out, double &rob_out)
{
    unsigned bad = 0;
    const EdgeRec er = smem_edge(tt.edges + e);
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
        const double err = ROW(E.p.rd, er.slot) - dist3m<M>(P0[0], P0[1], P0[2], Q[0], Q[1], Q[2], bad);
        const double Oe = ROW(E.p.ri, er.slot) * err;
        chi = err * Oe;
    } else if (er.kind == UWBGO_EDGE_PRIOR) {
        Pose Zinv, X, Dl;
        load_pose(T, er.a, X);
        load_Zinv(E.p.pZ, er.slot, Zinv);
        pose_mul(Zinv, X, Dl);
        double q[4], e6[6], Oe[6];
        R_to_quat_m<M>(Dl.R, q, bad);
        e6[0] = Dl.t[0]; e6[1] = Dl.t[1]; e6[2] = Dl.t[2];
        e6[3] = q[0]; e6[4] = q[1]; e6[5] = q[2];
        const double *O = E.p.pI + (size_t)er.slot * 36 * TILE;
        if (DIAG) { /* the non-zero terms of Omega e and e . Oe (see prior_terms_diag); NaN / inf -> dense arithmetic */
            double od[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) od[k] = ROW(O, 7 * k);
#pragma unroll
            for (int k = 0; k < 6; ++k) Oe[k] = od[k] * e6[k];
            chi = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) chi = chi + e6[k] * Oe[k];
            const double t = nonfinite_probe(od, 6, nonfinite_probe(e6, 6, 0.0));
            if (t != t) bad |= 1u;
        } else
            chi = chi2_6_t<false>(O, e6, Oe);
    } else {
        Pose Zinv, Xi, Xj;
        load_pose(T, er.a, Xi);
        load_pose(T, er.b, Xj);
        load_Zinv(E.p.sZ, er.slot, Zinv);
        double e6[6], Oe[6];
        se3_error_m<M>(Zinv, Xi, Xj, e6, bad);
        chi = chi2_6(E.p.sI, er.slot, e6, Oe);
    }
    chi_out = chi;
    rob_out = er.robust ? E.ck.rho0m<M>(chi, bad) : chi;
    return bad;
}

/* ---- F: the elimination chain on warp 0, the back-substitutions of every step on warp 1 ---------------------- */

/* factor_step<6> of uwbgo_block_solver.cuh cut in two.  Warp 0, pose i: S_i = H_ii + lambda I - G_c G_c^T, its
 * Cholesky factor (the diagonal slot keeps 1 / L_jj), z_i, and G_i = H_{i-1,i} L_i^-T; G_i, L_i, z_i go to the
 * hand-off buffer of the step's parity, where the next step (G_i, z_i) and warp 1 (all three) read them. */
/* 0 for every double arithmetic produces, but not provably so: an address that depends on it cannot be formed, and
 * its loads cannot be hoisted, before the value exists (ptxas otherwise lifts every shared-memory load of the step
 * to its top and spills what it loaded; a spill is an L2 round trip at the L1 size this kernel leaves) */
#ifndef UWBGO_GIT_TIES
#define UWBGO_GIT_TIES 0 /* experiment, bit 0: S columns behind the previous pivot, bit 1: G rows behind z, bit 2: G rows 3..5 behind
                          * rows 0..2 (fewer spills, but C4a 8,192 windows 9.9 ms with all three against 9.4 ms without) */
#endif
UWBGO_DI int after(double v) { return ((__double2hiint(v) & 0x7fffffff) == 0x7ff12345) ? 1 : 0; }

template <class M>
UWBGO_DI unsigned factor_main_step(const double *__restrict__ h, const double *__restrict__ Gc,
                                                         double *__restrict__ Gn, bool link, bool has_prev, double lambda)
{
    /* its own register frame (not inlined); returns bit 0 = a pivot was not positive, bit 1 = the branch-free roots
     * flagged an operand.  z of the child pose is read from the child's hand-off buffer. */
    bool ok = true;
    unsigned bad = 0;
    double L[21], z[6];
    int dep = 0;
    /* column j of S is formed right before the potrf eliminates it (same sums, fewer values alive) */
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double *hj = h + dep, *Gj = Gc + dep;
        double gj[6], col[6];
        if (link) {
#pragma unroll
            for (int k = 0; k < 6; ++k) gj[k] = ROW(Gj, j * 6 + k);
        }
#pragma unroll
        for (int r = j; r < 6; ++r) {
            double s = ROW(hj, up_idx(6, j, r));
            if (r == j) s = s + lambda;
            if (link) {
#pragma unroll
                for (int k = 0; k < 6; ++k) s = fma(-(r == j ? gj[k] : ROW(Gj, r * 6 + k)), gj[k], s);
            }
            col[r] = s;
        }
        double s = col[j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
        if (!(s > 0.0)) ok = false;
        const double inv = M::rsqrt_pivot(s, bad);
        L[lo_idx(j, j)] = inv;
#if UWBGO_GIT_TIES & 1
        dep = after(inv);
#endif
#pragma unroll
        for (int r = j + 1; r < 6; ++r) {
            double t = col[r];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
            L[lo_idx(r, j)] = t * inv;
        }
    }
    {
        const double *hz = h + dep, *Gz = Gc + dep;
        double zn[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (link) {
#pragma unroll
            for (int k = 0; k < 6; ++k) zn[k] = ROW(Gz, 57 + k);
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double s = ROW(hz, 57 + r);
            if (link) {
#pragma unroll
                for (int k = 0; k < 6; ++k) s = fma(-ROW(Gz, r * 6 + k), zn[k], s);
            }
#pragma unroll
            for (int k = 0; k < r; ++k) s = fma(-L[lo_idx(r, k)], z[k], s);
            z[r] = s * L[lo_idx(r, r)];
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) ROW(Gn, 57 + k) = z[k];
#if UWBGO_GIT_TIES & 2
    dep = after(z[5]);
#endif
    if (has_prev) {
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const double *hr = h + dep;
            double g[6];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
                double s = ROW(hr, 21 + r * 6 + cc);
#pragma unroll
                for (int k = 0; k < cc; ++k) s = fma(-g[k], L[lo_idx(cc, k)], s);
                g[cc] = s * L[lo_idx(cc, cc)];
            }
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) ROW(Gn, r * 6 + cc) = g[cc];
#if UWBGO_GIT_TIES & 4
            if (r == 2) dep = after(g[5]); /* rows 3..5 after rows 0..2 */
#endif
        }
    }
#pragma unroll
    for (int k = 0; k < 21; ++k) ROW(Gn, 36 + k) = L[k];
    return (ok ? 0u : 1u) | (bad ? 2u : 0u);
}

/* warp 1, pose i: c_i = L_i^-T z_i and M_i = L_i^-T G_i^T into the L record */
UWBGO_DI void factor_helper_step(const double *__restrict__ Gn, double *__restrict__ l, bool has_prev)
{
    double L[21], z[6], c[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) L[k] = ROW(Gn, 36 + k);
#pragma unroll
    for (int k = 0; k < 6; ++k) z[k] = ROW(Gn, 57 + k);
#pragma unroll
    for (int r = 5; r >= 0; --r) {
        double s = z[r];
#pragma unroll
        for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], c[k], s);
        c[r] = s * L[lo_idx(r, r)];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) ROW(l, k) = c[k];
    if (has_prev) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double g[6], m[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) g[r] = ROW(Gn, j * 6 + r);
#pragma unroll
            for (int r = 5; r >= 0; --r) {
                double s = g[r];
#pragma unroll
                for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], m[k], s);
                m[r] = s * L[lo_idx(r, r)];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) ROW(l, 6 + r * 6 + j) = m[r];
        }
    }
}

/* ---- F, cooperative form (UWBGO_GIT_COOP, default) -------------------------------------------------------------
 * The elimination step of a pose cut along its data dependences, so that only the potrf is one warp's chain:
 *   A  every warp: entries of S_i = H_ii + lambda I - G_c G_c^T and the G_c z_c part of z_i (27 sums of 6 FMAs,
 *      written over the staged H record in place: hd rows 0..20, b rows 57..62);
 *   B  warp 0: potrf of S_i (registers only) -> L_i into the hand-off buffer; the other warps meanwhile: c and the
 *      columns of M of pose i + 1 (what factor_helper_step did one pose behind) and the staging of H_{i-1};
 *   C  the other warps: the rows of G_i = H_{i-1,i} L_i^-T and z_i.
 * Every entry is produced by the operation sequence of factor_main_step / factor_helper_step (same operands, same
 * order), by one thread: the bits do not change, only who computes them. */
#ifndef UWBGO_GIT_COOP
#define UWBGO_GIT_COOP 1
#endif
#ifndef UWBGO_GIT_SPARSE_RINC
#define UWBGO_GIT_SPARSE_RINC 1 /* 0: dense rotation increments in the numeric Jacobians (A/B) */
#endif

/* phase A, row R of the upper triangle: S(j, R) for j <= R and the G_c z_c part of z_i[R]; G_c row R and z_c stay in
 * registers, the rows j < R stream through */
template <int R>
UWBGO_DI void coop_S_row(double *__restrict__ h, const double *__restrict__ Gc, bool link, double lambda)
{
    double acc[R + 2];
#pragma unroll
    for (int j = 0; j <= R; ++j) acc[j] = ROW(h, up_idx(6, j, R));
    acc[R] = acc[R] + lambda;
    acc[R + 1] = ROW(h, 57 + R);
    if (link) {
        double gr[6], zn[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            gr[k] = ROW(Gc, R * 6 + k);
            zn[k] = ROW(Gc, 57 + k);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double gj[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) gj[k] = ROW(Gc, j * 6 + k);
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[j] = fma(-gr[k], gj[k], acc[j]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[R] = fma(-gr[k], gr[k], acc[R]);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[R + 1] = fma(-gr[k], zn[k], acc[R + 1]);
    }
#pragma unroll
    for (int j = 0; j <= R; ++j) ROW(h, up_idx(6, j, R)) = acc[j];
    ROW(h, 57 + R) = acc[R + 1];
}

/* warps 1, 2, 3 take the row pairs (0, 5), (1, 4), (2, 3): nine sums and ~48 shared-memory loads each (every further
 * warp would load the rows of G_c again: a row is 256 bytes, two cycles of the SM's shared-memory pipe) */
UWBGO_DI void coop_form_S(double *__restrict__ h, const double *__restrict__ Gc, bool link, double lambda, int warp)
{
    if (warp == 1) {
        coop_S_row<5>(h, Gc, link, lambda);
        coop_S_row<0>(h, Gc, link, lambda);
    } else if (warp == 2) {
        coop_S_row<4>(h, Gc, link, lambda);
        coop_S_row<1>(h, Gc, link, lambda);
    } else if (warp == 3) {
        coop_S_row<3>(h, Gc, link, lambda);
        coop_S_row<2>(h, Gc, link, lambda);
    }
}

/* warp 0, phases B and C: the Cholesky factor of S_i (rows 0..20 of the staged record after phase A; the diagonal slot
 * keeps 1 / L_jj) into the hand-off buffer, then, behind the barrier that releases the G rows, z_i from its G_c z_c
 * part (rows 57..62) with L_i still in registers.  Straight-line code on scalars (generated; the loops of
 * factor_step<6> written out in their order): as loops over L[21] ptxas keeps the array in local memory.  Column
 * j + 1 of S is loaded behind the pivot operand of column j (see after()): it arrives during that pivot's root and
 * the 21 loads cannot all be lifted to the top.  fl: bit 0 = a pivot was not positive, bit 1 = the branch-free roots
 * flagged an operand */
template <class M>
UWBGO_DI void coop_potrf_z(const double *__restrict__ h, double *__restrict__ Gn, unsigned &fl)
{
    bool ok = true;
    unsigned bad = 0;
    double s;
    double c0 = ROW(h, 0), c1 = ROW(h, 1), c2 = ROW(h, 2), c3 = ROW(h, 3), c4 = ROW(h, 4), c5 = ROW(h, 5);
    /* column 0 */
    s = c0;
    const double *h1 = h + after(s);
    const double n1_1 = ROW(h1, 6), n1_2 = ROW(h1, 7), n1_3 = ROW(h1, 8), n1_4 = ROW(h1, 9), n1_5 = ROW(h1, 10);
    if (!(s > 0.0)) ok = false;
    const double l00 = M::rsqrt_pivot(s, bad);
    s = c1;
    const double l10 = s * l00;
    s = c2;
    const double l20 = s * l00;
    s = c3;
    const double l30 = s * l00;
    s = c4;
    const double l40 = s * l00;
    s = c5;
    const double l50 = s * l00;
    c1 = n1_1;
    c2 = n1_2;
    c3 = n1_3;
    c4 = n1_4;
    c5 = n1_5;
    /* column 1 */
    s = c1;
    s = fma(-l10, l10, s);
    const double *h2 = h + after(s);
    const double n2_2 = ROW(h2, 11), n2_3 = ROW(h2, 12), n2_4 = ROW(h2, 13), n2_5 = ROW(h2, 14);
    if (!(s > 0.0)) ok = false;
    const double l11 = M::rsqrt_pivot(s, bad);
    s = c2;
    s = fma(-l20, l10, s);
    const double l21 = s * l11;
    s = c3;
    s = fma(-l30, l10, s);
    const double l31 = s * l11;
    s = c4;
    s = fma(-l40, l10, s);
    const double l41 = s * l11;
    s = c5;
    s = fma(-l50, l10, s);
    const double l51 = s * l11;
    c2 = n2_2;
    c3 = n2_3;
    c4 = n2_4;
    c5 = n2_5;
    /* column 2 */
    s = c2;
    s = fma(-l20, l20, s);
    s = fma(-l21, l21, s);
    const double *h3 = h + after(s);
    const double n3_3 = ROW(h3, 15), n3_4 = ROW(h3, 16), n3_5 = ROW(h3, 17);
    if (!(s > 0.0)) ok = false;
    const double l22 = M::rsqrt_pivot(s, bad);
    s = c3;
    s = fma(-l30, l20, s);
    s = fma(-l31, l21, s);
    const double l32 = s * l22;
    s = c4;
    s = fma(-l40, l20, s);
    s = fma(-l41, l21, s);
    const double l42 = s * l22;
    s = c5;
    s = fma(-l50, l20, s);
    s = fma(-l51, l21, s);
    const double l52 = s * l22;
    c3 = n3_3;
    c4 = n3_4;
    c5 = n3_5;
    /* column 3 */
    s = c3;
    s = fma(-l30, l30, s);
    s = fma(-l31, l31, s);
    s = fma(-l32, l32, s);
    const double *h4 = h + after(s);
    const double n4_4 = ROW(h4, 18), n4_5 = ROW(h4, 19);
    if (!(s > 0.0)) ok = false;
    const double l33 = M::rsqrt_pivot(s, bad);
    s = c4;
    s = fma(-l40, l30, s);
    s = fma(-l41, l31, s);
    s = fma(-l42, l32, s);
    const double l43 = s * l33;
    s = c5;
    s = fma(-l50, l30, s);
    s = fma(-l51, l31, s);
    s = fma(-l52, l32, s);
    const double l53 = s * l33;
    c4 = n4_4;
    c5 = n4_5;
    /* column 4 */
    s = c4;
    s = fma(-l40, l40, s);
    s = fma(-l41, l41, s);
    s = fma(-l42, l42, s);
    s = fma(-l43, l43, s);
    const double *h5 = h + after(s);
    const double n5_5 = ROW(h5, 20);
    if (!(s > 0.0)) ok = false;
    const double l44 = M::rsqrt_pivot(s, bad);
    s = c5;
    s = fma(-l50, l40, s);
    s = fma(-l51, l41, s);
    s = fma(-l52, l42, s);
    s = fma(-l53, l43, s);
    const double l54 = s * l44;
    c5 = n5_5;
    /* column 5 */
    s = c5;
    s = fma(-l50, l50, s);
    s = fma(-l51, l51, s);
    s = fma(-l52, l52, s);
    s = fma(-l53, l53, s);
    s = fma(-l54, l54, s);
    if (!(s > 0.0)) ok = false;
    const double l55 = M::rsqrt_pivot(s, bad);
    ROW(Gn, 36) = l00;
    ROW(Gn, 37) = l10;
    ROW(Gn, 38) = l11;
    ROW(Gn, 39) = l20;
    ROW(Gn, 40) = l21;
    ROW(Gn, 41) = l22;
    ROW(Gn, 42) = l30;
    ROW(Gn, 43) = l31;
    ROW(Gn, 44) = l32;
    ROW(Gn, 45) = l33;
    ROW(Gn, 46) = l40;
    ROW(Gn, 47) = l41;
    ROW(Gn, 48) = l42;
    ROW(Gn, 49) = l43;
    ROW(Gn, 50) = l44;
    ROW(Gn, 51) = l50;
    ROW(Gn, 52) = l51;
    ROW(Gn, 53) = l52;
    ROW(Gn, 54) = l53;
    ROW(Gn, 55) = l54;
    ROW(Gn, 56) = l55;
    fl |= (ok ? 0u : 1u) | (bad ? 2u : 0u);
    double z0 = ROW(h, 57), z1 = ROW(h, 58), z2 = ROW(h, 59), z3 = ROW(h, 60), z4 = ROW(h, 61), z5 = ROW(h, 62);
    cta_bar();
    s = z0;
    z0 = s * l00;
    s = z1;
    s = fma(-l10, z0, s);
    z1 = s * l11;
    s = z2;
    s = fma(-l20, z0, s);
    s = fma(-l21, z1, s);
    z2 = s * l22;
    s = z3;
    s = fma(-l30, z0, s);
    s = fma(-l31, z1, s);
    s = fma(-l32, z2, s);
    z3 = s * l33;
    s = z4;
    s = fma(-l40, z0, s);
    s = fma(-l41, z1, s);
    s = fma(-l42, z2, s);
    s = fma(-l43, z3, s);
    z4 = s * l44;
    s = z5;
    s = fma(-l50, z0, s);
    s = fma(-l51, z1, s);
    s = fma(-l52, z2, s);
    s = fma(-l53, z3, s);
    s = fma(-l54, z4, s);
    z5 = s * l55;
    ROW(Gn, 57) = z0;
    ROW(Gn, 58) = z1;
    ROW(Gn, 59) = z2;
    ROW(Gn, 60) = z3;
    ROW(Gn, 61) = z4;
    ROW(Gn, 62) = z5;
}

/* warps 1, 2, 3, phase C: rows 2 (warp - 1) and 2 (warp - 1) + 1 of G_i = H_{i-1,i} L_i^-T (one load of L_i for both) */
UWBGO_DI void coop_G_rows(const double *__restrict__ h, double *__restrict__ Gn, int r0)
{
    double L[21], g[2][6];
#pragma unroll
    for (int k = 0; k < 21; ++k) L[k] = ROW(Gn, 36 + k);
    const double *src = h + (size_t)(21 + r0 * 6) * TILE;
    double *dst = Gn + (size_t)(r0 * 6) * TILE;
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) g[q][cc] = ROW(src, q * 6 + cc);
#pragma unroll
    for (int cc = 0; cc < 6; ++cc)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double s = g[q][cc];
#pragma unroll
            for (int k = 0; k < cc; ++k) s = fma(-g[q][k], L[lo_idx(cc, k)], s);
            g[q][cc] = s * L[lo_idx(cc, cc)];
        }
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) ROW(dst, q * 6 + cc) = g[q][cc];
}

/* task u of pose p's L record: u = 6: c_p = L_p^-T z_p; u < 6: column u of M_p = L_p^-T G_p^T */
UWBGO_DI void coop_helper_task(const double *__restrict__ Gn, double *__restrict__ l, int u, bool has_prev)
{
    if (u < 6 && !has_prev) return;
    double L[21], g[6], m[6];
#pragma unroll
    for (int k = 0; k < 21; ++k) L[k] = ROW(Gn, 36 + k);
    const double *src = Gn + (size_t)(u < 6 ? u * 6 : 57) * TILE;
#pragma unroll
    for (int r = 0; r < 6; ++r) g[r] = ROW(src, r);
#pragma unroll
    for (int r = 5; r >= 0; --r) {
        double s = g[r];
#pragma unroll
        for (int k = r + 1; k < 6; ++k) s = fma(-L[lo_idx(k, r)], m[k], s);
        m[r] = s * L[lo_idx(r, r)];
    }
    if (u == 6) {
#pragma unroll
        for (int k = 0; k < 6; ++k) ROW(l, k) = m[k];
    } else {
        double *dst = l + (size_t)(6 + u) * TILE;
#pragma unroll
        for (int r = 0; r < 6; ++r) ROW(dst, r * 6) = m[r];
    }
}

/* warps 1 .. NW-1 together stage `rows` consecutive rows of a tile array */
template <int NW>
UWBGO_DI void stage_rows_coop(double *dst, const double *src, int rows, int warp, int lane)
{
    const int chunks = rows * (TILE * (int)sizeof(double) / 16);
    for (int c = (warp - 1) * 32 + lane; c < chunks; c += (NW - 1) * 32)
        cp_async16(reinterpret_cast<char *>(dst) + 16 * c, reinterpret_cast<const char *>(src) + 16 * c);
}

/* warp 1 after the last step: x_i = c_i - M_i x_{i-1} in ascending order and computeScale() (gen_subst_scale,
 * chain case); x_i is left over c_i in the L record.  The L record and b of pose i + 1 are staged into shared
 * memory (two buffers of GIT_HAND rows, tile base `stage`) while pose i is substituted. */
UWBGO_DI double subst_scale_chain(const GenEnv &E, double *stage, int lane, bool ok, double lambda)
{
    const int N = E.tp->N;
    const double *LRt = E.p.LR - lane, *HBt = E.p.HB - lane;
    auto pf = [&](int i) {
        double *dst = stage + (size_t)(i & 1) * GIT_HAND * TILE;
        stage_rows(dst, LRt + (size_t)i * LR_GEN * TILE, LR_GEN, lane);
        stage_rows(dst + LR_GEN * TILE, HBt + ((size_t)i * HR_GEN + 57) * TILE, 6, lane);
        cp_commit();
    };
    __syncwarp(); /* the records were stored by the lanes of this warp */
    pf(0);
    double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double scale = 0.0;
    for (int i = 0; i < N; ++i) {
        if (i + 1 < N) {
            pf(i + 1);
            cp_wait<1>();
        } else
            cp_wait<0>();
        __syncwarp();
        const double *l = stage + (size_t)(i & 1) * GIT_HAND * TILE + lane;
        double *lp = E.p.LR + (size_t)i * LR_GEN * TILE;
        double x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { /* subst_step<6> */
            double v = ROW(l, r);
            if (i > 0) {
#pragma unroll
                for (int j = 0; j < 6; ++j) v = fma(-ROW(l, 6 + r * 6 + j), xp[j], v);
            }
            x[r] = v;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) xp[k] = ok ? x[k] : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) ROW(lp, k) = xp[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) scale = scale + xp[k] * (lambda * xp[k] + ROW(l, LR_GEN + k));
        __syncwarp(); /* the buffer is refilled two poses on */
    }
    return scale;
}

/* subst_scale_chain on four warps: warp 1 substitutes, warps 2..4 feed it.  Once the elimination is over the hand-off
 * and staging areas are one idle block of 4 x GIT_HAND rows, which holds a ring of five (L record | b) slots; each
 * feeding warp copies a third of a slot three poses ahead and releases pose i at one named barrier per pose (a slot
 * is refilled two barriers after its last read).  Alone, warp 1 spent 1,450 cycles per pose issuing its own copies
 * and waiting out the L2 round trip behind them */
UWBGO_DI void bar_subst() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
UWBGO_DI double subst_scale_split(const GenEnv &E, double *ring, int lane, int role, bool ok, double lambda)
{
    constexpr int RING = 5, DIST = 3, SROWS = LR_GEN + 6;
    static_assert(RING * SROWS <= 4 * GIT_HAND, "the ring lives in the hand-off + staging areas");
    static_assert(LR_GEN == 42 && SROWS == 48, "three feeding warps, sixteen rows each");
    const int N = E.tp->N;
    if (role > 0) {
        const double *LRt = E.p.LR - lane, *HBt = E.p.HB - lane;
        auto pf = [&](int i, int slot) {
            double *dst = ring + (size_t)slot * SROWS * TILE;
            if (role < 3)
                stage_rows(dst + (size_t)(16 * (role - 1)) * TILE, LRt + ((size_t)i * LR_GEN + 16 * (role - 1)) * TILE, 16, lane);
            else {
                stage_rows(dst + (size_t)32 * TILE, LRt + ((size_t)i * LR_GEN + 32) * TILE, LR_GEN - 32, lane);
                stage_rows(dst + (size_t)LR_GEN * TILE, HBt + ((size_t)i * HR_GEN + 57) * TILE, 6, lane);
            }
        };
        for (int i = 0; i < DIST; ++i) { /* (a group may be empty: the count per step stays the same) */
            if (i < N) pf(i, i);
            cp_commit();
        }
        int slot_pf = DIST;
        for (int i = 0; i < N; ++i) {
            if (i + DIST < N) pf(i + DIST, slot_pf);
            cp_commit();
            cp_wait<DIST>();
            bar_subst(); /* pose i is in its slot */
            slot_pf = slot_pf + 1 == RING ? 0 : slot_pf + 1;
        }
        cp_wait<0>();
        return 0.0;
    }
    double xp[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double scale = 0.0;
    int slot = 0;
    for (int i = 0; i < N; ++i) {
        bar_subst();
        const double *l = ring + (size_t)slot * SROWS * TILE + lane;
        double *lp = E.p.LR + (size_t)i * LR_GEN * TILE;
        double x[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) { /* subst_step<6> */
            double v = ROW(l, r);
            if (i > 0) {
#pragma unroll
                for (int j = 0; j < 6; ++j) v = fma(-ROW(l, 6 + r * 6 + j), xp[j], v);
            }
            x[r] = v;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) xp[k] = ok ? x[k] : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) ROW(lp, k) = xp[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) scale = scale + xp[k] * (lambda * xp[k] + ROW(l, LR_GEN + k));
        slot = slot + 1 == RING ? 0 : slot + 1;
    }
    return scale;
}

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
lm_general_items_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                        const __grid_constant__ DevWs ws, const int echi_smem)
{
    static_assert(NW >= 5 && NW <= 2 * GIT_HAND, "warp 0 eliminates, warps 1..3 form S and G, warps 1..4 substitute");
    __shared__ GitShared sh;
    extern __shared__ __align__(16) double dyn[]; /* topology | hand-off 2 x [63][32] | staging 2 x [63][32] | chi2 terms [2E][32] */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (tp.K <= GIT_MAX_SMEM_ANTENNAS)
        for (int k = threadIdx.x; k < 3 * tp.K; k += NW * 32) sh.ant[k] = ws.ant[k];
    if (warp == 0) { /* increment_R of uwbgo_math.cuh on (+-delta) e_k, k = threadIdx.x / 2 */
        bool sparse = true;
        if (threadIdx.x < 6) {
            double q[3] = {0.0, 0.0, 0.0};
            const int ax = threadIdx.x >> 1;
            q[ax] = (threadIdx.x & 1) ? -cfg.jdelta : cfg.jdelta;
            double *B = sh.rinc[threadIdx.x];
            increment_R(q, B);
            /* for delta = 1e-9 the increment is I plus the pair B[j][k] = -B[k][j] = -+2 delta (1 - 2 delta^2 rounds to 1);
             * tested on the bit patterns: unit diagonal, (+-)0 in row and column `ax`, the pair finite */
            const int j = (ax + 1) % 3, k = (ax + 2) % 3;
            auto is0 = [](double v) { return (__double_as_longlong(v) & 0x7fffffffffffffffLL) == 0; };
            sparse = B[0] == 1.0 && B[4] == 1.0 && B[8] == 1.0 && is0(B[3 * ax + j]) && is0(B[3 * ax + k]) && is0(B[3 * j + ax]) &&
                     is0(B[3 * k + ax]) && isfinite(B[3 * j + k]) && isfinite(B[3 * k + j]);
        }
        const bool all = __all_sync(0xffffffffu, sparse);
        if (lane == 0) sh.rinc_sparse = (all && UWBGO_GIT_SPARSE_RINC) ? 1 : 0;
    }
    if (threadIdx.x == 0) {
        sh.zero = 0;
        sh.ctr[0] = sh.ctr[1] = sh.ctr[2] = sh.ctr[3] = 0;
        sh.diag = 1;
    }
    /* topology tables first (16-byte aligned records), then the double-precision areas */
    GitTopo tt;
    size_t tbytes = 0;
    {
        unsigned char *tb = reinterpret_cast<unsigned char *>(dyn);
        const size_t o_edges = 0, o_ops = o_edges + sizeof(EdgeRec) * (size_t)tp.E, o_opb = o_ops + sizeof(int2) * (size_t)tp.n_ops,
                     o_se = o_opb + sizeof(int) * (size_t)(tp.N + 1);
        tbytes = (o_se + sizeof(int) * (size_t)tp.E + 15) & ~(size_t)15;
        EdgeRec *edges = reinterpret_cast<EdgeRec *>(tb + o_edges);
        int2 *ops = reinterpret_cast<int2 *>(tb + o_ops);
        int *opb = reinterpret_cast<int *>(tb + o_opb), *se = reinterpret_cast<int *>(tb + o_se);
        const int words = (int)(sizeof(EdgeRec) / 4) * tp.E;
        for (int k = threadIdx.x; k < words; k += NW * 32) reinterpret_cast<int *>(edges)[k] = __ldg(reinterpret_cast<const int *>(tp.edges) + k);
        for (int k = threadIdx.x; k < tp.n_ops; k += NW * 32) ops[k] = __ldg(reinterpret_cast<const int2 *>(tp.ops + k));
        for (int k = threadIdx.x; k <= tp.N; k += NW * 32) opb[k] = __ldg(tp.op_begin + k);
        for (int k = threadIdx.x; k < tp.E; k += NW * 32) se[k] = __ldg(tp.slot_edge + k);
        tt.edges = edges;
        tt.ops = ops;
        tt.op_begin = opb;
        tt.slot_edge = se;
    }
    double *const dsm = dyn + tbytes / sizeof(double);
    const int N = tp.N, NE = tp.E;
    const bool valid = (int64_t)blockIdx.x * TILE + lane < ws.W; /* workspaces are padded to whole tiles */

    /* the thread's view of the workspace, built where a phase needs it (a volatile zero in the lane index keeps
     * the fifteen pointers from living in registers across the phases that do not use them) */
    auto env = [&]() {
        const int ln = lane + *reinterpret_cast<volatile int *>(&sh.zero);
        GitEnv G;
        G.E.tp = &tp;
        G.E.cfg = &cfg;
        G.E.p = thread_ptrs<HR_GEN, LR_GEN>(tp, ws, (int64_t)blockIdx.x * TILE + ln);
        G.E.ant = (tp.K > 0 && tp.K <= GIT_MAX_SMEM_ANTENNAS) ? sh.ant : ws.ant;
        G.E.ck.init(cfg.kdelta);
        G.E.delta = cfg.jdelta;
        G.E.scalar = 1.0 / (2.0 * cfg.jdelta);
        G.rinc = &sh.rinc[0][0];
        G.rinc_sparse = sh.rinc_sparse != 0;
        G.Er = tp.Er;
        G.Ep = tp.Ep;
        G.Es = tp.Es;
        G.jrec = ws.jrec + (size_t)blockIdx.x * ((size_t)tp.Er * GR_RANGE + (size_t)tp.Ep * GR_PRIOR + (size_t)tp.Es * GR_SE3) * TILE + ln;
        return G;
    };
    /* per-edge chi2 terms: always to the workspace (uwbgo_result::edge_chi2 reads the last trial's there), and, when
     * they fit, also into the hand-off / staging area, which is idle between the C and D phases: the ordered sums of
     * D then read shared memory instead of making fifteen L2 round trips */
    auto echi_glob = [&]() {
        const int ln = lane + *reinterpret_cast<volatile int *>(&sh.zero);
        return ws.echi + ((size_t)blockIdx.x * tp.E * 2) * TILE + ln;
    };

    auto chi_phase = [&](const int *who, int sel) {
        const bool on = who[lane] != 0, diag = sh.diag != 0;
        const GitEnv G = env();
        double *echi = echi_glob(), *echi_s = dsm + lane;
        const int b = sh.cur[lane] ^ sel;
        const PoseBuf T{G.E.p.T(b), G.E.p.Rm(b)};
        for (int u = next_item(&sh.ctr[3], lane); u < NE; u = next_item(&sh.ctr[3], lane)) {
            const int e = tt.slot_edge[NE - 1 - u]; /* 6-D edges first: slots are ordered range | prior | se3 */
            if (on) {
                double chi, rob;
                /* (EdgeSE3 with the IEEE sequences at once, see the J phase) */
                if (tt.edges[e].kind == UWBGO_EDGE_SE3 ||
                    (diag ? item_chi<NbMathW, true>(G.E, tt, T, e, chi, rob) : item_chi<NbMathW, false>(G.E, tt, T, e, chi, rob)))
                    item_chi<IeeeMath, false>(G.E, tt, T, e, chi, rob);
                ROW(echi, 2 * e) = chi;
                ROW(echi, 2 * e + 1) = rob;
                if (echi_smem) {
                    ROW(echi_s, 2 * e) = chi;
                    ROW(echi_s, 2 * e + 1) = rob;
                }
            }
        }
    };
    auto chi_sum = [&](double &p, double &r) {
        const double *echi = echi_smem ? dsm + lane : echi_glob();
        double pp = 0.0, rr = 0.0;
        int e = 0;
        for (; e + 8 <= NE; e += 8) { /* loads first, then the two ordered sums */
            double v[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = ROW(echi, 2 * e + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                pp = pp + v[2 * k];
                rr = rr + v[2 * k + 1];
            }
        }
        for (; e < NE; ++e) {
            pp = pp + ROW(echi, 2 * e);
            rr = rr + ROW(echi, 2 * e + 1);
        }
        p = pp;
        r = rr;
    };
    auto publish = [&](const LmState &st) { /* warp 0 */
        sh.lin[lane] = (!st.done && st.need_lin) ? 1 : 0;
        sh.cur[lane] = st.cur;
        sh.act[lane] = st.done ? 0 : 1;
        const unsigned g = __ballot_sync(0xffffffffu, !st.done);
        const unsigned l = __ballot_sync(0xffffffffu, !st.done && st.need_lin);
        if (lane == 0) {
            sh.go = g != 0;
            sh.anylin = l != 0;
            sh.ctr[0] = sh.ctr[1] = sh.ctr[2] = sh.ctr[3] = 0; /* every queue is drained: the next round starts over */
        }
    };

    /* are the prior information matrices of this tile diagonal?  (one pass over them; see info_at) */
    __syncthreads();
    {
        const GitEnv G = env();
        bool ok = true;
        for (int sl = warp; sl < tp.Ep; sl += NW) {
            const double *O = G.E.p.pI + (size_t)sl * 36 * TILE;
#pragma unroll
            for (int k = 0; k < 36; ++k)
                if (k % 7 != 0 && __double_as_longlong(ROW(O, k)) != 0) ok = false;
        }
        if (!__all_sync(0xffffffffu, ok || !valid) && lane == 0) atomicAnd(&sh.diag, 0);
    }
    /* initial computeActiveErrors: every real window, buffer 0 */
    if (warp == 0) {
        sh.cur[lane] = 0;
        sh.act[lane] = valid ? 1 : 0;
    }
    __syncthreads();
    chi_phase(sh.act, 0);
    __syncthreads();
    if (warp == 0) {
        LmState st;
        st.lambda = 0.0; st.ni = 2.0; st.stale = 0.0; st.plainCur = 0.0; st.currentChi = 0.0; st.rho = 0.0;
        st.iterations = 0; st.trials_total = 0; st.flags = 0; st.qlast = 0; st.cur = 0; st.q = 0; st.it = 0;
        st.need_lin = true;
        st.done = !valid || cfg.max_iterations <= 0;
        if (valid) chi_sum(st.plainCur, st.currentChi);
        st.stale = st.plainCur;
        st.store(sh, lane);
        publish(st);
    }
    __syncthreads();

#ifdef UWBGO_GIT_TIMING
    long long tph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tq = git_clk();
    long long tf[4] = {0, 0, 0, 0}; /* cooperative F, warp 0: phase A, potrf, wait for C, prologue / epilogue */
#define GIT_TICK(k) do { long long tn_ = git_clk(); tph[k] += tn_ - tq; tq = tn_; } while (0)
#define GIT_FTICK(k) do { long long tn_ = git_clk(); tf[k] += tn_ - tq; tph[6] += tn_ - tq; tq = tn_; } while (0)
    long long tg[6] = {0, 0, 0, 0, 0, 0}, tw = 0; /* cooperative F, warp 1: work and barrier wait of the phases A, B, C */
#define GIT_WTICK0() do { tw = git_clk(); } while (0)
#define GIT_WTICK(k) do { long long tn_ = git_clk(); tg[k] += tn_ - tw; tw = tn_; } while (0)
#else
#define GIT_WTICK0()
#define GIT_WTICK(k)
#define GIT_TICK(k)
#define GIT_FTICK(k)
#endif
    const int n6 = tp.Ep + tp.Es;
    while (sh.go) {
        if (sh.anylin) {
            /* ---- J ---- */
            {
                const bool on = sh.lin[lane] != 0, diag = sh.diag != 0;
                const GitEnv G = env();
                const int b = sh.cur[lane];
                const PoseBuf T{G.E.p.T(b), G.E.p.Rm(b)};
                /* heaviest kind first: EdgeSE3, priors, range edges (vertex 0, then vertex 1).  Branch-free
                 * arithmetic first; an item whose operands it flags is evaluated again with the IEEE sequences (an
                 * item only writes its own record) */
                const int nS = 2 * tp.Es, nJ = nS + tp.Ep + tp.Er + tp.Er;
                for (int u = next_item(&sh.ctr[0], lane); u < nJ; u = next_item(&sh.ctr[0], lane)) {
                    if (!on) continue;
                    if (u < nS) {
                        /* (IEEE sequences at once: the quaternion of a near-identity rotation divides by 2 - ulp, the one
                         * divisor class the branch-free quotient refuses, and a twist chain is full of them.  NbMathW,
                         * which forms that quotient by the IEEE division on the spot, was tried here too: more spills,
                         * C4b 3.86 instead of 3.77 ms) */
                        if (u & 1)
                            item_se3<IeeeMath, 1>(G, tt, T, u >> 1);
                        else
                            item_se3<IeeeMath, 0>(G, tt, T, u >> 1);
                    }
                    else if (u < nS + tp.Ep) {
                        if (diag ? item_prior<NbMathW, true>(G, tt, T, u - nS) : item_prior<NbMathW, false>(G, tt, T, u - nS))
                            item_prior<IeeeMath, false>(G, tt, T, u - nS);
                    } else if (u < nS + tp.Ep + tp.Er) {
                        if (G.rinc_sparse ? item_range_v0<NbMath, true>(G, tt, T, u - nS - tp.Ep) : item_range_v0<NbMath>(G, tt, T, u - nS - tp.Ep))
                            item_range_v0<IeeeMath>(G, tt, T, u - nS - tp.Ep);
                    } else {
                        if (G.rinc_sparse ? item_range_v1<NbMath, true>(G, tt, T, u - nS - tp.Ep - tp.Er)
                                          : item_range_v1<NbMath>(G, tt, T, u - nS - tp.Ep - tp.Er))
                            item_range_v1<IeeeMath>(G, tt, T, u - nS - tp.Ep - tp.Er);
                    }
                }
            }
            GIT_TICK(8);
            __syncthreads();
            GIT_TICK(0);
            /* ---- H ---- */
            double md = 0.0;
            {
                const bool on = sh.lin[lane] != 0;
                const GitEnv G = env();
                /* the diagonal blocks, then the off-diagonal blocks */
                for (int u = next_item(&sh.ctr[1], lane); u < 2 * N; u = next_item(&sh.ctr[1], lane)) {
                    if (!on) continue;
                    if (u < N) {
                        const double m = item_h_diag(G, tt, u);
                        if (m > md) md = m;
                    } else
                        item_h_off(G, tt, u - N);
                }
            }
            dsm[warp * TILE + lane] = md; /* H -> F: max |H_kk| over the blocks each warp assembled; the hand-off area is idle */
            GIT_TICK(9);
            __syncthreads();
            GIT_TICK(1);
        }
        /* ---- F ---- */
#if UWBGO_GIT_COOP
        {
            if (warp == 0) {
                LmState st;
                st.load(sh, lane);
                if (!st.done && st.need_lin) {
                    st.stale = st.plainCur;
                    double maxdiag = 0.0;
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const double m = dsm[k * TILE + lane];
                        if (m > maxdiag) maxdiag = m;
                    }
                    if (st.it == 0) {
                        st.lambda = cfg.tau * maxdiag;
                        st.ni = 2.0;
                    }
                    st.rho = 0.0;
                    st.q = 0;
                    st.need_lin = false;
                    st.store(sh, lane);
                }
                sh.lam[lane] = st.lambda;
            }
            const GitEnv G = env();
            double *hand = dsm + lane, *stage = dsm + 2 * GIT_HAND * TILE;
            const double *HBt = G.E.p.HB - lane;
            /* one pass over the chain, three barriers per pose (phases A | B | C above) */
            auto sweep = [&](auto math) -> bool {
                using MATH = decltype(math);
                unsigned fl = 0;
                if (warp > 0) {
                    stage_rows_coop<NW>(stage + (size_t)((N - 1) & 1) * GIT_HAND * TILE, HBt + (size_t)(N - 1) * HR_GEN * TILE, HR_GEN, warp, lane);
                    cp_commit();
                    cp_wait<0>();
                }
                cta_bar(); /* (also: lambda is published, the maxima of the H phase have been read) */
                GIT_FTICK(3);
                const double lambda = sh.lam[lane];
                for (int i = N - 1; i >= 0; --i) {
                    double *hi = stage + (size_t)(i & 1) * GIT_HAND * TILE + lane;
                    double *Gc = hand + (size_t)((i + 1) & 1) * GIT_HAND * TILE, *Gn = hand + (size_t)(i & 1) * GIT_HAND * TILE;
                    GIT_WTICK0();
                    coop_form_S(hi, Gc, i + 1 < N, lambda, warp);
                    GIT_WTICK(0);
                    cta_bar();
                    GIT_WTICK(1);
                    GIT_FTICK(0);
                    if (warp == 0) {
                        coop_potrf_z<MATH>(hi, Gn, fl); /* (holds the barrier between phases B and C) */
                        GIT_FTICK(1);
                    } else {
                        if (i > 0) {
                            stage_rows_coop<NW>(stage + (size_t)((i - 1) & 1) * GIT_HAND * TILE, HBt + (size_t)(i - 1) * HR_GEN * TILE, HR_GEN, warp, lane);
                            cp_commit();
                        }
                        if (i + 1 < N)
                            for (int u = warp - 1; u < 7; u += NW - 1) coop_helper_task(Gc, G.E.p.LR + (size_t)(i + 1) * LR_GEN * TILE, u, true);
                        GIT_WTICK(2);
                        cta_bar();
                        GIT_WTICK(3);
                        if (warp <= 3 && i > 0) coop_G_rows(hi, Gn, 2 * (warp - 1));
                        cp_wait<0>();
                        GIT_WTICK(4);
                    }
                    cta_bar();
                    GIT_WTICK(5);
                    GIT_FTICK(2);
                }
                if (warp == 0) {
                    sh.tok[lane] = (fl & 1u) ? 0 : 1;
                    const bool redo = __any_sync(0xffffffffu, sh.act[lane] != 0 && (fl & 2u) != 0);
                    if (lane == 0) sh.redo = redo ? 1 : 0;
                } else
                    for (int u = warp - 1; u < 7; u += NW - 1) coop_helper_task(hand + (size_t)0 * GIT_HAND * TILE, G.E.p.LR, u, false);
                cta_bar();
                GIT_FTICK(3);
                return sh.redo != 0;
            };
            if (sweep(NbMath{})) sweep(IeeeMath{}); /* a trial only writes scratch: the repeat gives the IEEE bits */
            if (warp == 0) GIT_TICK(6);
            if (warp >= 1 && warp <= 4) {
                const double sc = subst_scale_split(G.E, dsm, lane, warp - 1, sh.tok[lane] != 0, sh.lam[lane]);
                if (warp == 1) sh.tscale[lane] = sc;
            }
        }
#else
        if (warp == 0) {
            double lambda;
            {
                LmState st;
                st.load(sh, lane);
                if (!st.done && st.need_lin) {
                    st.stale = st.plainCur;
                    double maxdiag = 0.0;
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const double m = dsm[k * TILE + lane];
                        if (m > maxdiag) maxdiag = m;
                    }
                    if (st.it == 0) {
                        st.lambda = cfg.tau * maxdiag;
                        st.ni = 2.0;
                    }
                    st.rho = 0.0;
                    st.q = 0;
                    st.need_lin = false;
                    st.store(sh, lane);
                }
                lambda = st.lambda;
            }
            sh.lam[lane] = lambda;
            __syncwarp(); /* every lane has read its maxima before the hand-off area is written */
            const GitEnv G = env();
            double *hand = dsm + lane, *stage = dsm + 2 * GIT_HAND * TILE;
            const double *HBt = G.E.p.HB - lane;
            const bool active = sh.act[lane] != 0;
            /* one pass over the chain; the H record of pose i - 1 is staged into shared memory while pose i is
             * eliminated.  Returns true when an active lane's operands left the range of the branch-free roots. */
            auto sweep = [&](auto math) -> bool {
                using MATH = decltype(math);
                unsigned fl = 0;
                bool redo = false;
                stage_rows(stage + (size_t)((N - 1) & 1) * GIT_HAND * TILE, HBt + (size_t)(N - 1) * HR_GEN * TILE, HR_GEN, lane);
                cp_commit();
                for (int i = N - 1; i >= 0; --i) {
                    if (i > 0) {
                        stage_rows(stage + (size_t)((i - 1) & 1) * GIT_HAND * TILE, HBt + (size_t)(i - 1) * HR_GEN * TILE, HR_GEN, lane);
                        cp_commit();
                        cp_wait<1>();
                    } else
                        cp_wait<0>();
                    __syncwarp();
                    fl |= factor_main_step<MATH>(stage + (size_t)(i & 1) * GIT_HAND * TILE + lane, hand + (size_t)((i + 1) & 1) * GIT_HAND * TILE,
                                                 hand + (size_t)(i & 1) * GIT_HAND * TILE, i + 1 < N, i > 0, lambda);
                    if (i == 0) {
                        sh.tok[lane] = (fl & 1u) ? 0 : 1;
                        redo = __any_sync(0xffffffffu, active && (fl & 2u) != 0);
                        if (lane == 0) sh.redo = redo ? 1 : 0;
                    }
                    bar_pair();
                }
                return redo;
            };
            if (sweep(NbMath{})) sweep(IeeeMath{}); /* a trial only writes scratch: the repeat gives the IEEE bits */
            GIT_TICK(6);
        } else if (warp == 1) {
            const GitEnv G = env();
            double *hand = dsm + lane, *stage = dsm + 2 * GIT_HAND * TILE;
            do {
                for (int i = N - 1; i >= 0; --i) {
                    bar_pair();
                    factor_helper_step(hand + (size_t)(i & 1) * GIT_HAND * TILE, G.E.p.LR + (size_t)i * LR_GEN * TILE, i > 0);
                }
            } while (*reinterpret_cast<volatile int *>(&sh.redo)); /* written before the last barrier of the sweep */
            sh.tscale[lane] = subst_scale_chain(G.E, stage, lane, sh.tok[lane] != 0, sh.lam[lane]);
        }
#endif
        __syncthreads();
        GIT_TICK(2);
        /* ---- U ---- */
        {
            const bool on = sh.act[lane] != 0;
            const GitEnv G = env();
            const int c = sh.cur[lane];
            const bool lin = sh.lin[lane] != 0;
            const PoseBuf Tc{G.E.p.T(c), G.E.p.Rm(c)}, Tn{G.E.p.T(c ^ 1), G.E.p.Rm(c ^ 1)};
            for (int i = next_item(&sh.ctr[2], lane); i < N; i = next_item(&sh.ctr[2], lane))
                if (on) item_update(G.E, i, Tc, Tn, lin);
        }
        GIT_TICK(11);
        __syncthreads();
        GIT_TICK(3);
        /* ---- C ---- */
        chi_phase(sh.act, 1);
        GIT_TICK(12);
        __syncthreads();
        GIT_TICK(4);
        /* ---- D ---- */
        if (warp == 0) {
            LmState st;
            st.load(sh, lane);
            GIT_TICK(13);
            if (!st.done) {
                const bool ok = sh.tok[lane] != 0;
                if (!ok) st.flags |= UWBGO_FLAG_CHOL_FAIL;
                double scale = sh.tscale[lane], tplain, tempChi;
                chi_sum(tplain, tempChi);
                GIT_TICK(14);
                st.stale = tplain;
                if (!ok) tempChi = DBL_MAX;
                scale = scale + 1e-3;
                st.rho = (st.currentChi - tempChi) / scale;
                const bool fin = isfinite(tempChi);
                if (!fin) st.flags |= UWBGO_FLAG_NONFINITE;
                if (st.rho > 0.0 && fin) {
                    double t = 2.0 * st.rho - 1.0;
                    double alpha = 1.0 - (t * t) * t;
                    alpha = (cfg.good_hi < alpha) ? cfg.good_hi : alpha;
                    double sf = (cfg.good_lo < alpha) ? alpha : cfg.good_lo;
                    st.lambda = st.lambda * sf;
                    st.ni = 2.0;
                    st.currentChi = tempChi;
                    st.plainCur = tplain;
                    st.cur ^= 1;
                } else {
                    st.lambda = st.lambda * st.ni;
                    st.ni = st.ni * 2.0;
                }
                ++st.q;
                ++st.trials_total;
                if (!(st.rho < 0.0 && st.q < cfg.max_trials)) { /* this iteration is over */
                    ++st.iterations;
                    st.qlast = st.q;
                    if (st.q == cfg.max_trials || st.rho == 0.0) {
                        st.flags |= UWBGO_FLAG_TERMINATED;
                        st.done = true;
                    } else if (++st.it >= cfg.max_iterations) {
                        st.done = true;
                    } else {
                        st.need_lin = true;
                    }
                }
                st.store(sh, lane);
            }
            publish(st);
            GIT_TICK(15);
        }
        __syncthreads();
        GIT_TICK(5);
    }
#ifdef UWBGO_GIT_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 2)
        printf("tile %d cycles (warp 0's own items + wait at the barrier): J %lld + %lld  H %lld + %lld  F %lld + %lld  U %lld + %lld  C %lld + %lld  D (load %lld, sums %lld, decision %lld, barrier %lld)\n",
               (int)blockIdx.x, tph[8], tph[0], tph[9], tph[1], tph[6], tph[2], tph[11], tph[3], tph[12], tph[4], tph[13], tph[14], tph[15], tph[5]);
#if UWBGO_GIT_COOP
    if (threadIdx.x == 0 && blockIdx.x < 2)
        printf("tile %d cooperative elimination, warp 0: phase A + barrier %lld, potrf %lld, waits out phases B / C %lld, prologue / epilogue %lld\n",
               (int)blockIdx.x, tf[0], tf[1], tf[2], tf[3]);
    if (threadIdx.x == 32 && blockIdx.x < 2)
        printf("tile %d cooperative elimination, warp 1: A work %lld + barrier %lld, B work %lld + barrier %lld, C work %lld + barrier %lld\n",
               (int)blockIdx.x, tg[0], tg[1], tg[2], tg[3], tg[4], tg[5]);
#endif
#endif

    if (warp == 0 && valid) {
        LmState st;
        st.load(sh, lane);
        double *chi2_out = ws.chi2 + (int64_t)blockIdx.x * 4 * TILE + lane;
        int32_t *status_out = ws.status + (int64_t)blockIdx.x * 4 * TILE + lane;
        ROW(chi2_out, 0) = st.plainCur;
        ROW(chi2_out, 1) = st.currentChi;
        ROW(chi2_out, 2) = st.stale;
        ROW(chi2_out, 3) = st.lambda;
        ROW(status_out, 0) = st.iterations;
        ROW(status_out, 1) = st.trials_total;
        ROW(status_out, 2) = st.flags;
        ROW(status_out, 3) = st.qlast;
    }
    {
        const GitEnv G = env();
        if (valid && sh.cur[lane]) { /* result always leaves in buffer 0 */
            for (int r = warp; r < N * 3; r += NW) ROW(G.E.p.T0, r) = ROW(G.E.p.T1, r);
            for (int r = warp; r < N * 9; r += NW) ROW(G.E.p.Rm0, r) = ROW(G.E.p.Rm1, r);
        }
    }
}

}  // namespace

size_t general_items_jrec_rows(const DevTopo &t) { return (size_t)t.Er * GR_RANGE + (size_t)t.Ep * GR_PRIOR + (size_t)t.Es * GR_SE3; }

#ifndef UWBGO_GIT_WARPS
#define UWBGO_GIT_WARPS 8
#endif
#ifndef UWBGO_GIT_MINB
#define UWBGO_GIT_MINB 2
#endif
#ifndef UWBGO_GIT_WIDE_WARPS
#define UWBGO_GIT_WIDE_WARPS 16
#endif

/* shared memory of one CTA: topology tables | hand-off 2 x 63 rows | staging 2 x 63 rows | per-edge chi2 terms
 * (the last only when two CTAs still fit an SM with them) */
static size_t git_topo_bytes(const DevTopo &t)
{
    return (sizeof(EdgeRec) * (size_t)t.E + sizeof(int2) * (size_t)t.n_ops + sizeof(int) * (size_t)(t.N + 1 + t.E) + 15) & ~(size_t)15;
}
constexpr size_t GIT_SMEM_TWO = 115000;     /* dynamic + static bytes at which two CTAs share an SM */
constexpr size_t GIT_SMEM_MAX = 200 * 1024; /* beyond this the CTA kernel takes the batch           */
constexpr size_t GIT_HAND_BYTES = 4 * (size_t)GIT_HAND * TILE * sizeof(double);

bool general_items_ok(const DevTopo &t, const DevWs &ws)
{
    return !t.fast && !t.tree && t.N >= 1 && ws.jrec != nullptr && ws.echi != nullptr && ws.LR != nullptr &&
           git_topo_bytes(t) + GIT_HAND_BYTES + sizeof(GitShared) <= GIT_SMEM_MAX;
}

cudaError_t launch_solve_general_items(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws, cudaStream_t st)
{
    const size_t echi = 2 * (size_t)topo.E * TILE * sizeof(double);
    const size_t base = git_topo_bytes(topo) + GIT_HAND_BYTES;
    /* (a shared-memory area of their own for the chi2 terms was measured slower: 39 KB more per CTA on C4a leave ~28 KB
     * of L1 for two CTAs, and every spilled register becomes an L2 round trip: 11.2 vs 10.0 ms at 8,192 windows,
     * 85.6 vs 77.6 ms at 65,536) */
    const int echi_smem = echi <= GIT_HAND_BYTES;
    const size_t sm = base;
    static bool configured[64] = {false}; /* the attribute is per device */
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    auto narrow = lm_general_items_kernel<UWBGO_GIT_WARPS, UWBGO_GIT_MINB>;
    auto wide = lm_general_items_kernel<UWBGO_GIT_WIDE_WARPS, 1>;
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GIT_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GIT_SMEM_MAX);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    /* up to one tile per SM: sixteen warps per tile, one CTA per SM (twice the warps in the item phases) */
    const char *w = getenv("UWBGO_GIT_WIDE");
    const bool use_wide = w ? atoi(w) != 0 : n_tiles(ws.W) <= sms[dev & 63];
    if (use_wide)
        wide<<<(unsigned)n_tiles(ws.W), UWBGO_GIT_WIDE_WARPS * 32, sm, st>>>(topo, cfg, ws, echi_smem);
    else
        narrow<<<(unsigned)n_tiles(ws.W), UWBGO_GIT_WARPS * 32, sm, st>>>(topo, cfg, ws, echi_smem);
    return cudaGetLastError();
}

}  // namespace uwbgo
