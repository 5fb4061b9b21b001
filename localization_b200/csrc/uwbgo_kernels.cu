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
 * Kernel paths (device code in uwbgo_fast.cuh / uwbgo_general.cuh / uwbgo_block_solver.cuh):
 *   CHAIN   the window Localization::addRangeEdge builds (per pose an anchor range edge, then the
 *           trajectory edge to its predecessor), R = I, no antenna offsets.  Rotation rows/columns
 *           of H and b are exactly zero there (the residual never reads R), so only translation
 *           3x3 blocks are carried -- bit-identical to carrying 6x6 blocks.  Straight-line sweeps;
 *           default kernel lm_chain_ws_kernel (two specialised warps per tile), single-warp variant
 *           lm_chain_kernel.
 *   FAST    same arithmetic for any UWB-only structure, table driven (lm_fast_kernel).
 *   GENERAL full 6x6 blocks: rotations, antenna offsets, EdgeSE3Prior, EdgeSE3, chains and
 *           forests (lm_general_kernel).
 * This file holds the LM driver, the __global__ entry points, the layout kernels and the
 * launchers.
 *
 * Compile with --fmad=false (see uwbgo_math.cuh).
 */
#include <cstdlib>
#include <cstring>
#include "uwbgo_fast.cuh"
#include "uwbgo_general.cuh"

namespace uwbgo {

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
    UWBGO_DI double linearize(int buf) const { return fast_linearize<SINK_NONE>(E, E.p.T(buf)); }
    UWBGO_DI bool trial(double lambda, int from, int to, double &scale, double &p, double &r) const
    {
        /* branch-free arithmetic first; operands outside its safe range (never in sane data) raise
         * `bad` and the trial -- which only writes L records and the trial estimates -- runs again
         * with the IEEE sequences */
        unsigned bad = 0;
        bool ok = chain_factor<NbMath>(E, E.p.T(from), lambda, &bad) && (lambda > 0.0);
        chain_solve_chi<NbMath>(E, ok, lambda, E.p.T(from), E.p.T(to), scale, p, r, &bad);
        if (bad) {
            ok = chain_factor<IeeeMath>(E, E.p.T(from), lambda) && (lambda > 0.0);
            chain_solve_chi<IeeeMath>(E, ok, lambda, E.p.T(from), E.p.T(to), scale, p, r);
        }
        return ok;
    }
};

template <>
struct Path<1> {
    FastEnv E;
    UWBGO_DI void chi(int buf, double &p, double &r) const { fast_chi_pass(E, E.p.T(buf), p, r); }
    UWBGO_DI double linearize(int buf) const { return fast_linearize<SINK_NONE>(E, E.p.T(buf)); }
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
                        int32_t *status_out, int &cur_out, int &last_rejected)
{
    last_rejected = 0;
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
            last_rejected = 0;
        } else {
            lambda = lambda * ni;
            ni = ni * 2.0;
            last_rejected = 1;
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

/* The result always leaves in buffer 0.  Normally that is a copy; when the caller wants the per-edge chi2 of
 * the LAST trial (uwbgo_result::edge_chi2) the buffers are swapped instead, so that a rejected last trial's
 * estimates survive in buffer 1, and stale_sel says which buffer the last trial's estimates are in. */
UWBGO_DI void finish_estimates(const DevWs &ws, int64_t tile, int lane, double *T0, double *T1, int rows, int cur,
                               int last_rejected)
{
    if (ws.stale_sel) {
        if (cur)
            for (int r = 0; r < rows; ++r) {
                const double a = ROW(T0, r), b = ROW(T1, r);
                ROW(T0, r) = b;
                ROW(T1, r) = a;
            }
        ws.stale_sel[tile * TILE + lane] = last_rejected;
    } else if (cur) {
        for (int r = 0; r < rows; ++r) ROW(T0, r) = ROW(T1, r);
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
    int cur, last_rej;
    lm_window<MODE>(P, cfg, ws.chi2 + tile * 4 * TILE + lane, ws.status + tile * 4 * TILE + lane, cur, last_rej);
    finish_estimates(ws, tile, lane, P.E.p.T0, P.E.p.T1, tp.N * 3, cur, last_rej);
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
    unsigned bad[2];         /* per warp: some active lane left the safe range of NbMath          */
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
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0, last_rej = 0;
    bool done = !valid || cfg.max_iterations <= 0;
    if (warp == 1) {
        fast_chi_pass(E, E.p.T0, plainCur, currentChi);
        stale = plainCur;
        const double maxdiag = fast_linearize<SINK_NONE>(E, E.p.T0);
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
        bool ok = true;
        double scale = 0.0;
        /* one trial = factor phase + substitution phase of both warps.  It only writes the L records,
         * the trial estimates and the shared hand-off buffers, so it can be repeated: first with the
         * branch-free arithmetic (NbMath), and -- when any active lane of the tile saw an operand
         * outside NbMath's safe range -- once more with the IEEE sequences. */
        auto trial = [&](auto math_tag) {
        using M = decltype(math_tag);
        unsigned bad = 0;
        ok = true;
        scale = 0.0;
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
                        const double err = dt - dist3m<M>(px, py, pz, cx, cy, cz, bad);
                        fast_jac_v0<M>(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, A, &bad);
                        fast_jac_v1<M>(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, B, &bad);
                        chain_weights<M>(E, err, it_, (rob & 2) != 0, Ow, omega_r, &bad);
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
                            const double err = da - dist3m<M>(cx, cy, cz, qx, qy, qz, bad);
                            const double chi = err * (ia * err);
                            p = p + chi;
                            r = r + ((tb.y & 1) ? E.ck.template rho0m<M>(chi, bad) : chi);
                        }
                        if (i > 0) {
                            const double err = dt - dist3m<M>(vx, vy, vz, cx, cy, cz, bad);
                            const double chi = err * (it_ * err);
                            p = p + chi;
                            r = r + ((tb.y & 2) ? E.ck.template rho0m<M>(chi, bad) : chi);
                        }
                        vx = cx; vy = cy; vz = cz;
                        da = nda; ia = nia; dt = ndt; it_ = nit; qx = nqx; qy = nqy; qz = nqz; tb = ntb;
                    }
                    ws_barrier();
                }
                sh.chi[0][lane] = p;
                sh.chi[1][lane] = r;
            }
        } else {
            /* ---------------- chain warp, factor phase: poses i = N-1 .. 0 ---------------- */
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
                        const double err = da - dist3m<M>(cx, cy, cz, qx, qy, qz, bad);
                        fast_jac_v0<M>(cx, cy, cz, qx, qy, qz, da, E.delta, E.scalar, J, &bad);
                        chain_weights<M>(E, err, ia, (rob & 1) != 0, Ow, omega_r, &bad);
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
                    factor_step<3, M>(h, l, true, i > 0, lambda, G, zn, ok, &bad);
                    chain_store_b(l, h);
                    cx = ncx; cy = ncy; cz = ncz; da = nda; ia = nia; qx = nqx; qy = nqy; qz = nqz; rob = nrob;
                    ws_barrier();
                }
                ok = ok && (lambda > 0.0);
            }
            /* ---------------- chain warp, substitution phase ---------------- */
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
        }
        const unsigned wb = __ballot_sync(0xffffffffu, act && bad != 0u);
        if (lane == 0) sh.bad[warp] = wb;
        ws_barrier(); /* chi2 and the range flags published */
        };
        trial(NbMath{});
        if (sh.bad[0] | sh.bad[1]) trial(IeeeMath{}); /* uniform over the CTA */
        if (warp == 0) {
            ws_barrier(); /* LM state published */
        } else {
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
                    last_rej = 0;
                    cur ^= 1;
                } else {
                    lambda = lambda * ni;
                    ni = ni * 2.0;
                    last_rej = 1;
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
        finish_estimates(ws, w / TILE, lane, E.p.T0, E.p.T1, N * 3, cur, last_rej);
        if (E.p.cnt) {
            for (int i = 0; i < N; ++i) {
                long long cc = (long long)E.p.cnt[(size_t)i * TILE] + (long long)iterations * __ldg(tp.num_calls + i) + trials_total;
                E.p.cnt[(size_t)i * TILE] = (int)(cc % cfg.orth_mod);
            }
        }
    }
}

}  // namespace uwbgo
#include "uwbgo_chain_tma.cuh"
#include "uwbgo_chain_tma4.cuh"
namespace uwbgo {

/* antenna lever arms: a handful of doubles every range edge reads -> shared memory, loaded once */
constexpr int MAX_SMEM_ANTENNAS = 16;
UWBGO_DI void gen_env_init(GenEnv &E, const DevTopo &tp, const DevCfg &cfg, const DevWs &ws, int64_t w,
                           double *s_ant)
{
    E.tp = &tp;
    E.cfg = &cfg;
    E.p = thread_ptrs<HR_GEN, LR_GEN>(tp, ws, w);
    E.ant = (tp.K > 0 && tp.K <= MAX_SMEM_ANTENNAS) ? s_ant : ws.ant;
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
}

__global__ void __launch_bounds__(CTA_THREADS)
lm_general_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                  const __grid_constant__ DevWs ws)
{
    __shared__ double s_ant[3 * MAX_SMEM_ANTENNAS];
    if (tp.K <= MAX_SMEM_ANTENNAS)
        for (int k = threadIdx.x; k < 3 * tp.K; k += CTA_THREADS) s_ant[k] = ws.ant[k];
    __syncthreads();
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    Path<0> P;
    gen_env_init(P.E, tp, cfg, ws, w, s_ant);
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    int cur, last_rej; /* (edge_chi2 / marginals of 6x6 windows come from the CTA kernel's scratch, not from here) */
    lm_window<0>(P, cfg, ws.chi2 + tile * 4 * TILE + lane, ws.status + tile * 4 * TILE + lane, cur, last_rej);
    if (cur) {
        for (int r = 0; r < tp.N * 3; ++r) ROW(P.E.p.T0, r) = ROW(P.E.p.T1, r);
        for (int r = 0; r < tp.N * 9; ++r) ROW(P.E.p.Rm0, r) = ROW(P.E.p.Rm1, r);
    }
}

}  // namespace uwbgo
#include "uwbgo_general_cta.cuh"
namespace uwbgo {

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
    fast_linearize<SINK_RECORDS>(E, E.p.T0);
    double *c = ws.chi2 + (w / TILE) * 2 * TILE + (w % TILE);
    ROW(c, 0) = p;
    ROW(c, 1) = r;
}

__global__ void __launch_bounds__(CTA_THREADS)
linearize_general_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                         const __grid_constant__ DevWs ws)
{
    __shared__ double s_ant[3 * MAX_SMEM_ANTENNAS];
    if (tp.K <= MAX_SMEM_ANTENNAS)
        for (int k = threadIdx.x; k < 3 * tp.K; k += CTA_THREADS) s_ant[k] = ws.ant[k];
    __syncthreads();
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    GenEnv E;
    gen_env_init(E, tp, cfg, ws, w, s_ant);
    double p, r;
    PoseBuf T0{E.p.T0, E.p.Rm0};
    gen_chi_pass(E, T0, p, r);
    gen_linearize(E, T0);
    double *c = ws.chi2 + (w / TILE) * 2 * TILE + (w % TILE);
    ROW(c, 0) = p;
    ROW(c, 1) = r;
}

/* The same stage with the work list of a tile spread out: one warp (= one CTA) per (tile, pose) builds that
 * pose's H record (gen_linearize_pose, lane = window); the chi2 passes -- one warp walks a tile's edges in
 * insertion order -- lead the grid.  Thread-per-window gives a 6x6 batch of a few thousand windows a handful
 * of warps per SM; this gives it N + 1 times as many.  Same device functions, same bits. */
__global__ void __launch_bounds__(32)
linearize_general_pose_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                              const __grid_constant__ DevWs ws)
{
    __shared__ double s_ant[3 * MAX_SMEM_ANTENNAS];
    const int lane = threadIdx.x;
    if (tp.K <= MAX_SMEM_ANTENNAS)
        for (int k = lane; k < 3 * tp.K; k += 32) s_ant[k] = ws.ant[k];
    __syncwarp();
    const int64_t tiles = n_tiles(ws.W);
    const bool chi = (int64_t)blockIdx.x < tiles;
    const int64_t item = (int64_t)blockIdx.x - tiles;
    const int64_t tile = chi ? (int64_t)blockIdx.x : item / tp.N;
    const int64_t w = tile * TILE + lane;
    if (w >= ws.W) return; /* padded lanes of the last tile */
    GenEnv E;
    gen_env_init(E, tp, cfg, ws, w, s_ant);
    const PoseBuf T0{E.p.T0, E.p.Rm0};
    if (chi) {
        double p, r;
        gen_chi_pass(E, T0, p, r);
        double *c = ws.chi2 + tile * 2 * TILE + lane;
        ROW(c, 0) = p;
        ROW(c, 1) = r;
    } else
        gen_linearize_pose<false>(E, T0, (int)(item % tp.N));
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
    if (j.mode == 4) { /* one shared row (UWBGO_SHARED_ANCHORS) to every window of a tile-layout array */
        const int64_t tile = blockIdx.x;
        const int c0 = blockIdx.y * 32;
        for (int y = threadIdx.y; y < 32; y += 8) {
            int c = c0 + y;
            if (c >= j.C) continue;
            static_cast<double *>(j.dst)[(tile * j.C + c) * TILE + threadIdx.x] = static_cast<const double *>(j.src)[c];
        }
        return;
    }
    if (j.mode == 5) { /* UWBGO_DIAG_INFO: src is [W][C / 36][6], the diagonals of 6x6 matrices; dst the tile-layout
                        * [tile][C][32] array of the full matrices, +0.0 off the diagonal */
        const int64_t tile = blockIdx.x;
        const int c0 = blockIdx.y * 32;
        const int64_t w = tile * TILE + threadIdx.x;
        for (int y = threadIdx.y; y < 32; y += 8) {
            int c = c0 + y;
            if (c >= j.C) continue;
            const int k = c % 36;
            double v = 0.0;
            if (k % 7 == 0 && w < jobs.W) v = static_cast<const double *>(j.src)[w * (int64_t)(j.C / 6) + (c / 36) * 6 + k / 7];
            static_cast<double *>(j.dst)[(tile * j.C + c) * TILE + threadIdx.x] = v;
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

/* ------------------------------------------------------------------------------------------ */
/* compact range form: the edge parameters of Localization::addRangeEdge / create_range_edge     */
/* (reference localization.cpp:316-319,331,338,350,608-627) built on the device, straight into   */
/* the measurement / information rows of the tile layout.  FP64 after an exact widening of the   */
/* float32 message fields; x^2 = x * x; (v_max * dt) / 3; one IEEE division: the bits of the host.*/
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
pack_range_msgs_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ RangeMsgsDev m, int64_t W,
                       double *__restrict__ rd, double *__restrict__ ri)
{
    const int64_t tile = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t w = tile * TILE + lane;
    for (int s = blockIdx.y * 8 + threadIdx.y; s < tp.Er; s += 8 * gridDim.y) {
        const int sub = __ldg(tp.range_sub + s);
        const int k = sub & 0x3fffffff;
        double d = 0.0, info = 0.0; /* pad lanes of the last tile: zero rows, as the transposing pack leaves them */
        if (w < W) {
            if (sub & 0x40000000) {
                const double mv = m.v_max * m.dt_pose[w * tp.Erp + k] / 3.0;
                info = 1.0 / (mv * mv);
            } else {
                const double derr = (double)m.distance_err[w * tp.Era + k];
                double cov = derr * derr;
                if (m.dt_anchor) {
                    const double mv = m.v_max * m.dt_anchor[w * tp.Era + k] / 3.0;
                    cov = cov + mv * mv;
                }
                d = (double)m.distance[w * tp.Era + k];
                info = 1.0 / cov;
            }
        }
        rd[(tile * tp.Er + s) * TILE + lane] = d;
        ri[(tile * tp.Er + s) * TILE + lane] = info;
    }
}

cudaError_t launch_pack_range_msgs(const DevTopo &topo, const RangeMsgsDev &m, int64_t W, double *rd, double *ri,
                                   cudaStream_t st)
{
    if (W <= 0 || topo.Er <= 0) return cudaSuccess;
    dim3 grid((unsigned)n_tiles(W), (unsigned)std::min(4, (topo.Er + 7) / 8));
    pack_range_msgs_kernel<<<grid, dim3(32, 8), 0, st>>>(topo, m, W, rd, ri);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* uwbgo_result::edge_chi2: edge->chi2() of every edge as g2o holds it after optimize() -- the    */
/* errors of the LAST trial (reference localization.cpp:172-181 would test these).                */
/*   6x6 windows: the CTA kernel's per-edge scratch still holds the last trial's terms.           */
/*   translation-only windows: recomputed from the estimate buffer the last trial wrote           */
/*   (DevWs::stale_sel), with the expression of fast_chi_pass / chain_solve_chi.                  */
/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(CTA_THREADS)
edge_chi2_out_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevWs ws, double *__restrict__ out)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    double *o = out + w * tp.E;
    if (!tp.fast) {
        const double *echi = ws.echi + (tile * tp.E * 2) * TILE + lane;
        for (int e = 0; e < tp.E; ++e) o[e] = ROW(echi, 2 * e);
        return;
    }
    const int sel = ws.stale_sel[tile * TILE + lane];
    const double *T = ws.T[sel] + (tile * (size_t)tp.N * 3) * TILE + lane;
    const double *anch = ws.anch + (tile * (size_t)tp.A * 3) * TILE + lane;
    const double *rd = ws.rd + (tile * (size_t)tp.Er) * TILE + lane, *ri = ws.ri + (tile * (size_t)tp.Er) * TILE + lane;
    for (int e = 0; e < tp.E; ++e) {
        const EdgeRec er = load_edge(tp.edges + e);
        const double *ta = T + (size_t)er.a * 3 * TILE;
        double qx, qy, qz;
        if (er.kind == UWBGO_EDGE_RANGE_ANCHOR) {
            qx = ROW(anch, er.b * 3); qy = ROW(anch, er.b * 3 + 1); qz = ROW(anch, er.b * 3 + 2);
        } else {
            const double *tb = T + (size_t)er.b * 3 * TILE;
            qx = ROW(tb, 0); qy = ROW(tb, 1); qz = ROW(tb, 2);
        }
        const double n = dist3(ROW(ta, 0), ROW(ta, 1), ROW(ta, 2), qx, qy, qz);
        const double err = ROW(rd, er.slot) - n;
        const double Oe = ROW(ri, er.slot) * err;
        o[e] = err * Oe;
    }
}

cudaError_t launch_edge_chi2_out(const DevTopo &topo, const DevWs &ws, double *edge_chi2, cudaStream_t st)
{
    if (ws.W <= 0 || topo.E <= 0) return cudaSuccess;
    edge_chi2_out_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, ws, edge_chi2);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* uwbgo_result::marginal: computeMarginals(spinv, newest vertex) (reference localization.cpp:    */
/* 185-189): the diagonal block of H^-1 of the newest pose, H = the H records of the last         */
/* buildSystem (no damping).  Elimination newest pose first as in the LM trials, lambda = 0:      */
/*   S_i = H_ii - sum over children G_c G_c^T = L_i L_i^T,  G_i = H_{p(i),i} L_i^-T,              */
/*   M_i = L_i^-T G_i^T; then along the path root -> newest:                                      */
/*   Sigma_root = S_root^-1,  Sigma_i = S_i^-1 + M_i Sigma_p(i) M_i^T.                            */
/* One thread per window; scratch [tile][N * 108][32] (G | M | S^-1 per pose).  Run once per      */
/* solve, after the LM kernel: plain loops, nothing here is on the hot path.                      */
/* ------------------------------------------------------------------------------------------ */
constexpr int MG_ROWS = 108;
size_t marginal_scratch_bytes(const DevTopo &topo, int64_t W)
{
    return topo.fast ? 0 : sizeof(double) * (size_t)n_tiles(W) * TILE * (size_t)topo.N * MG_ROWS;
}

__global__ void __launch_bounds__(CTA_THREADS)
marginal_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevWs ws, double *__restrict__ scratch,
                double *__restrict__ out, int32_t *__restrict__ okv)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    const int N = tp.N;
    double *o = out + w * 36;
    const int iterations = ROW(ws.status + tile * 4 * TILE + lane, 0);
    bool good = !tp.fast && iterations > 0; /* translation-only windows: the rotations are unobserved, H is singular */
    if (good) {
        const double *HB = ws.HB + (tile * (size_t)N * HR_GEN) * TILE + lane;
        double *mg = scratch + (tile * (size_t)N * MG_ROWS) * TILE + lane;
        for (int i = N - 1; i >= 0 && good; --i) {
            double S[36], L[36], Li[36];
            const double *h = HB + (size_t)i * HR_GEN * TILE;
            double *G = mg + (size_t)i * MG_ROWS * TILE, *M = G + 36 * TILE, *Sinv = G + 72 * TILE;
            for (int r = 0; r < 6; ++r)
                for (int c = 0; c <= r; ++c) S[6 * r + c] = ROW(h, up_idx(6, c, r));
            for (int ci = __ldg(tp.child_begin + i); ci < __ldg(tp.child_begin + i + 1); ++ci) { /* children, descending */
                const double *Gc = mg + (size_t)__ldg(tp.children + ci) * MG_ROWS * TILE;
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c <= r; ++c) {
                        double s = S[6 * r + c];
                        for (int k = 0; k < 6; ++k) s = fma(-ROW(Gc, 6 * r + k), ROW(Gc, 6 * c + k), s);
                        S[6 * r + c] = s;
                    }
            }
            for (int k = 0; k < 36; ++k) L[k] = 0.0;
            for (int j = 0; j < 6; ++j) {
                double s = S[6 * j + j];
                for (int k = 0; k < j; ++k) s = fma(-L[6 * j + k], L[6 * j + k], s);
                if (!(s > 0.0)) {
                    good = false;
                    break;
                }
                const double inv = 1.0 / sqrt(s);
                L[6 * j + j] = inv; /* the diagonal slot stores 1/L_jj */
                for (int r = j + 1; r < 6; ++r) {
                    double t = S[6 * r + j];
                    for (int k = 0; k < j; ++k) t = fma(-L[6 * r + k], L[6 * j + k], t);
                    L[6 * r + j] = t * inv;
                }
            }
            if (!good) break;
            for (int k = 0; k < 36; ++k) Li[k] = 0.0;
            for (int j = 0; j < 6; ++j) { /* Li = L^-1 (lower), then S^-1 = Li^T Li */
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
                    ROW(Sinv, 6 * r + c) = s;
                    ROW(Sinv, 6 * c + r) = s;
                }
            if (__ldg(tp.parent + i) >= 0) {
                double Gl[36], Ml[36];
                for (int r = 0; r < 6; ++r)
                    for (int cc = 0; cc < 6; ++cc) {
                        double s = ROW(h, 21 + r * 6 + cc);
                        for (int k = 0; k < cc; ++k) s = fma(-Gl[6 * r + k], L[6 * cc + k], s);
                        Gl[6 * r + cc] = s * L[6 * cc + cc];
                    }
                for (int j = 0; j < 6; ++j)
                    for (int r = 5; r >= 0; --r) {
                        double s = Gl[6 * j + r];
                        for (int k = r + 1; k < 6; ++k) s = fma(-L[6 * k + r], Ml[6 * k + j], s);
                        Ml[6 * r + j] = s * L[6 * r + r];
                    }
                for (int k = 0; k < 36; ++k) {
                    ROW(G, k) = Gl[k];
                    ROW(M, k) = Ml[k];
                }
            }
        }
        if (good) { /* the path from the newest pose up to its root, walked back down */
            int root = N - 1, len = 1;
            while (__ldg(tp.parent + root) >= 0) {
                root = __ldg(tp.parent + root);
                ++len;
            }
            double Sg[36];
            {
                const double *Sinv = mg + ((size_t)root * MG_ROWS + 72) * TILE;
                for (int k = 0; k < 36; ++k) Sg[k] = ROW(Sinv, k);
            }
            for (int step = len - 2; step >= 0; --step) {
                int i = N - 1;
                for (int k = 0; k < step; ++k) i = __ldg(tp.parent + i); /* the pose `step` links above the newest */
                const double *M = mg + ((size_t)i * MG_ROWS + 36) * TILE, *Sinv = M + 36 * TILE;
                double T[36], Nw[36];
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 6; ++c) {
                        double s = 0.0;
                        for (int j = 0; j < 6; ++j) s = fma(ROW(M, 6 * r + j), Sg[6 * j + c], s);
                        T[6 * r + c] = s;
                    }
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c <= r; ++c) {
                        double s = ROW(Sinv, 6 * r + c);
                        for (int j = 0; j < 6; ++j) s = fma(T[6 * r + j], ROW(M, 6 * c + j), s);
                        Nw[6 * r + c] = s;
                        Nw[6 * c + r] = s;
                    }
                for (int k = 0; k < 36; ++k) Sg[k] = Nw[k];
            }
            for (int k = 0; k < 36; ++k) o[k] = Sg[k];
        }
    }
    if (!good) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int k = 0; k < 36; ++k) o[k] = nan;
    }
    if (okv) okv[w] = good ? 1 : 0;
}

cudaError_t launch_marginal(const DevTopo &topo, const DevWs &ws, double *scratch, double *marginal,
                            int32_t *marginal_ok, cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    marginal_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, ws, scratch, marginal, marginal_ok);
    return cudaGetLastError();
}

/* dynamic shared memory of a FAST launch; anchors go to shared memory when 4 CTAs/SM still fit */
static size_t fast_smem_bytes(const DevTopo &topo, int threads, int *anchors_in_smem)
{
    size_t stash = sizeof(double) * STASH_PER_THREAD * threads;
    size_t anch = sizeof(double) * (size_t)topo.A * 3 * threads;
    /* budget: the CTAs that fill an SM (512 threads) must fit in its 227 KB */
    *anchors_in_smem = (topo.A > 0 && (stash + anch) * (512 / threads) <= 220 * 1024) ? 1 : 0;
    return stash + (*anchors_in_smem ? anch : 0);
}

#ifndef UWBGO_CHAIN_WARPS
#define UWBGO_CHAIN_WARPS 0 /* warps per tile of the CHAIN kernel: 2 (lm_chain_tma_kernel), 3 or 4 (lm_chain_tma4_kernel); 0 = by batch size */
#endif
/* Two warps per tile and eight tiles per SM retire the most tiles per second once the device is full; up to one
 * wave of the three-warp kernel (five tiles per SM) the kernel time is the latency of ONE tile, and the third
 * warp shortens it (8,192 windows: 1.91 instead of 2.11 ms; 2,048: 1.76 instead of 2.10 ms;
 * profiles/r02_chain_warps_ab.txt).  Developer switch: the environment variable UWBGO_CHAIN_WARPS overrides. */
static int chain_warps_per_tile(int64_t tiles)
{
    static int forced = -1, sms = 0;
    if (forced < 0) {
        const char *e = getenv("UWBGO_CHAIN_WARPS");
        forced = e ? atoi(e) : 0;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    if (forced) return forced;
    if (UWBGO_CHAIN_WARPS) return UWBGO_CHAIN_WARPS;
    return tiles <= (int64_t)UWBGO_TMA3_MINB * sms ? 3 : 2;
}

cudaError_t launch_solve(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws, cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    if (topo.fast) {
#ifndef UWBGO_CHAIN_WS
#define UWBGO_CHAIN_WS 1 /* 1: warp-specialised CHAIN kernel, 0: single-warp CHAIN kernel */
#endif
#ifndef UWBGO_CHAIN_TMA
#define UWBGO_CHAIN_TMA 1 /* 1: operands of the warp-specialised CHAIN kernel staged by the copy engine */
#endif
        if (topo.fast == 2 && UWBGO_CHAIN_WS && UWBGO_CHAIN_TMA) {
            const int nw = chain_warps_per_tile(n_tiles(ws.W_batch > ws.W ? ws.W_batch : ws.W));
            if (nw == 4)
                lm_chain_tma4_kernel<4><<<(unsigned)n_tiles(ws.W), 128, 0, st>>>(topo, cfg, ws);
            else if (nw == 3)
                lm_chain_tma4_kernel<3><<<(unsigned)n_tiles(ws.W), 96, 0, st>>>(topo, cfg, ws);
            else
                lm_chain_tma_kernel<<<(unsigned)n_tiles(ws.W), 64, 0, st>>>(topo, cfg, ws);
            return cudaGetLastError();
        }
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
    } else {
#ifndef UWBGO_GEN_CTA
#define UWBGO_GEN_CTA 1 /* 1: one CTA per tile, phases split over its warps; 0: one thread per window */
#endif
#ifndef UWBGO_GEN_ITEMS
#define UWBGO_GEN_ITEMS 1 /* 1: chains take the ITEM kernel (uwbgo_general_items.cu); UWBGO_GENERAL_KERNEL=cta in the
                           * environment selects the CTA kernel at run time (A/B runs, tests of both) */
#endif
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const unsigned tiles = (unsigned)n_tiles(ws.W);
        if (UWBGO_GEN_CTA && UWBGO_GEN_ITEMS && general_items_ok(topo, ws)) {
            /* chains take the ITEM kernel (eight warps per tile at 128 registers and two tiles per SM; sixteen warps and
             * one tile per SM up to one tile per SM) at every batch size: with the cooperative elimination it is ahead
             * of the 8-warp CTA kernel below for EdgeSE3 windows too (4,096 C4b windows: 2.23 vs 2.74 ms, 2,048: 1.63 vs
             * 2.43 ms; C4a 4,096: 4.72 vs 7.08 ms).  UWBGO_GENERAL_KERNEL=cta | items overrides the choice (A/B runs,
             * tests of both). */
            const char *sel = getenv("UWBGO_GENERAL_KERNEL");
            const bool items = sel ? strcmp(sel, "items") == 0 : true;
            if (items && !(sel && strcmp(sel, "cta") == 0)) return launch_solve_general_items(topo, cfg, ws, st);
        }
        if (UWBGO_GEN_CTA && ws.echi) {
#ifndef UWBGO_GCTA_WIDE
#define UWBGO_GCTA_WIDE 1 /* 8-warp CTAs for batches of at most one tile per SM */
#endif
            if (UWBGO_GCTA_WIDE && (int)tiles <= sms) {
                auto kern = lm_general_cta_kernel<8, 1>;
                if (gcta_dyn_smem(8) > 48 * 1024) {
                    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gcta_dyn_smem(8));
                    if (e != cudaSuccess) return e;
                }
                kern<<<tiles, 8 * 32, gcta_dyn_smem(8), st>>>(topo, cfg, ws);
            } else {
                auto kern = lm_general_cta_kernel<UWBGO_GCTA_WARPS, UWBGO_GCTA_MINB>;
                if (gcta_dyn_smem(UWBGO_GCTA_WARPS) > 48 * 1024) {
                    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)gcta_dyn_smem(UWBGO_GCTA_WARPS));
                    if (e != cudaSuccess) return e;
                }
                kern<<<tiles, UWBGO_GCTA_WARPS * 32, gcta_dyn_smem(UWBGO_GCTA_WARPS), st>>>(topo, cfg, ws);
            }
        }
        else
            lm_general_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, cfg, ws);
    }
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
    } else {
#ifndef UWBGO_LIN_GEN_POSE
#define UWBGO_LIN_GEN_POSE 1 /* 1: one warp per (tile, pose); 0: one thread per window */
#endif
        if (UWBGO_LIN_GEN_POSE)
            linearize_general_pose_kernel<<<(unsigned)(n_tiles(ws.W) * (topo.N + 1)), 32, 0, st>>>(topo, cfg, ws);
        else
            linearize_general_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, cfg, ws);
    }
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* Linearise stage, CHAIN windows, ONE kernel: tile-layout inputs -> public window-major blocks.  */
/* One warp owns (tile, run of LCF_RUN consecutive poses), lane = window, and walks its run from    */
/* the newest pose down with the straight-line chain_build of the LM kernels; the vertex-0 terms of */
/* the trajectory edge above the run are recomputed first (7 of the ~20 square roots of a pose).    */
/* Each pose's 18-number record goes through a padded shared-memory panel [19][33] (row 18 = zeros) */
/* and leaves as full 6x6 blocks, 16 bytes per lane: store A covers H_diag (lanes 0-17) and the     */
/* first 14 chunks of H_off, store B the rest of H_off and b.  DRAM traffic = the algorithmic bytes: */
/* inputs once, public arrays once, no H records in between.                                        */
/* What bounds it is the store stream, and what the store stream wants is measured in                */
/* scripts/micro/store_pattern.cu: 288-byte rows at a 14.4 KB stride reach the fill bandwidth        */
/* (7.2 TB/s) when the warps that are resident together cover neighbouring poses, 4.7 TB/s when      */
/* each warp walks 5 poses on its own (at any instant a window then has isolated rows 1,440 bytes    */
/* apart in flight).  Hence LCF_RUN = 1: consecutive warps take consecutive poses of a tile, at     */
/* the price of 27 instead of 20 square roots per pose (the kernel without its stores is 0.25 ms).  */
/* C3 on one B200: run 5 -> 0.67 ms, run 2 -> 0.55 ms, run 1 -> 0.51 ms.                            */
/* ------------------------------------------------------------------------------------------ */
#ifndef UWBGO_LCF_RUN
#define UWBGO_LCF_RUN 1
#endif
#ifndef UWBGO_LCF_WARPS
#define UWBGO_LCF_WARPS 1 /* warps per CTA; measured on C3: 1 -> 0.51 ms, 4 -> 0.54 ms, 8 -> 0.55 ms */
#endif
#ifndef UWBGO_LCF_MINB
#define UWBGO_LCF_MINB 16 /* 16 warps per SM at 128 registers; 102 / 85 registers spill: 1.05 / 1.20 ms */
#endif
constexpr int LCF_ROWS = HR_FAST + 1; /* panel rows: the record, then a row of zeros */

/* inputs of pose i for the run: chain_load plus the anchor of its range edge */
struct LcfIn {
    ChainIn in;
    double qx, qy, qz;
};
UWBGO_DI void lcf_load(const FastEnv &E, const double *__restrict__ T, int i, LcfIn &x)
{
    if (i > 0)
        chain_load<true>(E, T, i, x.in);
    else
        chain_load<false>(E, T, 0, x.in);
    const double *an = E.p.anch + (size_t)x.in.anchor * 3 * TILE;
    x.qx = ROW(an, 0); x.qy = ROW(an, 1); x.qz = ROW(an, 2);
}

/* which 16-byte chunk of a window's output this lane writes, per pose.  A window's blocks are
 * H_diag 36 doubles = 18 chunks, H_off 18 chunks, b 3 chunks = 39 chunks: store A covers chunks
 * 0..31 (lanes 0-17 H_diag, 18-31 H_off chunks 0-13), store B the rest (lanes 0-3 H_off chunks 14-17,
 * lanes 4-6 b).  r0 / r1 = panel rows of the chunk's two doubles (HR_FAST = the zero row). */
struct LcfLane {
    int a0, a1, b0, b1;
    int arr_a, arr_b; /* 0 H_diag, 1 H_off, 2 b, -1 idle */
    int off_a, off_b; /* double offset inside the block */
};
UWBGO_DI int lcf_row(int arr, int k)
{
    if (arr == 2) return k < 3 ? 15 + k : HR_FAST;
    const int r = k / 6, c = k % 6;
    if (r >= 3 || c >= 3) return HR_FAST;
    if (arr == 0) return r <= c ? up_idx(3, r, c) : up_idx(3, c, r);
    return 6 + 3 * r + c;
}
UWBGO_DI LcfLane lcf_lane(int lane)
{
    LcfLane L;
    L.arr_a = lane < 18 ? 0 : 1;
    L.off_a = lane < 18 ? 2 * lane : 2 * (lane - 18);
    L.arr_b = lane < 4 ? 1 : (lane < 7 ? 2 : -1);
    L.off_b = lane < 4 ? 2 * (14 + lane) : 2 * (lane - 4);
    L.a0 = lcf_row(L.arr_a, L.off_a);
    L.a1 = lcf_row(L.arr_a, L.off_a + 1);
    L.b0 = L.arr_b < 0 ? HR_FAST : lcf_row(L.arr_b, L.off_b);
    L.b1 = L.arr_b < 0 ? HR_FAST : lcf_row(L.arr_b, L.off_b + 1);
    return L;
}

/* per-edge chi2 | robustified chi2 of pose i's two edges into the scratch [tile][2E][32] (edge = data slot
 * in CHAIN windows: anchor edge of pose i at 2i - 1 (0 for pose 0), edge (i-1, i) at 2i) */
UWBGO_DI void lcf_put_chi(double *__restrict__ echi, int i, const double *chi4)
{
    const int sa = i == 0 ? 0 : 2 * i - 1;
    ROW(echi, 2 * sa) = chi4[0];
    ROW(echi, 2 * sa + 1) = chi4[1];
    if (i > 0) {
        ROW(echi, 4 * i) = chi4[2];
        ROW(echi, 4 * i + 1) = chi4[3];
    }
}

template <class M, bool CHI>
UWBGO_DI void lcf_run(const FastEnv &E, const double *__restrict__ T, double *__restrict__ echi, const int N, const int i_lo,
                      const int i_hi, const int64_t tile, const int lane, const int64_t W,
                      double (*panel)[33], double *__restrict__ H_diag, double *__restrict__ H_off,
                      double *__restrict__ b, unsigned &bad)
{
    LcfIn cur, nxt;
    lcf_load(E, T, i_hi, cur);
    const double *th = T + (size_t)i_hi * 3 * TILE;
    double cx = ROW(th, 0), cy = ROW(th, 1), cz = ROW(th, 2);
    double carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (i_hi + 1 < N) { /* vertex-0 terms of edge (i_hi, i_hi + 1) */
        const int j = i_hi + 1;
        const double *tp1 = T + (size_t)j * 3 * TILE;
        const double nx = ROW(tp1, 0), ny = ROW(tp1, 1), nz = ROW(tp1, 2);
        const double dt = ROW(E.p.rd, 2 * j), it = ROW(E.p.ri, 2 * j);
        const int rb = __ldg(&E.tp->chain[j].robust);
        if (i_hi > i_lo) lcf_load(E, T, i_hi - 1, nxt);
        const double err = dt - dist3m<M>(cx, cy, cz, nx, ny, nz, bad);
        fast_jac_v0<M>(cx, cy, cz, nx, ny, nz, dt, E.delta, E.scalar, carry, &bad);
        chain_weights<M>(E, err, it, (rb & 2) != 0, carry[3], carry[4], &bad);
    } else if (i_hi > i_lo)
        lcf_load(E, T, i_hi - 1, nxt);
    const LcfLane L = lcf_lane(lane);
    const int nw = (int)((W - tile * TILE) < TILE ? (W - tile * TILE) : TILE);
    for (int i = i_hi; i >= i_lo; --i) {
        double h[HR_FAST], chi4[4];
#ifdef UWBGO_LCF_NOLOAD
        for (int k = 0; k < HR_FAST; ++k) h[k] = (double)(i + k);
#elif defined(UWBGO_LCF_NOCOMPUTE)
        for (int k = 0; k < HR_FAST; ++k) h[k] = cx + k;
#else
        if (i > 0)
            chain_build_q<true, M, CHI>(E, cx, cy, cz, cur.qx, cur.qy, cur.qz, cur.in, carry, h, &bad, chi4);
        else
            chain_build_q<false, M, CHI>(E, cx, cy, cz, cur.qx, cur.qy, cur.qz, cur.in, carry, h, &bad, chi4);
        if (CHI) lcf_put_chi(echi, i, chi4);
#endif
        if (__any_sync(0xffffffffu, bad != 0)) return; /* the caller repeats the run with IEEE operations */
        __syncwarp();
#pragma unroll
        for (int k = 0; k < HR_FAST; ++k) panel[k][lane] = h[k];
        __syncwarp();
        /* next pose: its estimate came with this pose's inputs, the rest was fetched during the build */
        cx = cur.in.px; cy = cur.in.py; cz = cur.in.pz;
        cur = nxt;
        if (i - 2 >= i_lo) lcf_load(E, T, i - 2, nxt);
        const size_t wb = (size_t)tile * TILE;
        double *pa = L.arr_a == 0 ? H_diag + (wb * N + i) * 36 + L.off_a
                                  : (i > 0 ? H_off + (wb * (N - 1) + (i - 1)) * 36 + L.off_a : nullptr);
        double *pb = L.arr_b == 1 ? (i > 0 ? H_off + (wb * (N - 1) + (i - 1)) * 36 + L.off_b : nullptr)
                                  : (L.arr_b == 2 ? b + (wb * N + i) * 6 + L.off_b : nullptr);
        const size_t sa = L.arr_a == 0 ? (size_t)N * 36 : (size_t)(N - 1) * 36;
        const size_t sb = L.arr_b == 1 ? (size_t)(N - 1) * 36 : (size_t)N * 6;
#ifdef UWBGO_LCF_NOSTORE
        if (h[0] == 123.456)
#endif
#ifndef UWBGO_LCF_UNROLL
#define UWBGO_LCF_UNROLL 8
#endif
        UWBGO_PRAGMA(unroll UWBGO_LCF_UNROLL)
        for (int wl = 0; wl < nw; ++wl) {
            if (pa) {
                *reinterpret_cast<double2 *>(pa) = make_double2(panel[L.a0][wl], panel[L.a1][wl]);
                pa += sa;
            }
            if (pb) {
                *reinterpret_cast<double2 *>(pb) = make_double2(panel[L.b0][wl], panel[L.b1][wl]);
                pb += sb;
            }
        }
    }
}

/* CHI_TERMS: the run warps also write the per-edge chi2 terms (summed by chi_sum_kernel); otherwise, when
 * want_chi, the leading CTAs walk every tile's edges (fast_chi_pass) */
template <bool CHI_TERMS>
__global__ void __launch_bounds__(UWBGO_LCF_WARPS * 32, UWBGO_LCF_MINB)
linearize_chain_fused_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                             const __grid_constant__ DevWs ws, double *__restrict__ H_diag,
                             double *__restrict__ H_off, double *__restrict__ b, int runs, int want_chi,
                             int chi_every)
{
    __shared__ double sm[UWBGO_LCF_WARPS][LCF_ROWS][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sm[warp][HR_FAST][lane] = 0.0;
    /* the chi2 passes (one warp per tile) are long and get CTAs of their own: CTA b is a chi2 CTA iff
     * b % chi_every == 0 and b / chi_every < chi_ctas (chi_every = 1: they lead the grid) */
    const int64_t tiles = n_tiles(ws.W);
    const int64_t chi_ctas = want_chi ? (tiles + UWBGO_LCF_WARPS - 1) / UWBGO_LCF_WARPS : 0;
    const int64_t bid = blockIdx.x;
    int64_t tile;
    int run;
    if (chi_ctas > 0 && bid % chi_every == 0 && bid / chi_every < chi_ctas) {
        tile = (bid / chi_every) * UWBGO_LCF_WARPS + warp;
        run = -1;
        if (tile >= tiles) return;
    } else {
        int64_t before = 0; /* chi2 CTAs with a smaller block index */
        if (chi_ctas > 0) {
            before = bid / chi_every + 1;
            if (before > chi_ctas) before = chi_ctas;
        }
        const int64_t item = (bid - before) * UWBGO_LCF_WARPS + warp;
        if (item >= tiles * runs) return;
        tile = item / runs;
        run = (int)(item % runs);
    }
    const int N = tp.N;
    FastEnv E;
    E.tp = &tp;
    E.p = thread_ptrs<HR_FAST, LR_FAST>(tp, ws, tile * TILE + lane); /* workspaces are whole tiles */
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
    E.bs = 0;
    E.stash = nullptr;
    E.anch = E.p.anch;
    E.anch_stride = TILE;
    if (run < 0) { /* computeActiveErrors + chi2: lane = window, the sums run in insertion order */
        if (tile * TILE + lane < ws.W) {
            double p, r;
            fast_chi_pass(E, E.p.T0, p, r);
            double *c = ws.chi2 + tile * 2 * TILE + lane;
            ROW(c, 0) = p;
            ROW(c, 1) = r;
        }
        return;
    }
    const int i_lo = run * UWBGO_LCF_RUN;
    const int i_hi = (i_lo + UWBGO_LCF_RUN < N ? i_lo + UWBGO_LCF_RUN : N) - 1;
    unsigned bad = 0;
    double *echi = CHI_TERMS ? ws.echi + (tile * (size_t)tp.E * 2) * TILE + lane : nullptr;
    lcf_run<NbMath, CHI_TERMS>(E, E.p.T0, echi, N, i_lo, i_hi, tile, lane, ws.W, sm[warp], H_diag, H_off, b, bad);
    if (__any_sync(0xffffffffu, bad != 0)) {
        bad = 0;
        lcf_run<IeeeMath, CHI_TERMS>(E, E.p.T0, echi, N, i_lo, i_hi, tile, lane, ws.W, sm[warp], H_diag, H_off, b, bad);
    }
}

/* activeChi2 / activeRobustChi2 of the stage: the per-edge terms summed in insertion order */
__global__ void __launch_bounds__(CTA_THREADS)
chi_sum_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevWs ws)
{
    const int64_t w = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    if (w >= ws.W) return;
    const int64_t tile = w / TILE;
    const int lane = (int)(w % TILE);
    const double *e = ws.echi + (tile * (size_t)tp.E * 2) * TILE + lane;
    double p = 0.0, r = 0.0;
    int k = 0;
    for (; k + 8 <= tp.E; k += 8) {
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ROW(e, 2 * k + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p = p + v[2 * j];
            r = r + v[2 * j + 1];
        }
    }
    for (; k < tp.E; ++k) {
        p = p + ROW(e, 2 * k);
        r = r + ROW(e, 2 * k + 1);
    }
    double *c = ws.chi2 + tile * 2 * TILE + lane;
    ROW(c, 0) = p;
    ROW(c, 1) = r;
}

/* the output arrays must be 16-byte aligned (the blocks leave as 16-byte stores) */
bool linearize_chain_fused_ok(const DevTopo &topo, const double *H_diag, const double *H_off, const double *b)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(H_diag) | reinterpret_cast<uintptr_t>(H_off) |
                        reinterpret_cast<uintptr_t>(b);
    return topo.fast == 2 && (a & 15u) == 0;
}

cudaError_t launch_linearize_chain_fused(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws,
                                         double *H_diag, double *H_off, double *b, bool want_chi,
                                         cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    const int runs = (topo.N + UWBGO_LCF_RUN - 1) / UWBGO_LCF_RUN;
    const int64_t tiles = n_tiles(ws.W);
    const int64_t run_ctas = (tiles * runs + UWBGO_LCF_WARPS - 1) / UWBGO_LCF_WARPS;
#ifndef UWBGO_LCF_CHI_TERMS
#define UWBGO_LCF_CHI_TERMS 1 /* 1: per-edge chi2 terms from the run warps + chi_sum_kernel; 0: chi2 CTAs lead the grid */
#endif
    const bool terms = want_chi && UWBGO_LCF_CHI_TERMS && ws.echi;
    const int64_t chi_ctas = (want_chi && !terms) ? (tiles + UWBGO_LCF_WARPS - 1) / UWBGO_LCF_WARPS : 0;
    const int64_t ctas = run_ctas + chi_ctas;
    if (terms) {
        linearize_chain_fused_kernel<true><<<(unsigned)ctas, UWBGO_LCF_WARPS * 32, 0, st>>>(topo, cfg, ws, H_diag, H_off, b,
                                                                                            runs, 0, 1);
        chi_sum_kernel<<<window_blocks(ws.W), CTA_THREADS, 0, st>>>(topo, ws);
    } else
        linearize_chain_fused_kernel<false><<<(unsigned)ctas, UWBGO_LCF_WARPS * 32, 0, st>>>(topo, cfg, ws, H_diag, H_off, b,
                                                                                             runs, want_chi ? 1 : 0, 1);
    return cudaGetLastError();
}

/* H records (tile layout) -> public H_diag [W][N][36] (both triangles), H_off [W][N-1][36],
 * b [W][N][6].  One warp per (tile, pose): the record is expanded into a padded shared-memory
 * panel [78][33] with coalesced row reads, then written out window by window as contiguous
 * 288-byte / 48-byte runs. */
constexpr int EXPAND_WARPS = 2;
__global__ void __launch_bounds__(EXPAND_WARPS * 32)
expand_H_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevWs ws,
                double *__restrict__ H_diag, double *__restrict__ H_off, double *__restrict__ b)
{
    __shared__ double sm[EXPAND_WARPS][78][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = tp.N;
    const int64_t item = (int64_t)blockIdx.x * EXPAND_WARPS + warp;
    if (item >= n_tiles(ws.W) * N) return;
    const int64_t tile = item / N;
    const int i = (int)(item % N);
    double(*p)[33] = sm[warp];
    if (tp.fast) {
        const double *h = ws.HB + ((tile * N + i) * (size_t)HR_FAST) * TILE + lane;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const bool in = r < 3 && c < 3;
                p[6 * r + c][lane] = in ? (r <= c ? ROW(h, up_idx(3, r, c)) : ROW(h, up_idx(3, c, r))) : 0.0;
                p[36 + 6 * r + c][lane] = in ? ROW(h, 6 + 3 * r + c) : 0.0;
            }
#pragma unroll
        for (int r = 0; r < 6; ++r) p[72 + r][lane] = r < 3 ? ROW(h, 15 + r) : 0.0;
    } else {
        const double *h = ws.HB + ((tile * N + i) * (size_t)HR_GEN) * TILE + lane;
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                p[6 * r + c][lane] = r <= c ? ROW(h, up_idx(6, r, c)) : ROW(h, up_idx(6, c, r));
                p[36 + 6 * r + c][lane] = ROW(h, 21 + 6 * r + c);
            }
#pragma unroll
        for (int r = 0; r < 6; ++r) p[72 + r][lane] = ROW(h, 57 + r);
    }
    __syncwarp();
    for (int wl = 0; wl < TILE; ++wl) {
        const int64_t w = tile * TILE + wl;
        if (w >= ws.W) break;
        double *hd = H_diag + ((size_t)w * N + i) * 36;
        hd[lane] = p[lane][wl];
        if (lane < 4) hd[32 + lane] = p[32 + lane][wl];
        if (i > 0) {
            double *ho = H_off + ((size_t)w * (N - 1) + (i - 1)) * 36;
            ho[lane] = p[36 + lane][wl];
            if (lane < 4) ho[32 + lane] = p[68 + lane][wl];
        }
        if (lane < 6) b[((size_t)w * N + i) * 6 + lane] = p[72 + lane][wl];
    }
}

cudaError_t launch_expand_H(const DevTopo &topo, const DevWs &ws, double *H_diag, double *H_off,
                            double *b, cudaStream_t st)
{
    if (ws.W <= 0) return cudaSuccess;
    const int64_t items = n_tiles(ws.W) * topo.N;
    expand_H_kernel<<<(unsigned)((items + EXPAND_WARPS - 1) / EXPAND_WARPS), EXPAND_WARPS * 32, 0, st>>>(
        topo, ws, H_diag, H_off, b);
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

/* ------------------------------------------------------------------------------------------ */
/* Arithmetic self-test: NbMath against the IEEE operations on pseudo-random operands.           */
/* counts[0] operands tried, [1] results that differ in bits while NbMath did not raise its flag  */
/* (must be 0), [2] operands flagged (fall back to IEEE in the solver).                          */
/* mode 0: operands spread over the whole binary64 encoding space (every exponent, both signs,     */
/* NaN / inf / denormals); mode 1: magnitudes the solver sees (2^-60 .. 2^60).                    */
/* ------------------------------------------------------------------------------------------ */
UWBGO_DI unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
UWBGO_DI double selftest_operand(unsigned long long r, int mode)
{
    if (mode == 0) return __longlong_as_double((long long)r);
    unsigned long long mant = r & 0x000fffffffffffffULL;
    /* one operand in eight ends in a run of ones: significands 1.11..1 and their neighbours are where a
     * Newton reciprocal can miss the correctly rounded result by an ulp */
    if (((r >> 59) & 7ULL) == 0) mant |= 0x000fffffffffffffULL >> ((r >> 52) & 3ULL), mant &= ~((r >> 54) & 1ULL);
    const unsigned long long e = 1023ULL - 60ULL + ((r >> 52) % 121ULL);
    const unsigned long long sign = (r >> 63) << 63;
    return __longlong_as_double((long long)(sign | (e << 52) | mant));
}
UWBGO_DI bool same_bits(double a, double b)
{
    if (a != a && b != b) return true; /* NaN payloads are not part of the contract */
    return __double_as_longlong(a) == __double_as_longlong(b);
}
__global__ void __launch_bounds__(256) math_selftest_kernel(unsigned long long seed, int per_thread, int mode,
                                                           unsigned long long *counts)
{
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long tried = 0, wrong = 0, flagged = 0;
    for (int k = 0; k < per_thread; ++k) {
        const unsigned long long r0 = mix64(seed ^ (tid * 0x100000001b3ULL + (unsigned long long)k * 2ULL));
        const unsigned long long r1 = mix64(r0 ^ 0x5851f42d4c957f2dULL);
        const double a = selftest_operand(r0, mode), b = selftest_operand(r1, mode);
        const double xs = mode == 1 ? fabs(a) : a;
        unsigned bad;
        bad = 0;
        const double s = NbMath::sqrt_(xs, bad);
        if (bad) ++flagged; else if (!same_bits(s, sqrt(xs))) ++wrong;
        bad = 0;
        const double rc = NbMath::rcp(b, bad);
        if (bad) ++flagged; else if (!same_bits(rc, 1.0 / b)) ++wrong;
        bad = 0;
        const double q = NbMath::div(a, b, bad);
        if (bad) ++flagged; else if (!same_bits(q, a / b)) ++wrong;
        bad = 0;
        const double lg = NbMath::log_(xs, bad);
        if (bad) ++flagged; else if (!same_bits(lg, det_log(xs))) ++wrong;
        bad = 0;
        const double rp = NbMath::rsqrt_pivot(xs, bad); /* the Cholesky pivot of the WINDOW path */
        if (bad) ++flagged; else if (!same_bits(rp, 1.0 / sqrt(xs))) ++wrong;
        tried += 5;
    }
    atomicAdd(counts + 0, tried);
    atomicAdd(counts + 1, wrong);
    atomicAdd(counts + 2, flagged);
}

cudaError_t launch_math_selftest(unsigned long long seed, int blocks, int per_thread, int mode,
                                 unsigned long long *counts, cudaStream_t st)
{
    math_selftest_kernel<<<blocks, 256, 0, st>>>(seed, per_thread, mode, counts);
    return cudaGetLastError();
}

}  // namespace uwbgo
