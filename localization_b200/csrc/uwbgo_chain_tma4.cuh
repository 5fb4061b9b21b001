/*
 * uwbgo_chain_tma4.cuh — CHAIN path of the fused LM kernel with THREE or FOUR specialised warps per tile.
 *
 * lm_chain_tma_kernel (uwbgo_chain_tma.cuh) gives a tile of 32 windows two warps: the edge warp linearises
 * the trajectory edge of every pose, the chain warp linearises the anchor edge, assembles H, eliminates and
 * substitutes.  Its chain warp is the longer one (the edge warp waits out about a third of its time at the
 * step barrier), and a tile alone takes 2.1 ms: batches below one wave (the 8-GPU strong-scaling case,
 * 8,192 windows per GPU) run at the latency of one tile.  Here the step of a tile is cut finer:
 *   warp 0  P   trajectory edge (i-1, i): the Jacobian of vertex 0 (NW = 4) or of both vertices (NW = 3), the
 *               weights; in the substitution sweep the residuals of the trajectory edges and the two ORDERED
 *               chi2 sums
 *   warp 1  C   assembles the H record of pose i from what the other warps hand over, runs the elimination
 *               step and the substitution, owns the LM state.  Reads no estimate at all in the factor sweep
 *   warp 2  Q   anchor edge of pose i: Jacobian and weights (factor sweep), residual and rho0 (substitution)
 *   warp 3  PB  (NW = 4) the Jacobian of vertex 1 of the trajectory edge
 * P, Q and PB run one step ahead of C, as the edge warp does in the two-warp kernel; one named barrier per
 * step over the whole CTA.  Operand staging by the bulk-copy engine, the L records' bulk stores, the
 * branch-free arithmetic with the IEEE re-run and the LM bookkeeping are those of lm_chain_tma_kernel.
 * Every per-window operation and its order are unchanged: same bits.
 */
#ifndef UWBGO_CHAIN_TMA4_CUH
#define UWBGO_CHAIN_TMA4_CUH

#include "uwbgo_chain_tma.cuh"
#ifdef UWBGO_TMA4_TIMING
#include <cstdio>
#endif

namespace uwbgo {

struct alignas(128) Tma4Shared {
    union {
        struct {
            double T[4][2][3][TILE];      /* both estimate buffers of pose j in slot j & 3: read by P, a step later by Q / PB */
            double D[2][2][TILE];         /* [0] range of the anchor edge of pose i (Q), [1] of edge (i-1, i) (P) */
            double I[2][2][TILE];         /* their information                                      */
            double A[2][3][TILE];         /* Q: anchor of pose i                                    */
            double L[2][LR_FAST][TILE];   /* C: L record on its way out                             */
        } f;
        struct {
            double L[SUBST_RING][LR_FAST][TILE]; /* C: L record of pose k                           */
            double T[SUBST_RING][2][3][TILE];    /* C: both estimate buffers of pose k              */
            double D[SUBST_RING][2][TILE];       /* ranges of the anchor edge (Q) and of edge (k-2, k-1) (P) */
            double I[SUBST_RING][2][TILE];       /* their information                               */
            double A[SUBST_RING][3][TILE];       /* Q: anchor of pose k-1                           */
        } s;
    } in;
    double traj[2][8][TILE]; /* P (, PB) -> C: A(3), B(3), Ow, omega_r of edge (i-1, i), double buffered */
    double anc[2][5][TILE];  /* Q -> C: J(3), Ow, omega_r of the anchor edge of pose i                   */
    double tnew[2][3][TILE]; /* C -> P, Q: new estimate of pose i                                        */
    double qchi[2][2][TILE]; /* Q -> P: chi2 and its rho0 of the anchor edge of pose i                   */
    double chi[2][TILE];     /* P -> C: plain and robust chi2 of the trial                               */
    int act[TILE];
    int cur[TILE];
    unsigned bad[4];
    unsigned long long fbar[2];
    unsigned long long sbar[SUBST_RING];
};
static_assert(sizeof(((Tma4Shared *)nullptr)->in) >= TMA_STASH_DOUBLES * sizeof(double), "stash must fit in the staging area");

/* inputs of factor step k (k = 0 .. N): P, Q and PB work on pose i = N-1-k and its edges, C on pose N-k from
 * what they handed over in step k-1.  Rows 2 i - 1 (anchor edge of pose i) and 2 i (edge (i-1, i)) of the
 * range arrays are adjacent.  Step N stages nothing anybody reads (one row, so that the step has a
 * transaction to wait for like every other). */
UWBGO_DI void issue_factor_step4(const DevTopo &tp, const TileBase &g, Tma4Shared &sh, int N, int k, unsigned ev)
{
    const int b = (int)(ev & 1u);
    unsigned long long *bar = &sh.fbar[b];
    const int i = N - 1 - k, jT = N - 2 - k;
    const bool first = k == 0, last = k == N;
    const unsigned drows = last ? 1u : (i == 0 ? 1u : 2u);
    const unsigned bytes = ((jT >= 0 ? 6u : 0u) + (first ? 6u : 0u) + 2u * drows + (last ? 0u : 3u)) * ROW_BYTES;
    mbar_expect_tx(bar, bytes);
    if (first) {
        BULK_STREAM(sh.in.f.T[(N - 1) & 3][0], g.T0 + (size_t)(N - 1) * 3 * TILE, 3 * ROW_BYTES, bar);
        BULK_STREAM(sh.in.f.T[(N - 1) & 3][1], g.T1 + (size_t)(N - 1) * 3 * TILE, 3 * ROW_BYTES, bar);
    }
    if (jT >= 0) {
        BULK_STREAM(sh.in.f.T[jT & 3][0], g.T0 + (size_t)jT * 3 * TILE, 3 * ROW_BYTES, bar);
        BULK_STREAM(sh.in.f.T[jT & 3][1], g.T1 + (size_t)jT * 3 * TILE, 3 * ROW_BYTES, bar);
    }
    {
        /* pose 0 (and the idle last step): row 0 -> [0]; otherwise rows (2 i - 1, 2 i) -> [0], [1] */
        const int r0 = (last || i == 0) ? 0 : 2 * i - 1;
        BULK_STREAM(sh.in.f.D[b][0], g.rd + (size_t)r0 * TILE, drows * ROW_BYTES, bar);
        BULK_STREAM(sh.in.f.I[b][0], g.ri + (size_t)r0 * TILE, drows * ROW_BYTES, bar);
    }
    if (!last) {
        const int anchor = __ldg(&tp.chain[i].anchor);
        bulk_load(sh.in.f.A[b], g.anch + (size_t)anchor * 3 * TILE, 3 * ROW_BYTES, bar);
    }
}

/* inputs of substitution step k (k = 0 .. N) into ring slot `slot`: C works on pose k, P and Q on pose k-1 */
UWBGO_DI void issue_subst_step4(const DevTopo &tp, const TileBase &g, Tma4Shared &sh, int N, int k, int slot)
{
    unsigned long long *bar = &sh.sbar[slot];
    const bool doC = k < N, doP = k >= 1;
    const int j = k - 1;
    const unsigned prow = j == 0 ? 1u : 2u;
    const unsigned bytes = (doC ? (unsigned)(LR_FAST + 6) : 0u) * ROW_BYTES + (doP ? (2u * prow + 3u) : 0u) * ROW_BYTES;
    mbar_expect_tx(bar, bytes);
    if (doC) {
        bulk_load(sh.in.s.L[slot], g.LR + (size_t)k * LR_FAST * TILE, LR_FAST * ROW_BYTES, bar);
        BULK_STREAM(sh.in.s.T[slot][0], g.T0 + (size_t)k * 3 * TILE, 3 * ROW_BYTES, bar);
        BULK_STREAM(sh.in.s.T[slot][1], g.T1 + (size_t)k * 3 * TILE, 3 * ROW_BYTES, bar);
    }
    if (doP) {
        const int anchor = __ldg(&tp.chain[j].anchor);
        const int r0 = j == 0 ? 0 : 2 * j - 1;
        BULK_STREAM(sh.in.s.D[slot], g.rd + (size_t)r0 * TILE, prow * ROW_BYTES, bar);
        BULK_STREAM(sh.in.s.I[slot], g.ri + (size_t)r0 * TILE, prow * ROW_BYTES, bar);
        bulk_load(sh.in.s.A[slot], g.anch + (size_t)anchor * 3 * TILE, 3 * ROW_BYTES, bar);
    }
}

template <int NW>
UWBGO_DI void tma4_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
}

#ifndef UWBGO_TMA4_MINB
#define UWBGO_TMA4_MINB 4 /* four warps: 4 CTAs x 128 threads x 128 registers fill the register file */
#endif
#ifndef UWBGO_TMA3_MINB
#define UWBGO_TMA3_MINB 5 /* three warps: 5 CTAs x 96 threads x 136 registers */
#endif
template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? UWBGO_TMA4_MINB : UWBGO_TMA3_MINB)
lm_chain_tma4_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                     const __grid_constant__ DevWs ws)
{
    static_assert(NW == 3 || NW == 4, "three or four warps per tile");
    __shared__ Tma4Shared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * TILE + lane; /* tail lanes own zero-filled pad columns */
    const bool valid = w < ws.W;
    const int N = tp.N;
    const bool issuerF = warp == 2 && lane == 0; /* factor sweep: the anchor warp is the lightest  */
    const bool issuerS = warp == 1 && lane == 0; /* substitution sweep: the chain warp, which owns the L stores */
    FastEnv E;
    E.tp = &tp;
    E.p = thread_ptrs<HR_FAST, LR_FAST>(tp, ws, w);
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
    E.bs = TILE;
    E.stash = reinterpret_cast<double *>(&sh.in) + lane;
    E.anch = E.p.anch;
    E.anch_stride = TILE;
    TileBase g;
    {
        const size_t tile = blockIdx.x;
        g.T0 = ws.T[0] + tile * (size_t)N * 3 * TILE;
        g.T1 = ws.T[1] + tile * (size_t)N * 3 * TILE;
        g.rd = ws.rd + tile * (size_t)tp.Er * TILE;
        g.ri = ws.ri + tile * (size_t)tp.Er * TILE;
        g.anch = ws.anch + tile * (size_t)tp.A * 3 * TILE;
        g.LR = ws.LR + tile * (size_t)N * LR_FAST * TILE;
    }

    if (threadIdx.x == 0) {
        mbar_init(&sh.fbar[0], 1);
        mbar_init(&sh.fbar[1], 1);
#pragma unroll
        for (int m = 0; m < SUBST_RING; ++m) mbar_init(&sh.sbar[m], 1);
        mbar_fence_init();
    }
    /* LM state, chain warp only */
    double lambda = 0.0, ni = 2.0, stale = 0.0, plainCur = 0.0, currentChi = 0.0, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, q = 0, it = 0, cur = 0, last_rej = 0;
    bool done = !valid || cfg.max_iterations <= 0;
    double maxdiag = 0.0;
    if (warp == 1) {
        sh.act[lane] = valid ? 1 : 0;
        sh.cur[lane] = 0;
    }
    __syncthreads();
#ifdef UWBGO_TMA4_TIMING
    long long tw[6] = {0, 0, 0, 0, 0, 0}; /* factor: wait for operands, work, barrier; substitution: the same */
    long long tq_ = clock64();
#define T4_TICK(k) do { long long tn_ = clock64(); tw[k] += tn_ - tq_; tq_ = tn_; } while (0)
#else
#define T4_TICK(k)
#endif
    bool init = true; /* the first pass is the initial evaluation (x forced to 0), as in lm_chain_tma_kernel */
    unsigned ev = 0;
    int sslot = 0;
    unsigned spar = 0;
    for (;;) {
        const bool act = sh.act[lane] != 0;
        const int c = sh.cur[lane];
        if (__ballot_sync(0xffffffffu, act) == 0u) break;
        tma4_barrier<NW>(); /* everybody has read act / cur before the chain warp may overwrite them */
        double *const Tn = E.p.T(c ^ 1);
        bool ok = true;
        double scale = 0.0;
        auto trial = [&](auto math_tag) {
            using M = decltype(math_tag);
            unsigned bad = 0;
            ok = true;
            scale = 0.0;
            maxdiag = 0.0;
            const double lam = init ? 1.0 : lambda;
            /* ================= factor sweep: steps k = 0 .. N ================= */
            if (issuerF) issue_factor_step4(tp, g, sh, N, 0, ev);
            {
                double cx = 0.0, cy = 0.0, cz = 0.0; /* P: pose i of edge (i-1, i), carried from the step before */
                double G[9], zn[3], carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int m = 0; m < 9; ++m) G[m] = 0.0;
#pragma unroll
                for (int m = 0; m < 3; ++m) zn[m] = 0.0;
                for (int k = 0; k <= N; ++k) {
                    const int b = (int)((ev + (unsigned)k) & 1u);
                    if (issuerF && k < N) issue_factor_step4(tp, g, sh, N, k + 1, ev + (unsigned)k + 1u);
                    if (issuerS && k >= 2) bulk_wait_read<1>(); /* the store of step k-2 has left its L buffer */
                    T4_TICK(5);
                    mbar_wait(&sh.fbar[b], ((ev + (unsigned)k) >> 1) & 1u);
                    T4_TICK(0);
                    const int i = N - 1 - k; /* the pose P, Q and PB work on */
                    if (warp == 0) {
                        if (k == 0) {
                            const double(*tc)[TILE] = sh.in.f.T[(N - 1) & 3][c];
                            cx = tc[0][lane]; cy = tc[1][lane]; cz = tc[2][lane];
                        }
                        if (i >= 1) {
                            const double(*tp_)[TILE] = sh.in.f.T[(i - 1) & 3][c];
                            const double px = tp_[0][lane], py = tp_[1][lane], pz = tp_[2][lane];
                            const double dt = sh.in.f.D[b][1][lane], it_ = sh.in.f.I[b][1][lane];
                            const int rob = __ldg(&tp.chain[i].robust);
                            double A[3], B[3], Ow, omega_r, err;
                            if (NW == 4)
                                range_linearize<M, true, false>(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, err, A, nullptr, bad);
                            else
                                range_linearize<M, true, true>(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, err, A, B, bad);
                            chain_weights<M>(E, err, it_, (rob & 2) != 0, Ow, omega_r, &bad);
                            double(*o)[TILE] = sh.traj[k & 1];
                            o[0][lane] = A[0]; o[1][lane] = A[1]; o[2][lane] = A[2];
                            if (NW == 3) {
                                o[3][lane] = B[0]; o[4][lane] = B[1]; o[5][lane] = B[2];
                            }
                            o[6][lane] = Ow;   o[7][lane] = omega_r;
                            cx = px; cy = py; cz = pz;
                        }
                    } else if (warp == 3) {
                        if (i >= 1) { /* PB: the Jacobian of vertex 1 (pose i) of edge (i-1, i) */
                            const double(*tp_)[TILE] = sh.in.f.T[(i - 1) & 3][c];
                            const double(*tc)[TILE] = sh.in.f.T[i & 3][c];
                            const double dt = sh.in.f.D[b][1][lane];
                            double B[3], err;
                            range_linearize<M, false, true>(tp_[0][lane], tp_[1][lane], tp_[2][lane], tc[0][lane], tc[1][lane],
                                                            tc[2][lane], dt, E.delta, E.scalar, err, nullptr, B, bad);
                            double(*o)[TILE] = sh.traj[k & 1];
                            o[3][lane] = B[0]; o[4][lane] = B[1]; o[5][lane] = B[2];
                        }
                    } else if (warp == 2) {
                        if (i >= 0) { /* Q: the anchor edge of pose i */
                            const double(*tc)[TILE] = sh.in.f.T[i & 3][c];
                            const double px = tc[0][lane], py = tc[1][lane], pz = tc[2][lane];
                            const double qx = sh.in.f.A[b][0][lane], qy = sh.in.f.A[b][1][lane], qz = sh.in.f.A[b][2][lane];
                            const double da = sh.in.f.D[b][0][lane], ia = sh.in.f.I[b][0][lane];
                            const int rob = __ldg(&tp.chain[i].robust);
                            double J[3], Ow, omega_r, err;
                            range_linearize<M, true, false>(px, py, pz, qx, qy, qz, da, E.delta, E.scalar, err, J, nullptr, bad);
                            chain_weights<M>(E, err, ia, (rob & 1) != 0, Ow, omega_r, &bad);
                            double(*o)[TILE] = sh.anc[k & 1];
                            o[0][lane] = J[0]; o[1][lane] = J[1]; o[2][lane] = J[2];
                            o[3][lane] = Ow;   o[4][lane] = omega_r;
                        }
                    } else if (k >= 1) {
                        const int ic = N - k; /* C: pose ic, linearised by the other warps in step k-1 */
                        __syncwarp();         /* lane 0 has waited for the L buffer */
                        double h[HR_FAST];
#pragma unroll
                        for (int m = 0; m < HR_FAST; ++m) h[m] = 0.0;
                        {
                            const double(*a)[TILE] = sh.anc[(k - 1) & 1];
                            const double J[3] = {a[0][lane], a[1][lane], a[2][lane]};
                            chain_acc(J, a[3][lane], a[4][lane], h);
                        }
                        double nA[3] = {0.0, 0.0, 0.0}, nOw = 0.0, nOr = 0.0;
                        if (ic >= 1) { /* edge (ic-1, ic) */
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
                        if (init) {
                            double v;
                            v = fabs(h[0]); if (v > maxdiag) maxdiag = v;
                            v = fabs(h[3]); if (v > maxdiag) maxdiag = v;
                            v = fabs(h[5]); if (v > maxdiag) maxdiag = v;
                        }
                        double *l = &sh.in.f.L[k & 1][0][lane];
                        factor_step<3, M>(h, l, true, ic > 0, lam, G, zn, ok, &bad);
                        chain_store_b(l, h);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            bulk_store(g.LR + (size_t)ic * LR_FAST * TILE, sh.in.f.L[k & 1], LR_FAST * ROW_BYTES);
                            bulk_commit();
                        }
                    }
                    T4_TICK(1);
                    tma4_barrier<NW>();
                    T4_TICK(2);
                }
                ok = !init && ok && (lambda > 0.0);
            }
            ev += (unsigned)N + 1u;
            /* ================= substitution sweep: steps k = 0 .. N, fetched two steps ahead ================= */
            int islot = sslot;
            if (issuerS) {
                bulk_wait<0>(); /* every L record has landed before the first one is fetched back */
                issue_subst_step4(tp, g, sh, N, 0, islot);
                islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                if (N >= 1) {
                    issue_subst_step4(tp, g, sh, N, 1, islot);
                    islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                }
            }
            {
                double xp[3] = {0.0, 0.0, 0.0};         /* C: x of the previous pose                          */
                double p = 0.0, r = 0.0;                /* P: plain / robust chi2, summed in insertion order  */
                double vx = 0.0, vy = 0.0, vz = 0.0;    /* P: new estimate of pose k-2                        */
                double tchi = 0.0, trob = 0.0;          /* P: terms of edge (k-3, k-2), added a step later    */
                for (int k = 0; k <= N; ++k) {
                    const int b = sslot;
                    if (issuerS && k + 2 <= N) {
                        issue_subst_step4(tp, g, sh, N, k + 2, islot);
                        islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                    }
                    T4_TICK(5);
                    mbar_wait(&sh.sbar[b], (spar >> b) & 1u);
                    T4_TICK(3);
                    spar ^= 1u << b;
                    sslot = sslot == SUBST_RING - 1 ? 0 : sslot + 1;
                    if (warp == 1) {
                        if (k < N) {
                            const int i = k;
                            double l[LR_FAST];
#pragma unroll
                            for (int m = 0; m < LR_FAST; ++m) l[m] = sh.in.s.L[b][m][lane];
                            const double t0 = sh.in.s.T[b][c][0][lane], t1 = sh.in.s.T[b][c][1][lane], t2 = sh.in.s.T[b][c][2][lane];
                            subst_step<3>(l, i > 0, xp);
                            if (!ok) xp[0] = xp[1] = xp[2] = 0.0;
#pragma unroll
                            for (int m = 0; m < 3; ++m) scale = scale + xp[m] * (lam * xp[m] + l[12 + m]);
                            const double nx = xp[0] + t0, ny = xp[1] + t1, nz = xp[2] + t2;
                            double(*o)[TILE] = sh.tnew[i & 1];
                            o[0][lane] = nx; o[1][lane] = ny; o[2][lane] = nz;
                            if (act) {
                                double *to = Tn + (size_t)i * 3 * TILE;
                                ROW(to, 0) = nx; ROW(to, 1) = ny; ROW(to, 2) = nz;
                            }
                        }
                    } else if (warp == 2) {
                        if (k >= 1) { /* Q: anchor edge of pose i = k-1, produced in the previous step */
                            const int i = k - 1;
                            const int rob = __ldg(&tp.chain[i].robust);
                            const double(*tn)[TILE] = sh.tnew[i & 1];
                            const double cx = tn[0][lane], cy = tn[1][lane], cz = tn[2][lane];
                            const double da = sh.in.s.D[b][0][lane], ia = sh.in.s.I[b][0][lane];
                            const double qx = sh.in.s.A[b][0][lane], qy = sh.in.s.A[b][1][lane], qz = sh.in.s.A[b][2][lane];
                            const double err = da - dist3m<M>(cx, cy, cz, qx, qy, qz, bad);
                            const double chi = err * (ia * err);
                            sh.qchi[k & 1][0][lane] = chi;
                            sh.qchi[k & 1][1][lane] = (rob & 1) ? E.ck.template rho0m<M>(chi, bad) : chi;
                        }
                    } else if (warp == 0) {
                        /* P adds a step behind: the anchor terms of pose k-2 (Q, previous step), then the terms of
                         * edge (k-3, k-2) it computed itself in the previous step -- insertion order: per pose its
                         * anchor edge, then the trajectory edge to its predecessor */
                        if (k >= 2) {
                            const double(*qc)[TILE] = sh.qchi[(k - 1) & 1];
                            p = p + qc[0][lane];
                            r = r + qc[1][lane];
                            if (k >= 3) {
                                p = p + tchi;
                                r = r + trob;
                            }
                        }
                        if (k >= 1) {
                            const int i = k - 1;
                            const double(*tn)[TILE] = sh.tnew[i & 1];
                            const double cx = tn[0][lane], cy = tn[1][lane], cz = tn[2][lane];
                            if (i > 0) {
                                const int rob = __ldg(&tp.chain[i].robust);
                                const double dt = sh.in.s.D[b][1][lane], it_ = sh.in.s.I[b][1][lane];
                                const double err = dt - dist3m<M>(vx, vy, vz, cx, cy, cz, bad);
                                tchi = err * (it_ * err);
                                trob = (rob & 2) ? E.ck.template rho0m<M>(tchi, bad) : tchi;
                            }
                            vx = cx; vy = cy; vz = cz;
                        }
                    }
                    T4_TICK(4);
                    tma4_barrier<NW>();
                    T4_TICK(5);
                }
                if (warp == 0) { /* the terms of the newest pose N-1: its anchor edge (Q, step N), then edge (N-2, N-1) */
                    if (N >= 1) {
                        const double(*qc)[TILE] = sh.qchi[N & 1];
                        p = p + qc[0][lane];
                        r = r + qc[1][lane];
                        if (N >= 2) {
                            p = p + tchi;
                            r = r + trob;
                        }
                    }
                    sh.chi[0][lane] = p;
                    sh.chi[1][lane] = r;
                }
            }
            const unsigned wb = __ballot_sync(0xffffffffu, act && bad != 0u);
            if (lane == 0) sh.bad[warp] = wb;
            tma4_barrier<NW>(); /* chi2 and the range flags published */
        };
        trial(NbMath{});
        {
            unsigned anybad = sh.bad[0] | sh.bad[1] | sh.bad[2];
            if (NW == 4) anybad |= sh.bad[3];
            if (anybad) trial(IeeeMath{}); /* uniform over the CTA */
        }
        if (warp == 1 && init) {
            plainCur = sh.chi[0][lane];
            currentChi = sh.chi[1][lane];
            stale = plainCur;
            if (cfg.max_iterations > 0) lambda = cfg.tau * maxdiag;
            sh.act[lane] = done ? 0 : 1;
            fence_async_all();
        } else if (warp == 1) {
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
                    cur ^= 1; /* the trial buffer becomes the current estimate of this window */
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
                sh.act[lane] = done ? 0 : 1;
                sh.cur[lane] = cur;
            }
            fence_async_all(); /* the copy engine reads the trial estimates of this sweep in the next trial */
        }
        tma4_barrier<NW>(); /* LM state published */
        init = false;
    }
#ifdef UWBGO_TMA4_TIMING
    if (blockIdx.x == 0 && lane == 0)
        printf("tile 0 warp %d cycles: factor wait-operands %lld work %lld barrier %lld | subst wait-operands %lld work %lld | subst barrier + LM decision %lld (trials %d)\n",
               warp, tw[0], tw[1], tw[2], tw[3], tw[4], tw[5], trials_total);
#endif
    if (warp == 1 && valid) {
        const int64_t tile = w / TILE;
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
        finish_estimates(ws, tile, lane, E.p.T0, E.p.T1, N * 3, cur, last_rej);
        if (E.p.cnt) {
            for (int i = 0; i < N; ++i) {
                long long cc = (long long)E.p.cnt[(size_t)i * TILE] + (long long)iterations * __ldg(tp.num_calls + i) + trials_total;
                E.p.cnt[(size_t)i * TILE] = (int)(cc % cfg.orth_mod);
            }
        }
    }
}

}  // namespace uwbgo
#endif
