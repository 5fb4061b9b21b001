/*
 * uwbgo_chain_tma.cuh — CHAIN path of the fused LM kernel, warp-specialised, operands staged by the
 * bulk-copy engine (TMA, cp.async.bulk + mbarrier).
 *
 * One CTA = one tile of 32 windows = two warps, as in lm_chain_ws_kernel:
 *   warp 0 (P, "edge warp")   linearises the trajectory edge (i-1, i) of every pose and, in the
 *                             substitution phase, evaluates the residuals / chi2 at the new estimates;
 *   warp 1 (C, "chain warp")  linearises the anchor edge, assembles the H record, runs the elimination
 *                             chain and the substitution, owns the LM state.
 * What is new: neither warp issues a global load in the sweeps.  The tile layout arr[tile][row][lane]
 * makes everything one pose needs a handful of CONTIGUOUS runs (3 rows of T = 768 B, the L record =
 * 3840 B, ...), so lane 0 of the chain warp hands each step's inputs of BOTH warps to the copy engine
 * one step ahead (double buffered, one mbarrier per buffer, expect_tx = bytes of the step), and the L
 * records leave through a double-buffered shared-memory record and a bulk store.  The sweeps then read
 * shared memory with immediate offsets: no address arithmetic, no registers holding prefetched
 * operands, no scoreboard stalls on DRAM; a step's ~3.6k cycles hide the copy latency completely.
 *
 * Which of the two estimate buffers T[0] / T[1] is current differs from window to window (it flips
 * with every accepted trial), while the copy engine moves whole rows of 32 windows: both buffers'
 * rows of a pose are staged and every lane reads its own.
 *
 * Per-window arithmetic and its order are those of the single-warp CHAIN path: same bits.
 */
#ifndef UWBGO_CHAIN_TMA_CUH
#define UWBGO_CHAIN_TMA_CUH

#include "uwbgo_fast.cuh"

namespace uwbgo {

/* ---- mbarrier / bulk-copy primitives (PTX ISA 8.0+, sm_90+) ---- */
UWBGO_DI uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
UWBGO_DI void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
UWBGO_DI void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
UWBGO_DI void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
UWBGO_DI void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
/* global -> shared, completion counted on the mbarrier; bytes a multiple of 16, both ends 16-B aligned */
UWBGO_DI void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
/* the same with an L2 eviction policy (createpolicy) for the lines the copy touches */
UWBGO_DI void bulk_load_hint(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long pol)
{
    asm volatile(
        "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
        : "memory");
}
UWBGO_DI void bulk_store_hint(void *dst, const void *src, unsigned bytes, unsigned long long pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
                 "r"(smem_addr(src)), "r"(bytes), "l"(pol)
                 : "memory");
}
UWBGO_DI unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
UWBGO_DI unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#ifndef UWBGO_TMA_HINTS
#define UWBGO_TMA_HINTS 0 /* 1: L records evict_last on the way out, evict_first on the way back; 2: also the streamed operands evict_first */
#endif
/* shared -> global, tracked by the issuing thread's bulk groups */
UWBGO_DI void bulk_store(void *dst, const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
                 : "memory");
}
UWBGO_DI void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
UWBGO_DI void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
UWBGO_DI void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
/* order this thread's generic-proxy writes before later async-proxy (copy engine) accesses */
UWBGO_DI void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
UWBGO_DI void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

#define BULK_STREAM(dst, src, bytes, bar) \
    do { \
        if (UWBGO_TMA_HINTS >= 2) bulk_load_hint(dst, src, bytes, bar, l2_policy_evict_first()); \
        else bulk_load(dst, src, bytes, bar); \
    } while (0)

constexpr unsigned ROW_BYTES = TILE * sizeof(double); /* one row of a tile: 256 B */
constexpr int TMA_STASH_DOUBLES = 2 * FAST_MAX_CARRY * 5 * TILE;
constexpr int SUBST_RING = 3; /* substitution sweep: operands fetched two steps ahead */

struct alignas(128) TmaShared {
    /* staged operands; the two sweeps reuse the same bytes */
    union {
        struct {
            double T[4][2][3][TILE];      /* both estimate buffers of pose j in slot j & 3: read by P, two steps later by C */
            double D[2][2][TILE];         /* [0] range of edge (iP-1, iP) for P, [1] of the anchor edge of iC for C */
            double I[2][2][TILE];         /* their information                                      */
            double A[2][3][TILE];         /* C: anchor of pose iC                                   */
            double L[2][LR_FAST][TILE];   /* C: L record on its way out                             */
        } f;
        struct {
            double L[SUBST_RING][LR_FAST][TILE]; /* C: L record of pose k                           */
            double T[SUBST_RING][2][3][TILE];    /* C: both estimate buffers of pose k              */
            double D[SUBST_RING][2][TILE];       /* P: ranges of the anchor edge and edge (k-2, k-1) */
            double I[SUBST_RING][2][TILE];       /* P: their information                            */
            double A[SUBST_RING][3][TILE];       /* P: anchor of pose k-1                           */
        } s;
    } in;
    double traj[2][8][TILE]; /* P -> C: A(3), B(3), Ow, omega_r of edge (i-1, i), double buffered */
    double tnew[2][3][TILE]; /* C -> P: new estimate of pose i                                    */
    double chi[2][TILE];     /* P -> C: plain and robust chi2 of the trial                        */
    int act[TILE];           /* C -> P: window still being optimised                              */
    int cur[TILE];           /* C -> P: which estimate buffer of this window is current           */
    unsigned bad[2];         /* per warp: some active lane left the safe range of NbMath          */
    unsigned long long fbar[2];          /* factor sweep: one per step parity                     */
    unsigned long long sbar[SUBST_RING]; /* substitution sweep: one per ring slot                 */
#ifdef UWBGO_TMA_PAD /* experiment: fewer resident tiles per SM (a smaller working set in L2) at unchanged registers */
    char pad[UWBGO_TMA_PAD];
#endif
};
static_assert(sizeof(((TmaShared *)nullptr)->in) >= TMA_STASH_DOUBLES * sizeof(double), "stash must fit in the staging area");

UWBGO_DI void tma_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

/* tile-level (lane 0) views of the workspace: base of this tile's rows */
struct TileBase {
    const double *T0, *T1, *rd, *ri, *anch;
    double *LR;
};

/* inputs of factor step k (k = 0 .. N): P works on edge (iP-1, iP), iP = N-1-k; C on pose iC = N-k.
 * Rows 2 iC - 2 (edge (iP-1, iP)) and 2 iC - 1 (anchor edge of iC) of the range arrays are adjacent. */
UWBGO_DI void issue_factor_step(const DevTopo &tp, const TileBase &g, TmaShared &sh, int N, int k, unsigned ev)
{
    const int b = (int)(ev & 1u);
    unsigned long long *bar = &sh.fbar[b];
    const int iC = N - k, jT = N - 2 - k;
    const bool first = k == 0, last = k == N;
    const unsigned drows = (first || last) ? 1u : 2u;
    const unsigned bytes = ((jT >= 0 ? 6u : 0u) + (first ? 6u : 0u) + 2u * drows + (first ? 0u : 3u)) * ROW_BYTES;
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
        /* first: row 2N-2 -> [0]; last: row 0 -> [1]; otherwise rows (2 iC - 2, 2 iC - 1) -> [0], [1] */
        const int r0 = last ? 0 : 2 * iC - 2;
        BULK_STREAM(sh.in.f.D[b][last ? 1 : 0], g.rd + (size_t)r0 * TILE, drows * ROW_BYTES, bar);
        BULK_STREAM(sh.in.f.I[b][last ? 1 : 0], g.ri + (size_t)r0 * TILE, drows * ROW_BYTES, bar);
    }
    if (!first) {
        const int anchor = __ldg(&tp.chain[iC].anchor);
        bulk_load(sh.in.f.A[b], g.anch + (size_t)anchor * 3 * TILE, 3 * ROW_BYTES, bar);
    }
}

/* inputs of substitution step k (k = 0 .. N) into ring slot `slot`: C works on pose k, P on pose k-1 */
UWBGO_DI void issue_subst_step(const DevTopo &tp, const TileBase &g, TmaShared &sh, int N, int k, int slot)
{
    unsigned long long *bar = &sh.sbar[slot];
    const bool doC = k < N, doP = k >= 1;
    const int j = k - 1;
    const unsigned prow = j == 0 ? 1u : 2u;
    const unsigned bytes = (doC ? (unsigned)(LR_FAST + 6) : 0u) * ROW_BYTES + (doP ? (2u * prow + 3u) : 0u) * ROW_BYTES;
    mbar_expect_tx(bar, bytes);
    if (doC) {
        if (UWBGO_TMA_HINTS >= 1) bulk_load_hint(sh.in.s.L[slot], g.LR + (size_t)k * LR_FAST * TILE, LR_FAST * ROW_BYTES, bar, l2_policy_evict_first());
        else bulk_load(sh.in.s.L[slot], g.LR + (size_t)k * LR_FAST * TILE, LR_FAST * ROW_BYTES, bar);
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

#ifndef UWBGO_TMA_MINB
#define UWBGO_TMA_MINB 8
#endif
__global__ void __launch_bounds__(64, UWBGO_TMA_MINB)
lm_chain_tma_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                    const __grid_constant__ DevWs ws)
{
    __shared__ TmaShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * TILE + lane; /* tail lanes own zero-filled pad columns */
    const bool valid = w < ws.W;
    const int N = tp.N;
    const bool issuerF = warp == 0 && lane == 0; /* factor sweep: the edge warp is the lighter one   */
    const bool issuerS = warp == 1 && lane == 0; /* substitution sweep: the chain warp is, and it owns the L stores */
    FastEnv E;
    E.tp = &tp;
    E.p = thread_ptrs<HR_FAST, LR_FAST>(tp, ws, w);
    E.ck.init(cfg.kdelta);
    E.delta = cfg.jdelta;
    E.scalar = 1.0 / (2.0 * cfg.jdelta);
    E.bs = TILE;
    E.stash = reinterpret_cast<double *>(&sh.in) + lane; /* used by the initial linearisation only, before any copy is issued */
    E.anch = E.p.anch;
    E.anch_stride = TILE;
    /* tile bases from blockIdx only: the copy engine's addresses stay in uniform registers */
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
    /* The first pass of the loop below is not a trial but the initial evaluation, run through the same
     * two-warp machinery: its factor sweep yields max diag(H) (g2o's computeLambdaInit), its
     * substitution sweep is forced to x = 0 so that the edge warp's chi2 is the chi2 of the initial
     * estimate. */
    bool init = true;
    unsigned ev = 0;       /* factor steps so far: buffer = ev & 1, mbarrier parity = (ev >> 1) & 1 */
    int sslot = 0;         /* substitution ring: slot of the next step to consume ...               */
    unsigned spar = 0;     /* ... and the parity to wait for, one bit per slot                      */
    for (;;) {
        const bool act = sh.act[lane] != 0;
        const int c = sh.cur[lane]; /* the estimate buffer this window reads in this trial */
        if (__ballot_sync(0xffffffffu, act) == 0u) break;
        tma_barrier(); /* everybody has read act / cur before the chain warp may overwrite them */
        double *const Tn = E.p.T(c ^ 1); /* trial estimates */
        bool ok = true;
        double scale = 0.0;
        /* one trial = factor sweep + substitution sweep of both warps.  It only writes the L records,
         * the trial estimates and the shared hand-off buffers, so it can be repeated: first with the
         * branch-free arithmetic (NbMath) and -- when any active lane of the tile saw an operand
         * outside NbMath's safe range -- once more with the IEEE sequences. */
        auto trial = [&](auto math_tag) {
            using M = decltype(math_tag);
            unsigned bad = 0;
            ok = true;
            scale = 0.0;
            maxdiag = 0.0;
            const double lam = init ? 1.0 : lambda;
            /* ================= factor sweep: steps k = 0 .. N ================= */
            if (issuerF) issue_factor_step(tp, g, sh, N, 0, ev);
            {
                /* P state: pose i (c) of edge (i-1, i); C state: elimination carry */
                double cx = 0.0, cy = 0.0, cz = 0.0;
                double G[9], zn[3], carry[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int m = 0; m < 9; ++m) G[m] = 0.0;
#pragma unroll
                for (int m = 0; m < 3; ++m) zn[m] = 0.0;
                for (int k = 0; k <= N; ++k) {
                    const int b = (int)((ev + (unsigned)k) & 1u);
                    if (issuerF && k < N) issue_factor_step(tp, g, sh, N, k + 1, ev + (unsigned)k + 1u);
                    if (issuerS && k >= 2) bulk_wait_read<1>(); /* the store of step k-2 has left its L buffer */
                    mbar_wait(&sh.fbar[b], ((ev + (unsigned)k) >> 1) & 1u);
                    if (warp == 0) {
                        const int i = N - 1 - k; /* edge (i-1, i) */
                        if (k == 0) {
                            const double(*tc)[TILE] = sh.in.f.T[(N - 1) & 3][c];
                            cx = tc[0][lane]; cy = tc[1][lane]; cz = tc[2][lane];
                        }
                        if (i >= 1) {
                            const double(*tp_)[TILE] = sh.in.f.T[(i - 1) & 3][c];
                            const double px = tp_[0][lane], py = tp_[1][lane], pz = tp_[2][lane];
                            const double dt = sh.in.f.D[b][0][lane], it_ = sh.in.f.I[b][0][lane];
                            const int rob = __ldg(&tp.chain[i].robust);
                            double A[3], B[3], Ow, omega_r, err;
                            range_linearize<M, true, true>(px, py, pz, cx, cy, cz, dt, E.delta, E.scalar, err, A, B, bad);
                            chain_weights<M>(E, err, it_, (rob & 2) != 0, Ow, omega_r, &bad);
                            double(*o)[TILE] = sh.traj[k & 1];
                            o[0][lane] = A[0]; o[1][lane] = A[1]; o[2][lane] = A[2];
                            o[3][lane] = B[0]; o[4][lane] = B[1]; o[5][lane] = B[2];
                            o[6][lane] = Ow;   o[7][lane] = omega_r;
                            cx = px; cy = py; cz = pz;
                        }
                    } else if (k >= 1) {
                        const int i = N - k; /* pose i */
                        __syncwarp();        /* lane 0 has waited for the L buffer */
                        const double(*tc)[TILE] = sh.in.f.T[i & 3][c];
                        const double px = tc[0][lane], py = tc[1][lane], pz = tc[2][lane];
                        const double qx = sh.in.f.A[b][0][lane], qy = sh.in.f.A[b][1][lane], qz = sh.in.f.A[b][2][lane];
                        const double da = sh.in.f.D[b][1][lane], ia = sh.in.f.I[b][1][lane];
                        const int rob = __ldg(&tp.chain[i].robust);
                        double h[HR_FAST];
#pragma unroll
                        for (int m = 0; m < HR_FAST; ++m) h[m] = 0.0;
                        {
                            double J[3], Ow, omega_r, err;
                            range_linearize<M, true, false>(px, py, pz, qx, qy, qz, da, E.delta, E.scalar, err, J, nullptr, bad);
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
                        if (init) {
                            double v;
                            v = fabs(h[0]); if (v > maxdiag) maxdiag = v;
                            v = fabs(h[3]); if (v > maxdiag) maxdiag = v;
                            v = fabs(h[5]); if (v > maxdiag) maxdiag = v;
                        }
                        double *l = &sh.in.f.L[k & 1][0][lane];
                        factor_step<3, M>(h, l, true, i > 0, lam, G, zn, ok, &bad);
                        chain_store_b(l, h);
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (UWBGO_TMA_HINTS >= 1) bulk_store_hint(g.LR + (size_t)i * LR_FAST * TILE, sh.in.f.L[k & 1], LR_FAST * ROW_BYTES, l2_policy_evict_last());
                            else bulk_store(g.LR + (size_t)i * LR_FAST * TILE, sh.in.f.L[k & 1], LR_FAST * ROW_BYTES);
                            bulk_commit();
                        }
                    }
                    tma_barrier();
                }
                ok = !init && ok && (lambda > 0.0);
            }
            ev += (unsigned)N + 1u;
            /* ================= substitution sweep: steps k = 0 .. N, fetched two steps ahead ================= */
            int islot = sslot; /* issuer: slot of the next step to fetch */
            if (issuerS) {
                bulk_wait<0>(); /* every L record has landed before the first one is fetched back */
                issue_subst_step(tp, g, sh, N, 0, islot);
                islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                if (N >= 1) {
                    issue_subst_step(tp, g, sh, N, 1, islot);
                    islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                }
            }
            {
                double xp[3] = {0.0, 0.0, 0.0};         /* C: x of the previous pose        */
                double p = 0.0, r = 0.0;                /* P: plain / robust chi2           */
                double vx = 0.0, vy = 0.0, vz = 0.0;    /* P: new estimate of pose k-2      */
                for (int k = 0; k <= N; ++k) {
                    const int b = sslot;
                    if (issuerS && k + 2 <= N) {
                        issue_subst_step(tp, g, sh, N, k + 2, islot);
                        islot = islot == SUBST_RING - 1 ? 0 : islot + 1;
                    }
                    mbar_wait(&sh.sbar[b], (spar >> b) & 1u);
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
                    } else if (k >= 1) {
                        const int i = k - 1; /* pose i was produced in the previous step */
                        const int rob = __ldg(&tp.chain[i].robust);
                        const double(*tn)[TILE] = sh.tnew[i & 1];
                        const double cx = tn[0][lane], cy = tn[1][lane], cz = tn[2][lane];
                        {
                            const double da = sh.in.s.D[b][0][lane], ia = sh.in.s.I[b][0][lane];
                            const double qx = sh.in.s.A[b][0][lane], qy = sh.in.s.A[b][1][lane], qz = sh.in.s.A[b][2][lane];
                            const double err = da - dist3m<M>(cx, cy, cz, qx, qy, qz, bad);
                            const double chi = err * (ia * err);
                            p = p + chi;
                            r = r + ((rob & 1) ? E.ck.template rho0m<M>(chi, bad) : chi);
                        }
                        if (i > 0) {
                            const double dt = sh.in.s.D[b][1][lane], it_ = sh.in.s.I[b][1][lane];
                            const double err = dt - dist3m<M>(vx, vy, vz, cx, cy, cz, bad);
                            const double chi = err * (it_ * err);
                            p = p + chi;
                            r = r + ((rob & 2) ? E.ck.template rho0m<M>(chi, bad) : chi);
                        }
                        vx = cx; vy = cy; vz = cz;
                    }
                    tma_barrier();
                }
                if (warp == 0) {
                    sh.chi[0][lane] = p;
                    sh.chi[1][lane] = r;
                }
            }
            const unsigned wb = __ballot_sync(0xffffffffu, act && bad != 0u);
            if (lane == 0) sh.bad[warp] = wb;
            tma_barrier(); /* chi2 and the range flags published */
        };
        trial(NbMath{});
        if (sh.bad[0] | sh.bad[1]) trial(IeeeMath{}); /* uniform over the CTA */
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
        tma_barrier(); /* LM state published */
        init = false;
    }
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
