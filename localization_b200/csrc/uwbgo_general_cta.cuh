/*
 * uwbgo_general_cta.cuh — GENERAL path, one CTA per tile of 32 windows.
 *
 * lm_general_kernel gives one thread the whole window: at the batch sizes the general graphs come
 * in (thousands of windows of 15-20 poses with IMU / lidar / twist / pose edges) that is a handful
 * of warps per SM, each walking ~60 edges and 20 poses serially.  Here a tile of 32 windows gets a
 * CTA of NW warps and the LM trial is cut into phases at __syncthreads():
 *
 *   L  buildSystem      warp k linearises poses k, k+NW, ... (lane = window): one H record per
 *                       (window, pose), gathered in g2o insertion order (gen_linearize_pose)
 *   F  solve            warp 0: block elimination, substitution, computeScale (the serial chain of
 *                       the window; one lane per window); the other warps prefetch its H records
 *   U  update           warp k applies x to poses k, k+NW, ... (oplus into the trial buffer)
 *   C  computeActiveErrors  warp k evaluates edges k, k+NW, ...; chi2 and rho0 per edge to scratch
 *   D  decision         warp 0 sums the per-edge values in insertion order, accept / reject, lambda
 *
 * Per-window arithmetic and summation orders are those of lm_general_kernel (same device functions):
 * same bits.  The LM bookkeeping is the flat loop of lm_window (uwbgo_kernels.cu).
 */
#ifndef UWBGO_GENERAL_CTA_CUH
#define UWBGO_GENERAL_CTA_CUH
#ifdef UWBGO_GCTA_TIMING
#include <cstdio>
#endif

namespace uwbgo {

#ifndef UWBGO_GCTA_WARPS
#define UWBGO_GCTA_WARPS 4
#endif
#ifndef UWBGO_GCTA_MINB
#define UWBGO_GCTA_MINB 2
#endif
#ifndef UWBGO_GCTA_SMEM_ACC
#define UWBGO_GCTA_SMEM_ACC 0 /* 1: accumulators of the linearise phase in shared memory -- fewer spills (2.4 KB -> 0.6 KB
                                * of spill code) but slower: 11.2 vs 10.5 ms on C4a, 85.5 vs 76.9 ms at 65,536 windows */
#endif
constexpr size_t gcta_dyn_smem(int nw) { return UWBGO_GCTA_SMEM_ACC ? sizeof(double) * HR_GEN * nw * 32 : 0; }

constexpr int GCTA_MAX_WARPS = 8;
struct GctaShared {
    double ant[3 * MAX_SMEM_ANTENNAS];
    double maxd[GCTA_MAX_WARPS][TILE];   /* L -> F: max |H_kk| over the poses each warp linearised */
    int lin[TILE];                       /* D -> L: window starts an iteration (buildSystem)       */
    int cur[TILE];                       /* D -> all: which pose buffer holds the estimate         */
    int act[TILE];                       /* D -> all: window still being optimised                 */
    int go, anylin;
};

/* NW warps per tile: 4 (two CTAs per SM at 255 registers) is the default; when the batch has no more tiles
 * than the device has SMs, 8 warps and one CTA per SM halve the rounds of the L / U / C phases */
template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
lm_general_cta_kernel(const __grid_constant__ DevTopo tp, const __grid_constant__ DevCfg cfg,
                      const __grid_constant__ DevWs ws)
{
    __shared__ GctaShared sh;
#if UWBGO_GCTA_SMEM_ACC
    extern __shared__ double gcta_acc[]; /* [HR_GEN][NW * 32] */
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (tp.K <= MAX_SMEM_ANTENNAS)
        for (int k = threadIdx.x; k < 3 * tp.K; k += NW * 32) sh.ant[k] = ws.ant[k];
    const int64_t w = (int64_t)blockIdx.x * TILE + lane; /* workspaces are padded to whole tiles */
    GenEnv E;
    gen_env_init(E, tp, cfg, ws, w, sh.ant);
    double *echi = ws.echi + ((size_t)blockIdx.x * tp.E * 2) * TILE + lane;
    const int N = tp.N, NE = tp.E;
    auto buf = [&](int k) { return PoseBuf{E.p.T(k), E.p.Rm(k)}; };

    /* LM state of the lane's window; meaningful in warp 0 only */
    double lambda = 0.0, ni = 2.0, stale = 0.0, plainCur = 0.0, currentChi = 0.0, rho = 0.0;
    int iterations = 0, trials_total = 0, flags = 0, qlast = 0, cur = 0, q = 0, it = 0;
    bool need_lin = true, done = (w >= ws.W) || cfg.max_iterations <= 0;
    bool tok = true;     /* F -> D: the trial's factorisation succeeded */
    double tscale = 0.0; /* F -> D: computeScale() of the trial         */
    const bool valid = w < ws.W;

    /* phase C on the buffer `sel ^ cur` of every lane flagged in `who` */
    auto chi_phase = [&](const int *who, int sel) {
        if (who[lane]) {
            const PoseBuf T = buf(sh.cur[lane] ^ sel);
            /* edges in slot order, 6-D kinds first (slots: range | prior | se3): in insertion order the kinds of a
             * pose's edges repeat with a period that divides NW and every warp would get one kind only */
            for (int u = warp; u < NE; u += NW) {
                const int e = __ldg(tp.slot_edge + (NE - 1 - u));
                double chi, rob;
                if ((UWBGO_GCTA_PF & 2) && u + NW < NE) gen_edge_prefetch(E, T, __ldg(tp.slot_edge + (NE - 1 - u - NW)));
                gen_edge_chi(E, T, e, chi, rob);
                ROW(echi, 2 * e) = chi;
                ROW(echi, 2 * e + 1) = rob;
            }
        }
    };
    auto chi_sum = [&](double &p, double &r) {
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
    auto publish = [&]() { /* warp 0 */
        sh.lin[lane] = (!done && need_lin) ? 1 : 0;
        sh.cur[lane] = cur;
        sh.act[lane] = done ? 0 : 1;
        const unsigned g = __ballot_sync(0xffffffffu, !done);
        const unsigned l = __ballot_sync(0xffffffffu, !done && need_lin);
        if (lane == 0) {
            sh.go = g != 0;
            sh.anylin = l != 0;
        }
    };

    /* initial computeActiveErrors: every real window, buffer 0 */
    if (warp == 0) {
        sh.cur[lane] = 0;
        sh.act[lane] = valid ? 1 : 0;
    }
    __syncthreads();
    chi_phase(sh.act, 0);
    __syncthreads();
    if (warp == 0) {
        if (valid) chi_sum(plainCur, currentChi);
        stale = plainCur;
        publish();
    }
    __syncthreads();

#ifdef UWBGO_GCTA_TIMING
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tq = clock64();
#define GCTA_TICK(k) do { long long tn_ = clock64(); tph[k] += tn_ - tq; tq = tn_; } while (0)
#else
#define GCTA_TICK(k)
#endif
    while (sh.go) {
        if (sh.anylin) { /* phase L */
            double md = 0.0;
            if (sh.lin[lane]) {
                const PoseBuf T = buf(sh.cur[lane]);
                for (int i = warp; i < N; i += NW) {
#if UWBGO_GCTA_SMEM_ACC
                    /* the 63 accumulators of the pose's H record live in this thread's shared-memory
                     * column instead of 126 registers the linearisation of a 6-D edge does not have */
                    constexpr int ST = NW * 32;
                    double *col = gcta_acc + threadIdx.x;
                    double m = gen_linearize_pose_acc<false>(E, T, i, SmemAcc<ST>{col}, SmemAcc<ST>{col + 21 * ST},
                                                             SmemAcc<ST>{col + 57 * ST});
#else
                    double m = gen_linearize_pose<false>(E, T, i);
#endif
                    if (m > md) md = m;
                }
            }
            sh.maxd[warp][lane] = md;
            __syncthreads();
            GCTA_TICK(0);
        }
        if (warp == 0 && !done) { /* phase F */
            if (need_lin) {
                stale = plainCur;
                double maxdiag = 0.0;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    double m = sh.maxd[k][lane];
                    if (m > maxdiag) maxdiag = m;
                }
                if (it == 0) {
                    lambda = cfg.tau * maxdiag;
                    ni = 2.0;
                }
                rho = 0.0;
                q = 0;
                need_lin = false;
            }
            /* (the branch-free sqrt / reciprocal of the CHAIN kernels lose here: 11.4 vs 10.4 ms on C4a,
             * the 6x6 elimination step is already out of registers) */
            const bool ok = tp.tree ? factor_sweep_tree(tp, E.p.HB, E.p.LR, lambda)
                                    : factor_sweep<6>(E.p.HB, E.p.LR, N, lambda);
            if (!ok) flags |= UWBGO_FLAG_CHOL_FAIL;
            GCTA_TICK(1);
            tscale = gen_subst_scale(E, ok, lambda);
            tok = ok;
        } else if (warp != 0 && NW > 1) {
            /* the other warps pull the tile's H records towards L2 for the sweep of warp 0 */
            const char *hb = reinterpret_cast<const char *>(E.p.HB - lane);
            const size_t total = (size_t)N * HR_GEN * TILE * sizeof(double);
            for (size_t off = ((size_t)(warp - 1) * 32 + lane) * 128; off < total; off += (size_t)(NW - 1) * 32 * 128)
                prefetch_l2(hb + off);
        }
        __syncthreads();
        GCTA_TICK(2);
        if (sh.act[lane]) { /* phase U: estimate (+) x into the trial buffer, pose by pose */
            const int c = sh.cur[lane];
            const bool lin = sh.lin[lane] != 0;
            for (int i = warp; i < N; i += NW) gen_update_pose(E, i, buf(c), buf(c ^ 1), lin);
        }
        __syncthreads();
        GCTA_TICK(3);
        chi_phase(sh.act, 1); /* phase C at the trial estimates */
        __syncthreads();
        GCTA_TICK(4);
        if (warp == 0) { /* phase D */
            if (!done) {
                const bool ok = tok;
                double scale = tscale, tplain, tempChi;
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
            publish();
        }
        __syncthreads();
        GCTA_TICK(5);
    }
#ifdef UWBGO_GCTA_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 2)
        printf("tile %d cycles: L %lld  factor %lld  subst %lld  U %lld  C %lld  D %lld\n", (int)blockIdx.x, tph[0],
               tph[1], tph[2], tph[3], tph[4], tph[5]);
#endif

    if (warp == 0 && valid) {
        double *chi2_out = ws.chi2 + (int64_t)blockIdx.x * 4 * TILE + lane;
        int32_t *status_out = ws.status + (int64_t)blockIdx.x * 4 * TILE + lane;
        ROW(chi2_out, 0) = plainCur;
        ROW(chi2_out, 1) = currentChi;
        ROW(chi2_out, 2) = stale;
        ROW(chi2_out, 3) = lambda;
        ROW(status_out, 0) = iterations;
        ROW(status_out, 1) = trials_total;
        ROW(status_out, 2) = flags;
        ROW(status_out, 3) = qlast;
    }
    if (valid && sh.cur[lane]) { /* result always leaves in buffer 0 */
        for (int r = warp; r < N * 3; r += NW) ROW(E.p.T0, r) = ROW(E.p.T1, r);
        for (int r = warp; r < N * 9; r += NW) ROW(E.p.Rm0, r) = ROW(E.p.Rm1, r);
    }
}

}  // namespace uwbgo
#endif
