/*
 * uwbgo_internal.h — types shared by the kernels (uwbgo_kernels.cu) and the C-ABI host layer
 * (uwbgo_api.cu).  Not part of the public ABI (include/uwbgo.h).
 *
 * Device data layout ("tile layout").  One GPU thread solves one window; a warp therefore
 * owns a TILE of 32 consecutive windows.  Every per-window array lives in HBM as
 *     arr[tile][row][lane]          (lane = window % 32 fastest)
 * so that a warp-wide access to one row is one fully coalesced 256-byte transaction and a
 * tile's rows are one contiguous block that the warp streams front to back (linearise,
 * substitution sweep) or back to front (factor sweep).  The window-major arrays of the public
 * ABI are transposed into / out of this layout by the pack / unpack kernels.
 */
#ifndef UWBGO_INTERNAL_H
#define UWBGO_INTERNAL_H

#include <stdint.h>
#include <cuda_runtime.h>

#include "../../include/uwbgo.h"

namespace uwbgo {

constexpr int TILE = 32;          /* windows per tile = lanes per warp            */
constexpr int CTA_THREADS = 128;  /* 4 tiles per CTA                              */

/* rows per pose of the H / L records */
constexpr int HR_FAST = 18;       /* Hd upper 6 | link block 9 | b 3              */
constexpr int HR_GEN = 63;        /* Hd upper 21 | link block 36 | b 6            */
constexpr int LR_FAST = 15;       /* c 3 | M 9 | b 3   (x_i = c_i - M_i x_{i-1})  */
constexpr int LR_GEN = 42;        /* c 6 | M 36                                   */
constexpr int LR_TREE = 90;       /* c 6 | M 36 | G 36 | z 6 | x 6  (forest windows) */

/* One edge of the shared topology, in g2o insertion order. */
struct EdgeRec {
    int32_t kind;    /* UWBGO_EDGE_*                                                       */
    int32_t a, b;    /* vertex 0 pose index; vertex 1 pose index or anchor index           */
    int32_t slot;    /* per-kind data slot                                                 */
    int32_t ant;     /* antenna number of the vertex-0 offset (0 = none)                   */
    int32_t robust;  /* Cauchy kernel on/off                                               */
    int32_t base_a;  /* numeric-Jacobian oplus calls made on pose a before this edge's,    */
    int32_t base_b;  /*   resp. on pose b, within one linearisation (VertexSE3 counter)    */
    int32_t ant_b;   /* antenna number of the vertex-1 offset (0 = none)                   */
    int32_t pad[3];  /* records stay 16-byte aligned (load_edge reads them as int4)        */
};

/* Per-pose gather program for Hessian assembly: the edges touching pose i, in insertion
 * order, with the role pose i plays (0 = vertex 0 / owner, 1 = vertex 1). */
struct PoseOp {
    int32_t edge;
    int32_t role;
};

/* Schedule of the fused substitution + residual sweep: poses are produced in ascending order,
 * every edge is evaluated (in insertion order) as soon as both its poses exist. */
struct SchedOp {
    int32_t kind; /* 0 = produce pose idx, 1 = evaluate edge idx */
    int32_t idx;
};

/* CHAIN path: per-pose table of the window Localization::addRangeEdge builds (one anchor range
 * edge per pose, then the trajectory edge to its predecessor) */
struct ChainPose {
    int32_t anchor; /* anchor index of the pose's range edge                         */
    int32_t robust; /* bit 0: anchor edge has a kernel; bit 1: trajectory edge (k-1,k) */
};

struct DevTopo {
    int32_t N, A, K, E, Er, Ep, Es;
    int32_t fast;               /* 0 general 6x6 path; 1 translation-only path; 2 translation-only, */
                                /* standard chain (straight-line sweeps)                            */
    int32_t n_sched;
    int32_t n_ops;              /* length of `ops`                                           */
    int32_t tree;               /* 1: some pose's older neighbour is not its predecessor     */
    int32_t simple_chain;       /* 1: parent(i) = i - 1 for every i > 0 (no roots inside)    */
    const ChainPose *chain;     /* [N], fast == 2 only                                       */
    const int32_t *parent;      /* [N] the one older neighbour of pose j, -1 = none          */
    const int32_t *child_begin; /* [N+1] CSR over `children`                                 */
    const int32_t *children;    /* children of each pose, descending                         */
    const SchedOp *sched;       /* [n_sched = N + E]                                         */
    const EdgeRec *edges;       /* [E]                                                       */
    const PoseOp *ops;          /* concatenated per-pose op lists                            */
    const int32_t *op_begin;    /* [N+1]                                                     */
    const int32_t *num_calls;   /* [N] numeric oplus calls per linearisation on pose i       */
    const int32_t *slot_edge;   /* [Er + Ep + Es] edge index of every data slot: range slots, */
                                /* then prior slots, then se3 slots                          */
    const int32_t *range_sub;   /* [Er] per range slot: index of the edge among the edges of  */
                                /* its own kind; bit 30 set = RANGE_POSE (compact range form) */
    int32_t Era, Erp;           /* number of RANGE_ANCHOR / RANGE_POSE edges                  */
};

/* uwbgo_range_msgs with device pointers (window-major, as the caller passed them) */
struct RangeMsgsDev {
    const float *distance, *distance_err;
    const double *dt_anchor, *dt_pose;
    double v_max;
};

struct DevCfg {
    int32_t max_iterations, max_trials, orth_mod; /* orth_mod = orthogonalize_after + 1 */
    double tau, good_lo, good_hi, kdelta, jdelta;
};

/* tile-layout workspace of one launch (all device pointers) */
struct DevWs {
    int64_t W;          /* windows in this launch                     */
    double *T[2];       /* translations   [tile][N*3][32], two buffers */
    double *Rm[2];      /* rotations      [tile][N*9][32]  (general)  */
    int32_t *cnt;       /* oplus counters [tile][N][32]   (general)   */
    const double *anch; /* [tile][A*3][32]                            */
    const double *rd;   /* [tile][Er][32]                             */
    const double *ri;   /* [tile][Er][32]                             */
    const double *pZ;   /* [tile][Ep*12][32]                          */
    const double *pI;   /* [tile][Ep*36][32]                          */
    const double *sZ;   /* [tile][Es*12][32]                          */
    const double *sI;   /* [tile][Es*36][32]                          */
    double *HB;         /* H records      [tile][N*HR][32]            */
    double *LR;         /* L records      [tile][N*LR][32]            */
    const double *ant;  /* [K][3] antenna offsets, plain              */
    double *chi2;       /* [tile][4][32] (solve) or [tile][2][32]     */
    int32_t *status;    /* [tile][4][32]                              */
    double *echi;       /* [tile][E*2][32] per-edge chi2 | rho0 (general CTA kernel) */
    double *jrec;       /* [tile][general_items_jrec_rows][32] per-edge linearisation records (general ITEM kernel) */
    int64_t W_batch;    /* windows of the whole call when this launch is one chunk of a pipelined batch whose chunks
                         * run concurrently (0: this launch is the batch); picks the CHAIN kernel's warps per tile */
    int32_t *stale_sel; /* [tile][32] or NULL.  Non-NULL = the caller wants uwbgo_result::edge_chi2: the      */
                        /* translation-only kernels then leave the last trial's estimates in buffer stale_sel */
};

/* WINDOW path (uwbgo_window.cu, one CTA per window): the window-major arrays of the public ABI,
 * read and written in place -- device memory or mapped pinned host memory */
struct WinIo {
    int64_t W;
    const double *pose_t, *pose_R; /* pose_R NULL = identity rotations */
    const int32_t *cnt_in;         /* NULL = zeros                     */
    const double *anchors, *rd, *ri, *pZ, *pI, *sZ, *sI;
    const double *ant;             /* [K][3]                           */
    double *o_pose_t, *o_pose_R;   /* o_pose_R, o_cnt, o_chi2, o_status may be NULL */
    int32_t *o_cnt;
    double *o_chi2;
    int32_t *o_status;
    int32_t bd_ok;                 /* 1: range edges without lever arms and EdgeSE3Prior edges only, a simple chain: */
                                   /* the window is block-diagonal if its priors' information has no cross blocks    */
    int32_t t3;                    /* 1: range edges only, identity rotations, no lever arms, a simple chain: every */
                                   /* 6x6 block is zero outside its translation entries (3x3 solve, same bits)      */
};

/* pack / unpack job: transpose between window-major [W][C] and tile layout [tile][C][32] */
struct XposeJob {
    const void *src;
    void *dst;
    int32_t C;       /* columns per window                                  */
    int32_t elem;    /* element size: 8 (double) or 4 (int32)               */
    int32_t mode;    /* 0 transpose; 1 identity rotations (tile layout dst); 2 identity rotations */
                     /* (window-major dst); 3 zero fill (tile layout dst); 4 broadcast of one    */
                     /* shared row src[C] to every window (tile layout dst)                      */
    int32_t aux;
};
constexpr int MAX_XPOSE_JOBS = 12;
struct XposeJobs {
    int32_t n;
    int32_t pad;
    int64_t W;
    XposeJob job[MAX_XPOSE_JOBS];
};

/* launchers (uwbgo_kernels.cu); all asynchronous on `st`, return cudaGetLastError() */
cudaError_t launch_pack(const XposeJobs &jobs, cudaStream_t st);
/* compact range form -> measurement / information rows of the tile layout (create_range_edge on the device) */
cudaError_t launch_pack_range_msgs(const DevTopo &topo, const RangeMsgsDev &m, int64_t W, double *rd, double *ri,
                                   cudaStream_t st);
/* uwbgo_result::edge_chi2 [W][E] after launch_solve (window-major device array) */
cudaError_t launch_edge_chi2_out(const DevTopo &topo, const DevWs &ws, double *edge_chi2, cudaStream_t st);
/* uwbgo_result::marginal [W][36] / marginal_ok [W] after launch_solve; scratch: marginal_scratch_bytes() */
size_t marginal_scratch_bytes(const DevTopo &topo, int64_t W);
cudaError_t launch_marginal(const DevTopo &topo, const DevWs &ws, double *scratch, double *marginal,
                            int32_t *marginal_ok, cudaStream_t st);
cudaError_t launch_unpack(const XposeJobs &jobs, cudaStream_t st);
cudaError_t launch_solve(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws, cudaStream_t st);
cudaError_t launch_linearize(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws,
                             cudaStream_t st);
/* GENERAL path, ITEM kernel (uwbgo_general_items.cu): chains of 6x6 blocks, 8 warps per tile at 128 registers */
size_t general_items_jrec_rows(const DevTopo &topo);
bool general_items_ok(const DevTopo &topo, const DevWs &ws);
cudaError_t launch_solve_general_items(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws, cudaStream_t st);
/* WINDOW path: whole LM solve, one CTA per window, state in shared memory (uwbgo_window.cu) */
int window_path_candidates(int64_t W); /* LM trials evaluated speculatively per round for a batch of W */
size_t window_path_smem_bytes(const DevTopo &topo, int candidates);
cudaError_t launch_solve_window(const DevTopo &topo, const DevCfg &cfg, const WinIo &io, int candidates,
                                int device, cudaStream_t st);
/* CHAIN windows: linearise straight into the public window-major arrays, one kernel (+ chi2 pass) */
bool linearize_chain_fused_ok(const DevTopo &topo, const double *H_diag, const double *H_off, const double *b);
cudaError_t launch_linearize_chain_fused(const DevTopo &topo, const DevCfg &cfg, const DevWs &ws,
                                         double *H_diag, double *H_off, double *b, bool want_chi,
                                         cudaStream_t st);
/* expand H records (tile layout) into the public window-major H_diag/H_off/b arrays */
cudaError_t launch_expand_H(const DevTopo &topo, const DevWs &ws, double *H_diag, double *H_off,
                            double *b, cudaStream_t st);
/* (H + lambda I) x = b on window-major public arrays */
cudaError_t launch_factor_solve(int32_t N, int64_t W, const double *H_diag, const double *H_off,
                                const double *b, const double *lambda, double *x, int32_t *ok,
                                double *scratch, cudaStream_t st);
size_t factor_solve_scratch_bytes(int32_t N, int64_t W);
cudaError_t launch_fp64_peak(double *out, int iters, cudaStream_t st, int *blocks, int *threads);
cudaError_t launch_math_selftest(unsigned long long seed, int blocks, int per_thread, int mode,
                                 unsigned long long *counts, cudaStream_t st);

__host__ __device__ inline int64_t n_tiles(int64_t W) { return (W + TILE - 1) / TILE; }

}  // namespace uwbgo
#endif
