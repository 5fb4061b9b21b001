/*
 * uwbgo.h — C ABI of the B200-native batched sliding-window LM solver.
 *
 * This is the drop-in boundary for the one hot path of sair-lab/localization:
 * Localization::solve() (reference src/localization/localization.cpp:164-192), i.e.
 *   optimizer.initializeOptimization(); optimizer.optimize(iteration_max);
 * run over W independent windows at once.  A window is what the g2o graph holds at the
 * moment of that call: N moving VertexSE3 poses (src/localization/robot.cpp:39-55,75-110),
 * A fixed anchor vertices, and E edges in g2o insertion order
 * (EdgeSE3Range  src/types/types_edge_se3range.cpp:105-114,
 *  EdgeSE3Prior  localization.cpp:462-535, EdgeSE3 localization.cpp:438-459,560-605).
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no C++/torch types; nothing throws across the ABI
 *   - every function returns 0 on success or a negative UWBGO_E_* code;
 *     uwbgo_last_error() gives the text of the last failure on the calling thread
 *   - the caller owns every buffer it passes in; the library owns its device workspace
 *     inside the context; one context per GPU, used by one host thread at a time
 *   - all floating point is IEEE FP64, all indices int32, window-major C-contiguous arrays
 *   - there is NO CPU fallback: uwbgo_create() fails if no sm_100 device is present
 */
#ifndef UWBGO_H
#define UWBGO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UWBGO_ABI_VERSION 5

/* error codes */
#define UWBGO_OK              0
#define UWBGO_E_INVALID     (-1)  /* bad argument / inconsistent sizes            */
#define UWBGO_E_TOPOLOGY    (-2)  /* some pose has two different older neighbours  */
#define UWBGO_E_CUDA        (-3)  /* CUDA runtime error (text in last_error)      */
#define UWBGO_E_NODEVICE    (-4)  /* no usable sm_100 GPU                         */
#define UWBGO_E_NOMEM       (-5)

/* edge kinds (the four edge types Localization creates) */
#define UWBGO_EDGE_RANGE_ANCHOR 0 /* EdgeSE3Range(pose, fixed anchor)   localization.cpp:331,350 */
#define UWBGO_EDGE_RANGE_POSE   1 /* EdgeSE3Range(older pose, newer pose), measurement 0 localization.cpp:338 */
#define UWBGO_EDGE_PRIOR        2 /* EdgeSE3Prior on one pose (IMU / lidar)            localization.cpp:481,520 */
#define UWBGO_EDGE_SE3          3 /* EdgeSE3(older pose, newer pose): twist localization.cpp:588, pose :263 */

/*
 * Topology: the structure of the g2o graph, shared by every window of one batch
 * (all windows of a Monte-Carlo replay / parameter sweep / synthetic config have the same
 * structure; the host packer buckets windows by topology).  Edges are listed in g2o
 * insertion order, which is the order g2o accumulates them into H and b.
 *
 * Per-kind data slots are assigned by counting: the k-th RANGE_* edge of the list reads
 * range_d[w][k]; the k-th PRIOR edge reads prior_Z[w][k]; the k-th SE3 edge reads se3_Z[w][k].
 */
typedef struct uwbgo_topology {
    int32_t n_poses;      /* N  = robot/trajectory_length (localization.cpp:72)             */
    int32_t n_anchors;    /* A  = fixed Robot vertices (localization.cpp:97-98)             */
    int32_t n_antennas;   /* K  = entries of /uwb/antennaOffset (localization.cpp:111-123)  */
    int32_t n_edges;      /* E                                                              */
    const int32_t *edge_kind;   /* [E] UWBGO_EDGE_*                                          */
    const int32_t *edge_a;      /* [E] vertex 0: pose index 0..N-1 (0 = oldest)              */
    const int32_t *edge_b;      /* [E] vertex 1: anchor index (RANGE_ANCHOR); a NEWER pose    */
                                /*     index > edge_a (RANGE_POSE, SE3); ignored (PRIOR).     */
                                /*     Every pose may have at most ONE older pose neighbour   */
                                /*     (chains: its predecessor; addPoseEdge: its key vertex, */
                                /*     localization.cpp:258-267), else UWBGO_E_TOPOLOGY       */
    const int32_t *edge_ant;    /* [E] RANGE_*: antenna number of the vertex-0 offset,       */
                                /*     0 = identity, k>0 = ant_offsets[k-1] (localization.cpp:333). */
                                /*     Only the translation of an offset isometry reaches the      */
                                /*     residual: (X * O).translation() = R_X t_O + t_X             */
    const int32_t *edge_robust; /* [E] 1 = RobustKernelCauchy (delta 1), 0 = no kernel       */
    const int32_t *edge_ant_b;  /* [E] RANGE_*: antenna number of the vertex-1 offset, same      */
                                /*     numbering as edge_ant: offset[1] of EdgeSE3Range           */
                                /*     (types_edge_se3range.cpp:99-114), pidTo of                  */
                                /*     EdgeSE3RangeOffset (types_edge_se3range_offset.cpp:126-149). */
                                /*     Anchors are identity-rotation vertices, so on RANGE_ANCHOR  */
                                /*     the offset adds to the anchor position.  NULL = all 0       */
                                /*     (what Localization creates, localization.cpp:333)           */
} uwbgo_topology;

/*
 * Compact form of the range data: the FIELDS OF THE RANGE MESSAGES instead of the edge parameters
 * (optional, uwbgo_batch::range_msgs).  The edge parameters are then built on the device with the
 * arithmetic of Localization::addRangeEdge / create_range_edge (localization.cpp:316-319,331,338,350,
 * 608-627), bit for bit -- all in FP64 after an exact widening of the float32 message fields:
 *   k-th RANGE_ANCHOR edge:  measurement = distance[w][k]
 *                            cov = distance_err[w][k]^2                        (separate vertex, :318,331)
 *                                  + (v_max * dt_anchor[w][k] / 3)^2           (merged branch :350; only
 *                                                                              when dt_anchor != NULL)
 *   k-th RANGE_POSE edge:    measurement = 0,  cov = (v_max * dt_pose[w][k] / 3)^2     (:319,338)
 *   information = 1 / cov.   x^2 is x * x (what pow(x, 2) compiles to), v_max * dt / 3 = (v_max * dt) / 3.
 * Era / Erp = number of RANGE_ANCHOR / RANGE_POSE edges, each kind counted in insertion order.
 * Half the host->device bytes of range_d + range_info on the windows addRangeEdge builds.  Batches in
 * this form take the tile kernels (not the WINDOW path).
 */
typedef struct uwbgo_range_msgs {
    const float  *distance;      /* [W][Era]  UwbRange::distance     (float32 on the wire)        */
    const float  *distance_err;  /* [W][Era]  UwbRange::distance_err (float32 on the wire)        */
    const double *dt_anchor;     /* [W][Era]  stamp - requester's last stamp, or NULL             */
    const double *dt_pose;       /* [W][Erp]  stamp - the robot's last stamp; NULL iff Erp == 0   */
    double        v_max;         /* robot/maximum_velocity (localization.cpp:76)                  */
} uwbgo_range_msgs;

#define UWBGO_SHARED_ANCHORS 1   /* uwbgo_batch::shared: `anchors` is [A][3], one constellation for */
                                 /* every window of the batch (a fleet in one anchor field)         */
#define UWBGO_DIAG_INFO      2   /* uwbgo_batch::shared: prior_info is [W][Ep][6] and se3_info is   */
                                 /* [W][Es][6]: the DIAGONALS of the information matrices.  What    */
                                 /* Localization builds is diagonal (lidar: information(2,2),       */
                                 /* localization.cpp:478-479; IMU: the three rotation entries,      */
                                 /* :515-518); the 6x6 matrices are rebuilt on the device with +0.0 */
                                 /* elsewhere, so the solve sees the same bits at a sixth of the    */
                                 /* host->device bytes (C4a: 9.0 of 17.2 KB per window less).       */
                                 /* Batches in this form take the tile kernels                      */

/* Per-window numeric data, window-major.  Er/Ep/Es = number of RANGE_*, PRIOR, SE3 edges. */
typedef struct uwbgo_batch {
    int64_t n_windows;            /* W */
    const double  *pose_t;        /* [W][N][3]  initial translations                              */
    const double  *pose_R;        /* [W][N][9]  initial rotations, row-major; NULL = identity     */
    const int32_t *oplus_count;   /* [W][N]     VertexSE3::_numOplusCalls carried in; NULL = 0    */
    const double  *anchors;       /* [W][A][3]  fixed anchor positions                            */
    const double  *ant_offsets;   /* [K][3]     antenna lever arms (shared); NULL iff K == 0      */
    const double  *range_d;       /* [W][Er]    EdgeSE3Range measurement                          */
    const double  *range_info;    /* [W][Er]    EdgeSE3Range information (1x1)                    */
    const double  *prior_Z;       /* [W][Ep][12] EdgeSE3Prior measurement: R row-major (9), t (3) */
    const double  *prior_info;    /* [W][Ep][36] information, row-major ([W][Ep][6] with UWBGO_DIAG_INFO) */
    const double  *se3_Z;         /* [W][Es][12] EdgeSE3 measurement                              */
    const double  *se3_info;      /* [W][Es][36] information, row-major ([W][Es][6] with UWBGO_DIAG_INFO) */
    /* optional compact forms; zero / NULL = the arrays above are complete */
    const uwbgo_range_msgs *range_msgs; /* replaces range_d and range_info (both must then be NULL)  */
    int32_t shared;               /* bit mask of UWBGO_SHARED_*                                   */
    int32_t reserved;
} uwbgo_batch;

/* Solver constants; uwbgo_config_default() fills g2o's defaults at the pinned commit. */
typedef struct uwbgo_config {
    int32_t max_iterations;       /* optimizer/maximum_iteration (localization.cpp:65), default 20 */
    int32_t max_trials;           /* LM maxTrialsAfterFailure, 10                                  */
    int32_t orthogonalize_after;  /* VertexSE3::orthogonalizeAfter, 1000                           */
    int32_t reserved;
    double  tau;                  /* LM initial-lambda factor, 1e-5                                */
    double  good_step_lower;      /* 1/3                                                           */
    double  good_step_upper;      /* 2/3                                                           */
    double  kernel_delta;         /* RobustKernelCauchy delta, 1.0                                 */
    double  jacobian_delta;       /* numeric-Jacobian step of BaseBinaryEdge, 1e-9                 */
} uwbgo_config;

#define UWBGO_CHI2_STRIDE   4   /* chi2[w] = {plain chi2 at final estimate, robust chi2 at final   */
                                /*  estimate, g2o optimizer.chi2() as publish() reads it           */
                                /*  (localization.cpp:197: errors of the LAST trial, maybe         */
                                /*  rejected), final lambda}                                       */
#define UWBGO_STATUS_STRIDE 4   /* status[w] = {iterations run, total LM trials, flags, trials of  */
                                /*  the last iteration}                                            */
#define UWBGO_FLAG_CHOL_FAIL   1  /* some trial hit a non-positive pivot (trial rejected)          */
#define UWBGO_FLAG_TERMINATED  2  /* LM returned Terminate (max trials reached or rho == 0)        */
#define UWBGO_FLAG_NONFINITE   4  /* a trial produced a non-finite robust chi2                     */
#define UWBGO_FLAG_REJECTED    8  /* uwbgo_stream only: the robot's message was refused by the     */
                                  /* outlier gate, nothing was solved (localization.cpp:309-313)   */

typedef struct uwbgo_result {
    double  *pose_t;       /* [W][N][3]                                   */
    double  *pose_R;       /* [W][N][9]; may be NULL                      */
    int32_t *oplus_count;  /* [W][N];    may be NULL                      */
    double  *chi2;         /* [W][UWBGO_CHI2_STRIDE]                      */
    int32_t *status;       /* [W][UWBGO_STATUS_STRIDE]                    */
    /* Optional extras: what the commented-out tail of Localization::solve() would read
     * (localization.cpp:172-189).  NULL = not wanted.  Batches that ask for them take the tile kernels. */
    double  *edge_chi2;    /* [W][E]  edge->chi2() of every edge after optimize(), insertion order: the */
                           /*         errors of the LAST trial (accepted or not), i.e. the terms of    */
                           /*         chi2[w][2] -- their sum in edge order is chi2[w][2], bit for bit  */
                           /*         (the outlier pruning of localization.cpp:172-181 tests these)    */
    double  *marginal;     /* [W][36] computeMarginals(spinv, last_vertex) (localization.cpp:185-189):  */
                           /*         the 6x6 block of H^-1 that belongs to the NEWEST pose, row-major, */
                           /*         H = the system of the last buildSystem() without damping          */
    int32_t *marginal_ok;  /* [W]     1 = computed; 0 = H is not positive definite ("can't compute":    */
                           /*         e.g. UWB-only windows, whose rotations are unobserved); the        */
                           /*         marginal is then all NaN.  May be NULL                             */
} uwbgo_result;

typedef struct uwbgo_ctx uwbgo_ctx;

/* ---- lifetime ------------------------------------------------------------------------- */
int  uwbgo_abi_version(void);
void uwbgo_config_default(uwbgo_config *cfg);
/* Replaces the solver chain built in Localization::Localization (localization.cpp:44-52). */
int  uwbgo_create(int device, uwbgo_ctx **out);
void uwbgo_destroy(uwbgo_ctx *ctx);
const char *uwbgo_last_error(void);

/* Tuning of the host-pointer entry points: windows per pipeline chunk (rounded up to 32) and
 * number of concurrent stream lanes (1..16).  Defaults: 8192 windows, 8 lanes. */
int  uwbgo_set_pipeline(uwbgo_ctx *ctx, int64_t windows_per_chunk, int n_lanes);
/* Small batches -- the reference's own call pattern is ONE window per range message
 * (localization.cpp:371-375) -- take the WINDOW path: one CTA per window, the window's state in
 * shared memory, the arrays of uwbgo_batch read in place, one kernel launch per call.  Batches of up
 * to `max_windows` windows go that way (default 592; 0 switches the path off, a negative value
 * restores the default); larger ones, and windows whose state does not fit the shared memory of
 * one SM, take the tile kernels.  Results are the same bits either way. */
int  uwbgo_set_window_path(uwbgo_ctx *ctx, int64_t max_windows);
/* Page-locked host memory.  The host-pointer entry points accept any host memory; with buffers
 * from uwbgo_host_alloc their copies overlap the kernels of neighbouring chunks. */
void *uwbgo_host_alloc(size_t bytes);
void  uwbgo_host_free(void *p);

/* ---- the hot path ----------------------------------------------------------------------
 * Replaces optimizer.initializeOptimization(); optimizer.optimize(iteration_max)
 * (localization.cpp:168-170) for W windows.  HOST buffers in, HOST buffers out; the
 * host<->device copies are inside the call (pipelined in chunks over CUDA streams). */
int uwbgo_solve_batch(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                      const uwbgo_config *cfg, uwbgo_result *out);

/* Same, but every pointer inside `in` (except the struct itself and ant_offsets, which is
 * host memory) and `out` is a DEVICE pointer on the context's GPU, and the work is queued on
 * `stream` (a cudaStream_t; NULL = the default stream) without synchronising (exception: a call whose
 * ant_offsets differ from the previous call's waits for the device once, to replace the table). */
int uwbgo_solve_batch_device(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                             const uwbgo_config *cfg, uwbgo_result *out, void *stream);

/* ---- stages, exposed for parity tests and roofline measurement -------------------------
 * One linearisation at the given estimates: computeActiveErrors + buildSystem
 * (g2o BlockSolver; numeric Jacobians of BaseBinaryEdge for the range edges).
 *   H_diag [W][N][36]   diagonal 6x6 blocks, row-major, both triangles
 *   H_off  [W][N-1][36] H_off[j] = block (parent(j+1), j+1): rows of the older neighbour of
 *                       pose j+1 (chains: pose j), columns of pose j+1
 *   b      [W][N][6]
 *   chi2   [W][2]       {plain, robust} at the linearisation point
 * host pointers. */
int uwbgo_linearize_batch(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                          const uwbgo_config *cfg, double *H_diag, double *H_off, double *b,
                          double *chi2);
/* device-pointer variant, queued on `stream`.  Any alignment of the output arrays is accepted; with
 * 16-byte aligned H_diag / H_off / b the windows addRangeEdge builds take the one-kernel path. */
int uwbgo_linearize_batch_device(uwbgo_ctx *ctx, const uwbgo_topology *topo,
                                 const uwbgo_batch *in, const uwbgo_config *cfg, double *H_diag,
                                 double *H_off, double *b, double *chi2, void *stream);

/* Solve (H + lambda I) x = b for W block-tridiagonal systems (replaces
 * LinearSolverCholmod::solve, localization.h:84).  ok[w] = 0 when a pivot was not positive
 * (x[w] is then all zero).  Host pointers. */
int uwbgo_factor_solve_batch(uwbgo_ctx *ctx, int32_t n_poses, int64_t n_windows,
                             const double *H_diag, const double *H_off, const double *b,
                             const double *lambda, double *x, int32_t *ok);
int uwbgo_factor_solve_batch_device(uwbgo_ctx *ctx, int32_t n_poses, int64_t n_windows,
                                    const double *H_diag, const double *H_off, const double *b,
                                    const double *lambda, double *x, int32_t *ok, void *stream);

/* ---- resident fleet: the sliding windows of W robots stay on the device ---------------------- */
/*
 * The reference's own call pattern at fleet scale.  Localization::addRangeEdge (localization.cpp:297-376)
 * handles ONE range message per robot: a new vertex whose estimate is a copy of the newest one, the range
 * edge to the anchor (distance, cov = distance_err^2, :316-331), the zero-length trajectory edge to the
 * previous vertex (cov = (v_max dt / 3)^2, :319,338), the oldest vertex and its edges dropped from the
 * ring (robot.cpp:75-110), then solve() (:371-375).  A uwbgo_stream keeps the N-pose windows of W robots
 * in HBM between messages -- the estimates of the last solve and the message fields of the N range
 * edges -- so that a step moves only the new message to the device (16 bytes per robot) and the newest
 * pose, chi2 and status back (72 bytes), instead of whole windows (C3: 1,992 + 1,248 bytes).  The solve is
 * the one of uwbgo_solve_batch_device on the window the reference would hold: the same bits (tested
 * against windows shifted on the host).  Anchors are one constellation for the fleet.  Either all robots
 * range the same anchor in a step (one TDMA slot per anchor: _load / _step) or every robot has its own
 * anchor sequence (_load_robots / _step_robots: the anchor id is one more message field, 4 bytes per robot;
 * on the device the ids move with the poses and each window reads its own N anchor positions, so the graph
 * structure is one for the whole fleet and for every step).  Fleets of up to 592 robots are solved by the WINDOW
 * kernels (one CTA per robot), larger ones by the tile kernels.  UWB-only windows (uwb_only.yaml).
 */
typedef struct uwbgo_stream uwbgo_stream;
int  uwbgo_stream_create(uwbgo_ctx *ctx, int32_t n_poses, int32_t n_anchors, int64_t n_windows,
                         const double *anchors /* [A][3] */, double v_max, const uwbgo_config *cfg,
                         uwbgo_stream **out);
void uwbgo_stream_destroy(uwbgo_stream *s);
/* the windows as they stand (host arrays): estimates [W][N][3], the anchor of each pose's range edge [N],
 * distance / distance_err [W][N] (float32, as on the wire), stamp differences dt [W][N-1] */
int  uwbgo_stream_load(uwbgo_stream *s, const double *pose_t, const int32_t *anchor_of_pose,
                       const float *distance, const float *distance_err, const double *dt);
/* one range message per robot (host arrays [W]) from anchor `anchor`: shift the windows, append the new
 * vertex and its two edges, optimise, return the newest pose [W][3], chi2 [W][4] and status [W][4] (as
 * uwbgo_result; any may be NULL).  Returns when the results are in the host arrays.  Page-locked arrays
 * (uwbgo_host_alloc) are copied without staging. */
int  uwbgo_stream_step(uwbgo_stream *s, int32_t anchor, const float *distance, const float *distance_err,
                       const double *dt, double *newest_pose, double *chi2, int32_t *status);
/* the same with one anchor sequence PER ROBOT: anchor_of_pose is [W][N], a step's anchor ids are [W]
 * (UwbRange::responder_id mapped to the anchor's row, localization.cpp:305-306,331).  A stream is stepped the way
 * it was loaded (else UWBGO_E_INVALID); ids outside [0, A) are refused before anything is queued. */
int  uwbgo_stream_load_robots(uwbgo_stream *s, const double *pose_t, const int32_t *anchor_of_pose,
                              const float *distance, const float *distance_err, const double *dt);
int  uwbgo_stream_step_robots(uwbgo_stream *s, const int32_t *anchor, const float *distance,
                              const float *distance_err, const double *dt, double *newest_pose, double *chi2,
                              int32_t *status);
/* The outlier gate of addRangeEdge (localization.cpp:305-313; robot/distance_outlier, :78; the stream holds full
 * windows, so the gate's "window has filled" condition is met): a message whose range differs from the distance
 * between the robot's newest estimate and the anchor by more than distance_outlier is refused -- that robot's
 * window stays as it is and nothing is solved for it, as in the reference; the step returns its unchanged newest
 * pose, the chi2 of its last accepted message (zeros before the first) and status {0, 0, UWBGO_FLAG_REJECTED, 0}.
 * The caller keeps the stamp of a robot's last ACCEPTED message for the next dt, as Robot::last_header() does.
 * distance_outlier < 0 switches the gate off (the default).  Per-robot streams only (_load_robots): a refused
 * message would desynchronise a fleet-wide anchor sequence, so _step answers UWBGO_E_INVALID while the gate is on. */
int  uwbgo_stream_set_outlier_gate(uwbgo_stream *s, double distance_outlier);
/* all estimates of the windows as the last step left them, [W][N][3] (host array) */
int  uwbgo_stream_read(uwbgo_stream *s, double *pose_t);
/* kernels launched / duration of the LM kernel of the last step (ms; needs uwbgo_set_profiling) come
 * from the context: uwbgo_launch_count, uwbgo_last_kernel_ms */

/* ---- introspection ---------------------------------------------------------------------- */
/* number of kernels this library launched on ctx since creation (bench "gpu_launches") */
int64_t uwbgo_launch_count(const uwbgo_ctx *ctx);
/* path of the last solve on ctx: 0 = 6x6 tile kernel, 1 = translation-only tile kernel (R = I, zero
 * offsets), 2 = its straight-line instantiation for the window addRangeEdge builds, 3 = WINDOW path */
int     uwbgo_last_path(const uwbgo_ctx *ctx);
/* Kernel timing for roofline reports: with profiling on, every *_device call records CUDA events
 * on its stream immediately before and after its main kernel (the fused LM kernel of
 * uwbgo_solve_batch_device; the linearisation kernel plus the expansion of H to the public
 * full-block layout for uwbgo_linearize_batch_device);
 * uwbgo_last_kernel_ms waits for the last such kernel and returns its duration (-1 if none);
 * uwbgo_mean_kernel_ms averages the durations of the last `last_n` (<= 64) such calls. */
int     uwbgo_set_profiling(uwbgo_ctx *ctx, int on);
double  uwbgo_last_kernel_ms(uwbgo_ctx *ctx);
double  uwbgo_mean_kernel_ms(uwbgo_ctx *ctx, int last_n);
/* FP64 FMA micro-benchmark on the context's GPU: returns achieved FP64 FLOP/s (2 per FMA),
 * used as the measured FP64 roofline denominator (MEASURED_PEAKS.json has none). */
double  uwbgo_measure_fp64_peak(uwbgo_ctx *ctx, double *elapsed_ms);
/* Arithmetic self-test of the branch-free sqrt / reciprocal / division / logarithm sequences the
 * CHAIN kernels use (uwbgo_math.cuh, NbMath) against the IEEE operations, on about n_operands
 * pseudo-random operands.  mode 0: operands over the whole binary64 encoding space; mode 1:
 * magnitudes 2^-60..2^60.  counts[0] = results compared or flagged, counts[1] = results that
 * differ in bits from IEEE although the sequence did not flag its operand (must be 0),
 * counts[2] = operands flagged as outside the safe range (the solver re-runs those trials with
 * the IEEE operations). */
int     uwbgo_selftest_math(uwbgo_ctx *ctx, uint64_t seed, int64_t n_operands, int mode, int64_t counts[3]);

#ifdef __cplusplus
}
#endif
#endif /* UWBGO_H */
