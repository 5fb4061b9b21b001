/* uwbgo_stream: the sliding windows of a fleet resident in HBM (include/uwbgo.h, "resident fleet").
 *
 * The host-side caller of the hot path in the reference is Localization::addRangeEdge
 * (localization.cpp:297-376) on the Robot ring (robot.cpp:75-110): per range message a new vertex (estimate =
 * copy of the newest), two edges, the oldest vertex dropped, solve().  Here that bookkeeping runs on the
 * device for W robots at once: one small kernel shifts the estimates and the message fields by one pose and
 * appends the new message, the solve is uwbgo_solve_batch_device on the compact range form (edge parameters
 * built on the device, one anchor constellation), one small kernel gathers the newest poses.  Only the new
 * message goes in and the newest pose, chi2 and status come out.
 *
 * Layout: window-major device arrays (the layout of the ABI), two copies of the message fields (a step reads
 * one and writes the other), estimates in `res` (what the last solve left) and `in` (what the next one
 * starts from). */
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/uwbgo.h"

namespace {

__global__ void __launch_bounds__(256)
stream_shift_kernel(int64_t W, int N, const double *__restrict__ res, double *__restrict__ in,
                    const float *__restrict__ d_old, const float *__restrict__ e_old, const double *__restrict__ dt_old,
                    float *__restrict__ d_new, float *__restrict__ e_new, double *__restrict__ dt_new,
                    const float *__restrict__ msg_d, const float *__restrict__ msg_e, const double *__restrict__ msg_dt,
                    const int32_t *__restrict__ a_old, int32_t *__restrict__ a_new, const int32_t *__restrict__ msg_a,
                    const double *__restrict__ anchors, double *__restrict__ anch_w, const int32_t *__restrict__ rej)
{
    /* one thread per (window, pose): pose i of the new window is pose i + 1 of the old one; the new vertex
     * starts at the estimate of the newest (robot.cpp: new_vertex copies the last estimate) */
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * N) return;
    const int64_t w = t / N;
    const int i = (int)(t - w * N);
    if (rej && rej[w]) {
        /* message refused by the outlier gate (localization.cpp:309-313 returns before touching the graph): the
         * window stays as it is; stream_finish_kernel puts its estimates back after the batch solve */
#pragma unroll
        for (int k = 0; k < 3; ++k) in[t * 3 + k] = res[t * 3 + k];
        d_new[t] = d_old[t];
        e_new[t] = e_old[t];
        if (i + 1 < N) dt_new[w * (N - 1) + i] = dt_old[w * (N - 1) + i];
        const int32_t a = a_old[t];
        a_new[t] = a;
#pragma unroll
        for (int k = 0; k < 3; ++k) anch_w[t * 3 + k] = anchors[(size_t)a * 3 + k];
        return;
    }
    const int src = i + 1 < N ? i + 1 : N - 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) in[t * 3 + k] = res[(w * N + src) * 3 + k];
    d_new[t] = i + 1 < N ? d_old[w * N + i + 1] : msg_d[w];
    e_new[t] = i + 1 < N ? e_old[w * N + i + 1] : msg_e[w];
    if (i + 1 < N) dt_new[w * (N - 1) + i] = i + 2 < N ? dt_old[w * (N - 1) + i + 1] : msg_dt[w];
    if (a_new) {
        /* per-robot anchor sequences: the anchor id of every range edge moves with its pose, and the solve
         * reads the anchor of pose i of window w at anch_w[w][i] (a window-private "constellation" of N) */
        const int32_t a = i + 1 < N ? a_old[w * N + i + 1] : msg_a[w];
        a_new[t] = a;
#pragma unroll
        for (int k = 0; k < 3; ++k) anch_w[t * 3 + k] = anchors[(size_t)a * 3 + k];
    }
}

/* Small fleets: the edge parameters window-major, in the expanded form of the ABI (range_d / range_info [W][2N-1],
 * slots in insertion order: anchor edge of pose 0, then per pose k its anchor edge and the trajectory edge
 * (k-1, k)), so that the solve can take the WINDOW kernels (one CTA per robot).  The arithmetic of
 * Localization::addRangeEdge / create_range_edge as uwbgo.h states it for uwbgo_range_msgs: cov = err * err resp.
 * ((v_max dt) / 3)^2, information = 1 / cov, all in FP64 after the exact widening of the float32 fields */
__global__ void __launch_bounds__(256)
stream_expand_kernel(int64_t W, int N, const float *__restrict__ d, const float *__restrict__ e, const double *__restrict__ dt,
                     double v_max, double *__restrict__ rd, double *__restrict__ ri)
{
    const int Er = 2 * N - 1;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W * Er) return;
    const int64_t w = t / Er;
    const int s = (int)(t - w * Er);
    double meas, cov;
    if (s == 0 || (s & 1)) { /* anchor edge of pose k */
        const int k = (s + 1) >> 1;
        const double err = (double)e[w * N + k];
        meas = (double)d[w * N + k];
        cov = err * err;
    } else { /* trajectory edge (k - 1, k) */
        const int k = s >> 1;
        const double x = (v_max * dt[w * (N - 1) + (k - 1)]) / 3.0;
        meas = 0.0;
        cov = x * x;
    }
    rd[t] = meas;
    ri[t] = 1.0 / cov;
}

/* the outlier gate of Localization::addRangeEdge (localization.cpp:305-313): distance between the newest estimate
 * of the robot and the anchor (Eigen's norm of a 3-vector: sqrt((x^2 + y^2) + z^2)) against the measured range
 * (float32 on the wire, widened), refused when they differ by more than robot/distance_outlier */
__global__ void __launch_bounds__(256)
stream_gate_kernel(int64_t W, int N, const double *__restrict__ res, const double *__restrict__ anchors,
                   const int32_t *__restrict__ msg_a, const float *__restrict__ msg_d, double outlier, int32_t *__restrict__ rej)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const double *p = res + (w * N + (N - 1)) * 3, *q = anchors + (size_t)msg_a[w] * 3;
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    const double est = sqrt(dx * dx + dy * dy + dz * dz);
    rej[w] = fabs(est - (double)msg_d[w]) > outlier ? 1 : 0;
}

/* after the solve: the newest pose of every robot; a robot whose message was refused gets its window back (the
 * batch solve ran on it and is discarded: the reference does not call solve() for a refused message), the chi2 of
 * its last accepted message and status = {0, 0, UWBGO_FLAG_REJECTED, 0} */
__global__ void __launch_bounds__(256)
stream_finish_kernel(int64_t W, int N, double *__restrict__ res, const double *__restrict__ in, double *__restrict__ newest,
                     const int32_t *__restrict__ rej, double *__restrict__ chi2, double *__restrict__ chi2_prev,
                     int32_t *__restrict__ status)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    if (rej) {
        if (rej[w]) {
            for (int i = 0; i < N * 3; ++i) res[w * N * 3 + i] = in[w * N * 3 + i];
            for (int k = 0; k < UWBGO_CHI2_STRIDE; ++k) chi2[w * UWBGO_CHI2_STRIDE + k] = chi2_prev[w * UWBGO_CHI2_STRIDE + k];
            status[w * UWBGO_STATUS_STRIDE + 0] = 0;
            status[w * UWBGO_STATUS_STRIDE + 1] = 0;
            status[w * UWBGO_STATUS_STRIDE + 2] = UWBGO_FLAG_REJECTED;
            status[w * UWBGO_STATUS_STRIDE + 3] = 0;
        } else {
            for (int k = 0; k < UWBGO_CHI2_STRIDE; ++k) chi2_prev[w * UWBGO_CHI2_STRIDE + k] = chi2[w * UWBGO_CHI2_STRIDE + k];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) newest[w * 3 + k] = res[(w * N + (N - 1)) * 3 + k];
}

}  // namespace

struct uwbgo_stream {
    uwbgo_ctx *ctx = nullptr;
    int N = 0, A = 0;
    int64_t W = 0;
    double v_max = 0.0;
    uwbgo_config cfg{};
    std::vector<int32_t> anchor_of_pose;
    std::vector<int32_t> ek, ea, eb, eant, erob; /* topology arrays of the current window */
    char *dev = nullptr;                          /* one device allocation */
    double *res = nullptr, *in = nullptr, *dt[2] = {nullptr, nullptr}, *anchors = nullptr, *chi2 = nullptr, *newest = nullptr, *msg_dt = nullptr;
    float *d[2] = {nullptr, nullptr}, *e[2] = {nullptr, nullptr}, *msg_d = nullptr, *msg_e = nullptr;
    int32_t *status = nullptr;
    /* per-robot anchor sequences (uwbgo_stream_load_robots / _step_robots): one more allocation, made on first use */
    char *dev_r = nullptr;
    int32_t *aid[2] = {nullptr, nullptr}, *msg_a = nullptr;
    double *anch_w = nullptr, *chi2_prev = nullptr, *rd = nullptr, *ri = nullptr; /* rd / ri: small fleets only */
    int32_t *rej = nullptr;
    bool per_robot = false;
    bool api_per_robot = false; /* how the caller loaded it: a small fleet-wide stream is per_robot inside */
    bool gate = false;      /* outlier gate of addRangeEdge (uwbgo_stream_set_outlier_gate) */
    double outlier = 0.0;   /* robot/distance_outlier */
    int cur = 0;
    bool loaded = false;
    cudaStream_t st = nullptr;
};

namespace {

#define SCU(call)                                                         \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) {                                         \
            fprintf(stderr, "uwbgo_stream: %s: %s\n", #call, cudaGetErrorString(e__)); \
            return UWBGO_E_CUDA;                                          \
        }                                                                 \
    } while (0)

size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

/* per-robot fleets of up to this many robots hand the solve the expanded window-major form, which the WINDOW
 * kernels read (the library's default batch limit of that path, uwbgo_set_window_path) */
constexpr int64_t SMALL_FLEET = 592;

void build_topology(uwbgo_stream *s)
{
    /* the window addRangeEdge builds (localization.cpp:331-340): per pose its anchor edge, then the trajectory
     * edge to its predecessor; Cauchy kernel on both */
    s->ek.clear(); s->ea.clear(); s->eb.clear(); s->eant.clear(); s->erob.clear();
    for (int k = 0; k < s->N; ++k) {
        s->ek.push_back(UWBGO_EDGE_RANGE_ANCHOR); s->ea.push_back(k); s->eb.push_back(s->anchor_of_pose[k]);
        s->eant.push_back(0); s->erob.push_back(1);
        if (k > 0) {
            s->ek.push_back(UWBGO_EDGE_RANGE_POSE); s->ea.push_back(k - 1); s->eb.push_back(k);
            s->eant.push_back(0); s->erob.push_back(1);
        }
    }
}

}  // namespace

extern "C" {

int uwbgo_stream_create(uwbgo_ctx *ctx, int32_t n_poses, int32_t n_anchors, int64_t n_windows, const double *anchors,
                        double v_max, const uwbgo_config *cfg, uwbgo_stream **out)
{
    if (!out) return UWBGO_E_INVALID;
    *out = nullptr;
    if (!ctx || !anchors || !cfg || n_poses < 2 || n_anchors < 1 || n_windows < 1) return UWBGO_E_INVALID;
    auto s = new uwbgo_stream();
    s->ctx = ctx;
    s->N = n_poses;
    s->A = n_anchors;
    s->W = n_windows;
    s->v_max = v_max;
    s->cfg = *cfg;
    const size_t W = (size_t)n_windows, N = (size_t)n_poses;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += up256(bytes); return at; };
    const size_t o_res = take(W * N * 24), o_in = take(W * N * 24), o_dt0 = take(W * (N - 1) * 8), o_dt1 = take(W * (N - 1) * 8),
                 o_d0 = take(W * N * 4), o_d1 = take(W * N * 4), o_e0 = take(W * N * 4), o_e1 = take(W * N * 4),
                 o_anch = take((size_t)n_anchors * 24), o_chi = take(W * 32), o_stat = take(W * 16), o_new = take(W * 24),
                 o_md = take(W * 4), o_me = take(W * 4), o_mdt = take(W * 8);
    if (cudaMalloc(&s->dev, o) != cudaSuccess) {
        cudaGetLastError();
        delete s;
        return UWBGO_E_NOMEM;
    }
    s->res = reinterpret_cast<double *>(s->dev + o_res);
    s->in = reinterpret_cast<double *>(s->dev + o_in);
    s->dt[0] = reinterpret_cast<double *>(s->dev + o_dt0);
    s->dt[1] = reinterpret_cast<double *>(s->dev + o_dt1);
    s->d[0] = reinterpret_cast<float *>(s->dev + o_d0);
    s->d[1] = reinterpret_cast<float *>(s->dev + o_d1);
    s->e[0] = reinterpret_cast<float *>(s->dev + o_e0);
    s->e[1] = reinterpret_cast<float *>(s->dev + o_e1);
    s->anchors = reinterpret_cast<double *>(s->dev + o_anch);
    s->chi2 = reinterpret_cast<double *>(s->dev + o_chi);
    s->status = reinterpret_cast<int32_t *>(s->dev + o_stat);
    s->newest = reinterpret_cast<double *>(s->dev + o_new);
    s->msg_d = reinterpret_cast<float *>(s->dev + o_md);
    s->msg_e = reinterpret_cast<float *>(s->dev + o_me);
    s->msg_dt = reinterpret_cast<double *>(s->dev + o_mdt);
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMemcpy(s->anchors, anchors, (size_t)n_anchors * 24, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        uwbgo_stream_destroy(s);
        return UWBGO_E_CUDA;
    }
    *out = s;
    return 0;
}

void uwbgo_stream_destroy(uwbgo_stream *s)
{
    if (!s) return;
    if (s->st) {
        cudaStreamSynchronize(s->st);
        cudaStreamDestroy(s->st);
    }
    if (s->dev) cudaFree(s->dev);
    if (s->dev_r) cudaFree(s->dev_r);
    delete s;
}

}  // extern "C"

namespace {

int ensure_robot_arrays(uwbgo_stream *s)
{
    if (s->dev_r) return 0;
    const size_t W = (size_t)s->W, N = (size_t)s->N;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += up256(bytes); return at; };
    const size_t o_a0 = take(W * N * 4), o_a1 = take(W * N * 4), o_ma = take(W * 4), o_aw = take(W * N * 24),
                 o_rej = take(W * 4), o_cp = take(W * 32);
    const bool small = (int64_t)W <= SMALL_FLEET;
    const size_t o_rd = take(small ? W * (2 * N - 1) * 8 : 0), o_ri = take(small ? W * (2 * N - 1) * 8 : 0);
    if (cudaMalloc(&s->dev_r, o) != cudaSuccess) {
        cudaGetLastError();
        s->dev_r = nullptr;
        return UWBGO_E_NOMEM;
    }
    s->aid[0] = reinterpret_cast<int32_t *>(s->dev_r + o_a0);
    s->aid[1] = reinterpret_cast<int32_t *>(s->dev_r + o_a1);
    s->msg_a = reinterpret_cast<int32_t *>(s->dev_r + o_ma);
    s->anch_w = reinterpret_cast<double *>(s->dev_r + o_aw);
    s->rej = reinterpret_cast<int32_t *>(s->dev_r + o_rej);
    s->chi2_prev = reinterpret_cast<double *>(s->dev_r + o_cp);
    if (small) {
        s->rd = reinterpret_cast<double *>(s->dev_r + o_rd);
        s->ri = reinterpret_cast<double *>(s->dev_r + o_ri);
    }
    return 0;
}

/* anchor_of_pose: [N] (one sequence for the fleet) or, per_robot, [W][N] */
int load_impl(uwbgo_stream *s, bool per_robot, const double *pose_t, const int32_t *anchor_of_pose, const float *distance,
              const float *distance_err, const double *dt)
{
    if (!s || !pose_t || !anchor_of_pose || !distance || !distance_err || !dt) return UWBGO_E_INVALID;
    const size_t W = (size_t)s->W, N = (size_t)s->N;
    const size_t n_ids = per_robot ? W * N : N;
    for (size_t k = 0; k < n_ids; ++k)
        if (anchor_of_pose[k] < 0 || anchor_of_pose[k] >= s->A) return UWBGO_E_INVALID;
    if (per_robot) {
        int rc = ensure_robot_arrays(s);
        if (rc) return rc;
    }
    SCU(cudaStreamSynchronize(s->st));
    s->loaded = false;
    s->per_robot = per_robot;
    s->cur = 0;
    if (per_robot) {
        /* every window brings its own anchors: pose k reads "anchor k" of its window (anch_w[w][k]), so the
         * structure of the graph is the same for every robot and every step */
        s->anchor_of_pose.resize(N);
        for (size_t k = 0; k < N; ++k) s->anchor_of_pose[k] = (int32_t)k;
        SCU(cudaMemcpy(s->aid[0], anchor_of_pose, W * N * 4, cudaMemcpyHostToDevice));
        SCU(cudaMemset(s->chi2_prev, 0, W * 32)); /* what a robot reports if its very first message is refused */
    } else {
        s->anchor_of_pose.assign(anchor_of_pose, anchor_of_pose + N);
    }
    SCU(cudaMemcpy(s->res, pose_t, W * N * 24, cudaMemcpyHostToDevice));
    SCU(cudaMemcpy(s->d[0], distance, W * N * 4, cudaMemcpyHostToDevice));
    SCU(cudaMemcpy(s->e[0], distance_err, W * N * 4, cudaMemcpyHostToDevice));
    SCU(cudaMemcpy(s->dt[0], dt, W * (N - 1) * 8, cudaMemcpyHostToDevice));
    s->loaded = true;
    return 0;
}

/* anchor_w != NULL: one anchor id per robot (per_robot streams); else `anchor` for the whole fleet */
int step_impl(uwbgo_stream *s, int32_t anchor, const int32_t *anchor_w, const float *distance, const float *distance_err,
              const double *dt, double *newest_pose, double *chi2, int32_t *status)
{
    if (!s || !s->loaded || !distance || !distance_err || !dt) return UWBGO_E_INVALID;
    const bool per_robot = anchor_w != nullptr;
    if (per_robot != s->per_robot) return UWBGO_E_INVALID; /* the step must match the load */
    if (s->gate && !per_robot) return UWBGO_E_INVALID;     /* a refused message desynchronises a fleet-wide anchor sequence */
    const int32_t *rej = s->gate ? s->rej : nullptr;
    const size_t W = (size_t)s->W;
    const int N = s->N;
    if (per_robot) {
        for (size_t w = 0; w < W; ++w)
            if (anchor_w[w] < 0 || anchor_w[w] >= s->A) return UWBGO_E_INVALID;
    } else if (anchor < 0 || anchor >= s->A) {
        return UWBGO_E_INVALID;
    }
    /* the message straight from the caller's arrays (page-locked ones -- uwbgo_host_alloc -- make the copies
     * asynchronous; pageable ones work, the runtime stages them) */
    SCU(cudaMemcpyAsync(s->msg_d, distance, W * 4, cudaMemcpyHostToDevice, s->st));
    SCU(cudaMemcpyAsync(s->msg_e, distance_err, W * 4, cudaMemcpyHostToDevice, s->st));
    SCU(cudaMemcpyAsync(s->msg_dt, dt, W * 8, cudaMemcpyHostToDevice, s->st));
    if (per_robot) SCU(cudaMemcpyAsync(s->msg_a, anchor_w, W * 4, cudaMemcpyHostToDevice, s->st));
    if (rej) {
        stream_gate_kernel<<<(unsigned)((W + 255) / 256), 256, 0, s->st>>>((int64_t)W, N, s->res, s->anchors, s->msg_a, s->msg_d,
                                                                          s->outlier, s->rej);
        SCU(cudaGetLastError());
    }
    const int nxt = s->cur ^ 1;
    const int64_t threads = (int64_t)W * N;
    stream_shift_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s->st>>>(
        (int64_t)W, N, s->res, s->in, s->d[s->cur], s->e[s->cur], s->dt[s->cur], s->d[nxt], s->e[nxt], s->dt[nxt], s->msg_d,
        s->msg_e, s->msg_dt, per_robot ? s->aid[s->cur] : nullptr, per_robot ? s->aid[nxt] : nullptr, s->msg_a, s->anchors,
        s->anch_w, rej);
    SCU(cudaGetLastError());
    s->cur = nxt;
    if (!per_robot) {
        /* drop-oldest: the anchor pattern moves with the poses */
        for (int k = 0; k + 1 < N; ++k) s->anchor_of_pose[k] = s->anchor_of_pose[k + 1];
        s->anchor_of_pose[N - 1] = anchor;
    }
    build_topology(s);
    uwbgo_topology T{};
    T.n_poses = N;
    T.n_anchors = per_robot ? N : s->A;
    T.n_antennas = 0;
    T.n_edges = (int32_t)s->ek.size();
    T.edge_kind = s->ek.data();
    T.edge_a = s->ea.data();
    T.edge_b = s->eb.data();
    T.edge_ant = s->eant.data();
    T.edge_robust = s->erob.data();
    T.edge_ant_b = nullptr;
    uwbgo_range_msgs m{};
    m.distance = s->d[nxt];
    m.distance_err = s->e[nxt];
    m.dt_anchor = nullptr;
    m.dt_pose = s->dt[nxt];
    m.v_max = s->v_max;
    uwbgo_batch b{};
    b.n_windows = (int64_t)W;
    b.pose_t = s->in;
    b.anchors = per_robot ? s->anch_w : s->anchors;
    b.range_msgs = &m;
    b.shared = per_robot ? 0 : UWBGO_SHARED_ANCHORS;
    if (per_robot && s->rd) { /* a small fleet: one CTA per robot (uwbgo_window.cu) if the context's limit allows */
        const int64_t n = (int64_t)W * (2 * N - 1);
        stream_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s->st>>>((int64_t)W, N, s->d[nxt], s->e[nxt], s->dt[nxt], s->v_max,
                                                                            s->rd, s->ri);
        SCU(cudaGetLastError());
        b.range_msgs = nullptr;
        b.range_d = s->rd;
        b.range_info = s->ri;
    }
    uwbgo_result r{};
    r.pose_t = s->res;
    r.chi2 = s->chi2;
    r.status = s->status;
    int rc = uwbgo_solve_batch_device(s->ctx, &T, &b, &s->cfg, &r, s->st);
    if (rc) return rc;
    stream_finish_kernel<<<(unsigned)((W + 255) / 256), 256, 0, s->st>>>((int64_t)W, N, s->res, s->in, s->newest, rej, s->chi2,
                                                                        s->chi2_prev, s->status);
    SCU(cudaGetLastError());
    if (newest_pose) SCU(cudaMemcpyAsync(newest_pose, s->newest, W * 24, cudaMemcpyDeviceToHost, s->st));
    if (chi2) SCU(cudaMemcpyAsync(chi2, s->chi2, W * 32, cudaMemcpyDeviceToHost, s->st));
    if (status) SCU(cudaMemcpyAsync(status, s->status, W * 16, cudaMemcpyDeviceToHost, s->st));
    SCU(cudaStreamSynchronize(s->st));
    return 0;
}

}  // namespace

extern "C" {

int uwbgo_stream_load(uwbgo_stream *s, const double *pose_t, const int32_t *anchor_of_pose, const float *distance,
                      const float *distance_err, const double *dt)
{
    if (s && anchor_of_pose && s->W <= SMALL_FLEET) {
        /* a small fleet: the same anchor sequence for every robot, held per robot, so that the solve takes the
         * WINDOW kernels (the per-robot form has one graph structure for all steps and window-private anchors) */
        std::vector<int32_t> rows((size_t)s->W * s->N);
        for (int64_t w = 0; w < s->W; ++w)
            for (int k = 0; k < s->N; ++k) rows[(size_t)w * s->N + k] = anchor_of_pose[k];
        const int rc = load_impl(s, true, pose_t, rows.data(), distance, distance_err, dt);
        if (rc == 0) s->api_per_robot = false;
        return rc;
    }
    const int rc = load_impl(s, false, pose_t, anchor_of_pose, distance, distance_err, dt);
    if (rc == 0) s->api_per_robot = false;
    return rc;
}

int uwbgo_stream_load_robots(uwbgo_stream *s, const double *pose_t, const int32_t *anchor_of_pose, const float *distance,
                             const float *distance_err, const double *dt)
{
    const int rc = load_impl(s, true, pose_t, anchor_of_pose, distance, distance_err, dt);
    if (rc == 0) s->api_per_robot = true;
    return rc;
}

int uwbgo_stream_step(uwbgo_stream *s, int32_t anchor, const float *distance, const float *distance_err, const double *dt,
                      double *newest_pose, double *chi2, int32_t *status)
{
    if (!s || !s->loaded || s->api_per_robot || s->gate) return UWBGO_E_INVALID; /* (the gate: per-robot streams only) */
    if (s->per_robot) { /* small fleet held per robot: every robot hears this anchor */
        if (anchor < 0 || anchor >= s->A) return UWBGO_E_INVALID;
        const std::vector<int32_t> a((size_t)s->W, anchor);
        return step_impl(s, 0, a.data(), distance, distance_err, dt, newest_pose, chi2, status);
    }
    return step_impl(s, anchor, nullptr, distance, distance_err, dt, newest_pose, chi2, status);
}

int uwbgo_stream_step_robots(uwbgo_stream *s, const int32_t *anchor, const float *distance, const float *distance_err,
                             const double *dt, double *newest_pose, double *chi2, int32_t *status)
{
    if (!anchor || !s || !s->api_per_robot) return UWBGO_E_INVALID;
    return step_impl(s, 0, anchor, distance, distance_err, dt, newest_pose, chi2, status);
}

int uwbgo_stream_set_outlier_gate(uwbgo_stream *s, double distance_outlier)
{
    if (!s || distance_outlier != distance_outlier) return UWBGO_E_INVALID;
    s->gate = distance_outlier >= 0.0;
    s->outlier = distance_outlier;
    return 0;
}

int uwbgo_stream_read(uwbgo_stream *s, double *pose_t)
{
    if (!s || !s->loaded || !pose_t) return UWBGO_E_INVALID;
    SCU(cudaStreamSynchronize(s->st));
    SCU(cudaMemcpy(pose_t, s->res, (size_t)s->W * s->N * 24, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
