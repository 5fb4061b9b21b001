/*
 * uwbgo_api.cu — the C ABI of include/uwbgo.h: contexts, topology compilation, workspace
 * management and the chunked host<->device pipeline around the kernels in uwbgo_kernels.cu.
 *
 * There is no CPU path in this library: every entry point that computes needs an sm_100
 * device and fails with UWBGO_E_NODEVICE otherwise.
 */
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "uwbgo_internal.h"

using namespace uwbgo;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return UWBGO_E_CUDA;
}
#define CU(call)                                             \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

#ifndef UWBGO_GEN_CTA
#define UWBGO_GEN_CTA 1 /* same switch as in uwbgo_kernels.cu: the 6x6 LM kernel is the CTA-per-tile one */
#endif
#define UWBGO_GEN_CTA_BUILD UWBGO_GEN_CTA

/* grow-only device buffer */
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return fail(UWBGO_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

/* grow-only mapped pinned host buffer: the WINDOW path's kernel reads and writes it in place */
struct PinBuf {
    void *p = nullptr, *dp = nullptr; /* host address, device address */
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        release();
        size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 16);
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&dp, p, 0);
        if (e != cudaSuccess) {
            release();
            cudaGetLastError();
            return fail(UWBGO_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = dp = nullptr;
        cap = 0;
    }
};

/* largest batch the WINDOW path takes by default: four CTAs per SM; beyond that the tile kernels
 * (lane = window) have enough windows to fill their warps (scripts/window_crossover.py: ahead at 592 windows and
 * behind at 888 on every window shape, profiles/r02_window_crossover.txt) */
constexpr int64_t WIN_MAX_DEFAULT = 592;
constexpr size_t WIN_SMEM_LIMIT = 227 * 1024 - 256;

/* a compiled topology, resident on the device */
struct TopoEntry {
    std::vector<int32_t> key;
    DevTopo gen{};        /* edge table with antenna numbers                   */
    DevTopo fast{};       /* edge table with carry slots (fast-eligible only)  */
    bool fast_ok = false;
    bool chain_ok = false; /* standard addRangeEdge chain: straight-line sweeps */
    bool bd_ok = false;    /* range edges without lever arms + priors only on a simple chain (WINDOW path: */
                           /* block-diagonal elimination when the priors' information allows)              */
    void *dmem = nullptr;
    uint64_t stamp = 0;
};

/* one pipeline lane: a stream, its device-side window-major staging and its tile workspace */
struct Lane {
    cudaStream_t st = nullptr;
    DevBuf stage; /* window-major staging (host API only) */
    DevBuf tile;  /* tile-layout workspace */
    cudaEvent_t in_ready = nullptr; /* host pipeline: the inputs of the lane's chunk have arrived */
};

constexpr int MAX_LANES = 16;

}  // namespace

struct uwbgo_ctx {
    int device = 0;
    int n_lanes = 8;
    int64_t chunk = 8192;
    bool pipeline_set = false; /* uwbgo_set_pipeline was called: the caller's split is taken as is */
    Lane lane[MAX_LANES];
    DevBuf misc;          /* staging of the stand-alone factor/solve call, FP64 peak probe */
    DevBuf ant;
    std::vector<double> ant_host; /* the antenna table `ant` holds */
    int64_t win_max = WIN_MAX_DEFAULT; /* batches up to this size take the WINDOW path */
    int64_t batch_windows = 0;         /* host pipeline: windows of the call whose chunks are being launched */
    PinBuf pin;           /* mapped pinned staging of the WINDOW path (host API) */
    std::vector<std::unique_ptr<TopoEntry>> topos;
    uint64_t stamp = 0;
    int64_t launches = 0;
    int last_path = 0;
    cudaEvent_t ws_free = nullptr; /* completion of the last user-stream call on lane 0's workspace */
    bool ws_pending = false;
    bool profile = false;          /* bracket the main kernel of device-API calls with events */
    static constexpr int K_RING = 64; /* event pairs of the last profiled calls */
    cudaEvent_t k0[K_RING] = {}, k1[K_RING] = {};
    int64_t k_count = 0;           /* profiled calls since profiling was switched on */
};

namespace {

struct Slots {
    int Er = 0, Ep = 0, Es = 0;
};

int validate_topology(const uwbgo_topology *T, std::vector<int32_t> &slot, Slots &sl)
{
    if (!T) return fail(UWBGO_E_INVALID, "topology is NULL");
    if (T->n_poses < 1) return fail(UWBGO_E_INVALID, "n_poses must be >= 1");
    if (T->n_anchors < 0 || T->n_antennas < 0 || T->n_edges < 0)
        return fail(UWBGO_E_INVALID, "negative size in topology");
    if (T->n_edges > 0 && (!T->edge_kind || !T->edge_a || !T->edge_b || !T->edge_robust))
        return fail(UWBGO_E_INVALID, "edge arrays missing");
    slot.assign((size_t)T->n_edges, 0);
    for (int e = 0; e < T->n_edges; ++e) {
        int a = T->edge_a[e], b = T->edge_b[e];
        if (a < 0 || a >= T->n_poses) return fail(UWBGO_E_INVALID, "edge_a out of range");
        switch (T->edge_kind[e]) {
        case UWBGO_EDGE_RANGE_ANCHOR:
            if (b < 0 || b >= T->n_anchors) return fail(UWBGO_E_INVALID, "anchor index out of range");
            slot[e] = sl.Er++;
            break;
        case UWBGO_EDGE_RANGE_POSE:
            if (b <= a || b >= T->n_poses)
                return fail(UWBGO_E_TOPOLOGY, "pose-pose edge: vertex 1 must be a newer pose than vertex 0");
            slot[e] = sl.Er++;
            break;
        case UWBGO_EDGE_PRIOR:
            slot[e] = sl.Ep++;
            break;
        case UWBGO_EDGE_SE3:
            if (b <= a || b >= T->n_poses)
                return fail(UWBGO_E_TOPOLOGY, "pose-pose edge: vertex 1 must be a newer pose than vertex 0");
            slot[e] = sl.Es++;
            break;
        default:
            return fail(UWBGO_E_INVALID, "unknown edge kind");
        }
        if (T->edge_kind[e] <= UWBGO_EDGE_RANGE_POSE && T->edge_ant &&
            (T->edge_ant[e] < 0 || T->edge_ant[e] > T->n_antennas))
            return fail(UWBGO_E_INVALID, "edge_ant out of range");
        if (T->edge_kind[e] <= UWBGO_EDGE_RANGE_POSE && T->edge_ant_b &&
            (T->edge_ant_b[e] < 0 || T->edge_ant_b[e] > T->n_antennas))
            return fail(UWBGO_E_INVALID, "edge_ant_b out of range");
    }
    return 0;
}

/* initializeOptimization(): everything g2o derives from the graph structure alone */
int compile_topology(uwbgo_ctx *ctx, const uwbgo_topology *T, TopoEntry **out)
{
    std::vector<int32_t> slot;
    Slots sl;
    int rc = validate_topology(T, slot, sl);
    if (rc) return rc;
    const int N = T->n_poses, E = T->n_edges;
    std::vector<int32_t> key;
    key.reserve(5 + 6 * (size_t)E);
    key.push_back(N);
    key.push_back(T->n_anchors);
    key.push_back(T->n_antennas);
    key.push_back(E);
    for (int e = 0; e < E; ++e) {
        key.push_back(T->edge_kind[e]);
        key.push_back(T->edge_a[e]);
        key.push_back(T->edge_b[e]);
        key.push_back(T->edge_ant ? T->edge_ant[e] : 0);
        key.push_back(T->edge_robust[e] ? 1 : 0);
        key.push_back((T->edge_ant_b && T->edge_kind[e] <= UWBGO_EDGE_RANGE_POSE) ? T->edge_ant_b[e] : 0);
    }
    for (auto &t : ctx->topos)
        if (t->key == key) {
            t->stamp = ++ctx->stamp;
            *out = t.get();
            return 0;
        }

    /* forest rule: the newest-first elimination has no fill iff every pose has at most one older
     * neighbour (chains: its predecessor; pose edges: the key vertex of its keyframe) */
    std::vector<int32_t> parent((size_t)N, -1);
    bool tree = false;
    for (int e = 0; e < E; ++e) {
        const int k = T->edge_kind[e];
        if (k != UWBGO_EDGE_RANGE_POSE && k != UWBGO_EDGE_SE3) continue;
        const int a = T->edge_a[e], b = T->edge_b[e];
        if (parent[b] >= 0 && parent[b] != a)
            return fail(UWBGO_E_TOPOLOGY, "a pose with two different older neighbours needs a fill-in aware solver");
        parent[b] = a;
        if (a != b - 1) tree = true;
    }
    std::vector<int32_t> child_begin((size_t)N + 1, 0), children;
    for (int i = 0; i < N; ++i) {
        child_begin[i] = (int32_t)children.size();
        for (int c = N - 1; c > i; --c)
            if (parent[c] == i) children.push_back(c);
    }
    child_begin[N] = (int32_t)children.size();
    std::vector<EdgeRec> edges((size_t)E), fedges((size_t)E);
    std::vector<int32_t> calls((size_t)N, 0), carry((size_t)N, 0);
    std::vector<std::vector<PoseOp>> per_pose((size_t)N);
    bool fast_ok = true, bd_ok = true;
    for (int e = 0; e < E; ++e) {
        EdgeRec r{};
        r.kind = T->edge_kind[e];
        r.a = T->edge_a[e];
        r.b = T->edge_b[e];
        r.slot = slot[e];
        r.ant = (r.kind <= UWBGO_EDGE_RANGE_POSE && T->edge_ant) ? T->edge_ant[e] : 0;
        r.robust = T->edge_robust[e] ? 1 : 0;
        r.ant_b = (r.kind <= UWBGO_EDGE_RANGE_POSE && T->edge_ant_b) ? T->edge_ant_b[e] : 0;
        if (r.kind <= UWBGO_EDGE_RANGE_POSE) { /* BaseBinaryEdge numeric Jacobian: 12 oplus per free vertex */
            r.base_a = calls[r.a];
            calls[r.a] += 12;
            if (r.kind == UWBGO_EDGE_RANGE_POSE) {
                r.base_b = calls[r.b];
                calls[r.b] += 12;
            }
        }
        edges[e] = r;
        per_pose[r.a].push_back(PoseOp{e, 0});
        if (r.kind == UWBGO_EDGE_RANGE_POSE || r.kind == UWBGO_EDGE_SE3)
            per_pose[r.b].push_back(PoseOp{e, 1});
        EdgeRec f = r;
        if (r.kind > UWBGO_EDGE_RANGE_POSE || r.ant != 0 || r.ant_b != 0) fast_ok = false;
        if (r.kind == UWBGO_EDGE_SE3 || (r.kind <= UWBGO_EDGE_RANGE_POSE && (r.ant != 0 || r.ant_b != 0))) bd_ok = false;
        if (r.kind == UWBGO_EDGE_RANGE_POSE) {
            int k = carry[r.a]++;
            if (k >= 2 || r.b != r.a + 1) fast_ok = false;
            f.ant = (r.a & 1) * 2 + k; /* shared-memory carry slot of the vertex-1 terms */
        }
        fedges[e] = f;
    }
    /* the window Localization::addRangeEdge builds: per pose k its anchor range edge, then (k > 0)
     * the trajectory edge (k-1, k) -- localization.cpp:331-340 */
    std::vector<ChainPose> chain((size_t)N, ChainPose{0, 0});
    bool chain_ok = fast_ok && E == 2 * N - 1;
    for (int k = 0, e = 0; chain_ok && k < N; ++k) {
        const EdgeRec &ra = edges[e++];
        chain_ok = ra.kind == UWBGO_EDGE_RANGE_ANCHOR && ra.a == k;
        chain[k].anchor = ra.b;
        chain[k].robust = ra.robust ? 1 : 0;
        if (chain_ok && k > 0) {
            const EdgeRec &rt = edges[e++];
            chain_ok = rt.kind == UWBGO_EDGE_RANGE_POSE && rt.a == k - 1 && rt.b == k;
            chain[k].robust |= rt.robust ? 2 : 0;
        }
    }
    std::vector<PoseOp> ops;
    std::vector<int32_t> op_begin((size_t)N + 1, 0);
    for (int i = 0; i < N; ++i) {
        op_begin[i] = (int32_t)ops.size();
        ops.insert(ops.end(), per_pose[i].begin(), per_pose[i].end());
    }
    op_begin[N] = (int32_t)ops.size();
    /* fused substitution + residual sweep: produce poses in ascending order, evaluate every edge
     * (insertion order) as soon as both its poses exist */
    std::vector<SchedOp> sched;
    {
        int produced = 0;
        for (int e = 0; e < E; ++e) {
            int need = edges[e].a;
            if (edges[e].kind == UWBGO_EDGE_RANGE_POSE || edges[e].kind == UWBGO_EDGE_SE3)
                need = std::max(need, edges[e].b);
            while (produced <= need) sched.push_back(SchedOp{0, produced++});
            sched.push_back(SchedOp{1, e});
        }
        while (produced < N) sched.push_back(SchedOp{0, produced++});
    }

    /* data slot -> edge: range slots, then prior slots, then se3 slots (WINDOW path work lists) */
    std::vector<int32_t> slot_edge((size_t)std::max(E, 1), 0);
    for (int e = 0; e < E; ++e) {
        const int k = edges[e].kind;
        const int at = k <= UWBGO_EDGE_RANGE_POSE ? 0 : (k == UWBGO_EDGE_PRIOR ? sl.Er : sl.Er + sl.Ep);
        slot_edge[(size_t)at + edges[e].slot] = e;
    }

    /* compact range form: per range slot, the index of its edge among the edges of its own kind */
    std::vector<int32_t> range_sub((size_t)std::max(sl.Er, 1), 0);
    int era = 0, erp = 0;
    for (int e = 0; e < E; ++e) {
        if (edges[e].kind == UWBGO_EDGE_RANGE_ANCHOR) range_sub[(size_t)edges[e].slot] = era++;
        if (edges[e].kind == UWBGO_EDGE_RANGE_POSE) range_sub[(size_t)edges[e].slot] = 0x40000000 | erp++;
    }

    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_edges = 0;
    size_t o_fedges = o_edges + al(sizeof(EdgeRec) * std::max(E, 1));
    size_t o_ops = o_fedges + al(sizeof(EdgeRec) * std::max(E, 1));
    size_t o_begin = o_ops + al(sizeof(PoseOp) * std::max<size_t>(ops.size(), 1));
    size_t o_calls = o_begin + al(sizeof(int32_t) * (N + 1));
    size_t o_sched = o_calls + al(sizeof(int32_t) * N);
    size_t o_chain = o_sched + al(sizeof(SchedOp) * sched.size());
    size_t o_parent = o_chain + al(sizeof(ChainPose) * N);
    size_t o_cbegin = o_parent + al(sizeof(int32_t) * N);
    size_t o_children = o_cbegin + al(sizeof(int32_t) * (N + 1));
    size_t o_slotedge = o_children + al(sizeof(int32_t) * std::max<size_t>(children.size(), 1));
    size_t o_rsub = o_slotedge + al(sizeof(int32_t) * slot_edge.size());
    size_t total = o_rsub + al(sizeof(int32_t) * range_sub.size());
    std::vector<char> host(total, 0);
    memcpy(host.data() + o_rsub, range_sub.data(), sizeof(int32_t) * range_sub.size());
    memcpy(host.data() + o_slotedge, slot_edge.data(), sizeof(int32_t) * slot_edge.size());
    if (E) {
        memcpy(host.data() + o_edges, edges.data(), sizeof(EdgeRec) * E);
        memcpy(host.data() + o_fedges, fedges.data(), sizeof(EdgeRec) * E);
    }
    if (!ops.empty()) memcpy(host.data() + o_ops, ops.data(), sizeof(PoseOp) * ops.size());
    memcpy(host.data() + o_begin, op_begin.data(), sizeof(int32_t) * (N + 1));
    memcpy(host.data() + o_calls, calls.data(), sizeof(int32_t) * N);
    memcpy(host.data() + o_sched, sched.data(), sizeof(SchedOp) * sched.size());
    memcpy(host.data() + o_chain, chain.data(), sizeof(ChainPose) * N);
    memcpy(host.data() + o_parent, parent.data(), sizeof(int32_t) * N);
    memcpy(host.data() + o_cbegin, child_begin.data(), sizeof(int32_t) * (N + 1));
    if (!children.empty()) memcpy(host.data() + o_children, children.data(), sizeof(int32_t) * children.size());

    auto ent = std::make_unique<TopoEntry>();
    CU(cudaMalloc(&ent->dmem, total));
    cudaError_t ce = cudaMemcpy(ent->dmem, host.data(), total, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        cudaFree(ent->dmem);
        return fail_cuda(ce, "cudaMemcpy(topology)");
    }
    char *d = static_cast<char *>(ent->dmem);
    DevTopo g{};
    g.N = N;
    g.A = T->n_anchors;
    g.K = T->n_antennas;
    g.E = E;
    g.Er = sl.Er;
    g.Ep = sl.Ep;
    g.Es = sl.Es;
    g.fast = 0;
    g.edges = reinterpret_cast<const EdgeRec *>(d + o_edges);
    g.ops = reinterpret_cast<const PoseOp *>(d + o_ops);
    g.op_begin = reinterpret_cast<const int32_t *>(d + o_begin);
    g.num_calls = reinterpret_cast<const int32_t *>(d + o_calls);
    g.sched = reinterpret_cast<const SchedOp *>(d + o_sched);
    g.n_sched = (int32_t)sched.size();
    g.n_ops = (int32_t)ops.size();
    g.chain = reinterpret_cast<const ChainPose *>(d + o_chain);
    g.tree = tree ? 1 : 0;
    g.simple_chain = 1;
    for (int i = 1; i < N; ++i)
        if (parent[i] != i - 1) g.simple_chain = 0;
    g.parent = reinterpret_cast<const int32_t *>(d + o_parent);
    g.child_begin = reinterpret_cast<const int32_t *>(d + o_cbegin);
    g.children = reinterpret_cast<const int32_t *>(d + o_children);
    g.slot_edge = reinterpret_cast<const int32_t *>(d + o_slotedge);
    g.range_sub = reinterpret_cast<const int32_t *>(d + o_rsub);
    g.Era = era;
    g.Erp = erp;
    ent->gen = g;
    ent->fast = g;
    ent->fast.fast = chain_ok ? 2 : 1;
    ent->chain_ok = chain_ok;
    ent->fast.edges = reinterpret_cast<const EdgeRec *>(d + o_fedges);
    ent->fast_ok = fast_ok;
    ent->bd_ok = bd_ok && g.simple_chain;
    ent->key = std::move(key);
    ent->stamp = ++ctx->stamp;
    if (ctx->topos.size() >= 32) { /* evict the least recently used */
        size_t victim = 0;
        for (size_t k = 1; k < ctx->topos.size(); ++k)
            if (ctx->topos[k]->stamp < ctx->topos[victim]->stamp) victim = k;
        cudaDeviceSynchronize();
        cudaFree(ctx->topos[victim]->dmem);
        ctx->topos.erase(ctx->topos.begin() + (long)victim);
    }
    ctx->topos.push_back(std::move(ent));
    *out = ctx->topos.back().get();
    return 0;
}

int make_cfg(const uwbgo_config *c, DevCfg &d)
{
    if (!c) return fail(UWBGO_E_INVALID, "config is NULL");
    if (c->max_iterations < 0 || c->max_trials < 1 || c->orthogonalize_after < 0 ||
        c->orthogonalize_after > 1000000000)
        return fail(UWBGO_E_INVALID, "bad iteration / trial / orthogonalize_after setting");
    d.max_iterations = c->max_iterations;
    d.max_trials = c->max_trials;
    d.orth_mod = c->orthogonalize_after + 1;
    d.tau = c->tau;
    d.good_lo = c->good_step_lower;
    d.good_hi = c->good_step_upper;
    d.kdelta = c->kernel_delta;
    d.jdelta = c->jacobian_delta;
    return 0;
}

int check_batch(const TopoEntry &te, const uwbgo_batch *in)
{
    if (!in) return fail(UWBGO_E_INVALID, "batch is NULL");
    if (in->n_windows < 0) return fail(UWBGO_E_INVALID, "n_windows < 0");
    const DevTopo &g = te.gen;
    if (in->n_windows == 0) return 0;
    if (!in->pose_t) return fail(UWBGO_E_INVALID, "pose_t is NULL");
    if (g.A > 0 && !in->anchors) return fail(UWBGO_E_INVALID, "anchors is NULL");
    if (g.K > 0 && !in->ant_offsets) return fail(UWBGO_E_INVALID, "ant_offsets is NULL");
    if (in->shared & ~(UWBGO_SHARED_ANCHORS | UWBGO_DIAG_INFO)) return fail(UWBGO_E_INVALID, "unknown bit in batch.shared");
    if (in->range_msgs) {
        const uwbgo_range_msgs *m = in->range_msgs;
        if (in->range_d || in->range_info) return fail(UWBGO_E_INVALID, "range_msgs replaces range_d / range_info: pass one form");
        if (g.Era > 0 && (!m->distance || !m->distance_err)) return fail(UWBGO_E_INVALID, "range_msgs: distance / distance_err NULL");
        if (g.Erp > 0 && !m->dt_pose) return fail(UWBGO_E_INVALID, "range_msgs: dt_pose NULL");
    } else if (g.Er > 0 && (!in->range_d || !in->range_info))
        return fail(UWBGO_E_INVALID, "range data NULL");
    if (g.Ep > 0 && (!in->prior_Z || !in->prior_info)) return fail(UWBGO_E_INVALID, "prior data NULL");
    if (g.Es > 0 && (!in->se3_Z || !in->se3_info)) return fail(UWBGO_E_INVALID, "se3 data NULL");
    return 0;
}

/* carve the tile workspace of one launch out of a lane's buffer */
struct TileLayout {
    size_t bytes = 0;
    size_t off_T[2], off_R[2], off_cnt, off_anch, off_rd, off_ri, off_pZ, off_pI, off_sZ, off_sI, off_HB,
        off_LR, off_chi2, off_status, off_echi, off_sel, off_mg, off_jrec;
};
TileLayout tile_layout(const DevTopo &t, bool fast, int64_t W, bool want_cnt, bool want_LR, bool want_sel = false,
                       bool want_marginal = false)
{
    TileLayout L;
    const size_t tw = (size_t)n_tiles(W) * TILE; /* padded window count */
    size_t o = 0;
    auto take = [&](size_t rows, size_t elem) {
        size_t at = o;
        o += (rows * tw * elem + 255) & ~(size_t)255;
        return at;
    };
    const size_t N = (size_t)t.N;
    L.off_T[0] = take(N * 3, 8);
    L.off_T[1] = take(N * 3, 8);
    L.off_R[0] = fast ? 0 : take(N * 9, 8);
    L.off_R[1] = fast ? 0 : take(N * 9, 8);
    L.off_cnt = (fast && !want_cnt) ? 0 : take(N, 4);
    L.off_anch = take((size_t)t.A * 3, 8);
    L.off_rd = take((size_t)t.Er, 8);
    L.off_ri = take((size_t)t.Er, 8);
    L.off_pZ = take((size_t)t.Ep * 12, 8);
    L.off_pI = take((size_t)t.Ep * 36, 8);
    L.off_sZ = take((size_t)t.Es * 12, 8);
    L.off_sI = take((size_t)t.Es * 36, 8);
    /* FAST solves rebuild H inside the factor sweep and never store it */
    L.off_HB = (fast && want_LR) ? 0 : take(N * (fast ? HR_FAST : HR_GEN), 8);
    L.off_LR = want_LR ? take(N * (fast ? LR_FAST : (t.tree ? LR_TREE : LR_GEN)), 8) : 0;
    L.off_chi2 = take(4, 8);
    L.off_status = take(4, 4);
    /* per-edge chi2 terms: the general LM kernel, and the fused linearise stage of CHAIN windows */
    L.off_echi = ((!fast && want_LR) || (fast && !want_LR)) ? take((size_t)t.E * 2, 8) : 0;
    L.off_jrec = (!fast && want_LR && !t.tree) ? take(general_items_jrec_rows(t), 8) : 0;
    L.off_sel = want_sel ? take(1, 4) : 0;
    L.off_mg = (want_marginal && !fast) ? take(marginal_scratch_bytes(t, TILE) / (TILE * 8), 8) : 0;
    L.bytes = o;
    return L;
}

/* device-side core shared by every entry point: inputs/outputs are DEVICE window-major arrays.
 * mode 0: full LM solve.  mode 1: one linearisation (H_diag/H_off/b/chi2 out). */
struct StageOut {
    double *H_diag, *H_off, *b, *chi2;
};
int run_device(uwbgo_ctx *ctx, Lane &ln, const TopoEntry &te, const DevCfg &cfg,
               const uwbgo_batch *in, const double *d_ant, uwbgo_result *out, const StageOut *so,
               cudaStream_t st)
{
    const int64_t W = in->n_windows;
    if (W == 0) return 0;
    const bool fast = te.fast_ok && in->pose_R == nullptr;
    const DevTopo &tp = fast ? te.fast : te.gen;
    const bool want_cnt = (in->oplus_count != nullptr) || (out && out->oplus_count != nullptr);
    const bool want_echi = !so && out->edge_chi2 != nullptr, want_marg = !so && out->marginal != nullptr;
    if ((want_echi || want_marg) && !fast && !UWBGO_GEN_CTA_BUILD)
        return fail(UWBGO_E_INVALID, "edge_chi2 / marginal need the CTA-per-tile kernel (library built with -DUWBGO_GEN_CTA=0)");
    TileLayout L = tile_layout(tp, fast, W, want_cnt, so == nullptr, want_echi && fast, want_marg);
    int rc = ln.tile.reserve(L.bytes);
    if (rc) return rc;
    char *base = static_cast<char *>(ln.tile.p);
    DevWs ws{};
    ws.W = W;
    ws.W_batch = ctx->batch_windows;
    ws.T[0] = reinterpret_cast<double *>(base + L.off_T[0]);
    ws.T[1] = reinterpret_cast<double *>(base + L.off_T[1]);
    if (!fast) {
        ws.Rm[0] = reinterpret_cast<double *>(base + L.off_R[0]);
        ws.Rm[1] = reinterpret_cast<double *>(base + L.off_R[1]);
    }
    const bool have_cnt = !fast || want_cnt;
    ws.cnt = have_cnt ? reinterpret_cast<int32_t *>(base + L.off_cnt) : nullptr;
    ws.anch = reinterpret_cast<double *>(base + L.off_anch);
    ws.rd = reinterpret_cast<double *>(base + L.off_rd);
    ws.ri = reinterpret_cast<double *>(base + L.off_ri);
    ws.pZ = reinterpret_cast<double *>(base + L.off_pZ);
    ws.pI = reinterpret_cast<double *>(base + L.off_pI);
    ws.sZ = reinterpret_cast<double *>(base + L.off_sZ);
    ws.sI = reinterpret_cast<double *>(base + L.off_sI);
    ws.HB = reinterpret_cast<double *>(base + L.off_HB);
    ws.LR = so ? nullptr : reinterpret_cast<double *>(base + L.off_LR);
    ws.ant = d_ant;
    ws.chi2 = reinterpret_cast<double *>(base + L.off_chi2);
    ws.status = reinterpret_cast<int32_t *>(base + L.off_status);
    ws.echi = ((!fast && !so) || (fast && so)) ? reinterpret_cast<double *>(base + L.off_echi) : nullptr;
    ws.jrec = (!fast && !so && !tp.tree) ? reinterpret_cast<double *>(base + L.off_jrec) : nullptr;
    ws.stale_sel = (want_echi && fast) ? reinterpret_cast<int32_t *>(base + L.off_sel) : nullptr;

    XposeJobs pj{};
    pj.W = W;
    auto add = [&](XposeJobs &J, const void *src, void *dst, int C, int elem, int mode) {
        if (C <= 0) return;
        XposeJob &j = J.job[J.n++];
        j.src = src;
        j.dst = dst;
        j.C = C;
        j.elem = elem;
        j.mode = mode;
        j.aux = 0;
    };
    add(pj, in->pose_t, ws.T[0], tp.N * 3, 8, 0);
    if (!fast) {
        if (in->pose_R)
            add(pj, in->pose_R, ws.Rm[0], tp.N * 9, 8, 0);
        else
            add(pj, nullptr, ws.Rm[0], tp.N * 9, 8, 1);
    }
    if (have_cnt) {
        if (in->oplus_count)
            add(pj, in->oplus_count, ws.cnt, tp.N, 4, 0);
        else
            add(pj, nullptr, ws.cnt, tp.N, 4, 3);
    }
    add(pj, in->anchors, const_cast<double *>(ws.anch), tp.A * 3, 8, (in->shared & UWBGO_SHARED_ANCHORS) ? 4 : 0);
    if (!in->range_msgs) {
        add(pj, in->range_d, const_cast<double *>(ws.rd), tp.Er, 8, 0);
        add(pj, in->range_info, const_cast<double *>(ws.ri), tp.Er, 8, 0);
    }
    add(pj, in->prior_Z, const_cast<double *>(ws.pZ), tp.Ep * 12, 8, 0);
    /* UWBGO_DIAG_INFO: the matrices are rebuilt from their diagonals on the way into the tile layout */
    const int info_mode = (in->shared & UWBGO_DIAG_INFO) ? 5 : 0;
    add(pj, in->prior_info, const_cast<double *>(ws.pI), tp.Ep * 36, 8, info_mode);
    add(pj, in->se3_Z, const_cast<double *>(ws.sZ), tp.Es * 12, 8, 0);
    add(pj, in->se3_info, const_cast<double *>(ws.sI), tp.Es * 36, 8, info_mode);
    CU(launch_pack(pj, st));
    ctx->launches += 1;
    if (in->range_msgs && tp.Er > 0) { /* create_range_edge on the device */
        const uwbgo_range_msgs *m = in->range_msgs;
        RangeMsgsDev md{m->distance, m->distance_err, m->dt_anchor, m->dt_pose, m->v_max};
        CU(launch_pack_range_msgs(tp, md, W, const_cast<double *>(ws.rd), const_cast<double *>(ws.ri), st));
        ctx->launches += 1;
    }

    XposeJobs uj{};
    uj.W = W;
    const bool timed = ctx->profile && &ln == &ctx->lane[0];
    const int kslot = (int)(ctx->k_count % uwbgo_ctx::K_RING);
    if (timed) CU(cudaEventRecord(ctx->k0[kslot], st));
    if (so) {
#ifndef UWBGO_LIN_FUSED
#define UWBGO_LIN_FUSED 1 /* CHAIN windows: one fused linearise kernel instead of records + expansion */
#endif
        if (UWBGO_LIN_FUSED && fast && linearize_chain_fused_ok(tp, so->H_diag, so->H_off, so->b)) {
            CU(launch_linearize_chain_fused(tp, cfg, ws, so->H_diag, so->H_off, so->b, so->chi2 != nullptr, st));
            if (!so->chi2) ctx->launches -= 1; /* the fused kernel alone, or the fused kernel + chi_sum_kernel */
        } else {
            CU(launch_linearize(tp, cfg, ws, st));
            CU(launch_expand_H(tp, ws, so->H_diag, so->H_off, so->b, st));
        }
        if (timed) CU(cudaEventRecord(ctx->k1[kslot], st)); /* stage = linearise + expansion to the public layout */
        ctx->launches += 2;
        if (so->chi2) add(uj, ws.chi2, so->chi2, 2, 8, 0);
    } else {
        CU(launch_solve(tp, cfg, ws, st));
        if (timed) CU(cudaEventRecord(ctx->k1[kslot], st));
        ctx->launches += 1;
        add(uj, ws.T[0], out->pose_t, tp.N * 3, 8, 0);
        if (out->pose_R) {
            if (fast)
                add(uj, nullptr, out->pose_R, tp.N * 9, 8, 2);
            else
                add(uj, ws.Rm[0], out->pose_R, tp.N * 9, 8, 0);
        }
        if (out->oplus_count) add(uj, ws.cnt, out->oplus_count, tp.N, 4, 0);
        if (out->chi2) add(uj, ws.chi2, out->chi2, 4, 8, 0);
        if (out->status) add(uj, ws.status, out->status, 4, 4, 0);
        if (want_echi && tp.E > 0) {
            CU(launch_edge_chi2_out(tp, ws, out->edge_chi2, st));
            ctx->launches += 1;
        }
        if (want_marg) {
            CU(launch_marginal(tp, ws, fast ? nullptr : reinterpret_cast<double *>(base + L.off_mg), out->marginal,
                               out->marginal_ok, st));
            ctx->launches += 1;
        }
    }
    if (uj.n) {
        CU(launch_unpack(uj, st));
        ctx->launches += 1;
    }
    ctx->last_path = fast ? tp.fast : 0;
    if (timed) ctx->k_count += 1;
    return 0;
}

int upload_ant(uwbgo_ctx *ctx, const DevTopo &g, const uwbgo_batch *in, cudaStream_t st,
               const double **d_ant)
{
    *d_ant = nullptr;
    if (g.K <= 0) return 0;
    const size_t n = 3 * (size_t)g.K;
    /* ant_offsets is host memory in both API flavours.  The device copy is kept across calls: the
     * stream is synchronised only when the table CHANGES (the copy reads caller memory, and kernels
     * of earlier calls may still be reading the old table); in the steady state nothing is copied
     * and the *_device entry points queue their work without synchronising */
    if (ctx->ant.p && ctx->ant_host.size() == n && memcmp(ctx->ant_host.data(), in->ant_offsets, n * sizeof(double)) == 0) {
        *d_ant = static_cast<const double *>(ctx->ant.p);
        return 0;
    }
    CU(cudaDeviceSynchronize());
    int rc = ctx->ant.reserve(sizeof(double) * n);
    if (rc) return rc;
    ctx->ant_host.assign(in->ant_offsets, in->ant_offsets + n);
    CU(cudaMemcpyAsync(ctx->ant.p, ctx->ant_host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    *d_ant = static_cast<const double *>(ctx->ant.p);
    return 0;
}

/* WINDOW path eligibility: small batch, window state fits the shared memory of one SM */
bool window_path_ok(const uwbgo_ctx *ctx, const TopoEntry &te, int64_t W, const uwbgo_batch *in = nullptr,
                    const uwbgo_result *out = nullptr)
{
    if (in && (in->range_msgs || in->shared)) return false;                    /* compact inputs: tile kernels */
    if (out && (out->edge_chi2 || out->marginal)) return false;                /* extras: tile kernels          */
    return W > 0 && W <= ctx->win_max && window_path_smem_bytes(te.gen, 1) <= WIN_SMEM_LIMIT;
}

#ifndef UWBGO_WIN_NO_T3
#define UWBGO_WIN_NO_T3 0 /* A/B: 1 = translation-only windows take the 6x6 WINDOW kernel too */
#endif
/* WINDOW path on window-major arrays the device can address (device memory, or mapped pinned host
 * memory): ONE launch, no transposition */
int run_window(uwbgo_ctx *ctx, const TopoEntry &te, const DevCfg &cfg, const uwbgo_batch *in, const double *d_ant,
               uwbgo_result *out, cudaStream_t st)
{
    WinIo io{};
    io.W = in->n_windows;
    io.pose_t = in->pose_t;
    io.pose_R = in->pose_R;
    io.cnt_in = in->oplus_count;
    io.anchors = in->anchors;
    io.rd = in->range_d;
    io.ri = in->range_info;
    io.pZ = in->prior_Z;
    io.pI = in->prior_info;
    io.sZ = in->se3_Z;
    io.sI = in->se3_info;
    io.ant = d_ant;
    io.o_pose_t = out->pose_t;
    io.o_pose_R = out->pose_R;
    io.o_cnt = out->oplus_count;
    io.o_chi2 = out->chi2;
    io.o_status = out->status;
    /* UWBGO_WIN_BLOCKS=6 in the environment: every window on the full 6x6 blocks (A/B runs) */
    static const bool full_blocks = getenv("UWBGO_WIN_BLOCKS") && !strcmp(getenv("UWBGO_WIN_BLOCKS"), "6");
    const bool small_blocks = !UWBGO_WIN_NO_T3 && !full_blocks;
    io.t3 = (te.fast_ok && in->pose_R == nullptr && te.gen.simple_chain && small_blocks) ? 1 : 0;
    io.bd_ok = (te.bd_ok && !io.t3 && small_blocks) ? 1 : 0;
    const bool timed = ctx->profile;
    const int kslot = (int)(ctx->k_count % uwbgo_ctx::K_RING);
    if (timed) CU(cudaEventRecord(ctx->k0[kslot], st));
    int ks = window_path_candidates(io.W); /* as many speculative trials as fit the shared memory */
    while (ks > 1 && window_path_smem_bytes(te.gen, ks) > WIN_SMEM_LIMIT) ks /= 2;
    CU(launch_solve_window(te.gen, cfg, io, ks, ctx->device, st));
    if (timed) {
        CU(cudaEventRecord(ctx->k1[kslot], st));
        ctx->k_count += 1;
    }
    ctx->launches += 1;
    ctx->last_path = 3;
    return 0;
}

/* host arrays: one memcpy each into the mapped pinned block, one launch that reads and writes the
 * block in place, one stream synchronise, results copied out */
int host_window_path(uwbgo_ctx *ctx, const TopoEntry &te, const DevCfg &cfg, const uwbgo_batch *in, uwbgo_result *out)
{
    const DevTopo &g = te.gen;
    const size_t W = (size_t)in->n_windows, N = (size_t)g.N;
    size_t o = 0;
    auto take = [&](bool need, size_t bytes) {
        size_t at = o;
        if (need) o += (bytes + 63) & ~(size_t)63;
        return at;
    };
    const size_t o_ant = take(g.K > 0, (size_t)g.K * 24);
    const size_t o_t = take(true, W * N * 24), o_R = take(in->pose_R != nullptr, W * N * 72);
    const size_t o_cnt = take(in->oplus_count != nullptr, W * N * 4);
    const size_t o_an = take(g.A > 0, W * g.A * 24);
    const size_t o_rd = take(g.Er > 0, W * g.Er * 8), o_ri = take(g.Er > 0, W * g.Er * 8);
    const size_t o_pZ = take(g.Ep > 0, W * g.Ep * 96), o_pI = take(g.Ep > 0, W * g.Ep * 288);
    const size_t o_sZ = take(g.Es > 0, W * g.Es * 96), o_sI = take(g.Es > 0, W * g.Es * 288);
    const size_t r_t = take(true, W * N * 24), r_R = take(out->pose_R != nullptr, W * N * 72);
    const size_t r_cnt = take(out->oplus_count != nullptr, W * N * 4);
    const size_t r_chi = take(out->chi2 != nullptr, W * UWBGO_CHI2_STRIDE * 8);
    const size_t r_st = take(out->status != nullptr, W * UWBGO_STATUS_STRIDE * 4);
    int rc = ctx->pin.reserve(o);
    if (rc) return rc;
    char *h = static_cast<char *>(ctx->pin.p), *d = static_cast<char *>(ctx->pin.dp);
    auto put = [&](const void *src, size_t off, size_t bytes) {
        if (src && bytes) memcpy(h + off, src, bytes);
    };
    if (g.K > 0) put(in->ant_offsets, o_ant, (size_t)g.K * 24);
    put(in->pose_t, o_t, W * N * 24);
    put(in->pose_R, o_R, W * N * 72);
    put(in->oplus_count, o_cnt, W * N * 4);
    if (g.A > 0) put(in->anchors, o_an, W * g.A * 24);
    if (g.Er > 0) {
        put(in->range_d, o_rd, W * g.Er * 8);
        put(in->range_info, o_ri, W * g.Er * 8);
    }
    if (g.Ep > 0) {
        put(in->prior_Z, o_pZ, W * g.Ep * 96);
        put(in->prior_info, o_pI, W * g.Ep * 288);
    }
    if (g.Es > 0) {
        put(in->se3_Z, o_sZ, W * g.Es * 96);
        put(in->se3_info, o_sI, W * g.Es * 288);
    }
    uwbgo_batch db{};
    db.n_windows = in->n_windows;
    db.pose_t = reinterpret_cast<double *>(d + o_t);
    db.pose_R = in->pose_R ? reinterpret_cast<double *>(d + o_R) : nullptr;
    db.oplus_count = in->oplus_count ? reinterpret_cast<int32_t *>(d + o_cnt) : nullptr;
    db.anchors = reinterpret_cast<double *>(d + o_an);
    db.range_d = reinterpret_cast<double *>(d + o_rd);
    db.range_info = reinterpret_cast<double *>(d + o_ri);
    db.prior_Z = reinterpret_cast<double *>(d + o_pZ);
    db.prior_info = reinterpret_cast<double *>(d + o_pI);
    db.se3_Z = reinterpret_cast<double *>(d + o_sZ);
    db.se3_info = reinterpret_cast<double *>(d + o_sI);
    uwbgo_result dr{};
    dr.pose_t = reinterpret_cast<double *>(d + r_t);
    dr.pose_R = out->pose_R ? reinterpret_cast<double *>(d + r_R) : nullptr;
    dr.oplus_count = out->oplus_count ? reinterpret_cast<int32_t *>(d + r_cnt) : nullptr;
    dr.chi2 = out->chi2 ? reinterpret_cast<double *>(d + r_chi) : nullptr;
    dr.status = out->status ? reinterpret_cast<int32_t *>(d + r_st) : nullptr;
    cudaStream_t st = ctx->lane[0].st;
    if ((rc = run_window(ctx, te, cfg, &db, reinterpret_cast<const double *>(d + o_ant), &dr, st))) return rc;
    CU(cudaStreamSynchronize(st));
    memcpy(out->pose_t, h + r_t, W * N * 24);
    if (out->pose_R) memcpy(out->pose_R, h + r_R, W * N * 72);
    if (out->oplus_count) memcpy(out->oplus_count, h + r_cnt, W * N * 4);
    if (out->chi2) memcpy(out->chi2, h + r_chi, W * UWBGO_CHI2_STRIDE * 8);
    if (out->status) memcpy(out->status, h + r_st, W * UWBGO_STATUS_STRIDE * 4);
    return 0;
}

int user_stream_begin(uwbgo_ctx *ctx, cudaStream_t st)
{
    if (ctx->ws_pending) CU(cudaStreamWaitEvent(st, ctx->ws_free, 0));
    return 0;
}
int user_stream_end(uwbgo_ctx *ctx, cudaStream_t st)
{
    CU(cudaEventRecord(ctx->ws_free, st));
    ctx->ws_pending = true;
    return 0;
}

}  // namespace

/* ------------------------------------------------------------------------------------------ */
extern "C" {

int uwbgo_abi_version(void) { return UWBGO_ABI_VERSION; }

void uwbgo_config_default(uwbgo_config *cfg)
{
    if (!cfg) return;
    cfg->max_iterations = 20;       /* optimizer/maximum_iteration default, localization.cpp:65 */
    cfg->max_trials = 10;           /* g2o OptimizationAlgorithmLevenberg maxTrialsAfterFailure  */
    cfg->orthogonalize_after = 1000;/* g2o VertexSE3::orthogonalizeAfter                         */
    cfg->reserved = 0;
    cfg->tau = 1e-5;
    cfg->good_step_lower = 1.0 / 3.0;
    cfg->good_step_upper = 2.0 / 3.0;
    cfg->kernel_delta = 1.0;        /* RobustKernelCauchy default delta, localization.cpp:624    */
    cfg->jacobian_delta = 1e-9;     /* BaseBinaryEdge::linearizeOplus numeric step              */
}

const char *uwbgo_last_error(void) { return g_last_error.c_str(); }

int uwbgo_create(int device, uwbgo_ctx **out)
{
    if (!out) return fail(UWBGO_E_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(UWBGO_E_NODEVICE, "no CUDA device visible (this library has no CPU path)");
    }
    if (device < 0 || device >= n) return fail(UWBGO_E_NODEVICE, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(UWBGO_E_NODEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                          std::to_string(prop.minor) + ", kernels are built for sm_100a only");
    CU(cudaSetDevice(device));
    auto ctx = new uwbgo_ctx();
    ctx->device = device;
    e = cudaSuccess;
    for (int k = 0; k < MAX_LANES; ++k) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->lane[k].st, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->lane[k].in_ready, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            uwbgo_destroy(ctx);
            return fail_cuda(e, "stream/event creation");
        }
    }
    e = cudaEventCreateWithFlags(&ctx->ws_free, cudaEventDisableTiming);
    for (int k = 0; k < uwbgo_ctx::K_RING; ++k) {
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->k0[k]);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->k1[k]);
    }
    if (e != cudaSuccess) {
        uwbgo_destroy(ctx);
        return fail_cuda(e, "event creation");
    }
    *out = ctx;
    return 0;
}

void uwbgo_destroy(uwbgo_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int k = 0; k < MAX_LANES; ++k) {
        ctx->lane[k].stage.release();
        ctx->lane[k].tile.release();
        if (ctx->lane[k].st) cudaStreamDestroy(ctx->lane[k].st);
        if (ctx->lane[k].in_ready) cudaEventDestroy(ctx->lane[k].in_ready);
    }
    ctx->misc.release();
    ctx->ant.release();
    ctx->pin.release();
    for (auto &t : ctx->topos) cudaFree(t->dmem);
    if (ctx->ws_free) cudaEventDestroy(ctx->ws_free);
    for (int k = 0; k < uwbgo_ctx::K_RING; ++k) {
        if (ctx->k0[k]) cudaEventDestroy(ctx->k0[k]);
        if (ctx->k1[k]) cudaEventDestroy(ctx->k1[k]);
    }
    delete ctx;
}

int uwbgo_set_pipeline(uwbgo_ctx *ctx, int64_t windows_per_chunk, int n_lanes)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (windows_per_chunk < 32 || n_lanes < 1 || n_lanes > MAX_LANES)
        return fail(UWBGO_E_INVALID, "windows_per_chunk >= 32 and 1 <= n_lanes <= 16 required");
    ctx->chunk = (windows_per_chunk + 31) / 32 * 32;
    ctx->n_lanes = n_lanes;
    ctx->pipeline_set = true;
    return 0;
}

int uwbgo_set_window_path(uwbgo_ctx *ctx, int64_t max_windows)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    ctx->win_max = max_windows < 0 ? WIN_MAX_DEFAULT : max_windows;
    return 0;
}

/* pinned host memory for callers that want the host API to overlap copies with compute */
void *uwbgo_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        g_last_error = "cudaHostAlloc failed";
        return nullptr;
    }
    return p;
}
void uwbgo_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int uwbgo_solve_batch_device(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                             const uwbgo_config *cfg, uwbgo_result *out, void *stream)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (!out || (in && in->n_windows > 0 && !out->pose_t)) return fail(UWBGO_E_INVALID, "result.pose_t is NULL");
    CU(cudaSetDevice(ctx->device));
    TopoEntry *te = nullptr;
    int rc = compile_topology(ctx, topo, &te);
    if (rc) return rc;
    DevCfg dc;
    if ((rc = make_cfg(cfg, dc))) return rc;
    if ((rc = check_batch(*te, in))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = user_stream_begin(ctx, st))) return rc;
    const double *d_ant = nullptr;
    if ((rc = upload_ant(ctx, te->gen, in, st, &d_ant))) return rc;
    if (window_path_ok(ctx, *te, in->n_windows, in, out)) /* no workspace: nothing to order against other calls */
        return run_window(ctx, *te, dc, in, d_ant, out, st);
    if ((rc = run_device(ctx, ctx->lane[0], *te, dc, in, d_ant, out, nullptr, st))) return rc;
    return user_stream_end(ctx, st);
}

int uwbgo_linearize_batch_device(uwbgo_ctx *ctx, const uwbgo_topology *topo,
                                 const uwbgo_batch *in, const uwbgo_config *cfg, double *H_diag,
                                 double *H_off, double *b, double *chi2, void *stream)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    TopoEntry *te = nullptr;
    int rc = compile_topology(ctx, topo, &te);
    if (rc) return rc;
    if (!H_diag || !b || (te->gen.N > 1 && !H_off)) return fail(UWBGO_E_INVALID, "output array is NULL");
    DevCfg dc;
    if ((rc = make_cfg(cfg, dc))) return rc;
    if ((rc = check_batch(*te, in))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((rc = user_stream_begin(ctx, st))) return rc;
    const double *d_ant = nullptr;
    if ((rc = upload_ant(ctx, te->gen, in, st, &d_ant))) return rc;
    StageOut so{H_diag, H_off, b, chi2};
    if ((rc = run_device(ctx, ctx->lane[0], *te, dc, in, d_ant, nullptr, &so, st))) return rc;
    return user_stream_end(ctx, st);
}

int uwbgo_factor_solve_batch_device(uwbgo_ctx *ctx, int32_t n_poses, int64_t n_windows,
                                    const double *H_diag, const double *H_off, const double *b,
                                    const double *lambda, double *x, int32_t *ok, void *stream)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (n_poses < 1 || n_windows < 0) return fail(UWBGO_E_INVALID, "bad sizes");
    if (n_windows == 0) return 0;
    if (!H_diag || !b || !lambda || !x || (n_poses > 1 && !H_off))
        return fail(UWBGO_E_INVALID, "array is NULL");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if ((rc = user_stream_begin(ctx, st))) return rc;
    if ((rc = ctx->lane[0].tile.reserve(factor_solve_scratch_bytes(n_poses, n_windows)))) return rc;
    CU(launch_factor_solve(n_poses, n_windows, H_diag, H_off, b, lambda, x, ok,
                           static_cast<double *>(ctx->lane[0].tile.p), st));
    ctx->launches += 2;
    return user_stream_end(ctx, st);
}

/* ---- host-pointer flavours: chunked, multi-stream pipeline -------------------------------- */
namespace {
struct StageLayout {
    size_t bytes = 0;
    size_t in_pose_t, in_pose_R, in_cnt, in_anch, in_rd, in_ri, in_pZ, in_pI, in_sZ, in_sI;
    size_t in_md, in_me, in_mta, in_mtp; /* compact range form */
    size_t out_pose_t, out_pose_R, out_cnt, out_chi2, out_status, out_echi, out_marg, out_mok;
    size_t out_Hd, out_Ho, out_b;
};
}  // namespace

static int host_pipeline(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                         const uwbgo_config *cfg, uwbgo_result *out, double *H_diag, double *H_off,
                         double *b, double *chi2, bool linearize)
{
    CU(cudaSetDevice(ctx->device));
    TopoEntry *te = nullptr;
    int rc = compile_topology(ctx, topo, &te);
    if (rc) return rc;
    DevCfg dc;
    if ((rc = make_cfg(cfg, dc))) return rc;
    if ((rc = check_batch(*te, in))) return rc;
    const DevTopo &g = te->gen;
    const int64_t W = in->n_windows;
    if (W == 0) return 0;
    if (!linearize && window_path_ok(ctx, *te, W, in, out)) return host_window_path(ctx, *te, dc, in, out);
    if (ctx->ws_pending) { /* a device-API call may still own lane 0's workspace */
        CU(cudaEventSynchronize(ctx->ws_free));
        ctx->ws_pending = false;
    }
    const double *d_ant = nullptr;
    if ((rc = upload_ant(ctx, g, in, ctx->lane[0].st, &d_ant))) return rc;

    const size_t N = (size_t)g.N;
    const uwbgo_range_msgs *msgs = in->range_msgs;
    const bool shared_anch = (in->shared & UWBGO_SHARED_ANCHORS) != 0;
    const size_t info_n = (in->shared & UWBGO_DIAG_INFO) ? 6 : 36; /* doubles per information matrix on the wire */
    /* the LM kernel's duration is set by per-window latency, not by the chunk size, so a batch
     * that fits n_lanes chunks is split evenly and all its chunks run concurrently */
    int64_t chunk = std::min<int64_t>(ctx->chunk, (W + 31) / 32 * 32);
    if (W <= ctx->chunk * ctx->n_lanes) chunk = std::max<int64_t>(32, ((W + ctx->n_lanes - 1) / ctx->n_lanes + 31) / 32 * 32);
    /* 6x6 windows: the CTA-per-tile kernel holds 2 tiles per SM and a tile takes the whole kernel time
     * whatever the launch size, so up to one device-full of tiles a split only delays the last chunk's
     * start by the copies ahead of it (C4a, 8,192 windows: 14.0 ms in one chunk, 18.3 ms in eight) */
    const bool six = !(te->fast_ok && in->pose_R == nullptr);
    if (six && !ctx->pipeline_set && W <= 2 * 148 * TILE) chunk = (W + 31) / 32 * 32;
    StageLayout S;
    {
        size_t o = 0;
        auto take = [&](bool need, size_t per_window) {
            size_t at = o;
            if (need) o += ((size_t)chunk * per_window + 255) & ~(size_t)255;
            return at;
        };
        S.in_pose_t = take(true, N * 3 * 8);
        S.in_pose_R = take(in->pose_R != nullptr, N * 9 * 8);
        S.in_cnt = take(in->oplus_count != nullptr, N * 4);
        S.in_anch = take(g.A > 0, (size_t)g.A * 3 * 8); /* (shared anchors use the first A*3 doubles of it) */
        S.in_rd = take(g.Er > 0 && !msgs, (size_t)g.Er * 8);
        S.in_ri = take(g.Er > 0 && !msgs, (size_t)g.Er * 8);
        S.in_md = take(msgs && g.Era > 0, (size_t)g.Era * 4);
        S.in_me = take(msgs && g.Era > 0, (size_t)g.Era * 4);
        S.in_mta = take(msgs && g.Era > 0 && msgs->dt_anchor, (size_t)g.Era * 8);
        S.in_mtp = take(msgs && g.Erp > 0, (size_t)g.Erp * 8);
        S.in_pZ = take(g.Ep > 0, (size_t)g.Ep * 12 * 8);
        S.in_pI = take(g.Ep > 0, (size_t)g.Ep * info_n * 8);
        S.in_sZ = take(g.Es > 0, (size_t)g.Es * 12 * 8);
        S.in_sI = take(g.Es > 0, (size_t)g.Es * info_n * 8);
        if (linearize) {
            S.out_Hd = take(true, N * 36 * 8);
            S.out_Ho = take(N > 1, (N - 1) * 36 * 8);
            S.out_b = take(true, N * 6 * 8);
            S.out_chi2 = take(chi2 != nullptr, 2 * 8);
        } else {
            S.out_pose_t = take(true, N * 3 * 8);
            S.out_pose_R = take(out->pose_R != nullptr, N * 9 * 8);
            S.out_cnt = take(out->oplus_count != nullptr, N * 4);
            S.out_chi2 = take(out->chi2 != nullptr, 4 * 8);
            S.out_status = take(out->status != nullptr, 4 * 4);
            S.out_echi = take(out->edge_chi2 != nullptr && g.E > 0, (size_t)g.E * 8);
            S.out_marg = take(out->marginal != nullptr, 36 * 8);
            S.out_mok = take(out->marginal != nullptr && out->marginal_ok != nullptr, 4);
        }
        S.bytes = o;
    }
    const int n_lanes = (int)std::min<int64_t>(ctx->n_lanes, (W + chunk - 1) / chunk);
    for (int k = 0; k < n_lanes; ++k)
        if ((rc = ctx->lane[k].stage.reserve(S.bytes))) return rc;

    /* inside the chunk loop an error must not return: earlier chunks may still have copies in flight against the
     * caller's buffers.  Every error path falls through to the per-lane synchronisation below */
#define CUL(call)                                        \
    {                                                    \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) {                        \
            first_err = fail_cuda(e__, #call);           \
            break;                                       \
        }                                                \
    }
    int first_err = 0;
    int64_t c = 0;
    /* UWBGO_PIPE_TRACE=1: per-chunk event times of this call on stderr (diagnosis only) */
    static const bool trace = getenv("UWBGO_PIPE_TRACE") != nullptr;
    static const bool ordered_h2d = !(getenv("UWBGO_PIPE_H2D") && !strcmp(getenv("UWBGO_PIPE_H2D"), "lanes"));
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    if (trace) {
        for (int k = 0; k < n_lanes; ++k) cudaStreamSynchronize(ctx->lane[k].st);
        mark(ctx->lane[0].st);
    }
    /* Results bound for PAGEABLE host memory (a std::vector of the caller's; uwbgo_host_alloc gives page-locked
     * arrays): such a copy blocks the calling thread until it is done, i.e. until the chunk's kernel has finished,
     * so issued inside the loop it would keep the next chunk from even being launched and the chunks would run one
     * after the other (148 UWB-only windows of 50 poses: 9.1 ms instead of 2.0 ms).  While every chunk has a lane of
     * its own, those copies are issued after the loop, once all chunks are in flight */
    struct D2HJob {
        void *dst;
        const void *src;
        size_t bytes;
        cudaStream_t st;
    };
    std::vector<D2HJob> deferred;
    bool defer = false;
    if (!trace && (W + chunk - 1) / chunk <= n_lanes) {
        cudaPointerAttributes pa{};
        const void *probe = linearize ? static_cast<const void *>(H_diag) : static_cast<const void *>(out->pose_t);
        if (cudaPointerGetAttributes(&pa, probe) != cudaSuccess) {
            cudaGetLastError();
            defer = true;
        } else {
            defer = pa.type == cudaMemoryTypeUnregistered;
        }
    }
    ctx->batch_windows = W; /* the chunks run concurrently: kernels choose their shape by the whole batch */
    for (int64_t w0 = 0; w0 < W; w0 += chunk, ++c) {
        Lane &ln = ctx->lane[c % n_lanes];
        const int64_t wc = std::min<int64_t>(chunk, W - w0);
        char *sb = static_cast<char *>(ln.stage.p);
        cudaStream_t st = ln.st;
        /* Inputs cross the link in chunk order: issued independently from the lanes' streams, the copies of all
         * chunks share it, every chunk arrives late and the device idles meanwhile (C3: first chunk on the
         * device after 0.86 ms instead of 0.3 ms).  A lane's copies therefore wait for the event the previous
         * chunk's lane records after its own (a stream of its own for the inputs aliases a lane's hardware
         * queue once there are more streams than connections and stalls behind that lane's kernels) */
        cudaStream_t sin = st;
        if (ordered_h2d && c > 0) CUL(cudaStreamWaitEvent(st, ctx->lane[(c - 1) % n_lanes].in_ready, 0));
        auto h2d = [&](const void *src, size_t off, size_t per_window) -> cudaError_t {
            if (!src) return cudaSuccess;
            return cudaMemcpyAsync(sb + off, static_cast<const char *>(src) + (size_t)w0 * per_window,
                                   (size_t)wc * per_window, cudaMemcpyHostToDevice, sin);
        };
        auto d2h = [&](void *dst, size_t off, size_t per_window) -> cudaError_t {
            if (!dst) return cudaSuccess;
            if (defer) {
                deferred.push_back(D2HJob{static_cast<char *>(dst) + (size_t)w0 * per_window, sb + off, (size_t)wc * per_window, st});
                return cudaSuccess;
            }
            return cudaMemcpyAsync(static_cast<char *>(dst) + (size_t)w0 * per_window, sb + off,
                                   (size_t)wc * per_window, cudaMemcpyDeviceToHost, st);
        };
        CUL(h2d(in->pose_t, S.in_pose_t, N * 3 * 8));
        CUL(h2d(in->pose_R, S.in_pose_R, N * 9 * 8));
        CUL(h2d(in->oplus_count, S.in_cnt, N * 4));
        if (g.A > 0 && shared_anch) { /* one constellation: A*3 doubles per chunk instead of per window */
            CUL(cudaMemcpyAsync(sb + S.in_anch, in->anchors, (size_t)g.A * 3 * 8, cudaMemcpyHostToDevice, sin));
        } else {
            CUL(h2d(g.A > 0 ? in->anchors : nullptr, S.in_anch, (size_t)g.A * 3 * 8));
        }
        uwbgo_range_msgs dm{};
        if (msgs) {
            CUL(h2d(g.Era > 0 ? msgs->distance : nullptr, S.in_md, (size_t)g.Era * 4));
            CUL(h2d(g.Era > 0 ? msgs->distance_err : nullptr, S.in_me, (size_t)g.Era * 4));
            CUL(h2d(g.Era > 0 ? msgs->dt_anchor : nullptr, S.in_mta, (size_t)g.Era * 8));
            CUL(h2d(g.Erp > 0 ? msgs->dt_pose : nullptr, S.in_mtp, (size_t)g.Erp * 8));
            dm.distance = reinterpret_cast<float *>(sb + S.in_md);
            dm.distance_err = reinterpret_cast<float *>(sb + S.in_me);
            dm.dt_anchor = msgs->dt_anchor ? reinterpret_cast<double *>(sb + S.in_mta) : nullptr;
            dm.dt_pose = reinterpret_cast<double *>(sb + S.in_mtp);
            dm.v_max = msgs->v_max;
        } else {
            CUL(h2d(g.Er > 0 ? in->range_d : nullptr, S.in_rd, (size_t)g.Er * 8));
            CUL(h2d(g.Er > 0 ? in->range_info : nullptr, S.in_ri, (size_t)g.Er * 8));
        }
        CUL(h2d(g.Ep > 0 ? in->prior_Z : nullptr, S.in_pZ, (size_t)g.Ep * 12 * 8));
        CUL(h2d(g.Ep > 0 ? in->prior_info : nullptr, S.in_pI, (size_t)g.Ep * info_n * 8));
        CUL(h2d(g.Es > 0 ? in->se3_Z : nullptr, S.in_sZ, (size_t)g.Es * 12 * 8));
        CUL(h2d(g.Es > 0 ? in->se3_info : nullptr, S.in_sI, (size_t)g.Es * info_n * 8));
        if (ordered_h2d) CUL(cudaEventRecord(ln.in_ready, st));
        mark(st);
        uwbgo_batch db{};
        db.n_windows = wc;
        db.pose_t = reinterpret_cast<double *>(sb + S.in_pose_t);
        db.pose_R = in->pose_R ? reinterpret_cast<double *>(sb + S.in_pose_R) : nullptr;
        db.oplus_count = in->oplus_count ? reinterpret_cast<int32_t *>(sb + S.in_cnt) : nullptr;
        db.anchors = reinterpret_cast<double *>(sb + S.in_anch);
        db.ant_offsets = in->ant_offsets;
        db.range_d = msgs ? nullptr : reinterpret_cast<double *>(sb + S.in_rd);
        db.range_info = msgs ? nullptr : reinterpret_cast<double *>(sb + S.in_ri);
        db.range_msgs = msgs ? &dm : nullptr;
        db.shared = in->shared;
        db.prior_Z = reinterpret_cast<double *>(sb + S.in_pZ);
        db.prior_info = reinterpret_cast<double *>(sb + S.in_pI);
        db.se3_Z = reinterpret_cast<double *>(sb + S.in_sZ);
        db.se3_info = reinterpret_cast<double *>(sb + S.in_sI);
        if (linearize) {
            StageOut so{reinterpret_cast<double *>(sb + S.out_Hd),
                        reinterpret_cast<double *>(sb + S.out_Ho),
                        reinterpret_cast<double *>(sb + S.out_b),
                        chi2 ? reinterpret_cast<double *>(sb + S.out_chi2) : nullptr};
            rc = run_device(ctx, ln, *te, dc, &db, d_ant, nullptr, &so, st);
            if (rc) { first_err = rc; break; }
            CUL(d2h(H_diag, S.out_Hd, N * 36 * 8));
            if (N > 1) CUL(d2h(H_off, S.out_Ho, (N - 1) * 36 * 8));
            CUL(d2h(b, S.out_b, N * 6 * 8));
            CUL(d2h(chi2, S.out_chi2, 2 * 8));
        } else {
            uwbgo_result dr{};
            dr.pose_t = reinterpret_cast<double *>(sb + S.out_pose_t);
            dr.pose_R = out->pose_R ? reinterpret_cast<double *>(sb + S.out_pose_R) : nullptr;
            dr.oplus_count = out->oplus_count ? reinterpret_cast<int32_t *>(sb + S.out_cnt) : nullptr;
            dr.chi2 = out->chi2 ? reinterpret_cast<double *>(sb + S.out_chi2) : nullptr;
            dr.status = out->status ? reinterpret_cast<int32_t *>(sb + S.out_status) : nullptr;
            dr.edge_chi2 = (out->edge_chi2 && g.E > 0) ? reinterpret_cast<double *>(sb + S.out_echi) : nullptr;
            dr.marginal = out->marginal ? reinterpret_cast<double *>(sb + S.out_marg) : nullptr;
            dr.marginal_ok = (out->marginal && out->marginal_ok) ? reinterpret_cast<int32_t *>(sb + S.out_mok) : nullptr;
            rc = run_device(ctx, ln, *te, dc, &db, d_ant, &dr, nullptr, st);
            if (rc) { first_err = rc; break; }
            mark(st);
            CUL(d2h(out->pose_t, S.out_pose_t, N * 3 * 8));
            CUL(d2h(out->pose_R, S.out_pose_R, N * 9 * 8));
            CUL(d2h(out->oplus_count, S.out_cnt, N * 4));
            CUL(d2h(out->chi2, S.out_chi2, 4 * 8));
            CUL(d2h(out->status, S.out_status, 4 * 4));
            CUL(d2h(g.E > 0 ? out->edge_chi2 : nullptr, S.out_echi, (size_t)g.E * 8));
            CUL(d2h(out->marginal, S.out_marg, 36 * 8));
            CUL(d2h(out->marginal ? out->marginal_ok : nullptr, S.out_mok, 4));
            mark(st);
        }
    }
    ctx->batch_windows = 0;
    if (!first_err)
        for (const D2HJob &j : deferred) {
            cudaError_t e = cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyDeviceToHost, j.st);
            if (e != cudaSuccess) {
                first_err = fail_cuda(e, "cudaMemcpyAsync (results)");
                break;
            }
        }
    for (int k = 0; k < n_lanes; ++k) {
        cudaError_t e = cudaStreamSynchronize(ctx->lane[k].st);
        if (e != cudaSuccess && !first_err) first_err = fail_cuda(e, "cudaStreamSynchronize");
    }
    if (trace && !linearize && !first_err && tev.size() == 1 + 3 * (size_t)c) {
        fprintf(stderr, "pipe trace: %lld windows, chunk %lld, %d lanes (ms after the start: inputs on the device, results ready, results on the host)\n",
                (long long)W, (long long)chunk, n_lanes);
        for (int64_t k = 0; k < c; ++k) {
            float a = 0, b2 = 0, d = 0;
            cudaEventElapsedTime(&a, tev[0], tev[1 + 3 * k]);
            cudaEventElapsedTime(&b2, tev[0], tev[2 + 3 * k]);
            cudaEventElapsedTime(&d, tev[0], tev[3 + 3 * k]);
            fprintf(stderr, "  chunk %2lld  h2d %7.3f  solved %7.3f  d2h %7.3f\n", (long long)k, a, b2, d);
        }
    }
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
#undef CUL
    return first_err;
}

int uwbgo_solve_batch(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                      const uwbgo_config *cfg, uwbgo_result *out)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (!out || (in && in->n_windows > 0 && !out->pose_t)) return fail(UWBGO_E_INVALID, "result.pose_t is NULL");
    return host_pipeline(ctx, topo, in, cfg, out, nullptr, nullptr, nullptr, nullptr, false);
}

int uwbgo_linearize_batch(uwbgo_ctx *ctx, const uwbgo_topology *topo, const uwbgo_batch *in,
                          const uwbgo_config *cfg, double *H_diag, double *H_off, double *b,
                          double *chi2)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (!H_diag || !b || (topo && topo->n_poses > 1 && !H_off))
        return fail(UWBGO_E_INVALID, "output array is NULL");
    return host_pipeline(ctx, topo, in, cfg, nullptr, H_diag, H_off, b, chi2, true);
}

int uwbgo_factor_solve_batch(uwbgo_ctx *ctx, int32_t n_poses, int64_t n_windows,
                             const double *H_diag, const double *H_off, const double *b,
                             const double *lambda, double *x, int32_t *ok)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    if (n_poses < 1 || n_windows < 0) return fail(UWBGO_E_INVALID, "bad sizes");
    if (n_windows == 0) return 0;
    if (!H_diag || !b || !lambda || !x || (n_poses > 1 && !H_off))
        return fail(UWBGO_E_INVALID, "array is NULL");
    CU(cudaSetDevice(ctx->device));
    if (ctx->ws_pending) {
        CU(cudaEventSynchronize(ctx->ws_free));
        ctx->ws_pending = false;
    }
    const size_t N = (size_t)n_poses, W = (size_t)n_windows;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_Hd = 0, o_Ho = o_Hd + al(W * N * 36 * 8), o_b = o_Ho + al(W * (N - 1) * 36 * 8 + 8),
           o_l = o_b + al(W * N * 6 * 8), o_x = o_l + al(W * 8), o_ok = o_x + al(W * N * 6 * 8),
           total = o_ok + al(W * 4);
    int rc = ctx->misc.reserve(total);
    if (rc) return rc;
    char *m = static_cast<char *>(ctx->misc.p);
    cudaStream_t st = ctx->lane[0].st;
    CU(cudaMemcpyAsync(m + o_Hd, H_diag, W * N * 36 * 8, cudaMemcpyHostToDevice, st));
    if (N > 1) CU(cudaMemcpyAsync(m + o_Ho, H_off, W * (N - 1) * 36 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m + o_b, b, W * N * 6 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m + o_l, lambda, W * 8, cudaMemcpyHostToDevice, st));
    if ((rc = ctx->lane[0].tile.reserve(factor_solve_scratch_bytes(n_poses, n_windows)))) return rc;
    CU(launch_factor_solve(n_poses, n_windows, reinterpret_cast<double *>(m + o_Hd),
                           reinterpret_cast<double *>(m + o_Ho), reinterpret_cast<double *>(m + o_b),
                           reinterpret_cast<double *>(m + o_l), reinterpret_cast<double *>(m + o_x),
                           reinterpret_cast<int32_t *>(m + o_ok),
                           static_cast<double *>(ctx->lane[0].tile.p), st));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(x, m + o_x, W * N * 6 * 8, cudaMemcpyDeviceToHost, st));
    if (ok) CU(cudaMemcpyAsync(ok, m + o_ok, W * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int uwbgo_set_profiling(uwbgo_ctx *ctx, int on)
{
    if (!ctx) return fail(UWBGO_E_INVALID, "ctx is NULL");
    ctx->profile = on != 0;
    ctx->k_count = 0;
    return 0;
}

double uwbgo_mean_kernel_ms(uwbgo_ctx *ctx, int last_n)
{
    if (!ctx || ctx->k_count <= 0 || last_n <= 0) return -1.0;
    const int64_t n = std::min<int64_t>(std::min<int64_t>(last_n, ctx->k_count), uwbgo_ctx::K_RING);
    double sum = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        const int slot = (int)((ctx->k_count - 1 - j) % uwbgo_ctx::K_RING);
        if (cudaEventSynchronize(ctx->k1[slot]) != cudaSuccess) return -1.0;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->k0[slot], ctx->k1[slot]) != cudaSuccess) return -1.0;
        sum += (double)ms;
    }
    return sum / (double)n;
}

double uwbgo_last_kernel_ms(uwbgo_ctx *ctx) { return uwbgo_mean_kernel_ms(ctx, 1); }

int64_t uwbgo_launch_count(const uwbgo_ctx *ctx) { return ctx ? ctx->launches : 0; }
int uwbgo_last_path(const uwbgo_ctx *ctx) { return ctx ? ctx->last_path : 0; }

double uwbgo_measure_fp64_peak(uwbgo_ctx *ctx, double *elapsed_ms)
{
    if (!ctx) return 0.0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return 0.0;
    const int iters = 1 << 16;
    int blocks = 0, threads = 0;
    if (ctx->misc.reserve(sizeof(double) * 148 * 16 * 256)) return 0.0;
    cudaStream_t st = ctx->lane[0].st;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch_fp64_peak(static_cast<double *>(ctx->misc.p), 1 << 10, st, &blocks, &threads); /* warm-up */
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        launch_fp64_peak(static_cast<double *>(ctx->misc.p), iters, st, &blocks, &threads);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    ctx->launches += 4;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (cudaGetLastError() != cudaSuccess || !(best > 0.f)) return 0.0;
    if (elapsed_ms) *elapsed_ms = best;
    double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    return flops / (best * 1e-3);
}

int uwbgo_selftest_math(uwbgo_ctx *ctx, uint64_t seed, int64_t n_operands, int mode, int64_t counts[3])
{
    if (!ctx || !counts || n_operands <= 0 || (mode != 0 && mode != 1)) return fail(UWBGO_E_INVALID, "selftest_math: bad argument");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(UWBGO_E_CUDA, "cudaSetDevice failed");
    if (ctx->misc.reserve(3 * sizeof(unsigned long long))) return fail(UWBGO_E_NOMEM, "selftest_math: workspace");
    cudaStream_t st = ctx->lane[0].st;
    unsigned long long *d = static_cast<unsigned long long *>(ctx->misc.p);
    const int per_thread = 64, blocks = (int)std::min<int64_t>((n_operands + 256LL * per_thread * 4 - 1) / (256LL * per_thread * 4), 1 << 20);
    unsigned long long h[3] = {0, 0, 0};
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof h, st);
    if (e == cudaSuccess) e = launch_math_selftest(seed, blocks, per_thread, mode, d, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    ctx->launches += 1;
    if (e != cudaSuccess) return fail(UWBGO_E_CUDA, cudaGetErrorString(e));
    for (int k = 0; k < 3; ++k) counts[k] = (int64_t)h[k];
    return UWBGO_OK;
}

}  // extern "C"
