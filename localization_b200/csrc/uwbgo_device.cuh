/*
 * uwbgo_device.cuh — small device helpers shared by every kernel path: tile-layout row access,
 * prefetch, edge-table loads, the Cauchy kernel, the thread-private view of the workspace, and the
 * developer knobs of A/B builds.
 */
#ifndef UWBGO_DEVICE_CUH
#define UWBGO_DEVICE_CUH

#include "uwbgo_internal.h"
#include "uwbgo_math.cuh"

namespace uwbgo {

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */
#define ROW(p, r) ((p)[(size_t)(r) * TILE])

/* developer knobs for A/B builds (see DESIGN.md "prefetch") */
#ifndef UWBGO_FACTOR_PF
#define UWBGO_FACTOR_PF 1 /* 0 none, 1 register double buffer, 2 L2 prefetch */
#endif
#ifndef UWBGO_SOLVE_REGPF
#define UWBGO_SOLVE_REGPF 1 /* L record of the next pose prefetched into registers */
#endif
#ifndef UWBGO_L2PF_DIST
#define UWBGO_L2PF_DIST 2 /* > 0: prefetch.global.L2 this many records ahead of the sweeps */
#endif

#ifndef UWBGO_CHAIN_UNROLL
#define UWBGO_CHAIN_UNROLL 1 /* unroll factor of the CHAIN sweeps' pose loops */
#endif
#define UWBGO_PRAGMA_(x) _Pragma(#x)
#define UWBGO_PRAGMA(x) UWBGO_PRAGMA_(x)
#define UWBGO_CHAIN_UNROLL_PRAGMA UWBGO_PRAGMA(unroll UWBGO_CHAIN_UNROLL)

UWBGO_DI void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <int ROWS>
UWBGO_DI void prefetch_rows_l2(const double *p)
{
#pragma unroll
    for (int k = 0; k < ROWS; ++k) prefetch_l2(p + (size_t)k * TILE);
}

UWBGO_DI EdgeRec load_edge(const EdgeRec *e)
{
    const int4 *p = reinterpret_cast<const int4 *>(e);
    int4 u = __ldg(p), v = __ldg(p + 1);
    EdgeRec r;
    r.kind = u.x; r.a = u.y; r.b = u.z; r.slot = u.w;
    r.ant = v.x; r.robust = v.y; r.base_a = v.z; r.base_b = v.w;
    r.ant_b = 0; /* the translation-only paths never carry offsets; the general path reads it below */
    return r;
}

/* general path: the record with its vertex-1 antenna number */
UWBGO_DI EdgeRec load_edge_full(const EdgeRec *e)
{
    EdgeRec r = load_edge(e);
    r.ant_b = __ldg(&e->ant_b);
    return r;
}

struct Cauchy {
    double dsqr, dsqrReci;
    UWBGO_DI void init(double delta)
    {
        dsqr = delta * delta;
        dsqrReci = 1.0 / dsqr;
    }
    UWBGO_DI double rho0(double e2) const { return dsqr * det_log(dsqrReci * e2 + 1.0); }
    UWBGO_DI double rho1(double e2) const { return 1.0 / (dsqrReci * e2 + 1.0); }
    template <class M>
    UWBGO_DI double rho0m(double e2, unsigned &bad) const { return dsqr * M::log_(dsqrReci * e2 + 1.0, bad); }
    template <class M>
    UWBGO_DI double rho1m(double e2, unsigned &bad) const { return M::rcp(dsqrReci * e2 + 1.0, bad); }
};

/* thread-private view of the tile-layout workspace */
struct Ptrs {
    double *T0, *T1;   /* translations [N*3] rows, two buffers (selected with ?: so the struct */
    double *Rm0, *Rm1; /* rotations    [N*9] rows (GENERAL)      never needs a local-memory copy) */
    UWBGO_DI double *T(int k) const { return k ? T1 : T0; }
    UWBGO_DI double *Rm(int k) const { return k ? Rm1 : Rm0; }
    int32_t *cnt;
    const double *anch, *rd, *ri, *pZ, *pI, *sZ, *sI;
    double *HB, *LR;
};

template <int HR, int LRR>
UWBGO_DI Ptrs thread_ptrs(const DevTopo &tp, const DevWs &ws, int64_t w)
{
    int64_t tile = w / TILE;
    int lane = (int)(w % TILE);
    Ptrs p;
    p.T0 = ws.T[0] + (tile * (size_t)tp.N * 3) * TILE + lane;
    p.T1 = ws.T[1] + (tile * (size_t)tp.N * 3) * TILE + lane;
    p.Rm0 = ws.Rm[0] ? ws.Rm[0] + (tile * (size_t)tp.N * 9) * TILE + lane : nullptr;
    p.Rm1 = ws.Rm[1] ? ws.Rm[1] + (tile * (size_t)tp.N * 9) * TILE + lane : nullptr;
    p.cnt = ws.cnt ? ws.cnt + (tile * (size_t)tp.N) * TILE + lane : nullptr;
    p.anch = ws.anch ? ws.anch + (tile * (size_t)tp.A * 3) * TILE + lane : nullptr;
    p.rd = ws.rd ? ws.rd + (tile * (size_t)tp.Er) * TILE + lane : nullptr;
    p.ri = ws.ri ? ws.ri + (tile * (size_t)tp.Er) * TILE + lane : nullptr;
    p.pZ = ws.pZ ? ws.pZ + (tile * (size_t)tp.Ep * 12) * TILE + lane : nullptr;
    p.pI = ws.pI ? ws.pI + (tile * (size_t)tp.Ep * 36) * TILE + lane : nullptr;
    p.sZ = ws.sZ ? ws.sZ + (tile * (size_t)tp.Es * 12) * TILE + lane : nullptr;
    p.sI = ws.sI ? ws.sI + (tile * (size_t)tp.Es * 36) * TILE + lane : nullptr;
    p.HB = ws.HB + (tile * (size_t)tp.N * HR) * TILE + lane;
    p.LR = ws.LR ? ws.LR + (tile * (size_t)tp.N * (tp.tree ? LR_TREE : LRR)) * TILE + lane : nullptr;
    return p;
}

}  // namespace uwbgo
#endif
