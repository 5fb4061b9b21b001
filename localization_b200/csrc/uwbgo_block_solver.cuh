/*
 * uwbgo_block_solver.cuh — the linear solver of the LM trials: block Cholesky of H + lambda I over a
 * chain of poses, newest pose first, one window per thread (replaces LinearSolverCholmod::solve,
 * reference src/localization/localization.h:84).
 */
#ifndef UWBGO_BLOCK_SOLVER_CUH
#define UWBGO_BLOCK_SOLVER_CUH

#include "uwbgo_device.cuh"

namespace uwbgo {

/* ------------------------------------------------------------------------------------------ */
/* linear solver: block-tridiagonal Cholesky of H + lambda I, chain eliminated newest pose       */
/* first (replaces LinearSolverCholmod::solve).  D = 3 (FAST) or 6 (GENERAL).                   */
/*   H record of pose i:  Hd_i upper packed | H_{i-1,i} (rows i-1, cols i) | b_i                 */
/*   L record of pose i:  c_i | M_i      with the substitution  x_i = c_i - M_i x_{i-1}          */
/* The factor sweep walks the H records back to front and prefetches record i-1 into registers   */
/* while record i is being eliminated (D = 3), so the HBM latency of the stream hides behind     */
/* the sqrt/div dependency chain of the 3x3 potrf.                                               */
/* ------------------------------------------------------------------------------------------ */
template <int D>
struct Rec {
    static constexpr int TRI = D * (D + 1) / 2, SQ = D * D;
    static constexpr int H = TRI + SQ + D; /* rows of an H record */
    static constexpr int L = D + SQ;       /* rows of an L record */
};

template <int D>
UWBGO_DI void load_hrec(const double *__restrict__ h, double *r)
{
#pragma unroll
    for (int k = 0; k < Rec<D>::H; ++k) r[k] = ROW(h, k);
}

/* one elimination step on the H record held in `h`; G/zn carry G_i and z_{i+1} in, G_{i-1} and
 * z_i out */
template <int D, class M = IeeeMath>
UWBGO_DI void factor_step(const double *h, double *__restrict__ l, bool link, bool has_prev,
                          double lambda, double *G, double *zn, bool &ok, unsigned *badp = nullptr)
{
    unsigned bad_local = 0;
    unsigned &bad = badp ? *badp : bad_local;
    constexpr int TRI = Rec<D>::TRI, SQ = Rec<D>::SQ;
    double S[TRI], L[TRI], z[D], c[D];
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) {
            double s = h[up_idx(D, cc, r)];
            if (r == cc) s = s + lambda;
            if (link) {
#pragma unroll
                for (int k = 0; k < D; ++k) s = fma(-G[r * D + k], G[cc * D + k], s);
            }
            S[lo_idx(r, cc)] = s;
        }
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double s = S[lo_idx(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[lo_idx(j, k)], L[lo_idx(j, k)], s);
        if (!(s > 0.0)) ok = false;
        double inv = M::rcp(M::sqrt_(s, bad), bad);
        L[lo_idx(j, j)] = inv;
#pragma unroll
        for (int r = j + 1; r < D; ++r) {
            double t = S[lo_idx(r, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[lo_idx(r, k)], L[lo_idx(j, k)], t);
            L[lo_idx(r, j)] = t * inv;
        }
    }
#pragma unroll
    for (int r = 0; r < D; ++r) {
        double s = h[TRI + SQ + r];
        if (link) {
#pragma unroll
            for (int k = 0; k < D; ++k) s = fma(-G[r * D + k], zn[k], s);
        }
#pragma unroll
        for (int k = 0; k < r; ++k) s = fma(-L[lo_idx(r, k)], z[k], s);
        z[r] = s * L[lo_idx(r, r)];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) zn[k] = z[k];
#pragma unroll
    for (int r = D - 1; r >= 0; --r) {
        double s = z[r];
#pragma unroll
        for (int k = r + 1; k < D; ++k) s = fma(-L[lo_idx(k, r)], c[k], s);
        c[r] = s * L[lo_idx(r, r)];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) ROW(l, k) = c[k];
    if (has_prev) {
        double M[SQ];
#pragma unroll
        for (int r = 0; r < D; ++r)
#pragma unroll
            for (int cc = 0; cc < D; ++cc) {
                double s = h[TRI + r * D + cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) s = fma(-G[r * D + k], L[lo_idx(cc, k)], s);
                G[r * D + cc] = s * L[lo_idx(cc, cc)]; /* row r: entries k < cc are already new */
            }
#pragma unroll
        for (int j = 0; j < D; ++j)
#pragma unroll
            for (int r = D - 1; r >= 0; --r) {
                double s = G[j * D + r];
#pragma unroll
                for (int k = r + 1; k < D; ++k) s = fma(-L[lo_idx(k, r)], M[k * D + j], s);
                M[r * D + j] = s * L[lo_idx(r, r)];
            }
#pragma unroll
        for (int k = 0; k < SQ; ++k) ROW(l, D + k) = M[k];
    }
}

template <int D, class M = IeeeMath>
UWBGO_DI bool factor_sweep(const double *__restrict__ HB, double *__restrict__ LR, int N,
                           double lambda, unsigned *badp = nullptr)
{
    constexpr int SQ = Rec<D>::SQ, RH = Rec<D>::H, RL = Rec<D>::L;
    double G[SQ], zn[D];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < SQ; ++k) G[k] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) zn[k] = 0.0;
    if (D == 3 && UWBGO_FACTOR_PF == 1) {
        double ra[RH], rb[RH];
        int i = N - 1;
        load_hrec<D>(HB + (size_t)i * RH * TILE, ra);
        while (i >= 0) {
            if (i > 0) load_hrec<D>(HB + (size_t)(i - 1) * RH * TILE, rb);
            factor_step<D, M>(ra, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok, badp);
            --i;
            if (i < 0) break;
            if (i > 0) load_hrec<D>(HB + (size_t)(i - 1) * RH * TILE, ra);
            factor_step<D, M>(rb, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok, badp);
            --i;
        }
    } else {
        for (int i = N - 1; i >= 0; --i) {
            double r[RH];
            if (D == 3 && UWBGO_L2PF_DIST > 0 && i - UWBGO_L2PF_DIST >= 0)
                prefetch_rows_l2<RH>(HB + (size_t)(i - UWBGO_L2PF_DIST) * RH * TILE);
            load_hrec<D>(HB + (size_t)i * RH * TILE, r);
            factor_step<D, M>(r, LR + (size_t)i * RL * TILE, i + 1 < N, i > 0, lambda, G, zn, ok, badp);
        }
    }
    return ok;
}

/* x_i = c_i - M_i x_{i-1}; l = L record values (registers); xp holds x_{i-1} in, x_i out */
template <int D>
UWBGO_DI void subst_step(const double *l, bool link, double *xp)
{
    double x[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
        double s = l[r];
        if (link) {
#pragma unroll
            for (int j = 0; j < D; ++j) s = fma(-l[D + r * D + j], xp[j], s);
        }
        x[r] = s;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) xp[k] = x[k];
}

}  // namespace uwbgo
#endif
