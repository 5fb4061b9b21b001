// Microbenchmark: HBM write bandwidth of strided-segment stores (the output pattern of the linearise stage).
// Each warp owns 32 "windows"; window w of pose i has a SEG-byte segment at offset (w * N + i) * SEG_UNIT;
// variant R writes R consecutive poses (R * 288 B contiguous per window) per pass.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// mode 0: per (tile, pose): for each of 32 windows, lanes 0..17 write 16 B each (288 B row)
// mode 1: per (tile, run of R poses): for each window, all 32 lanes write 16 B chunks over R*288 bytes
template <int MODE>
__global__ void k(double2 *out, int N, int R, long tiles, int streaming)
{
    const int lane = threadIdx.x & 31;
    const long item = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int runs = (N + R - 1) / R;
    if (item >= tiles * runs) return;
    const long tile = item / runs;
    const int run = (int)(item % runs);
    const int i0 = run * R, i1 = min(N, i0 + R);
    const double2 v = make_double2((double)item, (double)lane);
    if (MODE == 0) {
        for (int i = i1 - 1; i >= i0; --i)
            for (int wl = 0; wl < 32; ++wl) {
                double2 *p = out + ((tile * 32 + wl) * N + i) * 18 + lane;
                if (lane < 18) { if (streaming) __stcs(p, v); else *p = v; }
            }
    } else {
        const int chunks = (i1 - i0) * 18;
        for (int wl = 0; wl < 32; ++wl) {
            double2 *p = out + ((tile * 32 + wl) * N + i0) * 18;
            for (int c = lane; c < chunks; c += 32) { if (streaming) __stcs(p + c, v); else p[c] = v; }
        }
    }
}

// mode 2: the four fragments the linearise kernel writes per (window, pose): H_diag 288 B, H_off 224 + 64 B, b 48 B;
// one pose per warp; launch bounds pin the occupancy (OCC CTAs of 32 threads per SM) like the real kernel's registers do
template <int OCC>
__global__ void __launch_bounds__(32, OCC) k4(double2 *Hd, double2 *Ho, double2 *B, int N, long tiles)
{
    extern __shared__ double dyn[]; // occupancy limiter
    const int lane = threadIdx.x & 31;
    const long item = blockIdx.x;
    if (item >= tiles * N) return;
    const long tile = item / N;
    const int i = (int)(item % N);
    const double2 v = make_double2((double)item, (double)lane);
#pragma unroll 8
    for (int wl = 0; wl < 32; ++wl) {
        const long w = tile * 32 + wl;
        double2 *pa = lane < 18 ? Hd + (w * N + i) * 18 + lane : Ho + (w * N + i) * 18 + (lane - 18);
        *pa = v;
        if (lane < 4) Ho[(w * N + i) * 18 + 14 + lane] = v;
        else if (lane < 7) B[(w * N + i) * 3 + (lane - 4)] = v;
    }
}

// mode 3: the same four-fragment output, but through the bulk-copy engine: each warp fills a window-major panel
// [32 windows][78 doubles] in shared memory, then every lane hands its window's three rows (288 + 288 + 48 B) to
// cp.async.bulk (shared -> global).  WARPS warps per CTA, panel per warp 19,968 B.
__device__ __forceinline__ unsigned saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) kbulk(double *Hd, double *Ho, double *B, int N, long tiles, int items_per_warp)
{
    extern __shared__ __align__(16) double panel_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *panel = panel_all + (size_t)warp * 32 * 78;
    double *mine = panel + lane * 78;
    for (int k = 0; k < 78; ++k) mine[k] = 0.0;
    const long first = ((long)blockIdx.x * WARPS + warp) * items_per_warp;
    for (int it = 0; it < items_per_warp; ++it) {
        const long item = first + it;
        if (item >= tiles * N) break;
        const long tile = item / N;
        const int i = (int)(item % N);
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        for (int k = 0; k < 21; ++k) mine[(k * 7) % 78] = (double)(item + k);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const long w = tile * 32 + lane;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Hd + (w * N + i) * 36), "r"(saddr(mine)), "r"(288) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(Ho + (w * N + i) * 36), "r"(saddr(mine + 36)), "r"(288) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(B + (w * N + i) * 6), "r"(saddr(mine + 72)), "r"(48) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    const int N = 50;
    const long W = 65536 * 2, tiles = W / 32;   // 2 x H_diag-sized = 1.89 GB
    const size_t bytes = (size_t)W * N * 288;
    double2 *out;
    CK(cudaMalloc(&out, bytes + (size_t)65536 * 50 * 48 + 4096));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        float best = 1e9;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
        }
        printf("%-44s %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
        return 0;
    };
    run("memset", [&] { cudaMemsetAsync(out, 0, bytes); });
    for (int st = 0; st < 2; ++st) {
        for (int R : {1, 5}) {
            char nm[96]; snprintf(nm, 96, "mode0 18-lane rows, run %d%s", R, st ? " (st.cs)" : "");
            int runs = (N + R - 1) / R; long warps = tiles * runs;
            run(nm, [&] { k<0><<<(unsigned)((warps * 32 + 255) / 256), 256>>>(out, N, R, tiles, st); });
        }
        for (int R : {1, 2, 5, 10, 25, 50}) {
            char nm[96]; snprintf(nm, 96, "mode1 32-lane, %d B contiguous per window%s", R * 288, st ? " (st.cs)" : "");
            int runs = (N + R - 1) / R; long warps = tiles * runs;
            run(nm, [&] { k<1><<<(unsigned)((warps * 32 + 255) / 256), 256>>>(out, N, R, tiles, st); });
        }
    }
    {
        const long W4 = 65536, tiles4 = W4 / 32;
        double2 *Hd = out, *Ho = out + (size_t)W4 * N * 18, *B = Ho + (size_t)W4 * N * 18;   // 2 x 0.94 GB + 0.16 GB < 1.89 GB
        const size_t b4 = (size_t)W4 * N * (288 + 288 + 48);
        auto run4 = [&](const char *name, auto launch) {
            float best = 1e9;
            for (int r = 0; r < 5; ++r) {
                cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
            }
            printf("%-44s %.3f ms  %.0f GB/s\n", name, best, b4 / best / 1e6);
        };
        // dynamic smem pins CTAs per SM: 227 KB / occ
        cudaFuncSetAttribute(k4<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        for (int occ : {32, 24, 16, 8}) {
            char nm[96]; snprintf(nm, 96, "4-fragment pattern, %d warps per SM", occ);
            size_t sm = occ == 32 ? 0 : (size_t)(220 * 1024 / occ) - 1024;
            run4(nm, [&] { k4<32><<<(unsigned)(tiles4 * N), 32, sm>>>(Hd, Ho, B, N, tiles4); });
        }
    }
    {
        const long W4 = 65536, tiles4 = W4 / 32;
        double *Hd = (double *)out, *Ho = Hd + (size_t)W4 * N * 36, *B = Ho + (size_t)W4 * N * 36;
        const size_t b4 = (size_t)W4 * N * (288 + 288 + 48);
        auto run4 = [&](const char *name, auto launch) {
            float best = 1e9;
            for (int r = 0; r < 5; ++r) {
                cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
            }
            printf("%-52s %.3f ms  %.0f GB/s\n", name, best, b4 / best / 1e6);
        };
        const size_t per_warp = 32 * 78 * 8;
        cudaFuncSetAttribute(kbulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_warp);
        cudaFuncSetAttribute(kbulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per_warp));
        for (int ipw : {1, 4, 16}) {
            char nm[96]; snprintf(nm, 96, "bulk store, 1-warp CTAs, %d items per warp", ipw);
            long warps = (tiles4 * N + ipw - 1) / ipw;
            run4(nm, [&] { kbulk<1><<<(unsigned)warps, 32, per_warp>>>(Hd, Ho, B, N, tiles4, ipw); });
            snprintf(nm, 96, "bulk store, 4-warp CTAs, %d items per warp", ipw);
            run4(nm, [&] { kbulk<4><<<(unsigned)((warps + 3) / 4), 128, 4 * per_warp>>>(Hd, Ho, B, N, tiles4, ipw); });
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
