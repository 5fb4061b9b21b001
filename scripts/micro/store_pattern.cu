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

int main()
{
    const int N = 50;
    const long W = 65536 * 2, tiles = W / 32;   // 2 x H_diag-sized = 1.89 GB
    const size_t bytes = (size_t)W * N * 288;
    double2 *out;
    CK(cudaMalloc(&out, bytes));
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
    CK(cudaDeviceSynchronize());
    return 0;
}
