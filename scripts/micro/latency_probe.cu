// Latency probe for the per-window CTA kernel design: dependent-issue latency of the FP64
// operations on the elimination chain, warp-level exchange primitives, and the launch + completion
// round trip of a single small kernel seen from the host (mapped pinned flag vs stream sync).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o latency_probe latency_probe.cu
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cuda_runtime.h>
#include "../../localization_b200/csrc/uwbgo_math.cuh"

using namespace uwbgo;

#define REP 256

template <int OP>
__global__ void chain_kernel(double *out, long long *cycles, double seed)
{
    double x = seed + threadIdx.x * 1e-3, y = 1.0000001, acc = 0.0;
    unsigned bad = 0;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) {
        if (OP == 0) x = fma(x, y, 1e-9);
        if (OP == 1) x = x + y;
        if (OP == 2) x = x * y;
        if (OP == 3) x = NbMath::sqrt_(x + 2.0, bad);
        if (OP == 4) x = NbMath::rcp(x + 2.0, bad);
        if (OP == 5) x = sqrt(x + 2.0);
        if (OP == 6) x = 1.0 / (x + 2.0);
        if (OP == 7) x = NbMath::rcp(NbMath::sqrt_(x + 2.0, bad), bad);
        if (OP == 8) x = 1.0 / sqrt(x + 2.0);
        if (OP == 9) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 10) x = NbMath::log_(x + 2.0, bad);
        if (OP == 11) x = NbMath::div(y, x + 2.0, bad);
    }
    long long t1 = clock64();
    acc += x + bad;
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void smem_exchange_kernel(double *out, long long *cycles)
{
    __shared__ double buf[64];
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) {
        buf[threadIdx.x & 31] = x;
        __syncwarp();
        x = buf[(threadIdx.x + 1) & 31] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NT>
__global__ void barrier_kernel(double *out, long long *cycles)
{
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) {
        asm volatile("bar.sync 1, %0;" ::"n"(NT));
        x += 1.0;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void syncthreads_kernel(double *out, long long *cycles)
{
    __shared__ double s[1024];
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < REP; ++i) {
        s[threadIdx.x] = x;
        __syncthreads();
        x = s[(threadIdx.x + 33) % blockDim.x] + 1.0;
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

// a small kernel that reads its input from mapped host memory and writes a result + flag there
__global__ void roundtrip_kernel(const double *in, double *outp, volatile int *flag, int n, int seq)
{
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += in[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
        outp[0] = s;
        __threadfence_system();
        *flag = seq;
    }
}

template <class F>
static void run_chain(const char *name, F launch, int per_iter = 1)
{
    double *out;
    long long *cyc;
    cudaMalloc(&out, 8 * 1024);
    cudaMalloc(&cyc, 8 * 16);
    launch(out, cyc);
    launch(out, cyc);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles per op\n", name, (double)h / REP / per_iter);
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaSetDevice(0);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
#define CH(OP, name) run_chain(name, [](double *o, long long *c) { chain_kernel<OP><<<1, 32>>>(o, c, 1.5); })
    CH(0, "DFMA dependent");
    CH(1, "DADD dependent");
    CH(2, "DMUL dependent");
    CH(3, "NbMath sqrt (+1 DADD)");
    CH(4, "NbMath rcp (+1 DADD)");
    CH(5, "IEEE sqrt (+1 DADD)");
    CH(6, "IEEE 1/x (+1 DADD)");
    CH(7, "NbMath rcp(sqrt) (+1 DADD)");
    CH(8, "IEEE 1/sqrt (+1 DADD)");
    CH(9, "SHFL f64 dependent");
    CH(10, "NbMath log (+1 DADD)");
    CH(11, "NbMath div (+1 DADD)");
    run_chain("smem store/syncwarp/load/syncwarp (+DADD)", [](double *o, long long *c) { smem_exchange_kernel<<<1, 32>>>(o, c); });
    run_chain("bar.sync 64 threads (+DADD)", [](double *o, long long *c) { barrier_kernel<64><<<1, 64>>>(o, c); });
    run_chain("bar.sync 128 threads (+DADD)", [](double *o, long long *c) { barrier_kernel<128><<<1, 128>>>(o, c); });
    run_chain("bar.sync 256 threads (+DADD)", [](double *o, long long *c) { barrier_kernel<256><<<1, 256>>>(o, c); });
    run_chain("smem + 2x __syncthreads, 128 threads", [](double *o, long long *c) { syncthreads_kernel<<<1, 128>>>(o, c); });
    run_chain("smem + 2x __syncthreads, 256 threads", [](double *o, long long *c) { syncthreads_kernel<<<1, 256>>>(o, c); });
    run_chain("smem + 2x __syncthreads, 512 threads", [](double *o, long long *c) { syncthreads_kernel<<<1, 512>>>(o, c); });

    // host round trip
    double *hin, *hout;
    int *hflag;
    cudaHostAlloc(&hin, 8 * 1024, cudaHostAllocMapped);
    cudaHostAlloc(&hout, 64, cudaHostAllocMapped);
    cudaHostAlloc(&hflag, 64, cudaHostAllocMapped);
    for (int i = 0; i < 1024; ++i) hin[i] = 1.0;
    double *din, *dout;
    int *dflag;
    cudaHostGetDevicePointer(&din, hin, 0);
    cudaHostGetDevicePointer(&dout, hout, 0);
    cudaHostGetDevicePointer(&dflag, hflag, 0);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    *hflag = 0;
    const int R = 2000;
    for (int mode = 0; mode < 3; ++mode) {
        double *dbuf_in, *dbuf_out;
        cudaMalloc(&dbuf_in, 8 * 1024);
        cudaMalloc(&dbuf_out, 64);
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 1; r <= R; ++r) {
            int seq = mode * R + r;
            if (mode == 0) { // mapped in/out, spin on flag
                roundtrip_kernel<<<1, 128, 0, st>>>(din, dout, dflag, 1024, seq);
                while (*(volatile int *)hflag != seq) {}
            } else if (mode == 1) { // mapped in/out, stream sync
                roundtrip_kernel<<<1, 128, 0, st>>>(din, dout, dflag, 1024, seq);
                cudaStreamSynchronize(st);
            } else { // memcpy in, kernel, memcpy out, sync
                cudaMemcpyAsync(dbuf_in, hin, 8 * 1024, cudaMemcpyHostToDevice, st);
                roundtrip_kernel<<<1, 128, 0, st>>>(dbuf_in, dbuf_out, dflag, 1024, seq);
                cudaMemcpyAsync(hout, dbuf_out, 8, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        double us = std::chrono::duration<double, std::micro>(t1 - t0).count() / R;
        const char *names[3] = {"mapped in/out + spin on mapped flag", "mapped in/out + cudaStreamSynchronize", "H2D copy + kernel + D2H copy + sync"};
        printf("host round trip, %-40s %8.2f us\n", names[mode], us);
    }
    return 0;
}
