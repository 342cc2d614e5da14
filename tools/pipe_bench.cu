// tools/pipe_bench.cu -- dependency-free issue-rate microbenchmark for the instructions the stage-2
// modular multiply-accumulate can be built from: IMAD.WIDE.U32 (32x32+64), IMAD (32x32+32 low),
// DFMA, and interleavings.  Gives R_mac, the INT32 roofline denominator of SURVEY 8(d).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_bench.cu -o tools/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ACC = 16;      // independent accumulators per thread
constexpr int ITERS = 4096;  // loop trips; each trip issues ACC ops of each kind

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned long long *out, unsigned a0, unsigned b0, double da, double db) {
    unsigned long long acc[ACC];
    unsigned lo[ACC];
    double d[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = threadIdx.x + i; lo[i] = threadIdx.x * 3 + i; d[i] = threadIdx.x + i; }
    unsigned a = a0 + threadIdx.x, b = b0 + blockIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0 || MODE == 3 || MODE == 4) acc[i] += (unsigned long long) (a + i) * (unsigned long long) b;   // IMAD.WIDE.U32
            if (MODE == 1) lo[i] = (a + i) * b + lo[i];                                                                  // IMAD
            if (MODE == 2 || MODE == 3) d[i] = fma(da, db + i, d[i]);                                                   // DFMA
            if (MODE == 4) { d[i] = fma(da, db + i, d[i]); d[i] = fma(db, da + i, d[i]); }                               // 1 WIDE : 2 DFMA
        }
        a += 7; b ^= a;
    }
    unsigned long long r = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) r += acc[i] + lo[i] + (unsigned long long) d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
double run(const char *name, double ops_per_trip) {
    int dev, sms;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * 8, threads = 256;
    unsigned long long *out;
    cudaMalloc(&out, sizeof(unsigned long long) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) k<MODE><<<blocks, threads>>>(out, 12345u, 777u, 1.000001, 0.999999);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int w = 0; w < reps; ++w) k<MODE><<<blocks, threads>>>(out, 12345u, 777u, 1.000001, 0.999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double) blocks * threads * ITERS * ACC * ops_per_trip * reps;
    double rate = ops / (ms * 1e-3);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("{\"bench\": \"%s\", \"ms\": %.3f, \"ops_per_s\": %.4e, \"per_clk_per_sm_at_max\": %.2f, \"sms\": %d, \"max_khz\": %d}\n",
           name, ms / reps, rate, rate / sms / (clk * 1e3), sms, clk);
    cudaFree(out);
    return rate;
}

int main() {
    run<0>("imad_wide_u32", 1);
    run<1>("imad_lo32", 1);
    run<2>("dfma", 1);
    run<3>("imad_wide+dfma(1:1), total ops", 2);
    run<4>("imad_wide+2dfma, total ops", 3);
    return 0;
}
