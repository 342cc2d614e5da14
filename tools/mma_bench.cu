// tools/mma_bench.cu -- issue-rate microbenchmark: legacy warp-level int8 MMA (mma.sync m16n8k32
// u8*u8+s32) against the CUDA-core candidates (IMAD.WIDE.U32, IMAD, DFMA) for the stage-2 residue MAC.
// Operands live in registers, so this is the pipe ceiling, not a GEMM.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__global__ void __launch_bounds__(256) k_mma(int *out, unsigned a0, unsigned b0) {
    constexpr int T = 8;  // independent accumulator tiles per warp
    int c[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) for (int j = 0; j < 4; ++j) c[t][j] = 0;
    unsigned a[4] = {a0 + threadIdx.x, a0 * 3 + threadIdx.x, a0 * 5, a0 * 7}, b[2] = {b0 + threadIdx.x, b0 * 3};
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int t = 0; t < T; ++t)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                         : "+r"(c[t][0]), "+r"(c[t][1]), "+r"(c[t][2]), "+r"(c[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int r = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) for (int j = 0; j < 4; ++j) r += c[t][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_alu(unsigned long long *out, unsigned a0, unsigned b0, double da) {
    constexpr int ACC = 16;
    unsigned long long acc[ACC];
    unsigned lo[ACC], x[ACC];
    double d[ACC], dx[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { acc[i] = threadIdx.x + i; lo[i] = i; d[i] = i; x[i] = a0 + i * 977 + threadIdx.x; dx[i] = da + i; }
    unsigned b = b0 + blockIdx.x;
    double db = (double) b0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0 || MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x[i]), "r"(b));
            if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
            if (MODE == 2 || MODE == 3) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx[i]), "d"(db));
            if (MODE == 4) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x[i]), "r"(b));
                             asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(dx[i]), "d"(db)); }
        }
        b += 3; db += 1.0;
    }
    unsigned long long r = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) r += acc[i] + lo[i] + (unsigned long long) d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
void timeit(const char *name, double ops, F launch) {
    int dev, sms, clk;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) launch();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int w = 0; w < reps; ++w) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = ops * reps / (ms * 1e-3);
    printf("{\"bench\": \"%s\", \"ms\": %.3f, \"ops_per_s\": %.4e, \"per_clk_per_sm_at_max_clock\": %.1f, \"err\": \"%s\"}\n", name, ms / reps, rate,
           rate / sms / (clk * 1e3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int dev, sms;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int blocks = sms * 4, threads = 256;
    void *out;
    cudaMalloc(&out, 8 * blocks * threads);
    double thr = (double) blocks * threads;
    timeit("mma.sync.m16n8k32.u8 (int8 MACs)", thr / 32 * ITERS * 8 * (16.0 * 8 * 32), [&] { k_mma<<<blocks, threads>>>((int *) out, 0x01020304u, 0x05060708u); });
    timeit("imad.wide.u32", thr * ITERS * 16, [&] { k_alu<0><<<blocks, threads>>>((unsigned long long *) out, 12345u, 777u, 1.5); });
    timeit("imad.lo.u32", thr * ITERS * 16, [&] { k_alu<1><<<blocks, threads>>>((unsigned long long *) out, 12345u, 777u, 1.5); });
    timeit("dfma", thr * ITERS * 16, [&] { k_alu<2><<<blocks, threads>>>((unsigned long long *) out, 12345u, 777u, 1.5); });
    timeit("imad.wide + dfma (ops = both)", thr * ITERS * 32, [&] { k_alu<3><<<blocks, threads>>>((unsigned long long *) out, 12345u, 777u, 1.5); });
    timeit("imad.lo + dfma (ops = both)", thr * ITERS * 32, [&] { k_alu<4><<<blocks, threads>>>((unsigned long long *) out, 12345u, 777u, 1.5); });
    return 0;
}
