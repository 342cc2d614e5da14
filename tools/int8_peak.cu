// int8_peak.cu -- measured dense int8 tensor-core peak of this GPU: cuBLAS s8 x s8 -> s32 GEMM (TN, 8192^3), best of 10 launches and the
// average of a 2 s back-to-back run, CUDA events.  The roofline denominator of the small-modulus stage 2 (bench.py reads the JSON line this
// prints).  Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/int8_peak.cu -lcublas -o tools/int8_peak
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CB(x) do { cublasStatus_t s_ = (x); if (s_ != CUBLAS_STATUS_SUCCESS) { fprintf(stderr, "%s: cublas status %d\n", #x, (int) s_); return 1; } } while (0)

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 8192;
    int8_t *A, *B;
    int32_t *C;
    CK(cudaMalloc(&A, (size_t) n * n));
    CK(cudaMalloc(&B, (size_t) n * n));
    CK(cudaMalloc(&C, (size_t) n * n * 4));
    CK(cudaMemset(A, 1, (size_t) n * n));
    CK(cudaMemset(B, 1, (size_t) n * n));
    cublasHandle_t h;
    CB(cublasCreate(&h));
    const int32_t one = 1, zero = 0;
    auto gemm = [&]() {
        return cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, CUDA_R_8I, n, B, CUDA_R_8I, n, &zero, C, CUDA_R_32I, n, CUBLAS_COMPUTE_32I,
                            CUBLAS_GEMM_DEFAULT_TENSOR_OP);
    };
    for (int i = 0; i < 3; ++i) CB(gemm());
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
        CK(cudaEventRecord(e0));
        CB(gemm());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    int reps = (int) (2000.0f / best) + 1;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) CB(gemm());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float total;
    CK(cudaEventElapsedTime(&total, e0, e1));
    const double ops = 2.0 * n * (double) n * n;
    printf("{\"what\": \"cuBLAS s8 x s8 -> s32 GEMM (TN), n = %d\", \"burst_tops\": %.1f, \"sustained_tops\": %.1f, \"best_ms\": %.4f, \"sustained_launches\": %d}\n", n,
           ops / (best * 1e-3) / 1e12, ops * reps / (total * 1e-3) / 1e12, best, reps);
    return 0;
}
