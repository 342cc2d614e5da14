// A caller written against the REFERENCE's names (what tests/blas/performance/test_gemm_performance.cu:71-185 does: rns_const_init,
// mp_const_init, cuda::mp_array_init / host2device, cuda::mp_gemm<...>, cuda::mp_dot<...>, device2host), compiled against
// include/mpres_compat.cuh instead of the reference headers.  Test infrastructure: reads mp_float_t records from a file, writes the results.
//   in:  int m, n, k;  records alpha, beta, A[m*k], B[k*n], C[m*n], x[k], y[k]        out: C[m*n], r[1]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define RNS_MODULI_SIZE 8
#define RNS_MODULI_VALUES {113812103, 113812105, 113812107, 113812109, 113812111, 113812117, 113812121, 113812123}   /* params.8_2double.h */
#include "mpres_compat.cuh"

static void put(mp_array_t &d, std::vector<mp_float_t> &h) { cuda::mp_array_init(d, h.size()); cuda::mp_array_host2device(d, h.data(), h.size()); }

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int dims[3];
    if (fread(dims, sizeof(int), 3, f) != 3) return 4;
    const int m = dims[0], n = dims[1], k = dims[2];
    auto rd = [&](size_t cnt) { std::vector<mp_float_t> v(cnt); if (fread(v.data(), sizeof(mp_float_t), cnt, f) != cnt) exit(5); return v; };
    std::vector<mp_float_t> alpha = rd(1), beta = rd(1), A = rd((size_t) m * k), B = rd((size_t) k * n), C = rd((size_t) m * n), x = rd(k), y = rd(k), r(1);
    fclose(f);
    rns_const_init();
    mp_const_init();
    if (mpres_compat_last_status() != 0) return 6;
    mp_array_t dA, dB, dC, dal, dbe, dbuf, dx, dy, dr, dbuf2;
    put(dA, A); put(dB, B); put(dC, C); put(dal, alpha); put(dbe, beta); put(dx, x); put(dy, y);
    cuda::mp_array_init(dbuf, (size_t) m * n); cuda::mp_array_init(dr, 1); cuda::mp_array_init(dbuf2, k);
    cuda::mp_gemm<32, 1, 128, 64, 16>(mblas_no_trans, mblas_no_trans, m, n, k, dal, dA, m, dB, k, dbe, dC, m, dbuf);
    if (mpres_compat_last_status() != 0) return 7;
    cuda::mp_dot<128, 64, 128, 64, 64>(k, dx, 1, dy, 1, dr, dbuf2);
    if (mpres_compat_last_status() != 0) return 8;
    cudaDeviceSynchronize();
    cuda::mp_array_device2host(C.data(), dC, C.size());
    cuda::mp_array_device2host(r.data(), dr, 1);
    f = fopen(argv[2], "wb");
    fwrite(C.data(), sizeof(mp_float_t), C.size(), f);
    fwrite(r.data(), sizeof(mp_float_t), 1, f);
    fclose(f);
    cuda::mp_scal<128, 64, 128>(k, dal, dx, 1);                        /* the level-1 entry points build through the shim too */
    cuda::mp_axpy<128, 64, 128>(k, dal, dx, 1, dy, 1, dbuf2);
    cuda::mp_rot<128, 64, 128>(k, dx, 1, dy, 1, dal, dbe, dbuf2, dbuf2);
    cuda::mp_ge_diag_scale<128, 64, 128>(mblas_right_side, m, n, dy, 1, dC, m);     /* n <= k elements of y as the diagonal */
    cuda::mp_ge_lr_scale<128, 64, 128>(m, n, dx, 1, dy, 1, dC, m);
    cuda::mp_asum<64, 64>(k, dx, 1, dr);                               /* ... and the norms and the two-stage SpMV */
    cuda::mp_norm<64, 64>(mblas_inf_norm, k, dx, 1, dr);
    cuda::mp_ge_norm<64, 64>(mblas_one_norm, m, n, dC, m, dr, dbuf2);
    {
        mp_collection_t as, cbuf;
        cuda::mp_collection_init(as, 1); cuda::mp_collection_init(cbuf, 1);
        cuda::mp_collection_host2device(as, x.data(), 1);
        int h_irp[2] = {0, 1}, h_ja[1] = {0}, *irp, *ja;
        cudaMalloc(&irp, sizeof(h_irp)); cudaMalloc(&ja, sizeof(h_ja));
        cudaMemcpy(irp, h_irp, sizeof(h_irp), cudaMemcpyHostToDevice); cudaMemcpy(ja, h_ja, sizeof(h_ja), cudaMemcpyHostToDevice);
        cuda::mp_spmv_mpmtx_csr2st<32, 64, 32, 64>(1, 1, 1, irp, ja, as, dx, dr, cbuf);
        cuda::mp_spmv_mpmtx_ell2st<32, 64, 32, 64>(1, 1, 1, ja, as, dx, dr, cbuf);
        cuda::mp_collection_clear(as); cuda::mp_collection_clear(cbuf);
    }
    if (mpres_compat_last_status() != 0) return 9;
    cuda::mp_array_clear(dA); cuda::mp_array_clear(dB); cuda::mp_array_clear(dC);
    printf("MP_PRECISION %d MP_H %d\n", MP_PRECISION, MP_H);
    return 0;
}
