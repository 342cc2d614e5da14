// mpres_b200.cu -- C-ABI (include/mpres_b200.h) over the sm_100a kernels.
//
// Host logic mirrors the argument checks and call structure of the reference's host templates
// (src/blas/gemm.cuh:69-167, src/blas/gemv.cuh:150-268, src/blas/dot.cuh:84-107) but every numeric
// step runs in this library's own kernels.  There is no CPU path.
#include "../../include/mpres_b200.h"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>

#include "ctx.hpp"
#include "kernels_ref_order.cuh"
#include "kernels_fast.cuh"

// a constants-only context (device < 0) cannot compute
#define NEED_DEVICE(c) do { if ((c) && (c)->device < 0) return -100; } while (0)

static_assert(sizeof(mpres_er_float_t) == 16 && sizeof(Er) == 16, "er_float_t layout (src/types.cuh:46-49)");

// host-side helpers of the lean upload (mpres_gemm_host): templates cannot live inside the extern "C" block
namespace {
struct BoundKey { long long e; unsigned long long m; };       // largest upper interval bound seen: (exponent incl. the double's, mantissa bits)
inline void bound_key_update(BoundKey &k, double frac, long long ex) {
    if (frac == 0) return;
    unsigned long long fb;
    const double a = frac < 0 ? -frac : frac;
    memcpy(&fb, &a, 8);
    const long long ue = (ex > 100000 ? 100000 : (ex < -100000 ? -100000 : ex)) + (long long) (fb >> 52);
    const unsigned long long fm = fb & 0xfffffffffffffull;
    if (ue > k.e || (ue == k.e && fm > k.m)) { k.e = ue; k.m = fm; }
}
template <int W>
void pack_lean_range(const unsigned long long *src, size_t rw, size_t Nw, size_t b, size_t e, unsigned long long *dst, BoundKey &k) {
    for (size_t i = b; i < e; ++i) {
        const unsigned long long *r = src + i * rw;
        unsigned long long *o = dst + i * (W + 3);
        for (int w = 0; w < W; ++w) o[w] = r[w];
        o[W] = r[Nw];                                  // sign, exponent
        const unsigned long long fr = r[Nw + 3], ex = r[Nw + 4];
        o[W + 1] = fr; o[W + 2] = ex;                  // eval[1]
        double f; long long x;
        memcpy(&f, &fr, 8); memcpy(&x, &ex, 8);
        bound_key_update(k, f, x);
    }
}
}  // namespace

extern "C" {

const char *mpres_version(void) { return "mpres-b200 0.1 (sm_100a)"; }

int mpres_finalize(mpres_ctx *c);

int mpres_init_moduli(mpres_ctx **out, const int *moduli, int n, int device) {
    if (!out || !moduli) return -1;
    if (device < 0) {
        // constants-only context (no device): lets the constant tables be checked on a CPU-only box.
        // Every compute entry point rejects it -- there is no CPU arithmetic in this library.
        mpres_ctx *c = new mpres_ctx();
        c->device = -1;
        int rc = compute_constants(moduli, n, c->hc);
        if (rc) { delete c; return rc - 10; }
        compute_small_consts(c->hc, c->sc);      // host tables only: lets the CPU tests check them
        *out = c;
        return 0;
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device >= ndev) return -2;
    mpres_ctx *c = new mpres_ctx();
    c->device = device;
    int rc = compute_constants(moduli, n, c->hc);
    if (rc) { delete c; return rc - 10; }
    DeviceGuard g(device);
    const HostConsts &h = c->hc;
    DevConsts *d = new DevConsts();
    memset(d, 0, sizeof(*d));
    d->N = h.N; d->log2M = h.log2M; d->mp_h = h.mp_h; d->mp_j = h.mp_j; d->ref_factor = h.ref_factor; d->precision = h.mp_precision;
    d->accuracy = h.accuracy;
    d->unit_low = {h.unit_low.frac, h.unit_low.exp}; d->unit_upp = {h.unit_upp.frac, h.unit_upp.exp};
    d->inv_low = {h.inv_low.frac, h.inv_low.exp}; d->inv_upp = {h.inv_upp.frac, h.inv_upp.exp};
    for (int i = 0; i < h.N; ++i) {
        d->moduli[i] = h.moduli[i]; d->part_inverse[i] = h.part_inverse[i]; d->barrett[i] = h.barrett[i];
        d->recip_rd[i] = h.recip_rd[i]; d->recip_ru[i] = h.recip_ru[i];
    }
    for (int j = 0; j < kThresh; ++j) {
        d->m_pow2[j] = h.m_pow2[j];
        for (int i = 0; i < h.N; ++i) {
            d->mi_pow2[j][i] = h.mi_pow2[(size_t) j * h.N + i];
            d->pow2_inv[j][i] = h.pow2_inv[(size_t) j * h.N + i];
        }
    }
    cudaError_t e;
    auto up = [&](int **dst, const std::vector<int> &src) {
        e = cudaMalloc(dst, src.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice);
        return e;
    };
    if (up(&c->d_pow2, h.pow2) != cudaSuccess || up(&c->d_inv_pow2, h.inv_pow2_ext) != cudaSuccess ||
        up(&c->d_mrc, h.mrc_inv) != cudaSuccess || up(&c->d_prefix, h.prefix_mod) != cudaSuccess ||
        up(&c->d_ext_w, h.ext_w) != cudaSuccess || up(&c->d_ext_t, h.ext_t) != cudaSuccess || up(&c->d_wpow2, h.wpow2) != cudaSuccess || up(&c->d_spow2, h.spow2) != cudaSuccess) { delete d; cudaGetLastError(); mpres_finalize(c); return (int) e; }
    d->pow2 = c->d_pow2; d->inv_pow2 = c->d_inv_pow2; d->mrc_inv = c->d_mrc; d->prefix_mod = c->d_prefix; d->ext_w = c->d_ext_w; d->ext_t = c->d_ext_t; d->ext_lazy = h.ext_lazy; d->wpow2 = c->d_wpow2; d->spow2 = c->d_spow2;
    for (int i = 0; i <= h.N; ++i) d->prefix_log2[i] = h.prefix_log2[i];
    // tables of the small-modulus stage 2
    d->small = nullptr;
    compute_small_consts(h, c->sc);
    if (c->sc.usable) {
        const SmallConsts &sc = c->sc;
        SmallDev *sd = new SmallDev();
        memset(sd, 0, sizeof(*sd));
        sd->usable = 1; sd->ext_cols = sc.ext_cols; sd->red_shift = sc.red_shift; sd->log2M_up = sc.log2M_up;
        for (int j = 0; j < 64; ++j) {
            const int pj = j < kSmallMax ? kSmallModuli[j] : 1;
            sd->p[j] = pj;
            sd->mu[j] = j < kSmallMax ? (unsigned) ((1ull << 32) / (unsigned long long) pj) : 0u;
            sd->rcp[j] = j < kSmallMax ? 1.0f / (float) pj : 0.f;
        }
        for (int j = 0; j <= kSmallMax; ++j) sd->prefix_log2[j] = sc.prefix_log2[j];
        for (int j = 0; j <= kSmallNinMax; ++j) sd->in_log2_milli[j] = sc.in_log2_milli[j];
        auto upb = [&](int slot, const void *src, size_t bytes) -> const void * {
            if (cudaMalloc(&c->d_small[slot], bytes) != cudaSuccess) return nullptr;
            if (cudaMemcpy(c->d_small[slot], src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
            return c->d_small[slot];
        };
        std::vector<uint32_t> cw64((size_t) 64 * 16, 0);
        memcpy(cw64.data(), sc.cw.data(), sc.cw.size() * sizeof(uint32_t));
        sd->inv = (const uint8_t *) upb(1, sc.inv.data(), sc.inv.size());
        sd->ext_b = (const uint8_t *) upb(2, sc.ext_b.data(), sc.ext_b.size());
        sd->cw = (const uint32_t *) upb(3, cw64.data(), cw64.size() * 4);
        sd->pws = (const uint8_t *) upb(4, sc.pws.data(), sc.pws.size());
        sd->in_mi = (const uint32_t *) upb(5, sc.in_mi.data(), sc.in_mi.size() * 4);
        sd->in_negmp = (const uint32_t *) upb(6, sc.in_negmp.data(), sc.in_negmp.size() * 4);
        sd->red_mu = (const uint32_t *) upb(7, sc.red_mu.data(), sc.red_mu.size() * 4);
        sd->bin_mi = (const uint32_t *) upb(8, sc.bin_mi.data(), sc.bin_mi.size() * 4);
        sd->bin_negmp = (const uint32_t *) upb(9, sc.bin_negmp.data(), sc.bin_negmp.size() * 4);
        const bool okf = upb(10, sc.full_mi.data(), sc.full_mi.size() * 4) && upb(11, sc.full_negm.data(), sc.full_negm.size() * 4) &&
                         upb(12, sc.full_m.data(), sc.full_m.size() * 4);
        const bool ok = okf && sd->inv && sd->ext_b && sd->cw && sd->pws && sd->in_mi && sd->in_negmp && sd->red_mu && sd->bin_mi && sd->bin_negmp;
        const void *dev = ok ? upb(0, sd, sizeof(SmallDev)) : nullptr;
        delete sd;
        if (!dev) { delete d; cudaGetLastError(); mpres_finalize(c); return (int) cudaErrorMemoryAllocation; }
        d->small = (const SmallDev *) dev;
    }
    e = cudaMalloc(&c->dconsts, sizeof(DevConsts));
    if (e == cudaSuccess) e = cudaMemcpy(c->dconsts, d, sizeof(DevConsts), cudaMemcpyHostToDevice);
    delete d;
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counter, (kCounterInts + kCounterExtra) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->d_counter, 0, (kCounterInts + kCounterExtra) * sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->h_sel, 8 * sizeof(int), cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); mpres_finalize(c); return (int) e; }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return 0;
}

int mpres_init(mpres_ctx **out, int moduli_size, int device) {
    for (const ModuliSet &s : kModuliSets)
        if (s.n == moduli_size) return mpres_init_moduli(out, s.values, s.n, device);
    return -3;
}

int mpres_finalize(mpres_ctx *c) {
    if (!c) return -1;
    if (c->device < 0) { delete c; return 0; }
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    mpres_ops_release(c);
    for (int i = 0; i < 24; ++i) if (c->ws[i]) cudaFree(c->ws[i]);
    for (int i = 0; i < 4; ++i) if (c->hs[i]) cudaStreamDestroy(c->hs[i]);
    for (int i = 0; i < 16; ++i) if (c->hev[i]) cudaEventDestroy(c->hev[i]);
    for (int i = 0; i < 16; ++i) if (c->d_small[i]) cudaFree(c->d_small[i]);
    if (c->serial_ev) cudaEventDestroy(c->serial_ev);
    for (cudaEvent_t e : c->prof_ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 6; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->h_sel) cudaFreeHost(c->h_sel);
    if (c->lean_host) cudaFreeHost(c->lean_host);
    cudaFree(c->d_pow2); cudaFree(c->d_inv_pow2); cudaFree(c->d_mrc); cudaFree(c->d_prefix); cudaFree(c->d_ext_w); cudaFree(c->d_ext_t); cudaFree(c->d_wpow2); cudaFree(c->d_spow2); cudaFree(c->dconsts); cudaFree(c->d_counter);
    delete c;
    return 0;
}

int mpres_moduli_size(const mpres_ctx *c) { return c ? c->hc.N : -1; }
int mpres_moduli_product_log2(const mpres_ctx *c) { return c ? c->hc.log2M : -1; }
int mpres_precision(const mpres_ctx *c) { return c ? c->hc.mp_precision : -1; }
int mpres_mp_h(const mpres_ctx *c) { return c ? c->hc.mp_h : 0; }
int mpres_mp_j(const mpres_ctx *c) { return c ? c->hc.mp_j : 0; }
int mpres_device(const mpres_ctx *c) { return c ? c->device : -1; }
size_t mpres_sizeof_mp_float(const mpres_ctx *c) { return c ? 4 * (size_t) c->hc.N + 40 : 0; }
int mpres_set_mode(mpres_ctx *c, int mode) { if (!c || mode < 0 || mode > 2) return -1; c->mode = mode; return 0; }
int mpres_get_mode(const mpres_ctx *c) { return c ? c->mode : -1; }
int mpres_set_stage2_kernel(mpres_ctx *c, int kind) {
    if (!c || kind < 0 || kind > 6) return -1;
    c->stage2 = (kind == MPRES_STAGE2_SMALL_TILED || kind == MPRES_STAGE2_SMALL_K64 || kind == MPRES_STAGE2_SMALL_T128) ? MPRES_STAGE2_SMALL : kind;
    c->small_persistent = kind == MPRES_STAGE2_SMALL_TILED ? 0 : 1;
    c->small_kb = kind == MPRES_STAGE2_SMALL_K64 ? 64 : 128;
    c->small_tj = kind == MPRES_STAGE2_SMALL ? 256 : 128;
    return 0;
}
int mpres_set_stage3_kernel(mpres_ctx *c, int kind) {
    if (!c || kind < 0 || kind > 4) return -1;
    c->norm_staged = kind == 4 ? 0 : 1;
    c->stage3 = kind == 1 ? 1 : 0;
    c->norm32 = kind == 2 ? 0 : 1;
    c->fuse_ext = kind == 3 ? 1 : 0;
    return 0;
}
int mpres_set_reduced_base(mpres_ctx *c, int on) { if (!c) return -1; c->reduced_base = on != 0; return 0; }
long mpres_last_base_size(mpres_ctx *c) {
    if (!c || c->device < 0) return -1;
    DeviceGuard g(c->device);
    int v = 0;
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -2;
    if (cudaMemcpy(&v, c->d_counter + 2, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    return v;
}
int mpres_last_small_base(mpres_ctx *c, int *moduli, int *input_moduli) {
    if (!c || c->device < 0) return -1;
    DeviceGuard g(c->device);
    int v[2] = {0, 0};
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -2;
    if (cudaMemcpy(v, c->d_counter + 4, 2 * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    if (moduli) *moduli = v[0];
    if (input_moduli) *input_moduli = v[1];
    return 0;
}
int mpres_last_host_upload_residues(const mpres_ctx *c) { return c ? c->last_host_lean : -1; }
int mpres_last_binary_rounding(const mpres_ctx *c) { return c ? (c->last_binary ? 1 : 0) : -1; }
int mpres_set_workspace_limit(mpres_ctx *c, size_t bytes) {
    if (!c) return -1;
    std::lock_guard<std::mutex> lk(c->mu);
    c->ws_limit = bytes;
    return 0;
}
long mpres_workspace_fallbacks(const mpres_ctx *c) { return c ? c->ws_fallbacks : -1; }
size_t mpres_workspace_bytes(const mpres_ctx *c) {
    size_t held = 0;
    if (c) for (int i = 0; i < 24; ++i) held += c->ws_size[i];
    return held;
}
int mpres_small_modulus(const mpres_ctx *c, int index) {
    if (!c || !c->sc.usable || index < 0 || index >= kSmallMax) return 0;
    return kSmallModuli[index];
}
long mpres_debug_read_workspace(mpres_ctx *c, int slot, size_t offset, void *host, size_t bytes) {
    if (!c || c->device < 0 || !host || slot < 0 || slot >= 12) return -1;
    if (!c->ws[slot] || offset + bytes > c->ws_size[slot]) return -2;
    DeviceGuard g(c->device);
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -3;
    if (cudaMemcpy(host, (const char *) c->ws[slot] + offset, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    return (long) bytes;
}
int mpres_set_vec_config(mpres_ctx *c, int cfg) { if (!c || cfg < 0 || cfg > 2) return -1; c->vec_config = cfg; return 0; }
int mpres_set_stage1_kernel(mpres_ctx *c, int kind) {
    if (!c || kind < 0 || kind > 3) return -1;
    c->stage1 = kind == 1 ? 1 : 0;
    c->minplus_sparse = (kind == 0 || kind == 3) ? 1 : 0;
    c->align_mma = kind == 3 ? 1 : 0;
    return 0;
}
long mpres_last_minplus_dense_count(mpres_ctx *c) {
    if (!c || c->device < 0) return -1;
    DeviceGuard g(c->device);
    int v[kCounterInts];
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -2;
    if (cudaMemcpy(v, c->d_counter, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    long t = 0;
    for (int p = 0; p < kMaxPanels; ++p) t += v[kCounterBlock * p + 6];
    return t;
}
long mpres_launch_count(const mpres_ctx *c) { return c ? c->launches.load() : -1; }

long mpres_last_fallback_count(mpres_ctx *c) {
    if (!c || c->device < 0) return -1;
    DeviceGuard g(c->device);
    int v[kCounterInts];
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -2;
    if (cudaMemcpy(v, c->d_counter, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    long t = 0;
    for (int p = 0; p < kMaxPanels; ++p) t += v[kCounterBlock * p + 0];
    return t;
}

long mpres_last_slow_count(mpres_ctx *c) {
    if (!c || c->device < 0) return -1;
    DeviceGuard g(c->device);
    int v[kCounterInts];
    if (cudaStreamSynchronize(c->last_stream) != cudaSuccess) return -2;
    if (cudaMemcpy(v, c->d_counter, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    long t = 0;
    for (int p = 0; p < kMaxPanels; ++p) t += v[kCounterBlock * p + 1];
    return t;
}

int mpres_set_profiling(mpres_ctx *c, int on) { if (!c) return -1; c->profiling = on != 0; c->ev_valid = false; return 0; }

// ms[0..2] = stage 1 (alignment), stage 2 (limb GEMM launches), stage 3 (normalise + epilogue + fallback)
// of the last fast-path mp_gemm; launches = number of stage-2 kernel launches.  Synchronises.
int mpres_last_stage_ms(mpres_ctx *c, float *ms, int *launches) {
    if (!c || !ms) return -1;
    if (!c->ev_valid) return -2;
    DeviceGuard g(c->device);
    CUDA_TRY(cudaEventSynchronize(c->ev[3]));
    for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(&ms[i], c->ev[i], c->ev[i + 1]));
    if (launches) *launches = c->last_stage2_launches;
    return 0;
}

/* "name=ms;name=ms;..." of the kernels of the last fast mp_gemm call (profiling on); returns the bytes written or < 0.  Synchronises. */
long mpres_last_kernel_ms(mpres_ctx *c, char *out, size_t cap) {
    if (!c || !out || cap == 0) return -1;
    if (c->prof_used < 2) return -2;
    DeviceGuard g(c->device);
    if (cudaEventSynchronize(c->prof_ev[c->prof_used - 1]) != cudaSuccess) return -3;
    size_t pos = 0;
    for (int i = 1; i < c->prof_used; ++i) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->prof_ev[i - 1], c->prof_ev[i]) != cudaSuccess) return -3;
        const int w = snprintf(out + pos, cap - pos, "%s%s=%.6f", i > 1 ? ";" : "", c->prof_name[i], ms);
        if (w < 0 || (size_t) w >= cap - pos) return -4;
        pos += (size_t) w;
    }
    return (long) pos;
}

long mpres_get_constant(const mpres_ctx *c, int which, void *out, size_t cap) {
    if (!c || !out) return -1;
    const HostConsts &h = c->hc;
    const void *src = nullptr;
    size_t bytes = 0;
    double tmp[5];
    switch (which) {
        case 0: src = h.moduli.data(); bytes = h.moduli.size() * 4; break;
        case 1: src = h.part_inverse.data(); bytes = h.part_inverse.size() * 4; break;
        case 2: src = h.pow2.data(); bytes = h.pow2.size() * 4; break;
        case 3: src = h.m_pow2.data(); bytes = h.m_pow2.size() * 4; break;
        case 4: src = h.mi_pow2.data(); bytes = h.mi_pow2.size() * 4; break;
        case 5: src = h.pow2_inv.data(); bytes = h.pow2_inv.size() * 4; break;
        case 6: src = h.mrc_inv.data(); bytes = h.mrc_inv.size() * 4; break;
        case 7: src = h.recip_rd.data(); bytes = h.recip_rd.size() * 8; break;
        case 8: src = h.recip_ru.data(); bytes = h.recip_ru.size() * 8; break;
        case 9: tmp[0] = h.accuracy; tmp[1] = h.unit_low.frac; tmp[2] = h.unit_upp.frac; tmp[3] = h.inv_low.frac; tmp[4] = h.inv_upp.frac;
                src = tmp; bytes = sizeof(tmp); break;
        case 10: tmp[0] = h.ref_factor; tmp[1] = (double) h.unit_low.exp; tmp[2] = (double) h.unit_upp.exp; tmp[3] = (double) h.inv_low.exp; tmp[4] = (double) h.inv_upp.exp;
                src = tmp; bytes = sizeof(tmp); break;
        default: return -2;
    }
    if (bytes > cap) return -3;
    memcpy(out, src, bytes);
    return (long) bytes;
}

/* ---- containers ---------------------------------------------------------------------------------- */

static int soa_alloc(mpres_ctx *c, int **digits, int **sign, int **exp, mpres_er_float_t **eval, size_t size) {
    const size_t n = size ? size : 1;
    CUDA_TRY(cudaMalloc(digits, n * c->hc.N * sizeof(int)));
    CUDA_TRY(cudaMalloc(sign, n * sizeof(int)));
    CUDA_TRY(cudaMalloc(exp, n * sizeof(int)));
    CUDA_TRY(cudaMalloc(eval, 2 * n * sizeof(mpres_er_float_t)));
    return 0;
}

int mpres_array_init(mpres_ctx *c, mpres_array_t *a, size_t size) {
    NEED_DEVICE(c);
    if (!c || !a) return -1;
    DeviceGuard g(c->device);
    memset(a, 0, sizeof(*a));
    int rc = soa_alloc(c, &a->digits, &a->sign, &a->exp, &a->eval, size);
    if (rc) return rc;
    CUDA_TRY(cudaMalloc(&a->buf, (size ? size : 1) * sizeof(mpres_int4)));
    CUDA_TRY(cudaMalloc(&a->len, sizeof(int)));
    int len = (int) size;
    CUDA_TRY(cudaMemcpy(a->len, &len, sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}
int mpres_array_clear(mpres_ctx *c, mpres_array_t *a) {
    NEED_DEVICE(c);
    if (!c || !a) return -1;
    DeviceGuard g(c->device);
    cudaFree(a->digits); cudaFree(a->sign); cudaFree(a->exp); cudaFree(a->eval); cudaFree(a->buf); cudaFree(a->len);
    memset(a, 0, sizeof(*a));
    return 0;
}
int mpres_collection_init(mpres_ctx *c, mpres_collection_t *a, size_t size) {
    NEED_DEVICE(c);
    if (!c || !a) return -1;
    DeviceGuard g(c->device);
    memset(a, 0, sizeof(*a));
    return soa_alloc(c, &a->digits, &a->sign, &a->exp, &a->eval, size);
}
int mpres_collection_clear(mpres_ctx *c, mpres_collection_t *a) {
    NEED_DEVICE(c);
    if (!c || !a) return -1;
    DeviceGuard g(c->device);
    cudaFree(a->digits); cudaFree(a->sign); cudaFree(a->exp); cudaFree(a->eval);
    memset(a, 0, sizeof(*a));
    return 0;
}

// AoS records <-> SoA on the device: one staging copy of the raw records, one kernel.
__global__ void k_aos_to_soa(int N, const char *recs, long long n, int *digits, int *sign, int *exp, Er *eval, long long len) {
    const long long rs = 4ll * N + 40;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n * (N + 6); t += (long long) gridDim.x * blockDim.x) {
        const long long i = t / (N + 6);
        const int f = (int) (t % (N + 6));
        const char *p = recs + i * rs;
        if (f < N) digits[i * N + f] = ((const int *) p)[f];
        else if (f == N) sign[i] = ((const int *) p)[N];
        else if (f == N + 1) exp[i] = ((const int *) p)[N + 1];
        else {
            const int q = f - N - 2;  // 0..3: lo.frac, lo.exp, up.frac, up.exp as 8-byte words
            const long long *src = (const long long *) (p + 4 * N + 8);
            long long *dst = (long long *) (eval + (q < 2 ? i : i + len));
            dst[q & 1] = src[q];
        }
    }
}
// lean records (what the fast mp_gemm path reads of an operand entry): nin residues | sign | exponent | upper interval bound
__global__ void k_lean_to_soa(int N, int nin, const char *recs, long long n, int *digits, int *sign, int *exp, Er *eval, long long len) {
    const long long ls = 4ll * nin + 24;
    const int F = nin + 4;                      // work items per record: residues, sign, exponent, two words of the bound
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n * F; t += (long long) gridDim.x * blockDim.x) {
        const long long i = t / F;
        const int f = (int) (t % F);
        const char *p = recs + i * ls;
        if (f < nin) digits[i * N + f] = ((const int *) p)[f];
        else if (f == nin) sign[i] = ((const int *) p)[nin];
        else if (f == nin + 1) exp[i] = ((const int *) p)[nin + 1];
        else ((long long *) (eval + i + len))[f - nin - 2] = ((const long long *) (p + 4 * nin + 8))[f - nin - 2];
    }
}
__global__ void k_soa_to_aos(int N, char *recs, long long n, const int *digits, const int *sign, const int *exp, const Er *eval, long long len) {
    const long long rs = 4ll * N + 40;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n * (N + 6); t += (long long) gridDim.x * blockDim.x) {
        const long long i = t / (N + 6);
        const int f = (int) (t % (N + 6));
        char *p = recs + i * rs;
        if (f < N) ((int *) p)[f] = digits[i * N + f];
        else if (f == N) ((int *) p)[N] = sign[i];
        else if (f == N + 1) ((int *) p)[N + 1] = exp[i];
        else {
            const int q = f - N - 2;
            long long *dst = (long long *) (p + 4 * N + 8);
            const long long *src = (const long long *) (eval + (q < 2 ? i : i + len));
            dst[q] = src[q & 1];
        }
    }
}

static int h2d_common(mpres_ctx *c, int *digits, int *sign, int *exp, mpres_er_float_t *eval, size_t len, const void *host, size_t size) {
    if (!c || !host) return -1;
    if (size == 0) return 0;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t rs = 4 * (size_t) c->hc.N + 40;
    void *stage;
    int rc = ws_reserve(c, 7, size * rs, &stage);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(stage, host, size * rs, cudaMemcpyHostToDevice));
    k_aos_to_soa<<<c->sm_count * 8, 256>>>(c->hc.N, (const char *) stage, (long long) size, digits, sign, exp, (Er *) eval, (long long) len);
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}
static int d2h_common(mpres_ctx *c, void *host, const int *digits, const int *sign, const int *exp, const mpres_er_float_t *eval, size_t len, size_t size) {
    if (!c || !host) return -1;
    if (size == 0) return 0;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    const size_t rs = 4 * (size_t) c->hc.N + 40;
    void *stage;
    int rc = ws_reserve(c, 7, size * rs, &stage);
    if (rc) return rc;
    k_soa_to_aos<<<c->sm_count * 8, 256>>>(c->hc.N, (char *) stage, (long long) size, digits, sign, exp, (const Er *) eval, (long long) len);
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(host, stage, size * rs, cudaMemcpyDeviceToHost));
    return 0;
}
static int read_len(const mpres_array_t *a, size_t *len) {
    int l = 0;
    CUDA_TRY(cudaMemcpy(&l, a->len, sizeof(int), cudaMemcpyDeviceToHost));
    *len = (size_t) l;
    return 0;
}

int mpres_array_host2device(mpres_ctx *c, mpres_array_t *dst, const void *host, size_t size) {
    NEED_DEVICE(c);
    if (!c || !dst) return -1;
    DeviceGuard g(c->device);
    size_t len;
    int rc = read_len(dst, &len);
    if (rc) return rc;
    if (size > len) return -2;
    return h2d_common(c, dst->digits, dst->sign, dst->exp, dst->eval, len, host, size);
}
int mpres_array_device2host(mpres_ctx *c, void *host, const mpres_array_t *src, size_t size) {
    NEED_DEVICE(c);
    if (!c || !src) return -1;
    DeviceGuard g(c->device);
    size_t len;
    int rc = read_len(src, &len);
    if (rc) return rc;
    if (size > len) return -2;
    return d2h_common(c, host, src->digits, src->sign, src->exp, src->eval, len, size);
}
int mpres_collection_host2device(mpres_ctx *c, mpres_collection_t *dst, const void *host, size_t size) {
    NEED_DEVICE(c);
    if (!c || !dst) return -1;
    return h2d_common(c, dst->digits, dst->sign, dst->exp, dst->eval, size, host, size);
}
int mpres_collection_device2host(mpres_ctx *c, void *host, const mpres_collection_t *src, size_t size) {
    NEED_DEVICE(c);
    if (!c || !src) return -1;
    return d2h_common(c, host, src->digits, src->sign, src->exp, src->eval, size, size);
}

int mpres_array_set_binary(mpres_ctx *c, mpres_array_t *dst, size_t offset, const int *sign, const int *exp,
                           const uint32_t *limbs, int nlimbs, size_t count, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !dst || !sign || !exp || !limbs || nlimbs < 1) return -1;
    if (count == 0) return 0;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    const int N = c->hc.N;
    MPRES_DISPATCH(N, {
        long long groups = (long long) count;
        int block = 256;
        long long blocks = std::min<long long>((groups * G + block - 1) / block, (long long) c->sm_count * 16);
        k_set_binary<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, view(dst), (long long) offset, sign, exp, limbs, nlimbs, (long long) count);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_probe(mpres_ctx *c, int op, void *r, const void *x, const void *y, const int *bits, size_t n, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !r || !x || op < 0 || op > 4) return -1;
    if ((op <= 1 && !y) || (op == 4 && !bits)) return -1;
    if (n == 0) return 0;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        int block = 128;
        long long blocks = ((long long) n * G + block - 1) / block;
        k_probe<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, op, (char *) r, (const char *) x, (const char *) y, bits, (long long) n);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

/* ---- GEMM ------------------------------------------------------------------------------------------ */

static int gemm_impl(mpres_ctx *c, int transa, int transb, int m, int n, int k, SoA alpha, SoA A, int lda, SoA B, int ldb,
                     SoA beta, SoA Cm, int ldc, const SoA *buffer, cudaStream_t st, FastShard *sh = nullptr) {
    // argument checks of src/blas/gemm.cuh:75-96 (the reference returns silently)
    if (m <= 0 || n <= 0 || k <= 0) return 0;
    const bool ta = transa != MPRES_NO_TRANS, tb = transb != MPRES_NO_TRANS;
    if (transa != MPRES_NO_TRANS && transa != MPRES_TRANS && transa != MPRES_CONJ_TRANS) return -2;
    if (transb != MPRES_NO_TRANS && transb != MPRES_TRANS && transb != MPRES_CONJ_TRANS) return -2;
    if (lda < std::max(1, ta ? k : m)) return -3;
    if (ldb < std::max(1, tb ? n : k)) return -4;
    if (ldc < std::max(1, m)) return -5;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    c->last_binary = false;
    c->last_fast_ok = false;
    c->last_nin = 0;
    const int N = c->hc.N;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, kCounterInts * sizeof(int), st));

    bool done = false;
    if (c->mode != MPRES_MODE_REFERENCE_ORDER) {
        rc = gemm_fast_full(c, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, st, &done, sh);
        if (rc == (int) cudaErrorMemoryAllocation && !sh) {
            // the fast path's workspaces did not fit: every reservation precedes the first write to C, so the call can still be served in
            // reference order, which needs one m x n scratch matrix like the reference itself (a sharded call cannot: its peers wait for this rank)
            CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, kCounterInts * sizeof(int), st));
            c->last_binary = false; c->last_fast_ok = false; c->last_nin = 0;
            c->ws_fallbacks++;
            done = false; rc = 0;
        }
        if (rc) return rc;
    }
    if (!done) {
        // reference order (a sharded call holds the complete B on every rank: its row block needs no exchange here)
        SoA S;
        int lds = m;
        if (buffer) S = *buffer;
        else { rc = ws_soa(c, 0, (size_t) m * n, &S); if (rc) return rc; }
        MPRES_DISPATCH(N, {
            int block = 128;
            long long groups = (long long) m * n;
            long long blocks = std::min<long long>((groups * G + block - 1) / block, (long long) c->sm_count * 64);
            k_gemm_ref_order<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, ta, tb, m, n, k, A, lda, B, ldb, S, lds, nullptr, nullptr);
        });
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        MPRES_DISPATCH(N, {
            int block = 128;
            long long groups = (long long) m * n;
            long long blocks = std::min<long long>((groups * G + block - 1) / block, (long long) c->sm_count * 32);
            k_gemm_epilogue<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, m, n, alpha, beta, S, lds, Cm, ldc);
        });
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
    }
    return call_end(c, st);
}

int mpres_gemm(mpres_ctx *c, int transa, int transb, int m, int n, int k, const mpres_array_t *alpha,
               const mpres_array_t *A, int lda, const mpres_array_t *B, int ldb, const mpres_array_t *beta,
               mpres_array_t *Cm, int ldc, mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !alpha || !A || !B || !beta || !Cm) return -1;
    SoA buf;
    if (buffer && buffer->digits) buf = view(buffer);
    return gemm_impl(c, transa, transb, m, n, k, view(alpha), view(A), lda, view(B), ldb, view(beta), view(Cm), ldc,
                     (buffer && buffer->digits) ? &buf : nullptr, (cudaStream_t) stream);
}
/* ---- GEMM over HOST operands: upload, compute and download pipelined by column panels ------------------------------
 * The reference's caller does mp_array_host2device x 3, mp_gemm, mp_array_device2host (tests/blas/test_gemm.cu); over PCIe the
 * copies are 20 x the compute, so the end-to-end rate is the bus rate.  Here C is cut into column panels: while panel j is
 * computed and its result travels back, panels j+1.. of B and C travel up -- the two PCIe directions run at the same time
 * and the device never waits for more than one panel. */

namespace {

constexpr int kHostRing = 3;
struct HostPipe {
    mpres_ctx *c;
    size_t rs, chunk;           // record size, records per staging chunk
    char *up_stage;             // kHostRing chunks
    char *lean_host = nullptr;  // pinned host slots (kHostRing x 64 MiB) the lean records are packed into
    int seq = 0;
    cudaStream_t s_up, s_unp, s_comp, s_down;
};

// records [off, off + cnt) of a host AoS array into the same positions of a device SoA (asynchronous: s_up copies, s_unp unpacks)
int host_upload(HostPipe &hp, const void *host, size_t off, size_t cnt, const SoA &dst) {
    mpres_ctx *c = hp.c;
    const int N = c->hc.N;
    while (cnt) {
        const size_t now = std::min(cnt, hp.chunk);
        const int slot = hp.seq % kHostRing;
        cudaEvent_t done = c->hev[slot], freed = c->hev[kHostRing + slot];
        char *stage = hp.up_stage + (size_t) slot * hp.chunk * hp.rs;
        CUDA_TRY(cudaStreamWaitEvent(hp.s_up, freed, 0));
        CUDA_TRY(cudaMemcpyAsync(stage, (const char *) host + off * hp.rs, now * hp.rs, cudaMemcpyHostToDevice, hp.s_up));
        CUDA_TRY(cudaEventRecord(done, hp.s_up));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_unp, done, 0));
        k_aos_to_soa<<<c->sm_count * 4, 256, 0, hp.s_unp>>>(N, stage, (long long) now, dst.digits + off * N, dst.sign + off, dst.exp + off, dst.eval + off,
                                                            dst.len_val);
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(freed, hp.s_unp));
        ++hp.seq;
        off += now; cnt -= now;
    }
    return 0;
}

// ---- lean upload of A and B ------------------------------------------------------------------------------------------------------
// The fast path reads of an operand entry its first n_in residues, sign, exponent and upper interval bound: 40 of the 168 bytes of a
// 424-bit record with the reference's benchmark inputs.  The host cores cut the records down (16 threads: 117 GB/s of source records on
// the GPU box, tools/host_pack_bench.cpp), so the PCIe link carries 4.2 instead of 8.5 GB per step at config 3.  n_in is guessed from a
// sample of the interval bounds, the packing pass sees every bound and reports the largest; gemm_impl's own choice is checked after every
// panel (HostPipe::lean_*), and whatever does not fit is uploaded in full.
// residues an entry with this bound needs (the rule of k_outer_info / k_choose_base, with one unit of slack), kSmallNinMax + 1 if more than the tables hold
int nin_for_bound(const mpres_ctx *c, const BoundKey &k) {
    if (k.e == LLONG_MIN) return 1;
    double mant;
    const unsigned long long mb = k.m | 0x3ff0000000000000ull;
    memcpy(&mant, &mb, 8);
    const double lx = (double) (k.e - 1023) + std::log2(mant);
    const double b = (lx + c->sc.log2M_up) * 1024.0;
    const long long xb = b > 1.0e8 ? 100000000 : (b < 0 ? 0 : (long long) std::ceil(b) + 4);
    const int cmax = std::min(kSmallNinMax, c->hc.N - 1);
    for (int q = 1; q <= cmax; ++q)
        if (c->sc.in_log2_milli[q] >= xb) return q;
    return kSmallNinMax + 1;
}
int host_threads() {
    static int t = 0;
    if (t == 0) {
        const char *env = getenv("MPRES_HOST_THREADS");
        t = env ? atoi(env) : (int) std::thread::hardware_concurrency();
        t = std::max(1, std::min(t, 64));
    }
    return t;
}
// records [0, cnt) at src (rs bytes each) -> lean records at dst; returns the largest bound.  nin even: everything moves as 8-byte words
// (records are 4N + 40 bytes with N even).
BoundKey pack_lean(const char *src, size_t cnt, int N, int nin, char *dst) {
    const size_t rw = (4 * (size_t) N + 40) / 8, Nw = (size_t) N / 2;
    const int T = (int) std::min<size_t>((size_t) host_threads(), std::max<size_t>(1, cnt / 4096));
    std::vector<BoundKey> keys((size_t) T, BoundKey{LLONG_MIN, 0ull});
    auto work = [&](int t) {
        BoundKey k{LLONG_MIN, 0ull};
        const size_t b = cnt * (size_t) t / T, e = cnt * (size_t) (t + 1) / T;
        const unsigned long long *s8 = (const unsigned long long *) src;
        unsigned long long *d8 = (unsigned long long *) dst;
        switch (nin / 2) {
            case 1: pack_lean_range<1>(s8, rw, Nw, b, e, d8, k); break;
            case 2: pack_lean_range<2>(s8, rw, Nw, b, e, d8, k); break;
            case 3: pack_lean_range<3>(s8, rw, Nw, b, e, d8, k); break;
            case 4: pack_lean_range<4>(s8, rw, Nw, b, e, d8, k); break;
            case 5: pack_lean_range<5>(s8, rw, Nw, b, e, d8, k); break;
            case 6: pack_lean_range<6>(s8, rw, Nw, b, e, d8, k); break;
            case 7: pack_lean_range<7>(s8, rw, Nw, b, e, d8, k); break;
            default: pack_lean_range<8>(s8, rw, Nw, b, e, d8, k); break;
        }
        keys[(size_t) t] = k;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    BoundKey k{LLONG_MIN, 0ull};
    for (const BoundKey &q : keys) if (q.e > k.e || (q.e == k.e && q.m > k.m)) k = q;
    return k;
}
// records [off, off + cnt) of a host AoS array into the same positions of a device SoA, as lean records with nin residues.  *need = residues
// the largest entry requires (> nin: the caller has to upload this range in full)
int host_upload_lean(HostPipe &hp, const void *host, size_t off, size_t cnt, const SoA &dst, int nin, int *need) {
    mpres_ctx *c = hp.c;
    const int N = c->hc.N;
    const size_t ls = 4 * (size_t) nin + 24;
    const size_t chunk = std::max<size_t>(1, std::min(hp.chunk * hp.rs, (size_t) 64 << 20) / ls);   // records per staging slot (device: hp.chunk records of hp.rs bytes; host: 64 MiB)
    BoundKey key{LLONG_MIN, 0ull};
    while (cnt) {
        const size_t now = std::min(cnt, chunk);
        const int slot = hp.seq % kHostRing;
        cudaEvent_t done = c->hev[slot], freed = c->hev[kHostRing + slot], hfree = c->hev[10 + slot];
        char *hstage = hp.lean_host + (size_t) slot * ((size_t) 64 << 20);
        char *stage = hp.up_stage + (size_t) slot * hp.chunk * hp.rs;
        CUDA_TRY(cudaEventSynchronize(hfree));                                   // the copy that last read this host slot has finished
        const BoundKey k2 = pack_lean((const char *) host + off * hp.rs, now, N, nin, hstage);
        if (k2.e > key.e || (k2.e == key.e && k2.m > key.m)) key = k2;
        CUDA_TRY(cudaStreamWaitEvent(hp.s_up, freed, 0));
        CUDA_TRY(cudaMemcpyAsync(stage, hstage, now * ls, cudaMemcpyHostToDevice, hp.s_up));
        CUDA_TRY(cudaEventRecord(done, hp.s_up));
        CUDA_TRY(cudaEventRecord(hfree, hp.s_up));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_unp, done, 0));
        k_lean_to_soa<<<c->sm_count * 4, 256, 0, hp.s_unp>>>(N, nin, stage, (long long) now, dst.digits + off * N, dst.sign + off, dst.exp + off, dst.eval + off,
                                                             dst.len_val);
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(freed, hp.s_unp));
        ++hp.seq;
        off += now; cnt -= now;
    }
    *need = nin_for_bound(c, key);
    return 0;
}
// n_in guessed from up to 4096 records spread over an operand
int sample_nin(const mpres_ctx *c, const void *host, size_t cnt) {
    const size_t rs = 4 * (size_t) c->hc.N + 40;
    BoundKey k{LLONG_MIN, 0ull};
    const size_t step = std::max<size_t>(1, cnt / 4096);
    for (size_t i = 0; i < cnt; i += step) {
        const char *r = (const char *) host + i * rs;
        double fr; long long ex;
        memcpy(&fr, r + 4 * c->hc.N + 24, 8); memcpy(&ex, r + 4 * c->hc.N + 32, 8);
        bound_key_update(k, fr, ex);
    }
    return nin_for_bound(c, k);
}

SoA soa_offset(const SoA &a, size_t off, int N) {
    SoA v = a;
    v.digits += off * N; v.sign += off; v.exp += off; v.eval += off;      // len_val (the offset of the upper bounds) is unchanged
    return v;
}

}  // namespace

// B comes from the host (Bd == nullptr) or is already resident on the device (Bd: e.g. gathered from the other ranks' slices)
static int gemm_host_impl(mpres_ctx *c, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, const SoA *Bd,
                          int ldb, const void *beta, const void *Cin, void *Cout, int ldc, int panels) {
    NEED_DEVICE(c);
    if (!c || !alpha || !A || (!B && !Bd) || !beta || !Cin || !Cout) return -1;
    if (m <= 0 || n <= 0 || k <= 0) return 0;                       // src/blas/gemm.cuh:75-78
    const bool ta = transa != MPRES_NO_TRANS, tb = transb != MPRES_NO_TRANS;
    if (transa != MPRES_NO_TRANS && transa != MPRES_TRANS && transa != MPRES_CONJ_TRANS) return -2;
    if (transb != MPRES_NO_TRANS && transb != MPRES_TRANS && transb != MPRES_CONJ_TRANS) return -2;
    if (lda < std::max(1, ta ? k : m)) return -3;
    if (ldb < std::max(1, tb ? n : k)) return -4;
    if (ldc < std::max(1, m)) return -5;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> hl(c->host_mu);
    const int N = c->hc.N;
    if (!c->host_ready) {
        for (int i = 0; i < 4; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&c->hs[i], cudaStreamNonBlocking));
        for (int i = 0; i < 16; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c->hev[i], cudaEventDisableTiming));
        c->host_ready = true;
    }
    // panels: whole columns of B and C are contiguous only when B is not transposed
    int np = panels;
    if (np <= 0) { const int w = ((n + 7) / 8 + 255) / 256 * 256; np = (n + w - 1) / w; }
    if (tb && !Bd) np = 1;
    np = std::max(1, std::min(np, n));
    const int wcols = (n + np - 1) / np;
    np = (n + wcols - 1) / wcols;

    const size_t extA = (size_t) lda * ((ta ? m : k) - 1) + (ta ? k : m), lenA = (size_t) lda * (ta ? m : k);
    const size_t extB = (size_t) ldb * ((tb ? k : n) - 1) + (tb ? n : k), lenB = (size_t) ldb * (tb ? k : n);
    const size_t lenC = (size_t) ldc * n;
    HostPipe hp;
    hp.c = c; hp.rs = 4 * (size_t) N + 40;
    hp.chunk = std::max<size_t>(1, ((size_t) 64 << 20) / hp.rs);
    hp.s_up = c->hs[0]; hp.s_unp = c->hs[1]; hp.s_comp = c->hs[2]; hp.s_down = c->hs[3];
    SoA dA, dB, dC, dS;
    int rc;
    void *p;
    // (a growing workspace frees and reallocates: cudaFree synchronises, no transfer of an earlier call is in flight here anyway)
    if ((rc = ws_soa(c, 12, lenA, &dA)) || (rc = ws_soa(c, 14, lenC, &dC)) || (rc = ws_soa(c, 17, 2, &dS))) return rc;
    if (Bd) dB = *Bd;
    else if ((rc = ws_soa(c, 13, lenB, &dB))) return rc;
    if ((rc = ws_reserve(c, 15, (size_t) kHostRing * hp.chunk * hp.rs, &p))) return rc;
    hp.up_stage = (char *) p;
    const size_t panel_recs = (size_t) ldc * wcols;
    if ((rc = ws_reserve(c, 16, 2 * panel_recs * hp.rs, &p))) return rc;
    char *down_stage = (char *) p;
    cudaEvent_t ev_ready = c->hev[6], ev_packed = c->hev[7], ev_dfree[2] = {c->hev[8], c->hev[9]};

    // lean upload of A and B (see host_upload_lean): only in the fast modes, on formats with the one-byte base
    int lean_nin = 0;                    // residues the lean uploads carry (0: everything goes up in full)
    bool A_full = true;
    const bool dbg = getenv("MPRES_DEBUG_LEAN") != nullptr;
    {
        const char *env = getenv("MPRES_HOST_LEAN");
        const char *envmin = getenv("MPRES_HOST_LEAN_MIN");                   // smallest operand (entries) worth the packing threads
        const bool want = c->mode != MPRES_MODE_REFERENCE_ORDER && c->sc.usable && c->stage2 == MPRES_STAGE2_SMALL && (!env || atoi(env) != 0) &&
                          (size_t) m * k >= (size_t) (envmin ? atoll(envmin) : 65536);
        if (want) {
            lean_nin = sample_nin(c, A, extA);
            if (!Bd) lean_nin = std::max(lean_nin, sample_nin(c, B, extB));
            lean_nin = (lean_nin + 1) & ~1;                                                         // whole 8-byte words per record
            if (lean_nin > kSmallNinMax || 4 * lean_nin + 24 > (int) hp.rs / 2) lean_nin = 0;      // no gain
        }
        if (lean_nin) {
            if (!c->lean_host) {
                if (cudaHostAlloc(&c->lean_host, (size_t) kHostRing * ((size_t) 64 << 20), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); c->lean_host = nullptr; lean_nin = 0; }
                else for (int i = 0; i < kHostRing; ++i) CUDA_TRY(cudaEventRecord(c->hev[10 + i], hp.s_up));
            }
            hp.lean_host = (char *) c->lean_host;
        }
    }
    c->last_host_lean = 0;
    if ((rc = host_upload(hp, alpha, 0, 1, dS))) return rc;
    { SoA b1 = soa_offset(dS, 1, N); if ((rc = host_upload(hp, beta, 0, 1, b1))) return rc; }
    // MPRES_HOST_C_AHEAD=n: the first n panels of C go before A, so that their copies keep the link busy while the host cores pack A (measured at
    // config 3: 137.3 ms against 127.2 ms without -- the lean chunks of A then queue behind 0.7 GB of C and the first panel starts later; default 0)
    const char *env_ca = getenv("MPRES_HOST_C_AHEAD");
    const int c_ahead = (lean_nin && env_ca) ? std::max(0, std::min(np, atoi(env_ca))) : 0;
    for (int j = 0; j < c_ahead; ++j) {
        const int j0 = j * wcols, nj = std::min(wcols, n - j0);
        if ((rc = host_upload(hp, Cin, (size_t) ldc * j0, (size_t) ldc * (nj - 1) + m, dC))) return rc;
    }
    if (lean_nin) {
        int need = 0;
        if ((rc = host_upload_lean(hp, A, 0, extA, dA, lean_nin, &need))) return rc;
        if (dbg) fprintf(stderr, "[mpres lean] sample says %d residues, A needs %d\n", lean_nin, need);
        A_full = false;
        if (need > lean_nin) { if ((rc = host_upload(hp, A, 0, extA, dA))) return rc; A_full = true; }      // the sample missed the largest entries
    } else if ((rc = host_upload(hp, A, 0, extA, dA))) return rc;
    bool B_lean_ok = true;               // the panels of B uploaded so far fit lean_nin
    auto upload_B = [&](size_t off, size_t cnt) -> int {
        if (!lean_nin) return host_upload(hp, B, off, cnt, dB);
        int need = 0, e = host_upload_lean(hp, B, off, cnt, dB, lean_nin, &need);
        if (e) return e;
        if (dbg) fprintf(stderr, "[mpres lean] B range %zu + %zu: needs %d residues (lean %d)\n", off, cnt, need, lean_nin);
        if (need > lean_nin) {
            B_lean_ok = false;
            if ((e = host_upload(hp, B, off, cnt, dB))) return e;
            if (!A_full) { if ((e = host_upload(hp, A, 0, extA, dA))) return e; A_full = true; }           // the call will read more residues of A too
            lean_nin = 0;                                                                                   // (no further attempts in this call)
        }
        return 0;
    };
    if (tb && !Bd && (rc = upload_B(0, extB))) return rc;
    cudaEvent_t ev_in[2] = {ev_ready, c->hev[13]};
    // the inputs of panel j: C first (its copy runs while the host cores pack the panel of B), then B
    auto issue_inputs = [&](int j) -> int {
        const int j0 = j * wcols, nj = std::min(wcols, n - j0);
        const size_t coff = (size_t) ldc * j0, ccnt = (size_t) ldc * (nj - 1) + m;
        int e = 0;
        if (j >= c_ahead && (e = host_upload(hp, Cin, coff, ccnt, dC))) return e;
        if (!tb && !Bd && (e = upload_B((size_t) ldb * j0, (size_t) ldb * (nj - 1) + k))) return e;
        CUDA_TRY(cudaEventRecord(ev_in[j & 1], hp.s_unp));
        return 0;
    };
    if ((rc = issue_inputs(0))) return rc;
    for (int j = 0; j < np; ++j) {
        const int j0 = j * wcols, nj = std::min(wcols, n - j0);
        const size_t boff = (size_t) ldb * j0;
        const size_t coff = (size_t) ldc * j0, ccnt = (size_t) ldc * (nj - 1) + m;
        // one panel ahead: the next panel's inputs travel (and are packed) while this one is multiplied -- gemm_impl waits on the host for its
        // base choice, which would otherwise leave the link idle
        if (j + 1 < np && (rc = issue_inputs(j + 1))) return rc;
        CUDA_TRY(cudaStreamWaitEvent(hp.s_comp, ev_in[j & 1], 0));
        // panel j of op(B): columns j0.. of a k x n matrix, or rows j0.. of an n x k one (whole B when it was uploaded in one piece)
        rc = gemm_impl(c, transa, transb, m, nj, k, dS, dA, lda, soa_offset(dB, tb ? (size_t) j0 : (size_t) ldb * j0, N), ldb, soa_offset(dS, 1, N),
                       soa_offset(dC, coff, N), ldc, nullptr, hp.s_comp);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        if (lean_nin && (!A_full || (!Bd && B_lean_ok))) {
            // the call chose its base and its residue count on the device (gemm_impl read them back): anything but the one-byte base with at
            // most lean_nin residues has read fields that were never uploaded -- upload everything that is still to be used and run the panel
            // again (C from the host: its results have not been downloaded yet)
            const bool ok = c->last_fast_ok && c->last_nin > 0 && c->last_nin <= lean_nin;
            if (dbg) fprintf(stderr, "[mpres lean] panel %d: the call read %d residues on the one-byte base %s\n", j, c->last_nin, c->last_fast_ok ? "(ok)" : "(NOT chosen)");
            if (!ok) {
                if (!A_full) { if ((rc = host_upload(hp, A, 0, extA, dA))) return rc; A_full = true; }
                if (!Bd) { if ((rc = host_upload(hp, B, tb ? 0 : boff, tb ? extB : extB - boff, dB))) return rc; }       // this panel and all later ones
                if ((rc = host_upload(hp, Cin, coff, ccnt, dC))) return rc;
                lean_nin = 0;
                CUDA_TRY(cudaEventRecord(ev_in[j & 1], hp.s_unp));
                CUDA_TRY(cudaStreamWaitEvent(hp.s_comp, ev_in[j & 1], 0));
                rc = gemm_impl(c, transa, transb, m, nj, k, dS, dA, lda, soa_offset(dB, tb ? (size_t) j0 : (size_t) ldb * j0, N), ldb, soa_offset(dS, 1, N),
                               soa_offset(dC, coff, N), ldc, nullptr, hp.s_comp);
                if (rc) { cudaDeviceSynchronize(); return rc; }
            } else {
                c->last_host_lean = lean_nin;
            }
        }
        char *stage = down_stage + (size_t) (j & 1) * panel_recs * hp.rs;
        CUDA_TRY(cudaStreamWaitEvent(hp.s_comp, ev_dfree[j & 1], 0));
        k_soa_to_aos<<<c->sm_count * 4, 256, 0, hp.s_comp>>>(N, stage, (long long) ccnt, dC.digits + coff * N, dC.sign + coff, dC.exp + coff, dC.eval + coff,
                                                             dC.len_val);
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev_packed, hp.s_comp));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_down, ev_packed, 0));
        CUDA_TRY(cudaMemcpyAsync((char *) Cout + coff * hp.rs, stage, ccnt * hp.rs, cudaMemcpyDeviceToHost, hp.s_down));
        CUDA_TRY(cudaEventRecord(ev_dfree[j & 1], hp.s_down));
    }
    CUDA_TRY(cudaStreamSynchronize(hp.s_down));
    CUDA_TRY(cudaStreamSynchronize(hp.s_comp));
    CUDA_TRY(cudaStreamSynchronize(hp.s_unp));
    return 0;
}

// every exit of the pipelined call, failures included, leaves no copy in flight on the caller's buffers
static int gemm_host_checked(mpres_ctx *c, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, const SoA *Bd,
                             int ldb, const void *beta, const void *Cin, void *Cout, int ldc, int panels) {
    const int rc = gemm_host_impl(c, transa, transb, m, n, k, alpha, A, lda, B, Bd, ldb, beta, Cin, Cout, ldc, panels);
    if (rc != 0 && c && c->device >= 0 && c->host_ready) {
        DeviceGuard g(c->device);
        for (int i = 0; i < 4; ++i) cudaStreamSynchronize(c->hs[i]);
        cudaGetLastError();
    }
    return rc;
}

int mpres_gemm_host(mpres_ctx *c, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const void *B, int ldb,
                    const void *beta, const void *Cin, void *Cout, int ldc, int panels) {
    if (!B) return -1;
    return gemm_host_checked(c, transa, transb, m, n, k, alpha, A, lda, B, nullptr, ldb, beta, Cin, Cout, ldc, panels);
}
int mpres_gemm_host_bdev(mpres_ctx *c, int transa, int transb, int m, int n, int k, const void *alpha, const void *A, int lda, const mpres_array_t *B, int ldb,
                         const void *beta, const void *Cin, void *Cout, int ldc, int panels) {
    if (!B) return -1;
    const SoA bd = view(B);
    return gemm_host_checked(c, transa, transb, m, n, k, alpha, A, lda, nullptr, &bd, ldb, beta, Cin, Cout, ldc, panels);
}

int mpres_gemm_coll(mpres_ctx *c, int transa, int transb, int m, int n, int k, const mpres_collection_t *alpha,
                    const mpres_collection_t *A, int lda, size_t lenA, const mpres_collection_t *B, int ldb, size_t lenB,
                    const mpres_collection_t *beta, mpres_collection_t *Cm, int ldc, size_t lenC, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !alpha || !A || !B || !beta || !Cm) return -1;
    return gemm_impl(c, transa, transb, m, n, k, view(alpha, 1), view(A, lenA), lda, view(B, lenB), ldb, view(beta, 1),
                     view(Cm, lenC), ldc, nullptr, (cudaStream_t) stream);
}

/* ---- row-sharded GEMM over the GPUs of one NVLink domain ------------------------------------------------------------------
 * The reference has no multi-GPU code; SURVEY 8(e) / BASELINE config 3: A and C split into row blocks, B needed by every rank.  One
 * process (or thread) per GPU.  A communicator owns this rank's receive buffer (one cudaMalloc: exchange slots, arrival flags, one
 * package slot per rank) and the peer mappings of the other ranks' buffers (CUDA IPC between processes, peer access inside one). */

struct mpres_shard {
    mpres_ctx *ctx = nullptr;
    int rank = 0, world = 1, n = 0, k_max = 0;
    long long nb_p = 0, k_p_max = 0;
    size_t pkg_stride = 0, comm_bytes = 0;
    char *comm = nullptr;
    char *peer_comm[kMaxPanels] = {nullptr};
    bool ipc_opened[kMaxPanels] = {false};
    bool connected = false;
    unsigned epoch = 0;
    cudaStream_t push[kMaxPanels] = {nullptr};
    cudaEvent_t ev_pkg = nullptr, ev_push[kMaxPanels] = {nullptr};
    int npush = 0;
};

namespace {
struct ShardHandle {          // what a rank publishes to the others (opaque to the caller)
    int pid, device, rank, world;
    unsigned long long bytes;
    void *raw;
    cudaIpcMemHandle_t mem;
};
constexpr size_t kShardXchgOff = 0, kShardFlagsOff = 1024, kShardPkgOff = 4096;
}  // namespace

size_t mpres_shard_handle_size(void) { return sizeof(ShardHandle); }

int mpres_shard_create(mpres_ctx *c, int rank, int world, int n, int k_max, mpres_shard **out) {
    NEED_DEVICE(c);
    if (!c || !out || world < 1 || world > kMaxPanels || rank < 0 || rank >= world || n <= 0 || k_max <= 0 || n % world != 0) return -1;
    DeviceGuard g(c->device);
    mpres_shard *s = new mpres_shard();
    s->ctx = c; s->rank = rank; s->world = world; s->n = n; s->k_max = k_max;
    s->nb_p = round_up(n / world, 256);
    s->k_p_max = round_up(k_max, 128);
    s->pkg_stride = (pkg_header_bytes(s->nb_p, s->k_p_max) + (size_t) kSmallMax * s->nb_p * s->k_p_max + 4095) & ~(size_t) 4095;
    s->comm_bytes = kShardPkgOff + (size_t) world * s->pkg_stride;
    cudaError_t e = cudaMalloc(&s->comm, s->comm_bytes);
    if (e == cudaSuccess) e = cudaMemset(s->comm, 0, kShardPkgOff);
    const char *env = getenv("MPRES_PUSH_STREAMS");
    s->npush = env ? atoi(env) : 2;
    s->npush = std::max(1, std::min(s->npush, std::max(1, world - 1)));
    for (int i = 0; i < s->npush && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&s->push[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_push[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_pkg, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaGetLastError(); mpres_shard_destroy(s); return (int) e; }
    s->peer_comm[rank] = s->comm;
    s->connected = world == 1;
    *out = s;
    return 0;
}

int mpres_shard_export(mpres_shard *s, void *handle_out) {
    if (!s || !handle_out) return -1;
    DeviceGuard g(s->ctx->device);
    ShardHandle h;
    memset(&h, 0, sizeof(h));
    h.pid = (int) getpid(); h.device = s->ctx->device; h.rank = s->rank; h.world = s->world; h.bytes = s->comm_bytes; h.raw = s->comm;
    CUDA_TRY(cudaIpcGetMemHandle(&h.mem, s->comm));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int mpres_shard_connect(mpres_shard *s, const void *handles) {
    if (!s || !handles) return -1;
    DeviceGuard g(s->ctx->device);
    const ShardHandle *hs = (const ShardHandle *) handles;
    for (int p = 0; p < s->world; ++p) {
        const ShardHandle &h = hs[p];
        if (h.rank != p || h.world != s->world || h.bytes != s->comm_bytes) return -8;     // the ranks were created with different shapes
        if (p == s->rank) continue;
        if (h.pid == (int) getpid()) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, s->ctx->device, h.device));
            if (!can) return -9;
            const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int) e;
            cudaGetLastError();
            s->peer_comm[p] = (char *) h.raw;
        } else {
            void *ptr = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h.mem, cudaIpcMemLazyEnablePeerAccess));
            s->peer_comm[p] = (char *) ptr;
            s->ipc_opened[p] = true;
        }
    }
    s->connected = true;
    return 0;
}

int mpres_shard_destroy(mpres_shard *s) {
    if (!s) return -1;
    DeviceGuard g(s->ctx->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < kMaxPanels; ++p) if (s->ipc_opened[p] && s->peer_comm[p]) cudaIpcCloseMemHandle(s->peer_comm[p]);
    for (int i = 0; i < kMaxPanels; ++i) { if (s->push[i]) cudaStreamDestroy(s->push[i]); if (s->ev_push[i]) cudaEventDestroy(s->ev_push[i]); }
    if (s->ev_pkg) cudaEventDestroy(s->ev_pkg);
    if (s->comm) cudaFree(s->comm);
    delete s;
    return 0;
}

int mpres_gemm_sharded(mpres_shard *s, int transa, int transb, int m_local, int n, int k, const mpres_array_t *alpha, const mpres_array_t *A, int lda,
                       const mpres_array_t *B, int ldb, const mpres_array_t *beta, mpres_array_t *Cm, int ldc, mpres_stream_t stream) {
    if (!s || !alpha || !A || !B || !beta || !Cm) return -1;
    if (!s->connected) return -10;
    if (n != s->n || k > s->k_max) return -11;
    mpres_ctx *c = s->ctx;
    FastShard sh;
    sh.rank = s->rank; sh.world = s->world;
    sh.epoch = ++s->epoch;
    sh.recv = s->comm + kShardPkgOff;
    sh.pkg_stride = s->pkg_stride;
    sh.flags = (unsigned *) (s->comm + kShardFlagsOff);
    for (int p = 0; p < s->world; ++p) {
        sh.peer_recv[p] = s->peer_comm[p] + kShardPkgOff;
        sh.peer_xchg[p] = (int *) (s->peer_comm[p] + kShardXchgOff);
        sh.peer_flags[p] = (unsigned *) (s->peer_comm[p] + kShardFlagsOff);
    }
    sh.npush = s->npush;
    for (int i = 0; i < s->npush; ++i) { sh.push[i] = s->push[i]; sh.ev_push[i] = s->ev_push[i]; }
    sh.ev_pkg = s->ev_pkg;
    return gemm_impl(c, transa, transb, m_local, n, k, view(alpha), view(A), lda, view(B), ldb, view(beta), view(Cm), ldc, nullptr, (cudaStream_t) stream,
                     s->world > 1 ? &sh : nullptr);
}

/* ---- GEMV ------------------------------------------------------------------------------------------ */

static int gemv_impl(mpres_ctx *c, int trans, int m, int n, SoA alpha, SoA A, int lda, SoA x, int incx, SoA beta, SoA y, int incy,
                     cudaStream_t st) {
    // src/blas/gemv.cuh:155-161
    if (m <= 0 || n <= 0) return 0;
    if (incx == 0 || incy == 0 || lda < std::max(1, m)) return -3;
    if (trans != MPRES_NO_TRANS && trans != MPRES_TRANS && trans != MPRES_CONJ_TRANS) return -2;
    const bool tr = trans != MPRES_NO_TRANS;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    const int N = c->hc.N;
    const int lenx = tr ? m : n, leny = tr ? n : m;
    SoA ax;
    rc = ws_soa(c, 1, (size_t) lenx, &ax);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, kCounterInts * sizeof(int), st));
    c->ev_valid = false;
    mv_mark(c, 0, st);
    MPRES_DISPATCH(N, {
        int block = 128;
        auto nb = [&](long long groups) { return (unsigned) std::min<long long>((groups * G + block - 1) / block, (long long) c->sm_count * 32); };
        // buffer1 = round(alpha * x), y = round(beta * y)          gemv.cuh:175-190
        k_vec_scale<G, R><<<nb(lenx), block, 0, st>>>(c->dconsts, lenx, ax, 1, x, incx, alpha);
        k_vec_scale<G, R><<<nb(leny), block, 0, st>>>(c->dconsts, leny, y, incy, y, incy, beta);
    });
    LAUNCHED(c); LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    bool done = false;
    if (c->mode != MPRES_MODE_REFERENCE_ORDER) {
        rc = gemv_fast(c, tr, m, n, A, lda, ax, y, incy, st, &done);
        if (rc) return rc;
    }
    if (!done) {
        MPRES_DISPATCH(N, {
            int block = 128;
            long long blocks = std::min<long long>(((long long) leny * G + block - 1) / block, (long long) c->sm_count * 64);
            k_gemv_ref_order<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, tr, m, n, A, lda, ax, y, incy, nullptr, nullptr);
        });
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
    }
    return call_end(c, st);
}

int mpres_gemv(mpres_ctx *c, int trans, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda,
               const mpres_array_t *x, int incx, const mpres_array_t *beta, mpres_array_t *y, int incy,
               mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer1; (void) buffer2;
    if (!c || !alpha || !A || !x || !beta || !y) return -1;
    return gemv_impl(c, trans, m, n, view(alpha), view(A), lda, view(x), incx, view(beta), view(y), incy, (cudaStream_t) stream);
}
int mpres_gemv_coll(mpres_ctx *c, int trans, int m, int n, const mpres_collection_t *alpha, const mpres_collection_t *A, int lda,
                    size_t lenA, const mpres_collection_t *x, int incx, size_t lenx, const mpres_collection_t *beta,
                    mpres_collection_t *y, int incy, size_t leny, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !alpha || !A || !x || !beta || !y) return -1;
    return gemv_impl(c, trans, m, n, view(alpha, 1), view(A, lenA), lda, view(x, lenx), incx, view(beta, 1), view(y, leny), incy,
                     (cudaStream_t) stream);
}

/* ---- SCAL, AXPY (the first of the other v1 operations) ------------------------------ */

int mpres_scal(mpres_ctx *c, int n, const mpres_array_t *alpha, mpres_array_t *x, int incx, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !alpha || !x) return -1;
    if (n <= 0 || incx <= 0) return 0;      // src/blas/scal.cuh:47-49 (silent return)
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_vec_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(x), incx, view(x), incx, view(alpha));
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_axpy(mpres_ctx *c, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, mpres_array_t *y, int incy,
               mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer;
    if (!c || !alpha || !x || !y) return -1;
    if (n <= 0) return 0;                   // src/blas/axpy.cuh:49-51
    if (incx == 0 || incy == 0) return -3;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_vec_axpy<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(alpha), view(x), incx, view(y), incy);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_waxpby(mpres_ctx *c, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, const mpres_array_t *beta,
                 const mpres_array_t *y, int incy, mpres_array_t *w, int incw, mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer;
    if (!c || !alpha || !x || !beta || !y || !w) return -1;
    if (n <= 0) return 0;                   // src/blas/waxpby.cuh:53-55
    if (incx == 0 || incy == 0 || incw == 0) return -3;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_vec_waxpby<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(alpha), view(x), incx, view(beta), view(y), incy, view(w), incw);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_ge_add(mpres_ctx *c, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda, const mpres_array_t *beta,
                 const mpres_array_t *B, int ldb, mpres_array_t *Cm, int ldc, mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer;
    if (!c || !alpha || !A || !beta || !B || !Cm) return -1;
    if (m <= 0 || n <= 0) return 0;         // src/blas/geadd.cuh:60-63
    if (lda < std::max(1, m) || ldb < std::max(1, m) || ldc < std::max(1, m)) return -3;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) m * n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_ge_add<G, R><<<nb, block, 0, st>>>(c->dconsts, m, n, view(alpha), view(A), lda, view(beta), view(B), ldb, view(Cm), ldc);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_ge_acc(mpres_ctx *c, int m, int n, const mpres_array_t *alpha, const mpres_array_t *A, int lda, const mpres_array_t *beta,
                 mpres_array_t *B, int ldb, mpres_array_t *buffer, mpres_stream_t stream) {
    return mpres_ge_add(c, m, n, alpha, A, lda, beta, B, ldb, B, ldb, buffer, stream);   // src/blas/geacc.cuh:57: B = alpha A + beta B
}

int mpres_ger(mpres_ctx *c, int m, int n, const mpres_array_t *alpha, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy,
              mpres_array_t *A, int lda, mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer1; (void) buffer2;
    if (!c || !alpha || !x || !y || !A) return -1;
    if (m < 0 || n < 0 || lda < std::max(1, m)) return -3;   // src/blas/ger.cuh:160-168
    if (incx == 0 || incy == 0) return -3;
    if (m == 0 || n == 0) return 0;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) m * n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_ger<G, R><<<nb, block, 0, st>>>(c->dconsts, m, n, view(alpha), view(x), incx, view(y), incy, view(A), lda);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int ge_scale_impl(mpres_ctx *c, int m, int n, const mpres_array_t *DL, int incdl, const mpres_array_t *DR, int incdr, mpres_array_t *A, int lda,
                         cudaStream_t st) {
    DeviceGuard g(c->device);
    const bool left = DL != nullptr, right = DR != nullptr;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) m * n * G + block - 1) / block, (long long) c->sm_count * 32);
        k_ge_diag_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, m, n, left ? view(DL) : view(A), incdl, right ? view(DR) : view(A), incdr, left, right,
                                                    view(A), lda);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int mpres_ge_diag_scale(mpres_ctx *c, int side, int m, int n, const mpres_array_t *D, int incd, mpres_array_t *A, int lda, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !D || !A) return -1;
    if (side != MPRES_LEFT_SIDE && side != MPRES_RIGHT_SIDE) return -3;
    if (m <= 0 || n <= 0) return 0;                         // src/blas/gediagscale.cuh:57-59
    if (incd == 0 || lda < std::max(1, m)) return -3;       // :61-63 (the reference returns silently)
    const bool left = side == MPRES_LEFT_SIDE;
    return ge_scale_impl(c, m, n, left ? D : nullptr, incd, left ? nullptr : D, incd, A, lda, (cudaStream_t) stream);
}

int mpres_ge_lr_scale(mpres_ctx *c, int m, int n, const mpres_array_t *DL, int incdl, const mpres_array_t *DR, int incdr, mpres_array_t *A, int lda,
                      mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !DL || !DR || !A) return -1;
    if (m <= 0 || n <= 0) return 0;                         // src/blas/gelrscale.cuh:59-61
    if (incdl == 0 || incdr == 0 || lda < std::max(1, m)) return -3;
    return ge_scale_impl(c, m, n, DL, incdl, DR, incdr, A, lda, (cudaStream_t) stream);
}

int mpres_rot(mpres_ctx *c, int n, mpres_array_t *x, int incx, mpres_array_t *y, int incy, const mpres_array_t *cs, const mpres_array_t *sn,
              mpres_array_t *buffer1, mpres_array_t *buffer2, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !x || !y || !cs || !sn) return -1;
    if (n <= 0) return 0;                                   // src/blas/rot.cuh:52-54
    if (incx == 0 || incy == 0) return -3;
    const bool unit = incx == 1 && incy == 1;
    if (!unit && (!buffer1 || !buffer2)) return -1;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        const int block = 128;
        const unsigned nb = (unsigned) std::min<long long>(((long long) n * G + block - 1) / block, (long long) c->sm_count * 32);
        if (unit) {
            k_vec_rot<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(x), view(y), view(cs), view(sn));
        } else {
            // the reference's sequence, step for step -- including its mp_scal(n, c, x, 1) calls, which scale the first n
            // CONTIGUOUS elements whatever incx is (src/blas/rot.cuh:79-82)
            k_vec_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(buffer1), 1, view(x), incx, view(sn));
            k_vec_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(buffer2), 1, view(y), incy, view(sn));
            k_vec_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(x), 1, view(x), 1, view(cs));
            k_vec_scale<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(y), 1, view(y), 1, view(cs));
            k_vec_addsub<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(x), incx, view(buffer2), 1, false);
            k_vec_addsub<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(y), incy, view(buffer1), 1, true);
        }
    });
    if (unit) { LAUNCHED(c); } else { for (int i = 0; i < 6; ++i) LAUNCHED(c); }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

/* ---- DOT ------------------------------------------------------------------------------------------- */

// partial (device AoS record) := sum x_i * y_i
static int dot_to_record(mpres_ctx *c, int n, SoA x, int incx, SoA y, int incy, char *rec_out, SoA out, cudaStream_t st) {
    const int N = c->hc.N;
    const size_t rs = 4 * (size_t) N + 40;
    bool done = false, tried = false;
    int rc = call_begin(c, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, kCounterInts * sizeof(int), st));
    c->ev_valid = false;
    if (c->mode != MPRES_MODE_REFERENCE_ORDER) {
        rc = dot_fast(c, n, x, incx, y, incy, rec_out, out, st, &done, &tried);
        if (rc) return rc;
    }
    if (done) return call_end(c, st);
    // reference order: always in REFERENCE_ORDER mode; in AUTO mode gated on the fast path's guard counter
    const int *gate = tried ? c->d_counter : nullptr;
    MPRES_DISPATCH(N, {
        const int block = 128;
        const long long gpb = block / G;
        long long blocks = std::min<long long>(((long long) n + gpb - 1) / gpb, (long long) c->sm_count * 8);
        const long long groups = blocks * gpb;
        void *parts;
        rc = ws_reserve(c, 2, (size_t) groups * rs, &parts);
        if (rc) return rc;
        k_dot_partial<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, n, x, incx, y, incy, (char *) parts, gate);
        k_tree_records<G, R><<<1, 256, 0, st>>>(c->dconsts, (char *) parts, groups, out, 0, rec_out, gate);
    });
    LAUNCHED(c); LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return call_end(c, st);
}

int mpres_dot(mpres_ctx *c, int n, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy, mpres_array_t *r,
              mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    (void) buffer;
    if (!c || !x || !y || !r) return -1;
    if (n <= 0) return 0;  // src/blas/dot.cuh:88-90
    if (incx == 0 || incy == 0) return -3;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    c->last_stream = (cudaStream_t) stream;
    return dot_to_record(c, n, view(x), incx, view(y), incy, nullptr, view(r), (cudaStream_t) stream);
}
int mpres_dot_coll(mpres_ctx *c, int n, const mpres_collection_t *x, int incx, size_t lenx, const mpres_collection_t *y, int incy,
                   size_t leny, mpres_collection_t *r, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !x || !y || !r) return -1;
    if (n <= 0) return 0;
    if (incx == 0 || incy == 0) return -3;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    c->last_stream = (cudaStream_t) stream;
    return dot_to_record(c, n, view(x, lenx), incx, view(y, leny), incy, nullptr, view(r, 1), (cudaStream_t) stream);
}
int mpres_axpy_dot(mpres_ctx *c, int n, const mpres_array_t *alpha, mpres_array_t *w, int incw, const mpres_array_t *v, int incv,
                   const mpres_array_t *u, int incu, mpres_array_t *r, mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !alpha || !w || !v || !u || !r) return -1;
    if (n <= 0) return 0;                   // src/blas/axpydot.cuh:36-38
    if (incw == 0 || incv == 0 || incu == 0) return -3;
    {
        DeviceGuard g(c->device);
        cudaStream_t st = (cudaStream_t) stream;
        MPRES_DISPATCH(c->hc.N, {
            const int block = 128;
            const unsigned nb = (unsigned) std::min<long long>(((long long) n * G + block - 1) / block, (long long) c->sm_count * 32);
            k_vec_wsub<G, R><<<nb, block, 0, st>>>(c->dconsts, n, view(alpha), view(v), incv, view(w), incw);
        });
        LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
    }
    return mpres_dot(c, n, u, incu, w, incw, r, buffer, stream);      // src/blas/axpydot.cuh:67
}
/* ---- sums of magnitudes and norms (src/blas/asum.cuh:41, genorm.cuh:142) ------------------------------------------------- */

// out[o] = sum_l |X(o, l)|: the exact-window accumulators when they apply, else the reference-order loops.  The caller holds the lock.
static int abs_sums(mpres_ctx *c, SoA X, long long so, long long sl, int nout, long long nterms, SoA out, cudaStream_t st) {
    const int N = c->hc.N;
    bool done = false;
    CUDA_TRY(cudaMemsetAsync(c->d_counter, 0, kCounterInts * sizeof(int), st));
    if (c->mode != MPRES_MODE_REFERENCE_ORDER && (sl == 1 || so == 1)) {
        int rc = abs_sums_fast(c, X, so, sl, nout, nterms, out, st, &done);
        if (rc) return rc;
    }
    if (!done) {
        if (nout == 1 && nterms > 4096) {
            // one long sum: partial sums per lane group, then a tree (the structure of src/mpreduct.cuh:120-149)
            const size_t rs = 4 * (size_t) N + 40;
            int rc = 0;
            SoA none;
            memset(&none, 0, sizeof(none));
            MPRES_DISPATCH(N, {
                const int block = 128;
                const long long gpb = block / G;
                const long long blocks = std::min<long long>((nterms + gpb - 1) / gpb, (long long) c->sm_count * 8);
                const long long groups = blocks * gpb;
                void *parts;
                rc = ws_reserve(c, 2, (size_t) groups * rs, &parts);
                if (rc) return rc;
                k_dot_partial<G, R><<<(unsigned) blocks, block, 0, st>>>(c->dconsts, nterms, X, (int) sl, none, 1, (char *) parts, nullptr);
                k_tree_records<G, R><<<1, 256, 0, st>>>(c->dconsts, (char *) parts, groups, out, 0, nullptr, nullptr);
            });
            LAUNCHED(c); LAUNCHED(c);
        } else {
            MPRES_DISPATCH(N, {
                const long long blocks = std::max<long long>(1, std::min<long long>(((long long) nout * G + 127) / 128, (long long) c->sm_count * 8));
                k_abs_sum_ref<G, R><<<(unsigned) blocks, 128, 0, st>>>(c->dconsts, X, so, sl, nout, nterms, out, 1, nullptr, nullptr);
            });
            LAUNCHED(c);
        }
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

int mpres_asum(mpres_ctx *c, int n, const mpres_array_t *x, int incx, mpres_array_t *r, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !x || !r) return -1;
    if (n <= 0 || incx <= 0) return 0;   // src/blas/asum.cuh:44-46
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = (cudaStream_t) stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    if ((rc = abs_sums(c, view(x), 0, incx, 1, n, view(r), st))) return rc;
    return call_end(c, st);
}

int mpres_ge_norm(mpres_ctx *c, int norm, int m, int n, const mpres_array_t *A, int lda, mpres_array_t *r, mpres_array_t *buffer, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !A || !r) return -1;
    if (norm != MPRES_ONE_NORM && norm != MPRES_INF_NORM) return -2;
    if (m <= 0 || n <= 0) return 0;      // src/blas/genorm.cuh:145-151
    if (lda < std::max(1, m)) return -3;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t st = (cudaStream_t) stream;
    int rc = call_begin(c, st);
    if (rc) return rc;
    c->last_stream = st;
    const bool one = norm == MPRES_ONE_NORM;         // one norm: largest column sum; infinity norm: largest row sum
    const int nout = one ? n : m;
    SoA sums;
    if (buffer && buffer->digits) sums = view(buffer);
    else if ((rc = ws_soa(c, 20, (size_t) nout, &sums))) return rc;
    if ((rc = abs_sums(c, view(A), one ? lda : 1, one ? 1 : lda, nout, one ? m : n, sums, st))) return rc;
    const SoA rv = view(r);
    if ((rc = mpres_internal_maxabs(c, nout, &sums, 1, &rv, st))) return rc;
    return call_end(c, st);
}

int mpres_dot_partial(mpres_ctx *c, int n, const mpres_array_t *x, int incx, const mpres_array_t *y, int incy, void *partial,
                      mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !x || !y || !partial) return -1;
    if (incx == 0 || incy == 0) return -3;
    DeviceGuard g(c->device);
    std::lock_guard<std::mutex> lk(c->mu);
    c->last_stream = (cudaStream_t) stream;
    if (n <= 0) { CUDA_TRY(cudaMemsetAsync(partial, 0, 4 * (size_t) c->hc.N + 40, (cudaStream_t) stream)); return 0; }
    SoA dummy{};
    return dot_to_record(c, n, view(x), incx, view(y), incy, (char *) partial, dummy, (cudaStream_t) stream);
}
int mpres_reduce_partials(mpres_ctx *c, const void *partials, int count, mpres_array_t *r, mpres_stream_t stream) {
    NEED_DEVICE(c);
    if (!c || !partials || !r || count < 0) return -1;
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t) stream;
    MPRES_DISPATCH(c->hc.N, {
        k_reduce_records<G, R><<<1, G, 0, st>>>(c->dconsts, (const char *) partials, count, view(r), 0, nullptr);
    });
    LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // extern "C"
