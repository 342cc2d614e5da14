// ctx.hpp -- the library context and small host helpers shared by the kernel launchers.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/mpres_b200.h"
#include "host_consts.hpp"
#include "mp_device.cuh"

namespace mpres {

// device tables of the small-modulus stage 2 (host_consts.hpp: SmallConsts; kernels_small.cuh)
struct SmallDev {
    int usable, ext_cols, red_shift, pad;
    int p[64];                     // moduli (1 beyond kSmallMax)
    unsigned mu[64];               // floor(2^32 / p)
    float rcp[64];                 // 1 / p
    int prefix_log2[kSmallMax + 2];
    int in_log2_milli[kSmallNinMax + 1];
    int pad2;
    double log2M_up;
    const uint8_t *inv;            // [kSmallMax + 1][64]
    const uint8_t *ext_b;          // [kSmallMax + 1][ext_cols][64]
    const uint32_t *cw;            // [64][16]
    const uint8_t *pws;            // [2 (kSmallShiftMax + 1)][64]
    const uint32_t *in_mi;         // [kSmallNinMax + 1][16][16]
    const uint32_t *in_negmp;      // [kSmallNinMax + 1][16]
    const uint32_t *red_mu;        // [N]
    const uint32_t *bin_mi;        // [kSmallMax + 1][kSmallMax][kBinW]
    const uint32_t *bin_negmp;     // [kSmallMax + 1][kBinW]
};

}  // namespace mpres

using namespace mpres;

struct mpres_ctx {
    int device = 0;
    int mode = MPRES_MODE_AUTO;
    int stage2 = MPRES_STAGE2_SMALL;  // which stage-2 kernel the fast path launches
    int reduced_base = 1;             // 1: run stages 1-2 on as many moduli as the exact sums need, then extend the base
    int stage1 = 0;                   // 0: vectorised alignment kernel, 1: round-1 kernel
    int align_mma = 0;                // small-modulus alignment: 1 = residues on the tensor cores (measured slower: 2.8 vs 2.3 ms), 0 = dp4a per entry and modulus
    int minplus_sparse = 1;           // (min,+) of the shift planes from candidate lists (0: dense DPX kernel)
    int stage3 = 0;                   // 0: entry-per-thread kernel + list, 1: residue-parallel tile kernel
    int small_tj = 256;               // persistent small-modulus kernel: rows of B' per tile (256: two MMAs share the A' operand; 128: one MMA, double-buffered accumulator)
    int small_kb = 128;               // persistent small-modulus kernel: bytes of K per stage (128: SWIZZLE_128B rows, 64: SWIZZLE_64B)
    int small_persistent = 1;         // small-modulus stage 2: persistent kernel with two TMEM accumulators (0: one tile per CTA)
    int fuse_ext = 0;                 // small-modulus path: 1 = base extension fused into the entry-per-thread normalisation kernel (measured slower: occupancy)
    int norm32 = 1;                   // entry-per-thread kernel: 32-bit Barrett products where every modulus has the same bit length <= 27
    int vec_config = 0;               // tile configuration of the mp_gemv / mp_dot kernels (A/B measurement)
    HostConsts hc;
    SmallConsts sc;                   // small-modulus base of the tensor-core stage 2
    void *d_small[16] = {nullptr};     // device copies of its tables (SmallDev first)
    DevConsts *dconsts = nullptr;
    int *d_pow2 = nullptr, *d_inv_pow2 = nullptr, *d_mrc = nullptr, *d_prefix = nullptr, *d_ext_w = nullptr, *d_ext_t = nullptr, *d_wpow2 = nullptr, *d_spow2 = nullptr;
    std::atomic<long> launches{0};
    // workspace pool (grown on demand, never freed per call)
    void *ws[24] = {nullptr};            // 12..17: device operands and staging rings of mpres_gemm_host; 18..23: mpres_ops.cu
    size_t ws_size[24] = {0};
    size_t ws_limit = 0;                 // mpres_set_workspace_limit: reservations beyond this many bytes in total fail like a device out of memory (0: no limit)
    long ws_fallbacks = 0;               // mp_gemm calls served in reference order because the fast path's workspaces did not fit
    cudaStream_t hs[4] = {nullptr};      // mpres_gemm_host: upload, unpack, compute, download
    cudaEvent_t hev[16] = {nullptr};
    bool host_ready = false;
    std::mutex host_mu;
    // device counters of the last call, kCounterBlock ints per column panel g (one panel unless the call is sharded):
    // [8g + 0] elements handed to the reference-order fallback, [8g + 1] elements listed for the residue-parallel normalisation,
    // [8g + 6] (min,+) pairs recomputed densely; block 0 also holds [2] n' of the reduced base and [4], [5] the small-base selection
    int *d_counter = nullptr;
    int *h_sel = nullptr;          // pinned: {n', -, one-byte moduli, input residues, error} read back after the base selection
    cudaStream_t last_stream = nullptr;
    // calls on one context share its workspaces and counters: a call on another stream than its predecessor's waits for it
    cudaEvent_t serial_ev = nullptr;
    cudaStream_t serial_stream = nullptr;
    bool serial_valid = false;
    // optional per-kernel timing (mpres_set_profiling): events recorded on the caller's stream after every kernel of the fast mp_gemm
    std::vector<cudaEvent_t> prof_ev;
    std::vector<const char *> prof_name;
    int prof_used = 0;
    int sm_count = 148;
    // optional per-stage timing (events recorded on the caller's stream)
    bool profiling = false;
    cudaEvent_t ev[6] = {nullptr};
    bool ev_valid = false;
    int last_stage2_launches = 0;
    void *lean_host = nullptr;           // pinned host slots of the lean uploads of mpres_gemm_host
    int last_host_lean = 0;              // residues per entry the last mpres_gemm_host uploaded of A / B (0: full records)
    bool last_fast_ok = false;           // the last fast mp_gemm ran on the one-byte base without the limb planes or the reference-order fallback
    int last_nin = 0;                    // ... reading this many residues per operand entry
    bool last_binary = false;            // the last fast mp_gemm rounded its exact sums in binary (kernels_bin.cuh: full-precision inputs)
    // opt-in shared-memory sizes (cudaFuncSetAttribute) are per device: remembered per context, not per process
    bool attr_bin = false, attr_ext = false, attr_fast = false, attr_align_mma = false, attr_align = false, attr_small = false, attr_umma = false;
    unsigned long long attr_norm = 0;    // bit NQ / 8: k_norm_fast<NQ, *, true> (staged residues of S)
    int norm_staged = 1;                 // k_norm_fast stages the residues of S in shared memory (mpres_set_stage3_kernel(4) = off)
    unsigned long long attr_fused = 0;   // bit NQ / 8: k_ext_norm_small<NQ, *>
    unsigned long long attr_bin2 = 0;    // bit NQ / 8: k_bin_norm2<NQ, ...>
    std::mutex mu;
};

constexpr int kMaxPanels = 16;            // column panels of one fast-path call (ranks of a sharded call)
constexpr int kCounterBlock = 8;
constexpr int kCounterInts = kCounterBlock * kMaxPanels;        // (+ kCounterExtra ints behind them: slices, slice width of the last call)
constexpr int kCounterExtra = 8;

// internal entry points shared between the translation units (not part of the C-ABI)
extern "C" int mpres_internal_maxabs(mpres_ctx *c, long long n, const mpres::SoA *x, int incx, const mpres::SoA *r, cudaStream_t st);
extern "C" void mpres_ops_release(mpres_ctx *c);

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int) e_; } while (0)

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int ws_reserve(mpres_ctx *c, int slot, size_t bytes, void **out) {
    if (c->ws_size[slot] < bytes) {
        if (c->ws[slot]) { cudaDeviceSynchronize(); cudaFree(c->ws[slot]); c->ws[slot] = nullptr; c->ws_size[slot] = 0; }
        size_t want = bytes + bytes / 8 + 256;
        if (c->ws_limit) {
            size_t held = 0;
            for (int i = 0; i < 24; ++i) held += c->ws_size[i];
            if (held + want > c->ws_limit) return (int) cudaErrorMemoryAllocation;
        }
        const cudaError_t e = cudaMalloc(&c->ws[slot], want);
        if (e != cudaSuccess) { c->ws[slot] = nullptr; cudaGetLastError(); return (int) e; }     // (the sticky last-error must not leak into the next call)
        c->ws_size[slot] = want;
    }
    *out = c->ws[slot];
    return 0;
}

SoA view(const mpres_array_t *a) {
    SoA v;
    v.digits = a->digits; v.sign = a->sign; v.exp = a->exp; v.eval = (Er *) a->eval; v.len_ptr = a->len; v.len_val = 0;
    return v;
}
SoA view(const mpres_collection_t *a, size_t len) {
    SoA v;
    v.digits = a->digits; v.sign = a->sign; v.exp = a->exp; v.eval = (Er *) a->eval; v.len_ptr = nullptr; v.len_val = (long long) len;
    return v;
}
// carve an SoA of `len` elements out of one workspace slot
int ws_soa(mpres_ctx *c, int slot, size_t len, SoA *out) {
    const size_t N = c->hc.N;
    size_t bytes = len * (4 * N + 4 + 4 + 32) + 64;
    void *p;
    int rc = ws_reserve(c, slot, bytes, &p);
    if (rc) return rc;
    char *b = (char *) p;
    out->eval = (Er *) b; b += len * 32;
    out->digits = (int *) b; b += len * 4 * N;
    out->sign = (int *) b; b += len * 4;
    out->exp = (int *) b;
    out->len_ptr = nullptr;
    out->len_val = (long long) len;
    return 0;
}

// Calls on one context share its workspaces and device counters.  call_begin makes the stream of this call wait for the previous
// call when that ran on another stream; call_end records the point the next call has to wait for.
inline int call_begin(mpres_ctx *c, cudaStream_t st) {
    if (c->serial_valid && c->serial_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, c->serial_ev, 0));
    return 0;
}
inline int call_end(mpres_ctx *c, cudaStream_t st) {
    if (!c->serial_ev) CUDA_TRY(cudaEventCreateWithFlags(&c->serial_ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(c->serial_ev, st));
    c->serial_stream = st;
    c->serial_valid = true;
    return 0;
}
// per-kernel timing marks (only while profiling is on)
inline void prof_reset(mpres_ctx *c, cudaStream_t st) {
    c->prof_used = 0;
    if (!c->profiling) return;
    if (c->prof_ev.empty()) { c->prof_ev.resize(64, nullptr); c->prof_name.resize(64, nullptr); }
    if (!c->prof_ev[0]) cudaEventCreate(&c->prof_ev[0]);
    cudaEventRecord(c->prof_ev[0], st);
    c->prof_name[0] = "begin";
    c->prof_used = 1;
}
inline void prof_mark(mpres_ctx *c, cudaStream_t st, const char *name) {
    if (!c->profiling || c->prof_used <= 0 || c->prof_used >= (int) c->prof_ev.size()) return;
    const int i = c->prof_used++;
    if (!c->prof_ev[i]) cudaEventCreate(&c->prof_ev[i]);
    cudaEventRecord(c->prof_ev[i], st);
    c->prof_name[i] = name;
}

inline int group_size(int N) { return N <= 8 ? 8 : N <= 16 ? 16 : 32; }

}  // namespace

// Dispatch on the (lanes per number, residues per lane) shape chosen from N.
#define MPRES_DISPATCH(N_, ...)                                    \
    do {                                                           \
        if ((N_) <= 8) { constexpr int G = 8, R = 1; __VA_ARGS__; }        \
        else if ((N_) <= 16) { constexpr int G = 16, R = 1; __VA_ARGS__; } \
        else if ((N_) <= 32) { constexpr int G = 32, R = 1; __VA_ARGS__; } \
        else if ((N_) <= 64) { constexpr int G = 32, R = 2; __VA_ARGS__; } \
        else { constexpr int G = 32, R = 4; __VA_ARGS__; }                 \
    } while (0)

#define LAUNCHED(c) ((c)->launches.fetch_add(1, std::memory_order_relaxed))

