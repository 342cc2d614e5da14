// gemm_fast.cuh -- host side of the exact-window fast path of mp_gemm: workspace carving, the device-side choice of the base, the three
// stages as kernel sequences, and their sharded form (rows of A / C split over the ranks of one NVLink domain, the B-side stage-1 results
// of every rank's column block delivered to all ranks by the copy engines while the tensor kernel already multiplies).
//
// A call is a sequence over COLUMN PANELS of B / C.  A single-GPU call has one panel.  In a sharded call (shard.cuh) rank r converts
// panel r (its n / G columns of B: window, alignment into the one-byte planes, shift plane and its candidate lists) into one
// contiguous "package", copies the package into the same position of every peer's receive buffer (cudaMemcpyAsync over NVLink, no SM
// involved) and raises that peer's arrival flag; the persistent tensor kernel walks the panels in ring order and waits inside the
// kernel for a panel's flag before its first TMA load of it.
#pragma once

static inline long long round_up(long long v, long long a) { return (v + a - 1) / a * a; }

namespace mpres {

__global__ void k_set_flag(unsigned *flag, unsigned epoch) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}
__global__ void k_wait_flag(const unsigned *flag, unsigned epoch) { wait_arrival(flag, epoch); }

}  // namespace mpres

// what a sharded call needs from its communicator (shard.cuh fills it); world == 1: a plain call
struct FastShard {
    int rank = 0, world = 1;
    unsigned epoch = 0;
    char *recv = nullptr;                  // local receive buffer: `world` packages, pkg_stride bytes apart
    size_t pkg_stride = 0;
    char *peer_recv[kMaxPanels] = {};      // the receive buffers of all ranks (peer-mapped; [rank] == recv)
    int *peer_xchg[kMaxPanels] = {};       // the exchange arrays of all ranks
    unsigned *flags = nullptr;             // local arrival flags [world]
    unsigned *peer_flags[kMaxPanels] = {};
    cudaStream_t push[kMaxPanels] = {};    // copy streams
    int npush = 0;
    cudaEvent_t ev_pkg = nullptr, ev_push[kMaxPanels] = {};
};

// B-side stage-1 results of one column panel
struct PanelPkg {
    OuterInfo *IB;
    int *thrB, *cposB, *cvalB;
    int16_t *SB, *SBT;
    uint8_t *QB;
};
static inline size_t pad1k(size_t v) { return (v + 1023) & ~(size_t) 1023; }
static inline size_t pkg_header_bytes(long long nb_p, long long k_p) {
    return pad1k((size_t) nb_p * sizeof(OuterInfo)) + pad1k((size_t) nb_p * 4) + 2 * pad1k((size_t) nb_p * kMcT * 4) + 2 * pad1k((size_t) nb_p * k_p * 2);
}
static inline PanelPkg pkg_carve(char *header, char *planes, long long nb_p, long long k_p) {
    PanelPkg p;
    char *q = header;
    p.IB = (OuterInfo *) q; q += pad1k((size_t) nb_p * sizeof(OuterInfo));
    p.thrB = (int *) q; q += pad1k((size_t) nb_p * 4);
    p.cposB = (int *) q; q += pad1k((size_t) nb_p * kMcT * 4);
    p.cvalB = (int *) q; q += pad1k((size_t) nb_p * kMcT * 4);
    p.SB = (int16_t *) q; q += pad1k((size_t) nb_p * k_p * 2);
    p.SBT = (int16_t *) q;
    p.QB = (uint8_t *) planes;
    return p;
}

namespace {

SoA soa_shift(const SoA &a, long long off, int N) {
    SoA v = a;
    v.digits += off * N; v.sign += off; v.exp += off; v.eval += off;      // len (the offset of the upper bounds) is unchanged
    return v;
}

}  // namespace

// ---- the limb-plane path (the format's own moduli as four byte limbs): chosen when the sums do not fit the one-byte base -----------------
// IA / IB: windows of all m rows / n columns; nprime (device) = moduli stage 2 runs on.
inline int gemm_fast_limb(mpres_ctx *c, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb, SoA alpha, SoA beta, SoA Cm, int ldc,
                          cudaStream_t st, OuterInfo *IA_in) {
    const int N = c->hc.N;
    const long long m_p = round_up(m, kBM), n_p = round_up(n, kBN), k_p = round_up(k, 128);
    const size_t bytesPA = (size_t) N * 4 * m_p * k_p, bytesPB = (size_t) N * 4 * n_p * k_p;
    const size_t bytesS = (size_t) N * n_p * m_p * 4;
    const size_t bytesSA = (size_t) m_p * k_p * 2, bytesSB = (size_t) n_p * k_p * 2, bytesD = (size_t) n_p * m_p * 2;
    const size_t bytesTab = (size_t) (3 * c->hc.log2M + 2) * N * sizeof(int);
    const size_t bytesTodo = (size_t) m * n * sizeof(long long);
    void *pPA, *pPB, *pS, *pMisc;
    int rc;
    if ((rc = ws_reserve(c, 3, bytesPA, &pPA))) return rc;
    if ((rc = ws_reserve(c, 4, bytesPB, &pPB))) return rc;
    if ((rc = ws_reserve(c, 5, bytesS, &pS))) return rc;
    // (slot 6 holds IA: the caller carved it there; this path's own scratch lives in slots 11 and 9)
    if ((rc = ws_reserve(c, 11, bytesSA + bytesSB + bytesD + pad1k((size_t) n_p * sizeof(OuterInfo)) + 2 * bytesTodo + bytesTab + 4096, &pMisc))) return rc;
    char *pm = (char *) pMisc;
    int16_t *SA = (int16_t *) pm; pm += bytesSA;
    int16_t *SB = (int16_t *) pm; pm += bytesSB;
    int16_t *D = (int16_t *) pm; pm += (bytesD + 1023) / 1024 * 1024;
    OuterInfo *IB = (OuterInfo *) pm; pm += pad1k((size_t) n_p * sizeof(OuterInfo));
    long long *todo = (long long *) pm; pm += bytesTodo;
    long long *slow = (long long *) pm; pm += bytesTodo;
    pm = (char *) (((uintptr_t) pm + 15) & ~(uintptr_t) 15);
    int *scal_tab = (int *) pm;
    OuterInfo *IA = IA_in;
    int *nprime = c->d_counter + 2;
    const long long soA = ta ? lda : 1, slA = ta ? 1 : lda;
    const long long soB = tb ? 1 : ldb, slB = tb ? ldb : 1;
    int launches = 0;
    // the windows of all columns (the caller may have looked at one column block only)
    k_outer_info<<<(unsigned) ((n * 32ll + 255) / 256), 256, 0, st>>>(c->dconsts, B, soB, slB, n, k, IB); ++launches;
    const size_t smem_align = (size_t) N * 4 * (kRun + 4);
    if (!c->attr_fast) {
        cudaFuncSetAttribute(k_align_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * (kRun + 4));
        cudaFuncSetAttribute(k_align_planes4, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 4 * (kRun + 4));
        cudaFuncSetAttribute(k_base_extend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) base_extend_smem(128));
        cudaFuncSetAttribute(k_limb_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
        cudaFuncSetAttribute(k_limb_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
        c->attr_fast = true;
    }
    if (N % 4 == 0 && N <= 128 && c->stage1 == 0) {
        k_align_planes4<<<dim3((unsigned) m_p, (unsigned) std::min<long long>(k_p / kRun, 4)), 256, smem_align, st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pPA, SA, m_p, k_p, nprime);
        k_align_planes4<<<dim3((unsigned) n_p, (unsigned) std::min<long long>(k_p / kRun, 4)), 256, smem_align, st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pPB, SB, n_p, k_p, nprime);
    } else {
        k_align_planes<<<dim3((unsigned) m_p, (unsigned) (k_p / kRun)), 256, smem_align, st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pPA, SA, m_p, k_p);
        k_align_planes<<<dim3((unsigned) n_p, (unsigned) (k_p / kRun)), 256, smem_align, st>>>(c->dconsts, B, soB, slB, n, k, IB, (uint8_t *) pPB, SB, n_p, k_p);
    }
    launches += 2;
    prof_mark(c, st, "k_align_planes");
    if (c->minplus_sparse && k_p >= 512) {
        // (min,+) from candidate lists (kernels_minplus.cuh)
        const long long m_t = round_up(m_p, 64), n_t = round_up(n_p, 64);
        const size_t bT = (size_t) k_p * (m_t + n_t) * 2, bD = (size_t) m_p * n_p * 2 * 2, bC = (size_t) (m_p + n_p) * kMcT * 8 + (size_t) (m_p + n_p) * 4;
        void *pMp;
        if ((rc = ws_reserve(c, 9, bT + bD + bC + (size_t) m * n * 8 + 256, &pMp))) return rc;
        char *q = (char *) pMp;
        int16_t *SAT = (int16_t *) q; q += (size_t) k_p * m_t * 2;
        int16_t *SBT = (int16_t *) q; q += (size_t) k_p * n_t * 2;
        int16_t *D1 = (int16_t *) q; q += (size_t) m_p * n_p * 2;
        int16_t *D2 = (int16_t *) q; q += (size_t) m_p * n_p * 2;
        int *cposA = (int *) q; q += (size_t) m_p * kMcT * 4;
        int *cvalA = (int *) q; q += (size_t) m_p * kMcT * 4;
        int *cposB = (int *) q; q += (size_t) n_p * kMcT * 4;
        int *cvalB = (int *) q; q += (size_t) n_p * kMcT * 4;
        int *thrA = (int *) q; q += (size_t) m_p * 4;
        int *thrB = (int *) q; q += (size_t) n_p * 4;
        q = (char *) (((uintptr_t) q + 15) & ~(uintptr_t) 15);
        long long *mplist = (long long *) q;
        k_mp_select<<<(unsigned) m, 256, 0, st>>>(SA, k_p, (int) k_p, m, cposA, cvalA, thrA);
        k_mp_select<<<(unsigned) n, 256, 0, st>>>(SB, k_p, (int) k_p, n, cposB, cvalB, thrB);
        k_mp_transpose<<<dim3((unsigned) (m_p / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SA, k_p, SAT, m_t);
        k_mp_transpose<<<dim3((unsigned) (n_p / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SB, k_p, SBT, n_t);
        k_mp_gather<<<(unsigned) m, 256, 0, st>>>(cposA, cvalA, SBT, n_t, m, (int) n_p, D1, n_p);
        k_mp_gather<<<(unsigned) n, 256, 0, st>>>(cposB, cvalB, SAT, m_t, n, (int) m_p, D2, m_p);
        k_mp_combine<<<dim3((unsigned) (m_p / 64), (unsigned) (n_p / 64)), 256, 0, st>>>(D1, n_p, D2, m_p, thrA, thrB, m, n, D, m_p, mplist, c->d_counter + 6);
        k_mp_fix<<<c->sm_count * 4, 256, 0, st>>>(SA, SB, k_p, D, m_p, mplist, c->d_counter + 6);
        launches += 8;
    } else {
        k_minplus<<<dim3((unsigned) (m_p / kMpTI), (unsigned) (n_p / kMpTJ)), 256, 0, st>>>(SA, SB, D, k_p, m_p, n_p); ++launches;
    }
    prof_mark(c, st, "k_minplus");
    if (c->profiling) { if (!c->ev[1]) cudaEventCreate(&c->ev[1]); cudaEventRecord(c->ev[1], st); }
    dim3 grid((unsigned) (n_p / kBN), (unsigned) (m_p / kBM), (unsigned) N);
    int gemm_launches = 0;
    for (long long kb = 0; kb < k_p; kb += 8064) {
        const int kl = (int) std::min<long long>(8064, k_p - kb);
        if (c->stage2 == MPRES_STAGE2_MMA_SYNC) {
            k_limb_gemm<0><<<grid, 256, kGemmSmem, st>>>(c->dconsts, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl, kb > 0, nprime);
            k_limb_gemm<1><<<grid, 256, kGemmSmem, st>>>(c->dconsts, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl, true, nprime);
            gemm_launches += 2;
        } else {
            if ((rc = launch_limb_umma(c, c->stage2 != MPRES_STAGE2_UMMA_UNSTACKED, (const uint8_t *) pPA, (const uint8_t *) pPB, (int *) pS, m_p, n_p, k_p, kb, kl,
                                       kb > 0, st))) return rc;
            gemm_launches += 1;
        }
    }
    prof_mark(c, st, "k_limb_umma");
    if (c->profiling) { if (!c->ev[2]) cudaEventCreate(&c->ev[2]); cudaEventRecord(c->ev[2], st); }
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    bool have_fast = c->stage3 == 0;
    switch (N) { case 8: case 16: case 24: case 32: case 40: case 48: case 56: case 64: break; default: have_fast = false; }
    const bool f32 = c->sc.usable && c->sc.red_shift >= 24 && c->sc.red_shift <= 27 && c->norm32;
    if (c->reduced_base && N % 4 == 0) {
        const dim3 gx((unsigned) ((m + kExtThreads - 1) / kExtThreads), (unsigned) std::min(n, 256));
        k_base_extend<<<gx, kExtThreads, base_extend_smem(N), st>>>(c->dconsts, m, n, (int *) pS, m_p, n_p, c->d_counter + 2);
        ++launches;
        prof_mark(c, st, "k_base_extend");
    }
    auto norm_fast = [&](auto tag) {
        constexpr int NQ = decltype(tag)::value;
        const unsigned g3 = (unsigned) ((long long) ((m + kNormFastThreads - 1) / kNormFastThreads) * n);
        const int rowsT = 3 * c->hc.log2M + 2;
        k_scalar_tables<<<(rowsT * NQ + 255) / 256, 256, 0, st>>>(c->dconsts, alpha, beta, scal_tab);
        const size_t sm_cds = (size_t) kNormFastThreads * (NQ + 1) * sizeof(int);
        auto launch_norm = [&](auto kern, size_t smem) {
            kern<<<g3, kNormFastThreads, smem, st>>>(c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc,
                                                     scal_tab, todo, c->d_counter, slow, c->d_counter + 1, allow_fb, nullptr);
        };
        if (c->norm_staged) {
            const size_t sm_st = sm_cds + (size_t) NQ * kNormFastThreads * sizeof(int);
            if (sm_st > 48 * 1024 && !(c->attr_norm >> (NQ / 8) & 1ull)) {
                cudaFuncSetAttribute(k_norm_fast<NQ, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                cudaFuncSetAttribute(k_norm_fast<NQ, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                c->attr_norm |= 1ull << (NQ / 8);
            }
            if (f32) launch_norm(k_norm_fast<NQ, true, true>, sm_st); else launch_norm(k_norm_fast<NQ, false, true>, sm_st);
        } else {
            if (f32) launch_norm(k_norm_fast<NQ, true, false>, sm_cds); else launch_norm(k_norm_fast<NQ, false, false>, sm_cds);
        }
        launches += 2;
    };
    if (have_fast) {
        switch (N) {
            case 8: norm_fast(std::integral_constant<int, 8>{}); break;
            case 16: norm_fast(std::integral_constant<int, 16>{}); break;
            case 24: norm_fast(std::integral_constant<int, 24>{}); break;
            case 32: norm_fast(std::integral_constant<int, 32>{}); break;
            case 40: norm_fast(std::integral_constant<int, 40>{}); break;
            case 48: norm_fast(std::integral_constant<int, 48>{}); break;
            case 56: norm_fast(std::integral_constant<int, 56>{}); break;
            case 64: norm_fast(std::integral_constant<int, 64>{}); break;
            default: have_fast = false;
        }
        prof_mark(c, st, "k_norm_fast");
    }
    MPRES_DISPATCH(N, {
        if (have_fast) {
            k_norm_list<G, R><<<c->sm_count * 8, 256, 0, st>>>(c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc,
                                                              slow, c->d_counter + 1);
        } else {
            constexpr int kNormTile = 256 / G;
            const unsigned g3 = (unsigned) ((long long) ((m + kNormTile - 1) / kNormTile) * n);
            k_normalize_epilogue<G, R><<<g3, 256, (size_t) N * (kNormTile + 1) * 4, st>>>(
                c->dconsts, m, n, k, (const int *) pS, D, m_p, n_p, IA, IB, alpha, beta, Cm, ldc, todo, c->d_counter, allow_fb);
        }
        ++launches;
        if (allow_fb) {
            k_gemm_todo<G, R><<<c->sm_count * 8, 128, 0, st>>>(c->dconsts, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, todo, c->d_counter);
            ++launches;
        }
    });
    prof_mark(c, st, "k_norm_list+k_gemm_todo");
    c->last_stage2_launches = gemm_launches;
    for (int i = 0; i < launches + gemm_launches; ++i) LAUNCHED(c);
    return 0;
}

// ---- sharded call, flat layout -------------------------------------------------------------------------------------------------------
// When the ranks' column blocks are whole 256-column tiles (nb % 256 == 0) the B-side stage-1 results of ALL blocks live in ONE set of
// arrays in every rank's receive buffer -- windows [n], candidate lists [n][64], shift plane [n][k_p], one-byte planes [P][n][k_p] -- and a
// rank copies its block of each array into the same position of every peer's buffer (copy engines over NVLink; the planes as one 2-D
// copy).  Everything downstream is then the single-GPU code on column SEGMENTS: the panels are multiplied and normalised in groups (ring
// order from the rank's own block), so the normalisation of the first group runs while the blocks of the next are still arriving, and
// every stage-3 kernel runs once per segment instead of once per panel.
struct FlatLayout { size_t oIB, oThr, oCpos, oCval, oSB, oQB, total; };
static inline FlatLayout flat_layout(long long n, long long k_p, int planes) {
    FlatLayout f;
    size_t o = 0;
    f.oIB = o; o += pad1k((size_t) n * sizeof(OuterInfo));
    f.oThr = o; o += pad1k((size_t) n * 4);
    f.oCpos = o; o += pad1k((size_t) n * kMcT * 4);
    f.oCval = o; o += pad1k((size_t) n * kMcT * 4);
    f.oSB = o; o += pad1k((size_t) n * k_p * 2);
    f.oQB = o; o += (size_t) planes * n * k_p;
    f.total = o;
    return f;
}

inline int gemm_fast_flat(mpres_ctx *c, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb,
                          SoA alpha, SoA beta, SoA Cm, int ldc, cudaStream_t st, bool *done, FastShard *sh) {
    *done = false;
    const int N = c->hc.N;
    const int W = sh->world, rank = sh->rank;
    const int nb = n / W;
    const long long m_p = round_up(m, kBM), m_ps = round_up(m, kSN), k_p = round_up(k, 128);
    if (k_p > 32000 * 128ll) return 0;
    const bool small_on = c->stage2 == MPRES_STAGE2_SMALL && c->sc.usable;
    bool sparse_mp = c->minplus_sparse && k_p >= 512;
    int rc;
    const FlatLayout fl = flat_layout(n, k_p, kSmallMax);
    if (fl.total > (size_t) W * sh->pkg_stride) return -7;
    char *rb = sh->recv;
    OuterInfo *IB = (OuterInfo *) (rb + fl.oIB);
    int *thrB = (int *) (rb + fl.oThr), *cposB = (int *) (rb + fl.oCpos), *cvalB = (int *) (rb + fl.oCval);
    int16_t *SB = (int16_t *) (rb + fl.oSB);
    uint8_t *QB = (uint8_t *) (rb + fl.oQB);
    // ---- workspace that does not depend on the base ----
    const size_t bytesIA = pad1k((size_t) m_ps * sizeof(OuterInfo)), bytesSA = pad1k((size_t) m_ps * k_p * 2);
    const size_t bytesPart = pad1k((size_t) std::max<long long>(m_ps, nb) * sizeof(OuterPart));
    const size_t bytesTab = pad1k((size_t) (3 * c->hc.log2M + 2) * N * sizeof(int));
    const size_t bytesCandA = 2 * pad1k((size_t) m_ps * kMcT * 4) + pad1k((size_t) m_ps * 4);
    const size_t bytesList = pad1k((size_t) m * n * sizeof(long long));
    const size_t bytesD = pad1k((size_t) n * m_p * 2);
    void *pMisc;
    if ((rc = ws_reserve(c, 6, bytesIA + bytesPart + bytesSA + bytesTab + bytesCandA + 3 * bytesList + 3 * bytesD + 4096, &pMisc))) return rc;
    char *pm = (char *) pMisc;
    OuterInfo *IA = (OuterInfo *) pm; pm += bytesIA;
    OuterPart *part = (OuterPart *) pm; pm += bytesPart;
    int16_t *SA = (int16_t *) pm; pm += bytesSA;
    int *scal_tab = (int *) pm; pm += bytesTab;
    int *cposA = (int *) pm; pm += pad1k((size_t) m_ps * kMcT * 4);
    int *cvalA = (int *) pm; pm += pad1k((size_t) m_ps * kMcT * 4);
    int *thrA = (int *) pm; pm += pad1k((size_t) m_ps * 4);
    long long *todo = (long long *) pm; pm += bytesList;
    long long *slow = (long long *) pm; pm += bytesList;
    long long *mplist = (long long *) pm; pm += bytesList;
    int16_t *D = (int16_t *) pm; pm += bytesD;
    int16_t *D1 = (int16_t *) pm; pm += bytesD;
    int16_t *D2 = (int16_t *) pm;
    int *nprime = c->d_counter + 2, *sel = c->d_counter + 4;
    const long long soA = ta ? lda : 1, slA = ta ? 1 : lda;
    const long long soB = tb ? 1 : ldb, slB = tb ? ldb : 1;
    auto mark = [&](int i) { if (c->profiling) { if (!c->ev[i]) cudaEventCreate(&c->ev[i]); cudaEventRecord(c->ev[i], st); } };
    prof_reset(c, st);
    mark(0);
    int launches = 0;
    const long long col_own = (long long) rank * nb;
    const SoA Bown = soa_shift(B, col_own * soB, N);
    auto outer_info = [&](const SoA &X, long long so, long long sl, int outer, int inner, OuterInfo *info) {
        if (so == 1 && sl != 1 && inner >= 64) {
            const int lineb = (outer + 31) / 32;
            int chunks = std::max(1, std::min((c->sm_count * 8 + lineb - 1) / lineb, inner / 32));
            const int chunk = (inner + chunks - 1) / chunks;
            chunks = (inner + chunk - 1) / chunk;
            k_outer_part_init<<<(outer + 255) / 256, 256, 0, st>>>(part, outer);
            k_outer_part<<<dim3((unsigned) lineb, (unsigned) chunks), 256, 0, st>>>(X, sl, outer, inner, chunk, part);
            k_outer_part_final<<<(outer + 255) / 256, 256, 0, st>>>(c->dconsts, part, outer, info);
            launches += 2;
        } else if (sl == 1 && outer < c->sm_count * 16 && inner >= 2048) {
            // few long lines: several blocks per line
            const int chunks = std::max(1, std::min((c->sm_count * 16 + outer - 1) / outer, inner / 1024));
            const int chunk = (inner + chunks - 1) / chunks;
            k_outer_part_init<<<(outer + 255) / 256, 256, 0, st>>>(part, outer);
            k_outer_part_l<<<dim3((unsigned) outer, (unsigned) ((inner + chunk - 1) / chunk)), 256, 0, st>>>(X, so, outer, inner, chunk, part);
            k_outer_part_final<<<(outer + 255) / 256, 256, 0, st>>>(c->dconsts, part, outer, info);
            launches += 2;
        } else {
            k_outer_info<<<(unsigned) ((outer * 32ll + 255) / 256), 256, 0, st>>>(c->dconsts, X, so, sl, outer, inner, info);
        }
        ++launches;
    };
    // the previous call's copies have left this rank's block of the arrays (safeguard: the rendezvous below already implies it)
    for (int si = 0; si < sh->npush && si < W - 1; ++si) CUDA_TRY(cudaStreamWaitEvent(st, sh->ev_push[si], 0));
    outer_info(A, soA, slA, m, k, IA);
    outer_info(Bown, soB, slB, nb, k, IB + col_own);
    prof_mark(c, st, "k_outer_info");
    Xchg x;
    memset(&x, 0, sizeof(x));
    x.world = W; x.rank = rank;
    x.epoch = sh->epoch; x.parity = (int) (sh->epoch & 1u);
    for (int p = 0; p < W; ++p) x.peer[p] = sh->peer_xchg[p];
    x.err = c->d_counter + 3;
    k_choose_base<<<1, 256, 0, st>>>(c->dconsts, IA, m, IB + col_own, nb, k, c->reduced_base, small_on ? 1 : 0, nprime, sel, x);
    ++launches;
    CUDA_TRY(cudaMemcpyAsync(c->h_sel, c->d_counter + 2, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int P = c->h_sel[2];
    if (c->h_sel[1]) return -50;                                  // a rank did not reach the call
    const bool binary = P > 0 && c->h_sel[5] > c->hc.log2M - 2;
    c->last_binary = binary;
    c->last_fast_ok = P > 0;
    c->last_nin = c->h_sel[3];
    prof_mark(c, st, "k_choose_base");
    if (P <= 0) {
        rc = gemm_fast_limb(c, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, st, IA);
        if (rc) return rc;
        mark(3);
        c->ev_valid = c->profiling;
        for (int i = 0; i < launches; ++i) LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        *done = true;
        return 0;
    }
    if (binary) sparse_mp = false;
    // ---- workspace that depends on the base ----
    void *pQA, *pS8, *pS = nullptr, *pT = nullptr;
    if ((rc = ws_reserve(c, 8, (size_t) P * m_ps * k_p, &pQA))) return rc;
    if ((rc = ws_reserve(c, 10, (size_t) P * n * m_ps, &pS8))) return rc;
    if (!binary && (rc = ws_reserve(c, 5, (size_t) N * n * m_p * 4, &pS))) return rc;
    if (sparse_mp && (rc = ws_reserve(c, 11, (size_t) k_p * (m_ps + n) * 2 + 512, &pT))) return rc;
    int16_t *SAT = (int16_t *) pT, *SBT = sparse_mp ? SAT + (size_t) k_p * m_ps : nullptr;
    if (!c->attr_ext) {
        cudaFuncSetAttribute(k_ext_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        cudaFuncSetAttribute(k_ext_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        c->attr_ext = true;
    }
    if (!c->attr_align) { cudaFuncSetAttribute(k_align_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(false)); c->attr_align = true; }
    // ---- stage 1: the rank's column block of B first (it has to travel), then its rows of A ----
    {
        const unsigned gB = (unsigned) std::min<long long>((nb / kASo) * (k_p / kASl), (long long) c->sm_count * align_small_blocks(false, false));
        const unsigned gA = (unsigned) std::min<long long>((m_ps / kASo) * (k_p / kASl), (long long) c->sm_count * align_small_blocks(false, false));
        k_align_small<false><<<gB, 256, align_small_smem(false), st>>>(c->dconsts, Bown, soB, slB, nb, k, IB + col_own, QB + col_own * k_p, SB + col_own * k_p,
                                                                       nb, k_p, sel, n);
        ++launches;
        prof_mark(c, st, "k_align_small(B)");
        if (sparse_mp) {
            k_mp_select<<<(unsigned) nb, 256, 0, st>>>(SB + col_own * k_p, k_p, (int) k_p, nb, cposB + col_own * kMcT, cvalB + col_own * kMcT, thrB + col_own);
            ++launches;
            prof_mark(c, st, "k_mp_select(B)");
        }
        CUDA_TRY(cudaEventRecord(sh->ev_pkg, st));
        k_align_small<false><<<gA, 256, align_small_smem(false), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
        ++launches;
        prof_mark(c, st, "k_align_small(A)");
        if (sparse_mp) {
            k_mp_select<<<(unsigned) m, 256, 0, st>>>(SA, k_p, (int) k_p, m, cposA, cvalA, thrA);
            k_mp_transpose<<<dim3((unsigned) (m_ps / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SA, k_p, SAT, m_ps);
            launches += 2;
            prof_mark(c, st, "k_mp_select+transpose(A)");
        }
        // the block's share of every array to every peer, peers in the order they will need it (rank - 1 multiplies it second)
        for (int d = 1; d < W; ++d) {
            const int peer = (rank - d + W) % W;
            const int si = (d - 1) % sh->npush;
            cudaStream_t ps = sh->push[si];
            if (d - 1 < sh->npush) CUDA_TRY(cudaStreamWaitEvent(ps, sh->ev_pkg, 0));
            char *dst = sh->peer_recv[peer];
            auto push = [&](size_t off, size_t per_col) -> cudaError_t {
                const size_t o = off + (size_t) col_own * per_col;
                return cudaMemcpyAsync(dst + o, rb + o, (size_t) nb * per_col, cudaMemcpyDeviceToDevice, ps);
            };
            CUDA_TRY(push(fl.oIB, sizeof(OuterInfo)));
            if (sparse_mp) { CUDA_TRY(push(fl.oThr, 4)); CUDA_TRY(push(fl.oCpos, kMcT * 4)); CUDA_TRY(push(fl.oCval, kMcT * 4)); }
            if (!binary) CUDA_TRY(push(fl.oSB, (size_t) k_p * 2));
            {
                const size_t o = fl.oQB + (size_t) col_own * k_p;
                if ((size_t) n * k_p < ((size_t) 1 << 31)) {
                    CUDA_TRY(cudaMemcpy2DAsync(dst + o, (size_t) n * k_p, rb + o, (size_t) n * k_p, (size_t) nb * k_p, (size_t) P, cudaMemcpyDeviceToDevice, ps));
                } else {                                   // beyond the pitch limit of 2-D copies: plane by plane
                    for (int z = 0; z < P; ++z)
                        CUDA_TRY(cudaMemcpyAsync(dst + o + (size_t) z * n * k_p, rb + o + (size_t) z * n * k_p, (size_t) nb * k_p, cudaMemcpyDeviceToDevice, ps));
                }
            }
            k_set_flag<<<1, 1, 0, ps>>>(sh->peer_flags[peer] + rank, sh->epoch);
            ++launches;
        }
        for (int si = 0; si < sh->npush && si < W - 1; ++si) CUDA_TRY(cudaEventRecord(sh->ev_push[si], sh->push[si]));
    }
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    bool have_fast = c->stage3 == 0;
    switch (N) { case 8: case 16: case 24: case 32: case 40: case 48: case 56: case 64: break; default: have_fast = false; }
    const bool f32 = c->sc.usable && c->sc.red_shift >= 24 && c->sc.red_shift <= 27 && c->norm32;
    if (have_fast && !binary) {
        const int rowsT = 3 * c->hc.log2M + 2;
        k_scalar_tables<<<(rowsT * N + 255) / 256, 256, 0, st>>>(c->dconsts, alpha, beta, scal_tab);
        ++launches;
    }
    mark(1);
    // ---- groups of panels in ring order: multiply, then normalise the group's column segments ----
    int NG = std::min(W, 4);        // measured on 8 and 4 B200: 2.31 / 2.93 ms with four groups against 2.41 / 3.08 with two and 2.68 / 3.14 with one
    if (const char *env = getenv("MPRES_SHARD_GROUPS")) NG = std::max(1, std::min(atoi(env), W));
    int gemm_launches = 0, seg_index = 0;
    bool first_seg = true;
    for (int gi = 0; gi < NG; ++gi) {
        const int pb = (int) ((long long) W * gi / NG), pe = (int) ((long long) W * (gi + 1) / NG);
        if (pe <= pb) continue;
        SmallPanels pan;
        pan.count = pe - pb; pan.first = (rank + pb) % W; pan.own = rank; pan.epoch = sh->epoch; pan.flags = sh->flags;
        pan.s8_panel = (long long) nb * m_ps; pan.ring = W; pan.plane_rows = n; pan.slices = 1; pan.pair = -1;
        for (long long kb = 0; kb < k_p; kb += kSmallKChunk) {
            const int kl = (int) std::min<long long>(kSmallKChunk, k_p - kb);
            if ((rc = launch_small_umma(c, P, (const uint8_t *) pQA, QB, (long long) nb * k_p, (uint8_t *) pS8, m_ps, nb, k_p, kb, kl, kb > 0, sel, pan, st, n))) return rc;
            gemm_launches += 1;
        }
        if (gi == 0) { prof_mark(c, st, "k_small_umma_p"); mark(2); }
        // the group's panels as maximal runs of consecutive column blocks
        int pi = pb;
        while (pi < pe) {
            const int g0 = (rank + pi) % W;
            int cnt = 1;
            while (pi + cnt < pe && g0 + cnt < W) ++cnt;                   // (ring order wraps at W: a new segment starts at block 0)
            const long long col0 = (long long) g0 * nb;
            const int nc = cnt * nb;
            for (int q = 0; q < cnt; ++q)
                if (g0 + q != rank) { k_wait_flag<<<1, 1, 0, st>>>(sh->flags + g0 + q, sh->epoch); ++launches; }
            const SoA Cg = soa_shift(Cm, col0 * ldc, N);
            const SoA Bg = soa_shift(B, col0 * soB, N);
            int *cnt_s = c->d_counter + kCounterBlock * seg_index;
            long long *todo_s = todo + (size_t) col0 * m, *slow_s = slow + (size_t) col0 * m, *mpl_s = mplist + (size_t) col0 * m;
            int16_t *Ds = D + col0 * m_p;
            uint8_t *S8s = (uint8_t *) pS8 + col0 * m_ps;
            if (binary) {
                if ((rc = bin_norm_segment(c, first_seg, m, nc, S8s, m_ps, n, sel, IA, IB + col0, alpha, beta, Cg, ldc, st, &launches))) return rc;
                if (first_seg) prof_mark(c, st, "k_bin_norm");
            } else {
                int *Ss = (int *) pS + col0 * m_p;
                if (sparse_mp) {
                    int16_t *D1s = D1 + col0 * m_p, *D2s = D2 + col0 * m_p;
                    k_mp_transpose<<<dim3((unsigned) (nc / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SB + col0 * k_p, k_p, SBT + col0, n);
                    k_mp_gather<<<(unsigned) m, 256, 0, st>>>(cposA, cvalA, SBT + col0, n, m, nc, D1s, nc);
                    k_mp_gather<<<(unsigned) nc, 256, 0, st>>>(cposB + col0 * kMcT, cvalB + col0 * kMcT, SAT, m_ps, nc, (int) m_p, D2s, m_p);
                    k_mp_combine<<<dim3((unsigned) (m_p / 64), (unsigned) (nc / 64)), 256, 0, st>>>(D1s, nc, D2s, m_p, thrA, thrB + col0, m, nc, Ds, m_p, mpl_s, cnt_s + 6);
                    k_mp_fix<<<c->sm_count * 4, 256, 0, st>>>(SA, SB + col0 * k_p, k_p, Ds, m_p, mpl_s, cnt_s + 6);
                    launches += 5;
                } else {
                    k_minplus<<<dim3((unsigned) (m_p / kMpTI), (unsigned) (nc / kMpTJ)), 256, 0, st>>>(SA, SB + col0 * k_p, Ds, k_p, m_p, nc);
                    ++launches;
                }
                if (first_seg) prof_mark(c, st, "k_mp_transpose(B)+gather+combine+fix");
                {
                    const unsigned gx = (unsigned) std::min<long long>((m_p / kXT) * nc, (long long) c->sm_count * 4);
                    const size_t sm = ext_small_smem(c->sc.ext_cols, N);
                    if (c->sc.red_shift) k_ext_small<true><<<gx, kXT, sm, st>>>(c->dconsts, m, nc, S8s, m_p, m_ps, n, Ss, n, sel);
                    else k_ext_small<false><<<gx, kXT, sm, st>>>(c->dconsts, m, nc, S8s, m_p, m_ps, n, Ss, n, sel);
                    ++launches;
                    if (first_seg) prof_mark(c, st, "k_ext_small");
                }
                auto norm_fast = [&](auto tag) {
                    constexpr int NQ = decltype(tag)::value;
                    const unsigned g3 = (unsigned) ((long long) ((m + kNormFastThreads - 1) / kNormFastThreads) * nc);
                    const size_t sm_cds = (size_t) kNormFastThreads * (NQ + 1) * sizeof(int);
                    auto launch_norm = [&](auto kern, size_t smem) {
                        kern<<<g3, kNormFastThreads, smem, st>>>(c->dconsts, m, nc, k, (const int *) Ss, Ds, m_p, n, IA, IB + col0, alpha, beta, Cg, ldc,
                                                                 scal_tab, todo_s, cnt_s, slow_s, cnt_s + 1, allow_fb, nullptr);
                    };
                    if (c->norm_staged) {
                        const size_t sm_st = sm_cds + (size_t) NQ * kNormFastThreads * sizeof(int);
                        if (sm_st > 48 * 1024 && !(c->attr_norm >> (NQ / 8) & 1ull)) {
                            cudaFuncSetAttribute(k_norm_fast<NQ, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                            cudaFuncSetAttribute(k_norm_fast<NQ, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                            c->attr_norm |= 1ull << (NQ / 8);
                        }
                        if (f32) launch_norm(k_norm_fast<NQ, true, true>, sm_st); else launch_norm(k_norm_fast<NQ, false, true>, sm_st);
                    } else {
                        if (f32) launch_norm(k_norm_fast<NQ, true, false>, sm_cds); else launch_norm(k_norm_fast<NQ, false, false>, sm_cds);
                    }
                    ++launches;
                };
                bool hf = have_fast;
                if (hf) {
                    switch (N) {
                        case 8: norm_fast(std::integral_constant<int, 8>{}); break;
                        case 16: norm_fast(std::integral_constant<int, 16>{}); break;
                        case 24: norm_fast(std::integral_constant<int, 24>{}); break;
                        case 32: norm_fast(std::integral_constant<int, 32>{}); break;
                        case 40: norm_fast(std::integral_constant<int, 40>{}); break;
                        case 48: norm_fast(std::integral_constant<int, 48>{}); break;
                        case 56: norm_fast(std::integral_constant<int, 56>{}); break;
                        case 64: norm_fast(std::integral_constant<int, 64>{}); break;
                        default: hf = false;
                    }
                    if (first_seg) prof_mark(c, st, "k_norm_fast");
                }
                MPRES_DISPATCH(N, {
                    if (hf) {
                        k_norm_list<G, R><<<c->sm_count * 8, 256, 0, st>>>(c->dconsts, m, nc, k, (const int *) Ss, Ds, m_p, n, IA, IB + col0, alpha, beta, Cg, ldc,
                                                                          slow_s, cnt_s + 1);
                    } else {
                        constexpr int kNormTile = 256 / G;
                        const unsigned g3 = (unsigned) ((long long) ((m + kNormTile - 1) / kNormTile) * nc);
                        k_normalize_epilogue<G, R><<<g3, 256, (size_t) N * (kNormTile + 1) * 4, st>>>(
                            c->dconsts, m, nc, k, (const int *) Ss, Ds, m_p, n, IA, IB + col0, alpha, beta, Cg, ldc, todo_s, cnt_s, allow_fb);
                    }
                    ++launches;
                    if (allow_fb) {
                        k_gemm_todo<G, R><<<c->sm_count * 8, 128, 0, st>>>(c->dconsts, ta, tb, m, nc, k, A, lda, Bg, ldb, alpha, beta, Cg, ldc, todo_s, cnt_s);
                        ++launches;
                    }
                });
                if (first_seg) prof_mark(c, st, "k_norm_list+k_gemm_todo");
            }
            first_seg = false;
            ++seg_index;
            pi += cnt;
        }
    }
    mark(3);
    prof_mark(c, st, "end");
    c->ev_valid = c->profiling;
    c->last_stage2_launches = gemm_launches;
    for (int i = 0; i < launches + gemm_launches; ++i) LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return 0;
}

// ---- the fast path -----------------------------------------------------------------------------------------------------------------------
// *done tells the caller whether C is final (false: the call has to run in reference order).  sh != nullptr: sharded call -- m, A, Cm are
// this rank's row block, B is the rank's complete copy of B (the rank converts columns [rank n / world, (rank + 1) n / world) of it and
// receives the other panels' packages; the complete copy serves the limb-plane path and the reference-order fallback).
inline int gemm_fast_full(mpres_ctx *c, bool ta, bool tb, int m, int n, int k, SoA A, int lda, SoA B, int ldb,
                          SoA alpha, SoA beta, SoA Cm, int ldc, cudaStream_t st, bool *done, FastShard *sh = nullptr) {
    *done = false;
    const int N = c->hc.N;
    const int W = sh ? sh->world : 1, rank = sh ? sh->rank : 0;      // W column panels
    if (W > kMaxPanels || n % W != 0) return -6;
    const int nb = n / W;                                   // columns per panel
    if (sh && W > 1 && nb % 256 == 0 && c->stage2 == MPRES_STAGE2_SMALL && c->small_persistent && c->small_kb == 128 && c->small_tj == 256 &&
        !c->align_mma && !c->fuse_ext) {
        const char *env = getenv("MPRES_SHARD_FLAT");
        if (!env || atoi(env) != 0) return gemm_fast_flat(c, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, st, done, sh);
    }
    const long long m_p = round_up(m, kBM), m_ps = round_up(m, kSN), k_p = round_up(k, 128), nb_p = round_up(nb, 256);
    if (k_p > 32000 * 128ll) return 0;
    const bool small_on = c->stage2 == MPRES_STAGE2_SMALL && c->sc.usable;
    bool sparse_mp = c->minplus_sparse && k_p >= 512;
    int rc;
    // ---- workspace that does not depend on the base: windows, shift plane of A', candidate lists of A', per-panel lists and delta planes ----
    const size_t hdr = pkg_header_bytes(nb_p, k_p);
    const size_t bytesIA = pad1k((size_t) m_ps * sizeof(OuterInfo)), bytesSA = pad1k((size_t) m_ps * k_p * 2);
    const size_t bytesTab = pad1k((size_t) (3 * c->hc.log2M + 2) * N * sizeof(int));
    const size_t bytesCandA = 2 * pad1k((size_t) m_ps * kMcT * 4) + pad1k((size_t) m_ps * 4);
    const size_t bytesList = pad1k((size_t) m * nb * sizeof(long long));          // per panel, three of them: todo, slow, (min,+) pairs
    const size_t bytesD = pad1k((size_t) nb_p * m_p * 2);                          // per panel, three of them: delta, D1, D2
    const size_t perPanel = 3 * bytesList + 3 * bytesD;
    const size_t bytesPart = pad1k((size_t) std::max<long long>(m_ps, nb_p) * sizeof(OuterPart));
    void *pMisc;
    if ((rc = ws_reserve(c, 6, bytesIA + bytesPart + bytesSA + bytesTab + bytesCandA + (size_t) W * perPanel + (sh ? 0 : hdr) + 4096, &pMisc))) return rc;
    char *pm = (char *) pMisc;
    OuterInfo *IA = (OuterInfo *) pm; pm += bytesIA;
    OuterPart *part = (OuterPart *) pm; pm += bytesPart;
    int16_t *SA = (int16_t *) pm; pm += bytesSA;
    int *scal_tab = (int *) pm; pm += bytesTab;
    int *cposA = (int *) pm; pm += pad1k((size_t) m_ps * kMcT * 4);
    int *cvalA = (int *) pm; pm += pad1k((size_t) m_ps * kMcT * 4);
    int *thrA = (int *) pm; pm += pad1k((size_t) m_ps * 4);
    char *panel_ws = pm; pm += (size_t) W * perPanel;
    char *local_hdr = pm;                                    // single-GPU call: the header of the one package
    int *nprime = c->d_counter + 2, *sel = c->d_counter + 4;

    // element (o, l) index strides: op(A)(i, l) and op(B)(l, j)
    const long long soA = ta ? lda : 1, slA = ta ? 1 : lda;
    const long long soB = tb ? 1 : ldb, slB = tb ? ldb : 1;
    auto mark = [&](int i) { if (c->profiling) { if (!c->ev[i]) cudaEventCreate(&c->ev[i]); cudaEventRecord(c->ev[i], st); } };
    prof_reset(c, st);
    mark(0);
    int launches = 0;
    // own panel: header in the receive buffer (sharded) or in the misc workspace; the planes pointer is known once P is
    char *own_hdr = sh ? sh->recv + (size_t) rank * sh->pkg_stride : local_hdr;
    PanelPkg own = pkg_carve(own_hdr, nullptr, nb_p, k_p);
    const SoA Bown = soa_shift(B, (long long) rank * nb * soB, N);
    // exponent base and window of every line: a warp per line when the line runs along the contiguous direction of the input, else the
    // chunked kernel with lanes along the lines
    auto outer_info = [&](const SoA &X, long long so, long long sl, int outer, int inner, OuterInfo *info) {
        if (so == 1 && sl != 1 && inner >= 64) {
            const int lineb = (outer + 31) / 32;
            int chunks = std::max(1, std::min((c->sm_count * 8 + lineb - 1) / lineb, inner / 32));
            const int chunk = (inner + chunks - 1) / chunks;
            chunks = (inner + chunk - 1) / chunk;
            k_outer_part_init<<<(outer + 255) / 256, 256, 0, st>>>(part, outer);
            k_outer_part<<<dim3((unsigned) lineb, (unsigned) chunks), 256, 0, st>>>(X, sl, outer, inner, chunk, part);
            k_outer_part_final<<<(outer + 255) / 256, 256, 0, st>>>(c->dconsts, part, outer, info);
            launches += 2;
        } else if (sl == 1 && outer < c->sm_count * 16 && inner >= 2048) {
            // few long lines: several blocks per line
            const int chunks = std::max(1, std::min((c->sm_count * 16 + outer - 1) / outer, inner / 1024));
            const int chunk = (inner + chunks - 1) / chunks;
            k_outer_part_init<<<(outer + 255) / 256, 256, 0, st>>>(part, outer);
            k_outer_part_l<<<dim3((unsigned) outer, (unsigned) ((inner + chunk - 1) / chunk)), 256, 0, st>>>(X, so, outer, inner, chunk, part);
            k_outer_part_final<<<(outer + 255) / 256, 256, 0, st>>>(c->dconsts, part, outer, info);
            launches += 2;
        } else {
            k_outer_info<<<(unsigned) ((outer * 32ll + 255) / 256), 256, 0, st>>>(c->dconsts, X, so, sl, outer, inner, info);
        }
    };
    outer_info(A, soA, slA, m, k, IA);
    outer_info(Bown, soB, slB, nb, k, own.IB);
    prof_mark(c, st, "k_outer_info");
    Xchg x;
    memset(&x, 0, sizeof(x));
    x.world = W; x.rank = rank;
    if (sh) {
        x.epoch = sh->epoch; x.parity = (int) (sh->epoch & 1u);
        for (int p = 0; p < W; ++p) x.peer[p] = sh->peer_xchg[p];
        x.err = c->d_counter + 3;
    }
    // (2: the significands may be cut into slices when the sums exceed the one-byte base -- single-GPU calls of a format with a binary epilogue)
    const int small_mode = !small_on ? 0 : ((!sh && bin2_variant(c) && c->small_persistent && !c->align_mma) ? 2 : 1);
    k_choose_base<<<1, 256, 0, st>>>(c->dconsts, IA, m, own.IB, nb, k, c->reduced_base, small_mode, nprime, sel, x);
    launches += 3;
    // the host needs the choice: which kernels to launch, how many planes to reserve and to send
    CUDA_TRY(cudaMemcpyAsync(c->h_sel, c->d_counter + 2, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_sel + 6, sel + kSelSlices, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int P = c->h_sel[2], nin = c->h_sel[3];
    const int slices = std::max(1, c->h_sel[6]), ND = 2 * slices - 1;
    // full-precision inputs: the exact sums exceed the number format (window guard of kernels_norm.cuh) but not the one-byte base --
    // stage 3 rebuilds them in binary and rounds once (kernels_bin.cuh); the (min,+) exponent product is not needed then
    const bool binary = P > 0 && c->h_sel[5] > c->hc.log2M - 2;
    c->last_binary = binary;
    c->last_fast_ok = P > 0;
    c->last_nin = c->h_sel[3];
    if (sh && c->h_sel[1]) return -50;                       // a rank did not reach the call
    prof_mark(c, st, "k_choose_base");
    (void) nin;
    if (P <= 0) {
        rc = gemm_fast_limb(c, ta, tb, m, n, k, A, lda, B, ldb, alpha, beta, Cm, ldc, st, IA);
        if (rc) return rc;
        mark(3);
        c->ev_valid = c->profiling;
        for (int i = 0; i < launches; ++i) LAUNCHED(c);
        CUDA_TRY(cudaGetLastError());
        *done = true;
        return 0;
    }
    if (binary) sparse_mp = false;
    // ---- workspace that depends on the base ----
    void *pQA, *pS8, *pS, *pSAT = nullptr, *pPl = nullptr;
    if ((rc = ws_reserve(c, 8, (size_t) slices * P * m_ps * k_p, &pQA))) return rc;
    const size_t s8_panel = (size_t) ND * P * nb_p * m_ps, s_panel = (size_t) N * nb_p * m_p * 4;
    if ((rc = ws_reserve(c, 10, (size_t) W * s8_panel, &pS8))) return rc;
    if (binary) pS = nullptr;
    else if ((rc = ws_reserve(c, 5, (size_t) W * s_panel, &pS))) return rc;
    if (sparse_mp && (rc = ws_reserve(c, 11, (size_t) k_p * m_ps * 2 + 256, &pSAT))) return rc;
    if (!sh && (rc = ws_reserve(c, 9, (size_t) slices * P * nb_p * k_p, &pPl))) return rc;
    const size_t planes_bytes = (size_t) slices * P * nb_p * k_p;
    if (sh && hdr + planes_bytes > sh->pkg_stride) return -7;
    own.QB = sh ? (uint8_t *) (own_hdr + hdr) : (uint8_t *) pPl;
    auto panel_pkg = [&](int g) { return sh ? pkg_carve(sh->recv + (size_t) g * sh->pkg_stride, sh->recv + (size_t) g * sh->pkg_stride + hdr, nb_p, k_p) : own; };
    int16_t *SAT = (int16_t *) pSAT;

    // ---- stage 1: alignment into the one-byte base ----
    if (!c->attr_ext) {
        cudaFuncSetAttribute(k_ext_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        cudaFuncSetAttribute(k_ext_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ext_small_smem(512, 128));
        c->attr_ext = true;
    }
    {
        const int resident = align_small_blocks(c->align_mma != 0, slices > 1);
        const unsigned gA = (unsigned) std::min<long long>((m_ps / kASo) * (k_p / kASl), (long long) c->sm_count * resident);
        const unsigned gB = (unsigned) std::min<long long>((nb_p / kASo) * (k_p / kASl), (long long) c->sm_count * resident);
        // the rank's column block first: its package has to travel (the copies of the previous call have left the package by now: the
        // peers' flags of that call were raised behind them and every rank has passed this call's rendezvous; the wait is a safeguard)
        if (sh) for (int si = 0; si < sh->npush && si < W - 1; ++si) CUDA_TRY(cudaStreamWaitEvent(st, sh->ev_push[si], 0));
        if (c->align_mma) {
            if (!c->attr_align_mma) { cudaFuncSetAttribute(k_align_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(true)); c->attr_align_mma = true; }
            k_align_small<true><<<gB, 256, align_small_smem(true), st>>>(c->dconsts, Bown, soB, slB, nb, k, own.IB, own.QB, own.SB, nb_p, k_p, sel);
        } else {
            if (!c->attr_align) {
                cudaFuncSetAttribute(k_align_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(false));
                cudaFuncSetAttribute(k_align_small<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) align_small_smem(false));
                c->attr_align = true;
            }
            if (slices > 1) k_align_small<false, true><<<gB, 256, align_small_smem(false), st>>>(c->dconsts, Bown, soB, slB, nb, k, own.IB, own.QB, own.SB, nb_p, k_p, sel);
            else k_align_small<false><<<gB, 256, align_small_smem(false), st>>>(c->dconsts, Bown, soB, slB, nb, k, own.IB, own.QB, own.SB, nb_p, k_p, sel);
        }
        prof_mark(c, st, "k_align_small(B)");
        if (sparse_mp) {
            k_mp_select<<<(unsigned) nb, 256, 0, st>>>(own.SB, k_p, (int) k_p, nb, own.cposB, own.cvalB, own.thrB);
            k_mp_transpose<<<dim3((unsigned) (nb_p / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(own.SB, k_p, own.SBT, nb_p);
            launches += 2;
            prof_mark(c, st, "k_mp_select+transpose(B)");
        }
        if (sh) {
            // the package is complete: hand it to the copy engines, peers in the order they will need it (rank - 1 multiplies it second)
            CUDA_TRY(cudaEventRecord(sh->ev_pkg, st));
            const size_t bytes = hdr + planes_bytes;
            for (int d = 1; d < W; ++d) {
                const int peer = (rank - d + W) % W;
                const int si = (d - 1) % sh->npush;
                cudaStream_t ps = sh->push[si];
                if (d - 1 < sh->npush) CUDA_TRY(cudaStreamWaitEvent(ps, sh->ev_pkg, 0));
                CUDA_TRY(cudaMemcpyAsync(sh->peer_recv[peer] + (size_t) rank * sh->pkg_stride, own_hdr, bytes, cudaMemcpyDeviceToDevice, ps));
                k_set_flag<<<1, 1, 0, ps>>>(sh->peer_flags[peer] + rank, sh->epoch);
                ++launches;
            }
            for (int si = 0; si < sh->npush && si < W - 1; ++si) CUDA_TRY(cudaEventRecord(sh->ev_push[si], sh->push[si]));
        }
        if (c->align_mma) k_align_small<true><<<gA, 256, align_small_smem(true), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
        else if (slices > 1) k_align_small<false, true><<<gA, 256, align_small_smem(false), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
        else k_align_small<false><<<gA, 256, align_small_smem(false), st>>>(c->dconsts, A, soA, slA, m, k, IA, (uint8_t *) pQA, SA, m_ps, k_p, sel);
        launches += 2;
        prof_mark(c, st, "k_align_small(A)");
        if (sparse_mp) {
            k_mp_select<<<(unsigned) m, 256, 0, st>>>(SA, k_p, (int) k_p, m, cposA, cvalA, thrA);
            k_mp_transpose<<<dim3((unsigned) (m_ps / 64), (unsigned) (k_p / 64)), 256, 0, st>>>(SA, k_p, SAT, m_ps);
            launches += 2;
            prof_mark(c, st, "k_mp_select+transpose(A)");
        }
    }
    const bool allow_fb = c->mode == MPRES_MODE_AUTO;
    bool have_fast = c->stage3 == 0;
    switch (N) { case 8: case 16: case 24: case 32: case 40: case 48: case 56: case 64: break; default: have_fast = false; }
    const bool f32 = c->sc.usable && c->sc.red_shift >= 24 && c->sc.red_shift <= 27 && c->norm32;
    const bool fused = have_fast && c->fuse_ext;
    if (have_fast && !binary) {
        const int rowsT = 3 * c->hc.log2M + 2;
        k_scalar_tables<<<(rowsT * N + 255) / 256, 256, 0, st>>>(c->dconsts, alpha, beta, scal_tab);
        ++launches;
    }
    struct PanelWs { long long *todo, *slow, *mplist; int16_t *D, *D1, *D2; int *cnt; };
    auto panel_carve = [&](int g) {
        PanelWs w;
        char *q = panel_ws + (size_t) g * perPanel;
        w.todo = (long long *) q; q += bytesList;
        w.slow = (long long *) q; q += bytesList;
        w.mplist = (long long *) q; q += bytesList;
        w.D = (int16_t *) q; q += bytesD;
        w.D1 = (int16_t *) q; q += bytesD;
        w.D2 = (int16_t *) q;
        w.cnt = c->d_counter + kCounterBlock * g;
        return w;
    };
    // (min,+) product of the shift planes of one panel: delta(i, j) for the m x nb block
    auto minplus_panel = [&](int g) {
        const PanelPkg pk = panel_pkg(g);
        const PanelWs w = panel_carve(g);
        if (sparse_mp) {
            k_mp_gather<<<(unsigned) m, 256, 0, st>>>(cposA, cvalA, pk.SBT, nb_p, m, (int) nb_p, w.D1, nb_p);
            k_mp_gather<<<(unsigned) nb, 256, 0, st>>>(pk.cposB, pk.cvalB, SAT, m_ps, nb, (int) m_p, w.D2, m_p);
            k_mp_combine<<<dim3((unsigned) (m_p / 64), (unsigned) (nb_p / 64)), 256, 0, st>>>(w.D1, nb_p, w.D2, m_p, thrA, pk.thrB, m, nb, w.D, m_p, w.mplist, w.cnt + 6);
            k_mp_fix<<<c->sm_count * 4, 256, 0, st>>>(SA, pk.SB, k_p, w.D, m_p, w.mplist, w.cnt + 6);
            launches += 4;
        } else {
            k_minplus<<<dim3((unsigned) (m_p / kMpTI), (unsigned) (nb_p / kMpTJ)), 256, 0, st>>>(SA, pk.SB, w.D, k_p, m_p, nb_p);
            ++launches;
        }
    };
    // ---- stage 2: one u8 GEMM per one-byte modulus, all panels in one persistent launch ----
    SmallPanels pan;
    pan.count = W; pan.first = rank; pan.own = rank; pan.epoch = sh ? sh->epoch : 0u; pan.flags = sh ? sh->flags : nullptr; pan.s8_panel = (long long) s8_panel;
    pan.ring = 0; pan.plane_rows = 0; pan.slices = slices; pan.pair = -1;
    const uint8_t *PB0 = sh ? (const uint8_t *) (sh->recv + hdr) : own.QB;
    const long long pb_panel = sh ? (long long) sh->pkg_stride : (long long) planes_bytes;
    int gemm_launches = 0;
    auto stage2 = [&]() -> int {
        // sliced significands: all slice pairs of a sum in ONE K loop (pairs x 255^2 x chunk < 2^31).  Measured at config 3 with p-bit inputs:
        // 31.9 ms; one launch per pair index with the launches adding up in S8 (MPRES_SLICE_PASSES=1: two planes per modulus in flight
        // instead of six) took 56.6 ms -- nine short K loops per tile set expose the epilogue and the pipeline fill nine times
        const bool passes = slices > 1 && getenv("MPRES_SLICE_PASSES") && atoi(getenv("MPRES_SLICE_PASSES")) != 0;
        const long long chunk = (slices > 1 && !passes) ? (kSmallKChunk / slices) / 128 * 128 : kSmallKChunk;
        for (int pr = 0; pr < (passes ? slices : 1); ++pr) {
            pan.pair = passes ? pr : -1;
            for (long long kb = 0; kb < k_p; kb += chunk) {
                const int kl = (int) std::min<long long>(chunk, k_p - kb);
                int r2 = launch_small_umma(c, P, (const uint8_t *) pQA, PB0, pb_panel, (uint8_t *) pS8, m_ps, nb_p, k_p, kb, kl, kb > 0 || pr > 0, sel, pan, st);
                if (r2) return r2;
                gemm_launches += 1;
            }
        }
        return 0;
    };
    if (!sh) {
        if (!binary) minplus_panel(0);
        prof_mark(c, st, "k_mp_gather+combine+fix");
        mark(1);
        if ((rc = stage2())) return rc;
        prof_mark(c, st, "k_small_umma_p");
        mark(2);
    } else {
        // sharded: the multiplication starts on the rank's own panel while the others arrive; the (min,+) products follow
        mark(1);
        if ((rc = stage2())) return rc;
        prof_mark(c, st, "k_small_umma_p");
        mark(2);
    }
    // ---- stage 3 per panel: base extension, normalisation, alpha / beta epilogue ----
    for (int pi = 0; pi < W; ++pi) {
        const int g = (rank + pi) % W;
        const PanelPkg pk = panel_pkg(g);
        const PanelWs w = panel_carve(g);
        const SoA Cg = soa_shift(Cm, (long long) g * nb * ldc, N);
        const SoA Bg = soa_shift(B, (long long) g * nb * soB, N);
        uint8_t *S8g = (uint8_t *) pS8 + (size_t) g * s8_panel;
        int *Sg = (int *) ((char *) pS + (size_t) g * s_panel);
        if (sh) {
            if (g != rank) { k_wait_flag<<<1, 1, 0, st>>>(sh->flags + g, sh->epoch); ++launches; }
            if (!binary) minplus_panel(g);
            if (pi == 0) prof_mark(c, st, "k_mp_gather+combine+fix");
        }
        if (binary) {
            if ((rc = bin_norm_segment(c, pi == 0, m, nb, S8g, m_ps, nb_p, sel, IA, pk.IB, alpha, beta, Cg, ldc, st, &launches, slices, w.todo, w.cnt))) return rc;
            if (pi == 0) prof_mark(c, st, "k_bin_norm");
            if (slices > 1) {
                MPRES_DISPATCH(N, {
                    k_gemm_todo<G, R><<<c->sm_count * 8, 128, 0, st>>>(c->dconsts, ta, tb, m, nb, k, A, lda, Bg, ldb, alpha, beta, Cg, ldc, w.todo, w.cnt);
                });
                ++launches;
            }
            continue;
        }
        if (!fused) {
            const unsigned gx = (unsigned) std::min<long long>((m_p / kXT) * nb, (long long) c->sm_count * 4);     // persistent: four blocks per SM
            const size_t sm = ext_small_smem(c->sc.ext_cols, N);
            if (c->sc.red_shift) k_ext_small<true><<<gx, kXT, sm, st>>>(c->dconsts, m, nb, S8g, m_p, m_ps, nb_p, Sg, nb_p, sel);
            else k_ext_small<false><<<gx, kXT, sm, st>>>(c->dconsts, m, nb, S8g, m_p, m_ps, nb_p, Sg, nb_p, sel);
            ++launches;
            if (pi == 0) prof_mark(c, st, "k_ext_small");
        }
        auto norm_fast = [&](auto tag) {
            constexpr int NQ = decltype(tag)::value;
            const unsigned g3 = (unsigned) ((long long) ((m + kNormFastThreads - 1) / kNormFastThreads) * nb);
            const size_t sm_cds = (size_t) kNormFastThreads * (NQ + 1) * sizeof(int);
            if (fused) {
                // base extension and normalisation in one kernel
                const unsigned gx = (unsigned) ((m_p / kXT) * nb);
                const size_t sm = ext_small_smem(c->sc.ext_cols, NQ) + (ext_norm_cds_aliased(NQ) ? 0 : sm_cds);
                if (!(c->attr_fused >> (NQ / 8) & 1ull)) {
                    cudaFuncSetAttribute(k_ext_norm_small<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
                    cudaFuncSetAttribute(k_ext_norm_small<NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);
                    c->attr_fused |= 1ull << (NQ / 8);
                }
                if (f32)
                    k_ext_norm_small<NQ, true><<<gx, kXT, sm, st>>>(c->dconsts, m, nb, k, S8g, m_p, m_ps, nb_p, Sg, nb_p, sel, w.D, IA, pk.IB, alpha, beta, Cg, ldc,
                                                                   scal_tab, w.todo, w.cnt, w.slow, w.cnt + 1, allow_fb);
                else
                    k_ext_norm_small<NQ, false><<<gx, kXT, sm, st>>>(c->dconsts, m, nb, k, S8g, m_p, m_ps, nb_p, Sg, nb_p, sel, w.D, IA, pk.IB, alpha, beta, Cg, ldc,
                                                                    scal_tab, w.todo, w.cnt, w.slow, w.cnt + 1, allow_fb);
                ++launches;
                return;
            }
            auto launch_norm = [&](auto kern, size_t smem) {
                kern<<<g3, kNormFastThreads, smem, st>>>(c->dconsts, m, nb, k, (const int *) Sg, w.D, m_p, nb_p, IA, pk.IB, alpha, beta, Cg, ldc,
                                                         scal_tab, w.todo, w.cnt, w.slow, w.cnt + 1, allow_fb, nullptr);
            };
            if (c->norm_staged) {
                const size_t sm_st = sm_cds + (size_t) NQ * kNormFastThreads * sizeof(int);
                if (sm_st > 48 * 1024 && !(c->attr_norm >> (NQ / 8) & 1ull)) {
                    cudaFuncSetAttribute(k_norm_fast<NQ, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                    cudaFuncSetAttribute(k_norm_fast<NQ, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm_st);
                    c->attr_norm |= 1ull << (NQ / 8);
                }
                if (f32) launch_norm(k_norm_fast<NQ, true, true>, sm_st); else launch_norm(k_norm_fast<NQ, false, true>, sm_st);
            } else {
                if (f32) launch_norm(k_norm_fast<NQ, true, false>, sm_cds); else launch_norm(k_norm_fast<NQ, false, false>, sm_cds);
            }
            ++launches;
        };
        bool hf = have_fast;
        if (hf) {
            switch (N) {
                case 8: norm_fast(std::integral_constant<int, 8>{}); break;
                case 16: norm_fast(std::integral_constant<int, 16>{}); break;
                case 24: norm_fast(std::integral_constant<int, 24>{}); break;
                case 32: norm_fast(std::integral_constant<int, 32>{}); break;
                case 40: norm_fast(std::integral_constant<int, 40>{}); break;
                case 48: norm_fast(std::integral_constant<int, 48>{}); break;
                case 56: norm_fast(std::integral_constant<int, 56>{}); break;
                case 64: norm_fast(std::integral_constant<int, 64>{}); break;
                default: hf = false;
            }
            if (pi == 0) prof_mark(c, st, fused ? "k_ext_norm_small" : "k_norm_fast");
        }
        MPRES_DISPATCH(N, {
            if (hf) {
                k_norm_list<G, R><<<c->sm_count * 8, 256, 0, st>>>(c->dconsts, m, nb, k, (const int *) Sg, w.D, m_p, nb_p, IA, pk.IB, alpha, beta, Cg, ldc,
                                                                  w.slow, w.cnt + 1);
            } else {
                constexpr int kNormTile = 256 / G;
                const unsigned g3 = (unsigned) ((long long) ((m + kNormTile - 1) / kNormTile) * nb);
                k_normalize_epilogue<G, R><<<g3, 256, (size_t) N * (kNormTile + 1) * 4, st>>>(
                    c->dconsts, m, nb, k, (const int *) Sg, w.D, m_p, nb_p, IA, pk.IB, alpha, beta, Cg, ldc, w.todo, w.cnt, allow_fb);
            }
            ++launches;
            if (allow_fb) {
                k_gemm_todo<G, R><<<c->sm_count * 8, 128, 0, st>>>(c->dconsts, ta, tb, m, nb, k, A, lda, Bg, ldb, alpha, beta, Cg, ldc, w.todo, w.cnt);
                ++launches;
            }
        });
        if (pi == 0) prof_mark(c, st, "k_norm_list+k_gemm_todo");
    }
    mark(3);
    prof_mark(c, st, "end");
    c->ev_valid = c->profiling;
    c->last_stage2_launches = gemm_launches;
    for (int i = 0; i < launches + gemm_launches; ++i) LAUNCHED(c);
    CUDA_TRY(cudaGetLastError());
    *done = true;
    return 0;
}
