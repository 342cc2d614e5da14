// host_consts.hpp -- every precomputed constant of the RNS floating-point format, derived from the
// moduli alone with a small exact integer type (no GMP / MPFR on the product path).
//
// Replaces rns_const_init (reference src/rns.cuh:324-442) and mp_const_init
// (src/arith/arith_utils.cuh:44-85); definitions are tabulated in SURVEY.md Appendix A.  Bit-for-bit
// agreement with the reference's GMP/MPFR-derived values is enforced by tests/test_constants.py.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace mpres {

constexpr int kScalingThreshold = 30;        // RNS_P2_SCALING_THRESHOLD (src/params.h:42)
constexpr double kEvalRelError = 0.0000001;  // RNS_EVAL_RELATIVE_ERROR (src/params.h:47)
constexpr int kMaxModuli = 128;

// Predefined moduli sets of src/params/32-bit-n-double-moduli/params.<N>_<n>double.h (data of the
// number format).  Generated table, see moduli_sets.inc.
struct ModuliSet { int n; const int *values; };
#include "moduli_sets.inc"

// Non-negative arbitrary-size integer, little-endian 32-bit limbs; just what the constants need.
class BigUInt {
public:
    std::vector<uint32_t> limb;
    BigUInt(uint64_t v = 0) { while (v) { limb.push_back((uint32_t) v); v >>= 32; } }
    void mul_small(uint32_t f) {
        uint64_t carry = 0;
        for (auto &l : limb) { uint64_t t = (uint64_t) l * f + carry; l = (uint32_t) t; carry = t >> 32; }
        if (carry) limb.push_back((uint32_t) carry);
    }
    // returns remainder, *this becomes the quotient
    uint32_t div_small(uint32_t d) {
        uint64_t rem = 0;
        for (size_t i = limb.size(); i-- > 0;) {
            uint64_t cur = (rem << 32) | limb[i];
            limb[i] = (uint32_t) (cur / d);
            rem = cur % d;
        }
        while (!limb.empty() && limb.back() == 0) limb.pop_back();
        return (uint32_t) rem;
    }
    int bit_length() const {
        if (limb.empty()) return 0;
        return (int) (32 * (limb.size() - 1)) + (32 - __builtin_clz(limb.back()));
    }
    uint64_t low_bits(int j) const {  // value mod 2^j, j <= 63
        uint64_t v = 0;
        if (!limb.empty()) v = limb[0];
        if (limb.size() > 1) v |= (uint64_t) limb[1] << 32;
        return j >= 64 ? v : v & ((1ull << j) - 1);
    }
    uint64_t top_bits(int count) const {  // the `count` (<= 64) most significant bits
        int L = bit_length();
        uint64_t v = 0;
        for (int b = L - 1; b >= L - count && b >= 0; --b) v = (v << 1) | ((limb[b / 32] >> (b % 32)) & 1u);
        return v;
    }
    bool any_bits_below(int pos) const {  // any set bit at positions [0, pos)
        for (int b = 0; b < pos; ++b) if ((limb[b / 32] >> (b % 32)) & 1u) return true;
        return false;
    }
};

inline int64_t inverse_mod(int64_t x, int64_t m) {  // extended Euclid; gcd(x, m) must be 1
    int64_t a = ((x % m) + m) % m, b = m, u = 1, v = 0;
    while (b) { int64_t q = a / b, t = a - q * b; a = b; b = t; t = u - q * v; u = v; v = t; }
    return a == 1 ? ((u % m) + m) % m : 0;
}

struct ErHost { double frac; long exp; };

struct HostConsts {
    int N = 0, log2M = 0, mp_precision = 0, mp_h = 0, mp_j = 0, ref_factor = 0;
    double accuracy = 0;
    ErHost unit_low{}, unit_upp{}, inv_low{}, inv_upp{};
    std::vector<int> moduli, part_inverse, pow2, m_pow2, mi_pow2, pow2_inv, mrc_inv, inv_pow2_ext;
    // reduced-base fast path (not in the reference): floor(log2(m_0 * ... * m_{n-1})) for n = 0..N and
    // prefix_mod[i * N + q] = (m_0 * ... * m_{i-1}) mod m_q for i = 0..N
    std::vector<int> prefix_log2, prefix_mod;
    // CRT base extension from the first c moduli (c = 1..N-1), M' = m_0 ... m_{c-1}, M'_i = M' / m_i:
    //   ext_w[c * N + i] = (M'_i)^-1 mod m_i  (i < c),   ext_t[(c * N + q) * N + i] = M'_i mod m_q  (i < c)
    std::vector<int> ext_w, ext_t;
    std::vector<int> spow2;   // [2 (log2M+1) + 1][N]  row 2s: 2^s, row 2s+1: -2^s, last row: zeros (mp_gemv / mp_dot term multipliers)
    std::vector<int> wpow2;   // [log2M+1][N]  w_i * 2^j mod m_i (interval evaluation of a magnified number in one multiplication)
    int ext_lazy = 0;   // 1 if N * max(m)^2 < 2^63: the extension sums need no intermediate reduction
    std::vector<double> recip_rd, recip_ru;
    std::vector<uint64_t> barrett;  // floor((2^64 - 1) / m_i)
};

// returns 0 on success, negative on unusable moduli
inline int compute_constants(const int *mods, int N, HostConsts &c) {
    if (N < 2 || N > kMaxModuli) return -1;        // (the reference ships sets of up to 160 moduli; this library stops at 128)
    if (N & 1) return -5;                             // odd N: the reference's mp_float_t is padded to 4N + 44 bytes; only even counts are laid out here
    for (int i = 0; i < N; ++i) {
        if (mods[i] < 3 || (mods[i] & 1) == 0) return -2;  // power-of-two scaling needs odd moduli
        for (int j = 0; j < i; ++j) if (std::__gcd(mods[i], mods[j]) != 1) return -3;
    }
    c.N = N;
    c.moduli.assign(mods, mods + N);
    BigUInt M(1);
    for (int i = 0; i < N; ++i) M.mul_small((uint32_t) mods[i]);
    const int L = M.bit_length();
    if (L <= 64) return -4;
    c.log2M = L - 1;  // RNS_MODULI_PRODUCT_LOG2

    // w_i = (M / m_i)^-1 mod m_i                                        rns.cuh:339-346
    c.part_inverse.resize(N);
    for (int i = 0; i < N; ++i) {
        int64_t p = 1;
        for (int j = 0; j < N; ++j) if (j != i) p = p * (mods[j] % mods[i]) % mods[i];
        c.part_inverse[i] = (int) inverse_mod(p, mods[i]);
    }
    // 2^j mod m_i, j = 0..log2M                                          rns.cuh:352-358
    c.pow2.resize((size_t) (c.log2M + 1) * N);
    c.inv_pow2_ext.resize((size_t) (c.log2M + 1) * N);
    for (int i = 0; i < N; ++i) {
        int64_t v = 1, iv = 1, half = (mods[i] + 1) / 2;
        for (int j = 0; j <= c.log2M; ++j) {
            c.pow2[(size_t) j * N + i] = (int) v;
            c.inv_pow2_ext[(size_t) j * N + i] = (int) iv;
            v = v * 2 % mods[i];
            iv = iv * half % mods[i];
        }
    }
    // M mod 2^j, M_i mod 2^j, (2^j)^-1 mod m_i for j = 1..30             rns.cuh:360-382
    c.m_pow2.resize(kScalingThreshold);
    c.mi_pow2.resize((size_t) kScalingThreshold * N);
    c.pow2_inv.resize((size_t) kScalingThreshold * N);
    for (int j = 0; j < kScalingThreshold; ++j) c.m_pow2[j] = (int) M.low_bits(j + 1);
    for (int i = 0; i < N; ++i) {
        BigUInt Mi = M;
        Mi.div_small((uint32_t) mods[i]);
        for (int j = 0; j < kScalingThreshold; ++j) {
            c.mi_pow2[(size_t) j * N + i] = (int) Mi.low_bits(j + 1);
            c.pow2_inv[(size_t) j * N + i] = c.inv_pow2_ext[(size_t) (j + 1) * N + i];
        }
    }
    // interval-evaluation constants                                       rns.cuh:385-410
    c.accuracy = 4 * std::pow(2.0, 1 - 53) * N * std::log2((double) N) * (1 + kEvalRelError / 2) / kEvalRelError;
    c.ref_factor = (int) std::floor(std::log2(1 / (2 * c.accuracy)));
    // The reference divides by M rounded DOWN to 53 bits (mpfr_init default precision, rns.cuh:336).
    const uint64_t mant = M.top_bits(53);  // in [2^52, 2^53)
    const int s = L - 53;
    {
        unsigned __int128 num = (unsigned __int128) 1 << 105;
        uint64_t q = (uint64_t) (num / mant);
        bool inexact = (num % mant) != 0;
        auto pack = [&](uint64_t qq) {
            ErHost e;
            if (qq == (1ull << 53)) { e.frac = 0.5; e.exp = -52 - s + 1; }
            else { e.frac = (double) qq / 9007199254740992.0; e.exp = -52 - s; }
            return e;  // mpfr_get_d_2exp convention: frac in [0.5, 1)
        };
        c.unit_low = pack(q);
        c.unit_upp = pack(q + (inexact ? 1 : 0));
    }
    // (M-1)/M = 1 - 1/M: for log2 M > 54 the 53-bit roundings are 1 - 2^-53 (down) and 1.0 (up)
    c.inv_low = {1.0 - std::ldexp(1.0, -53), 0};
    c.inv_upp = {0.5, 1};
    // 1/m_i widened by the reference's host emulation of directed rounding   rns.cuh:412-415,
    // dinterval.cuh:142-154, bitwise.cuh:51-61
    c.recip_rd.resize(N);
    c.recip_ru.resize(N);
    {
        const double eps = std::ldexp(1.0, -53), phi1 = eps * (1 + 2 * eps), eta = std::ldexp(1.0, -1074);
        for (int i = 0; i < N; ++i) {
            volatile double q = 1.0 / (double) mods[i];
            volatile double e = phi1 * std::fabs(q);
            e = e + eta;
            c.recip_rd[i] = q - e;
            c.recip_ru[i] = q + e;
        }
    }
    // m_i^-1 mod m_j for j > i                                            rns.cuh:417-426
    c.mrc_inv.assign((size_t) N * N, 0);
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) c.mrc_inv[(size_t) i * N + j] = (int) inverse_mod(mods[i], mods[j]);
    c.barrett.resize(N);
    for (int i = 0; i < N; ++i) c.barrett[i] = ~0ull / (uint64_t) mods[i];
    // MP_PRECISION, MP_H, MP_J on the 53-bit truncated M                  arith_utils.cuh:44-61
    c.mp_precision = (L - 1) / 2 - 1;
    {
        // smallest h with 4^h >= Md  (Md = mant * 2^s)
        int h = (L + 1) / 2;             // 4^h >= 2^L > Md
        // 4^(h-1) >= Md  <=>  2h-2 >= L, or 2h-2 == L-1 and Md == 2^(L-1)
        while (true) {
            int e = 2 * (h - 1);
            bool ge = e >= L || (e == L - 1 && mant == (1ull << 52));
            if (!ge) break;
            --h;
        }
        c.mp_h = -h;
    }
    c.mp_j = -1;
    // prefix products of the moduli: sizes (for choosing how many moduli an exact sum needs) and residues
    // (mixed-radix weights for the base extension)
    c.prefix_log2.assign(N + 1, 0);
    c.prefix_mod.assign((size_t) (N + 1) * N, 0);
    {
        BigUInt P(1);
        std::vector<int64_t> r(N, 1);
        for (int i = 0; i <= N; ++i) {
            c.prefix_log2[i] = P.bit_length() - 1;
            for (int q = 0; q < N; ++q) c.prefix_mod[(size_t) i * N + q] = (int) (r[q] % mods[q]);
            if (i < N) {
                P.mul_small((uint32_t) mods[i]);
                for (int q = 0; q < N; ++q) r[q] = r[q] % mods[q] * (mods[i] % mods[q]) % mods[q];
            }
        }
    }
    c.spow2.assign((size_t) (2 * (c.log2M + 1) + 1) * N, 0);
    for (int j = 0; j <= c.log2M; ++j)
        for (int i = 0; i < N; ++i) {
            const int v = c.pow2[(size_t) j * N + i];
            c.spow2[(size_t) (2 * j) * N + i] = v;
            c.spow2[(size_t) (2 * j + 1) * N + i] = v ? mods[i] - v : 0;
        }
    c.wpow2.resize((size_t) (c.log2M + 1) * N);
    for (int j = 0; j <= c.log2M; ++j)
        for (int i = 0; i < N; ++i) c.wpow2[(size_t) j * N + i] = (int) ((int64_t) c.pow2[(size_t) j * N + i] * c.part_inverse[i] % mods[i]);
    c.ext_w.assign((size_t) N * N, 0);
    c.ext_t.assign((size_t) N * N * N, 0);
    {
        std::vector<int64_t> suf(N + 1);
        for (int cc = 1; cc < N; ++cc)
            for (int q = 0; q < N; ++q) {
                const int64_t mq = mods[q];
                suf[cc] = 1;
                for (int i = cc - 1; i >= 0; --i) suf[i] = suf[i + 1] * (mods[i] % mq) % mq;   // m_i ... m_{cc-1}
                for (int i = 0; i < cc; ++i) {
                    const int64_t t = (int64_t) c.prefix_mod[(size_t) i * N + q] * suf[i + 1] % mq;
                    c.ext_t[((size_t) cc * N + q) * N + i] = (int) t;
                    if (q == i) c.ext_w[(size_t) cc * N + i] = (int) inverse_mod(t, mq);
                }
            }
        int64_t mx = 0;
        for (int i = 0; i < N; ++i) mx = mods[i] > mx ? mods[i] : mx;
        c.ext_lazy = ((long double) N * (long double) mx * (long double) mx < 9.0e18L) ? 1 : 0;
    }
    return 0;
}


// ---- small-modulus base of the tensor-core stage 2 (not in the reference) ------------------------------
// The exact per-entry sums S of the fast mp_gemm path are integers with |S| < 2^need.  Any set of pairwise
// coprime moduli whose product exceeds 4 * 2^need determines them, so stage 2 runs on moduli that fit ONE
// unsigned byte (256, 251, 243 = 3^5, 241, ...): one int8 tensor-core GEMM carries ~7.6 bits of S, against
// sixteen limb x limb GEMMs per 26.6-bit reference modulus (1.7 bits per GEMM).  The inputs reach that base
// through their binary representation (CRT over the first n_in reference moduli, exact because the
// significands are known to be small), the outputs return to the reference moduli by a CRT base extension.
constexpr int kSmallMax = 54;          // small moduli available (product ~ 2^362)
constexpr int kSmallK = 64;            // K extent of the base-extension MMA: residues, rank, zero padding
constexpr int kSmallNinMax = 16;       // reference moduli the input conversion may read
constexpr int kSmallShiftMax = 400;    // alignment shifts the +-2^s table covers (> log2 of the product)
constexpr int kBinW = 12;              // 32-bit words of a sum rebuilt in binary from the one-byte base (2^384 > 54 M')
static const int kSmallModuli[kSmallMax] = {
    256, 251, 243, 241, 239, 233, 229, 227, 223, 211, 199, 197, 193, 191, 181, 179, 173, 169, 167, 163, 157, 151, 149, 139, 137, 131, 127,
    125, 121, 113, 109, 107, 103, 101, 97,  89,  83,  79,  73,  71,  67,  61,  59,  53,  49,  47,  43,  41,  37,  31,  29,  23,  19,  17};

struct SmallConsts {
    bool usable = false;                 // N % 4 == 0, every reference modulus below 2^27
    int ext_cols = 0;                    // columns of the extension operand: 128 per block of 32 reference moduli
    int red_shift = 0;                   // > 0: all reference moduli have the same bit length k >= 24 (fast reduction of the extension sums)
    std::vector<int> prefix_log2;        // [kSmallMax + 1]  floor(log2(p_0 ... p_{c-1}))
    std::vector<uint8_t> inv;            // [kSmallMax + 1][64]   (M'_c / p_i)^-1 mod p_i, i < c
    std::vector<uint8_t> ext_b;          // [kSmallMax + 1][ext_cols][64]  limbs of (M'_c / p_i) mod m_q (K index i < c) and of
                                         //                                m_q - M'_c mod m_q (K index 63, multiplies the rank)
    std::vector<uint32_t> cw;            // [kSmallMax][16]  byte b of word w = 256^(4 w + b) mod p_j
    std::vector<uint8_t> pws;            // [2 (kSmallShiftMax + 1)][64]  row 2 s: 2^s mod p_j, row 2 s + 1: -2^s mod p_j
    std::vector<uint32_t> in_mi;         // [kSmallNinMax + 1][16][16]  words of (m_0 ... m_{c-1}) / m_i
    std::vector<uint32_t> in_negmp;      // [kSmallNinMax + 1][16]      words of 2^(32 c) - m_0 ... m_{c-1}
    std::vector<int> in_log2_milli;      // [kSmallNinMax + 1]  floor(1024 log2(m_0 ... m_{c-1})) - 1
    std::vector<uint32_t> red_mu;        // [N]  floor(2^(k + 30) / m_q), k = red_shift: one-correction 32-bit Barrett constant (barrett_k in mp_device.cuh)
    // binary reconstruction of a sum from its one-byte residues (kernels_bin.cuh): 32-bit words of M'_c / p_i and of 2^(32 kBinW) - M'_c
    std::vector<uint32_t> bin_mi;        // [kSmallMax + 1][kSmallMax][kBinW]
    std::vector<uint32_t> bin_negmp;     // [kSmallMax + 1][kBinW]
    // binary reconstruction of a number of the format (CRT over all N moduli): words of M / m_i, of 2^(32 full_nw) - M and of M
    int full_nw = 0;                     // words of M + 1
    std::vector<uint32_t> full_mi, full_negm, full_m;   // [N][full_nw], [full_nw], [full_nw]
    double log2M_up = 0;                 // log2(M), rounded up a little
};

// column of the extension operand that holds limb `limb` of reference modulus q (see k_ext_small: a thread of the
// m16n8k32 accumulator fragment owns columns 2t, 2t+1 of an 8-column tile; limbs 0,1 and 2,3 of one modulus sit at the
// same position of two neighbouring tiles)
inline int small_ext_col(int q, int limb) { return (q >> 5) * 128 + ((q & 31) >> 2) * 16 + (limb >> 1) * 8 + (q & 3) * 2 + (limb & 1); }

inline double biguint_log2(const BigUInt &v) {
    const int L = v.bit_length();
    if (L == 0) return 0;
    const int take = L < 53 ? L : 53;
    return std::log2((double) v.top_bits(take)) + (double) (L - take);
}

inline void compute_small_consts(const HostConsts &c, SmallConsts &s) {
    const int N = c.N, P = kSmallMax;
    int mx = 0, kmin = 32, kmax = 0;
    for (int i = 0; i < N; ++i) {
        mx = std::max(mx, c.moduli[i]);
        const int k = 32 - __builtin_clz((unsigned) c.moduli[i]);
        kmin = std::min(kmin, k); kmax = std::max(kmax, k);
    }
    s.usable = (N % 4 == 0) && mx < (1 << 27) && N >= 4;
    if (!s.usable) return;
    s.red_shift = (kmin == kmax && kmin >= 24 && kmin <= 27) ? kmin : 0;
    s.red_mu.assign(N, 0);
    if (s.red_shift)
        for (int i = 0; i < N; ++i) s.red_mu[i] = (uint32_t) ((1ull << (s.red_shift + 30)) / (uint64_t) c.moduli[i]);
    s.ext_cols = ((N + 31) / 32) * 128;
    s.prefix_log2.assign(P + 1, 0);
    {
        BigUInt prod(1);
        for (int cc = 0; cc <= P; ++cc) {
            s.prefix_log2[cc] = prod.bit_length() - 1;
            if (cc < P) prod.mul_small((uint32_t) kSmallModuli[cc]);
        }
    }
    s.inv.assign((size_t) (P + 1) * 64, 0);
    s.ext_b.assign((size_t) (P + 1) * s.ext_cols * 64, 0);
    for (int cc = 1; cc <= P; ++cc) {
        for (int i = 0; i < cc; ++i) {
            const int pi = kSmallModuli[i];
            int64_t prod = 1;
            for (int j = 0; j < cc; ++j) if (j != i) prod = prod * (kSmallModuli[j] % pi) % pi;
            s.inv[(size_t) cc * 64 + i] = (uint8_t) inverse_mod(prod, pi);
        }
        std::vector<int64_t> pre(cc + 1), suf(cc + 1);
        for (int q = 0; q < N; ++q) {
            const int64_t mq = c.moduli[q];
            pre[0] = 1;
            for (int i = 0; i < cc; ++i) pre[i + 1] = pre[i] * (kSmallModuli[i] % mq) % mq;
            suf[cc] = 1;
            for (int i = cc - 1; i >= 0; --i) suf[i] = suf[i + 1] * (kSmallModuli[i] % mq) % mq;
            for (int i = 0; i <= cc; ++i) {
                uint32_t v;
                if (i < cc) v = (uint32_t) (pre[i] * suf[i + 1] % mq);
                else { const int64_t mp = pre[cc]; v = (uint32_t) (mp ? mq - mp : 0); }
                const int kidx = i < cc ? i : kSmallK - 1;   // the rank sits in the last K slot
                for (int limb = 0; limb < 4; ++limb)
                    s.ext_b[((size_t) cc * s.ext_cols + small_ext_col(q, limb)) * 64 + kidx] = (uint8_t) (v >> (8 * limb));
            }
        }
    }
    s.cw.assign((size_t) P * 16, 0);
    for (int j = 0; j < P; ++j) {
        const int p = kSmallModuli[j];
        int v = 1 % p;
        for (int b = 0; b < 64; ++b) {
            s.cw[(size_t) j * 16 + b / 4] |= (uint32_t) v << (8 * (b % 4));
            v = v * 256 % p;
        }
    }
    s.pws.assign((size_t) 2 * (kSmallShiftMax + 1) * 64, 0);
    for (int j = 0; j < P; ++j) {
        const int p = kSmallModuli[j];
        int v = 1 % p;
        for (int sh = 0; sh <= kSmallShiftMax; ++sh) {
            s.pws[(size_t) (2 * sh) * 64 + j] = (uint8_t) v;
            s.pws[(size_t) (2 * sh + 1) * 64 + j] = (uint8_t) (v ? p - v : 0);
            v = v * 2 % p;
        }
    }
    s.in_mi.assign((size_t) (kSmallNinMax + 1) * 16 * 16, 0);
    s.in_negmp.assign((size_t) (kSmallNinMax + 1) * 16, 0);
    s.in_log2_milli.assign(kSmallNinMax + 1, -1);
    for (int cc = 1; cc <= kSmallNinMax && cc < N; ++cc) {
        BigUInt prod(1);
        for (int i = 0; i < cc; ++i) prod.mul_small((uint32_t) c.moduli[i]);
        s.in_log2_milli[cc] = (int) std::floor(1024.0 * biguint_log2(prod)) - 1;
        // 2^(32 cc) - prod, cc words
        {
            uint64_t borrow = 0;
            for (int w = 0; w < cc; ++w) {
                const uint64_t pw = w < (int) prod.limb.size() ? prod.limb[w] : 0;
                const uint64_t sub = pw + borrow;
                const uint32_t r = (uint32_t) (0ull - sub);
                borrow = (sub != 0) ? 1 : 0;
                s.in_negmp[(size_t) cc * 16 + w] = r;
            }
        }
        for (int i = 0; i < cc; ++i) {
            BigUInt mi = prod;
            mi.div_small((uint32_t) c.moduli[i]);
            for (int w = 0; w < cc && w < (int) mi.limb.size(); ++w) s.in_mi[((size_t) cc * 16 + i) * 16 + w] = mi.limb[w];
        }
    }
    {
        BigUInt M(1);
        for (int i = 0; i < N; ++i) M.mul_small((uint32_t) c.moduli[i]);
        s.log2M_up = biguint_log2(M) + 1e-9;
    }
    {
        BigUInt M(1);
        for (int i = 0; i < N; ++i) M.mul_small((uint32_t) c.moduli[i]);
        s.full_nw = (int) M.limb.size() + 1;
        s.full_mi.assign((size_t) N * s.full_nw, 0);
        s.full_negm.assign(s.full_nw, 0);
        s.full_m.assign(s.full_nw, 0);
        for (int w = 0; w < (int) M.limb.size(); ++w) s.full_m[w] = M.limb[w];
        uint64_t borrow = 0;
        for (int w = 0; w < s.full_nw; ++w) {
            const uint64_t sub = (uint64_t) s.full_m[w] + borrow;
            s.full_negm[w] = (uint32_t) (0ull - sub);
            borrow = sub != 0 ? 1 : 0;
        }
        for (int i = 0; i < N; ++i) {
            BigUInt q = M;
            q.div_small((uint32_t) c.moduli[i]);
            for (int w = 0; w < (int) q.limb.size(); ++w) s.full_mi[(size_t) i * s.full_nw + w] = q.limb[w];
        }
    }
    s.bin_mi.assign((size_t) (P + 1) * P * kBinW, 0);
    s.bin_negmp.assign((size_t) (P + 1) * kBinW, 0);
    {
        BigUInt prod(1);
        for (int cc = 1; cc <= P; ++cc) {
            prod.mul_small((uint32_t) kSmallModuli[cc - 1]);
            uint64_t borrow = 0;
            for (int w = 0; w < kBinW; ++w) {
                const uint64_t pw = w < (int) prod.limb.size() ? prod.limb[w] : 0;
                const uint64_t sub = pw + borrow;
                s.bin_negmp[(size_t) cc * kBinW + w] = (uint32_t) (0ull - sub);
                borrow = (sub != 0) ? 1 : 0;
            }
            for (int i = 0; i < cc; ++i) {
                BigUInt mi = prod;
                mi.div_small((uint32_t) kSmallModuli[i]);
                for (int w = 0; w < kBinW && w < (int) mi.limb.size(); ++w) s.bin_mi[((size_t) cc * P + i) * kBinW + w] = mi.limb[w];
            }
        }
    }
}

}  // namespace mpres
