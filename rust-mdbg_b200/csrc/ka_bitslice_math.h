// ka_bitslice_math.h -- arithmetic of the bit-sliced K-A variant (host + device, no warp primitives).
//
// The density test of Read::extract (src/read.rs:183,196-208: keep canonical ntHash <= hash_bound)
// is evaluated for 32 HPC positions per instruction instead of one:
//
//   * a base is the 2-bit code (ASCII >> 1) & 3 (A0 C1 T2 G3); the HPC string of a tile is held as
//     two bit planes a (code bit 0) and b (code bit 1), bit k of word n = HPC position 32n+k;
//   * bit q of H[base] (ntHash seed, crate nthash) is a boolean function of (a, b), i.e. ONE LOP3 on
//     plane words; bit 63 of the forward hash of 32 consecutive windows is the XOR over j < l of such
//     functions of the planes shifted by j (funnel shifts);
//   * the next bit down follows from the rolling identity of ntHash read across bit planes:
//         fh(i+1) = rol(fh(i),1) ^ rol(H[s_i], l) ^ H[s_{i+l}]
//       => F_{b-1}[i] = F_b[i+1] ^ Hbit_{b-l}(s_i) ^ Hbit_b(s_{i+l})                (3 instructions)
//         rh(i+1) = ror(rh(i),1) ^ ror(RC[s_i],1) ^ rol(RC[s_{i+l}], l-1)
//       => R_{b+1}[i] = R_b[i+1] ^ RCbit_{b+1}(s_i) ^ RCbit_{b+1-l}(s_{i+l});
//   * hash <= bound with bound < 2^(64-T) implies that the top T bits of fh or of rh are all zero:
//     the OR of the top T planes of each strand is a superset test (2^-T per strand); survivors are
//     re-evaluated exactly in 64 bits from 4-base tables.
//
// Everything here is plain integer C++ so that tests/model/ can run the very same code on the CPU.
#pragma once
#include <stdint.h>

#include "mdbg_common.cuh"

namespace mdbg {
namespace bs {

MDBG_HD uint32_t popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
// (hi:lo) >> s, low word; s in [0, 31]
MDBG_HD uint32_t fsr(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}
// (hi:lo) << s, high word; s in [0, 31]
MDBG_HD uint32_t fsl(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}
MDBG_HD uint32_t low_mask(uint32_t n) { return n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u); }

// ---- seeds by 2-bit code ------------------------------------------------------------------
// (constexpr functions are __host__ only for nvcc unless marked: called from device code unmarked
// they compile -- inside templates even without a diagnostic -- but to garbage)
#if defined(__CUDACC__)
#define MDBG_HDC __host__ __device__ constexpr
#else
#define MDBG_HDC constexpr
#endif
MDBG_HDC uint64_t seed_fwd(uint32_t code) { return code == 0 ? NT_A : code == 1 ? NT_C : code == 2 ? NT_T : NT_G; }
MDBG_HDC uint64_t seed_rc(uint32_t code) { return code == 0 ? NT_T : code == 1 ? NT_G : code == 2 ? NT_A : NT_C; }
// truth table of bit q (mod 64) of the seed over the code (index = code = b<<1 | a)
MDBG_HDC uint32_t tt_fwd(int q) {
    const int s = ((q % 64) + 64) % 64;
    return (uint32_t)(((seed_fwd(0) >> s) & 1u) | (((seed_fwd(1) >> s) & 1u) << 1) |
                      (((seed_fwd(2) >> s) & 1u) << 2) | (((seed_fwd(3) >> s) & 1u) << 3));
}
MDBG_HDC uint32_t tt_rc(int q) {
    const int s = ((q % 64) + 64) % 64;
    return (uint32_t)(((seed_rc(0) >> s) & 1u) | (((seed_rc(1) >> s) & 1u) << 1) |
                      (((seed_rc(2) >> s) & 1u) << 2) | (((seed_rc(3) >> s) & 1u) << 3));
}
// acc ^ f(a, b) for the boolean function with truth table tt; with tt a compile-time constant
// (unrolled loops) this is one LOP3.
MDBG_HD uint32_t apply_tt(uint32_t tt, uint32_t acc, uint32_t a, uint32_t b) {
    uint32_t f = 0;
    if (tt & 1u) f ^= ~a & ~b;
    if (tt & 2u) f ^= a & ~b;
    if (tt & 4u) f ^= ~a & b;
    if (tt & 8u) f ^= a & b;
    return acc ^ f;
}

// ---- ASCII -> bit planes --------------------------------------------------------------------
// The code bits sit at bit 1 (a) and bit 2 (b) of every byte.  Eight bases per multiply: two words are
// merged nibble-wise (bases 0..3 keep their low nibbles, bases 4..7 move theirs into the high nibbles:
// one shift on the FMA pipe + one LOP3), so a plane's eight bits sit at bits 1 / 5 (2 / 6) of the four
// bytes, and ONE multiply gathers them into the top byte in base order (all partial products land on
// distinct bits: no carries).  Four such bytes are one plane word of 32 bases (three PRMT).
constexpr uint32_t GATHER_A8 = 0x00810204u;  // bit 8j+1 -> 24+j, bit 8j+5 -> 28+j
constexpr uint32_t GATHER_B8 = 0x00408102u;  // bit 8j+2 -> 24+j, bit 8j+6 -> 28+j
MDBG_HD void plane_pair(uint32_t w0, uint32_t w1, uint32_t& pa, uint32_t& pb) {
    const uint32_t n = (w0 & 0x0F0F0F0Fu) | ((w1 << 4) & 0xF0F0F0F0u);
    pa = (n & 0x22222222u) * GATHER_A8;
    pb = (n & 0x44444444u) * GATHER_B8;
}
// word whose byte i is the top byte of x_i
MDBG_HD uint32_t top_bytes(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(__byte_perm(x0, x1, 0x0073), __byte_perm(x2, x3, 0x0073), 0x5410);
#else
    return (x0 >> 24) | ((x1 >> 24) << 8) | ((x2 >> 24) << 16) | (x3 & 0xFF000000u);
#endif
}
// v >> n for n in [1, 32] (32 gives 0)
MDBG_HD uint32_t shr_clamp(uint32_t v, uint32_t n) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_rc(v, 0u, n);
#else
    return n >= 32u ? 0u : (v >> n);
#endif
}

// Alphabet check, word-parallel (same identities as ka_minimizers.cu): a byte is one of A C G T iff
// bits 7,5,3 are 0, bit 6 is 1, b0 != b4 and b4 == (b2 & ~b1); evaluated at bit 4 of every byte
// from left-shifted copies and OR-accumulated.
struct BadAcc { uint32_t k, y, x; };
MDBG_HD void bad_accumulate(BadAcc& b, uint32_t w) {
    uint32_t s4 = w << 4, s3 = w << 3, s2 = w << 2;
    b.k |= w ^ 0x40404040u;
    b.y |= ~(s4 ^ w);
    b.x |= (w ^ (s2 & ~s3));
}
MDBG_HD uint32_t bad_of(const BadAcc& b) { return (b.k & 0xE8E8E8E8u) | ((b.y | b.x) & 0x10101010u); }
MDBG_HD bool is_acgt(uint32_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// ---- compaction (parallel suffix "compress", Hacker's Delight 7-4) of two planes by one mask ----
MDBG_HD void pext_pair(uint32_t m, uint32_t& x, uint32_t& y) {
    x &= m;
    y &= m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2;
        mp ^= mp << 4;
        mp ^= mp << 8;
        mp ^= mp << 16;
        const uint32_t mv = mp & m;
        m = (m ^ mv) | (mv >> (1 << i));
        uint32_t t = x & mv;
        x = (x ^ t) | (t >> (1 << i));
        t = y & mv;
        y = (y ^ t) | (t >> (1 << i));
        mk &= ~mp;
    }
}

// Prefix XOR of a word (bit i = parity of bits 0..i): XOR of the word shifted by 0..31.  Three shifts per level
// ({0,1,2} x {0,3,6} x {0,9,18} x {0,27} covers 0..53): four 3-input LOP3 on the ALU pipe and seven shifts on the
// FMA pipe (IMAD.SHL), one ALU instruction and one level of dependency fewer than the five shift-XOR doublings.
MDBG_HD uint32_t prefix_xor(uint32_t v) {
    v = v ^ (v << 1) ^ (v << 2);
    v = v ^ (v << 3) ^ (v << 6);
    v = v ^ (v << 9) ^ (v << 18);
    return v ^ (v << 27);
}

// The same network in two parts: rounds 0..3 (moves by 1, 2, 4, 8), then round 4 (moves by 16), which does
// nothing unless some base has 16 or more dropped bases below it -- `mk` says so (callers vote on it).
struct PextState { uint32_t m, mk; };
MDBG_HD PextState pext_pair_rounds4(uint32_t m, uint32_t& x, uint32_t& y) {
    x &= m;
    y &= m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t mp = prefix_xor(mk);
        const uint32_t mv = mp & m;
        m = (m ^ mv) | (mv >> (1 << i));
        uint32_t t = x & mv;
        x = (x ^ t) | (t >> (1 << i));
        t = y & mv;
        y = (y ^ t) | (t >> (1 << i));
        mk &= ~mp;
    }
    return PextState{m, mk};
}
MDBG_HD void pext_pair_round5(const PextState& st, uint32_t& x, uint32_t& y) {
    const uint32_t mv = prefix_xor(st.mk) & st.m;
    uint32_t t = x & mv;
    x = (x ^ t) | (t >> 16);
    t = y & mv;
    y = (y ^ t) | (t >> 16);
}

// position of the k-th (0-based) set bit of m; k < popc(m)
MDBG_HD uint32_t select_bit(uint32_t m, uint32_t k) {
    uint32_t pos = 0, c;
    c = popc32(m & 0xFFFFu); if (k >= c) { k -= c; m >>= 16; pos += 16; }
    c = popc32(m & 0xFFu);   if (k >= c) { k -= c; m >>= 8;  pos += 8; }
    c = popc32(m & 0xFu);    if (k >= c) { k -= c; m >>= 4;  pos += 4; }
    c = popc32(m & 0x3u);    if (k >= c) { k -= c; m >>= 2;  pos += 2; }
    c = m & 1u;              if (k >= c) { pos += 1; }
    return pos;
}

// ---- the filter: 32 HPC positions [s, s+32) in (a0, b0), the next 32 in (a1, b1) ---------------
// Returns a word whose bit k (k <= 32-T) is set when the top T bits of fh(s+k) or of rh(s+k) are all
// zero (bits above 32-T are undefined).  Positions past the end of the data read as code 0: the
// identities above hold for any continuation, the caller drops windows that leave the read.
template <int L, int T>
MDBG_HD uint32_t filter_window(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    uint32_t F = 0, R = 0;       // F_63 and R_{64-T}, built directly from their l terms
#pragma unroll
    for (int j = 0; j < L; j++) {
        const uint32_t aj = fsr(a0, a1, j), bj = fsr(b0, b1, j);
        F = apply_tt(tt_fwd(63 - (L - 1) + j), F, aj, bj);
        R = apply_tt(tt_rc((64 - T) - j), R, aj, bj);
    }
    const uint32_t aL = fsr(a0, a1, L), bL = fsr(b0, b1, L);
    uint32_t orF = F, orR = R;
#pragma unroll
    for (int t = 1; t < T; t++) {
        const int b = 64 - t;          // F_b -> F_{b-1}
        F >>= 1;
        F = apply_tt(tt_fwd(b - L), F, a0, b0);
        F = apply_tt(tt_fwd(b), F, aL, bL);
        orF |= F;
        const int c = 64 - T + t;      // R_{c-1} -> R_c
        R >>= 1;
        R = apply_tt(tt_rc(c), R, a0, b0);
        R = apply_tt(tt_rc(c - L), R, aL, bL);
        orR |= R;
    }
    return ~(orF & orR);
}

// The same filter on 64 positions [s, s+64): planes in (a0, a1, a2), (b0, b1, b2).  The low half takes its
// chain look-ahead from the high half (funnel shift) and stays valid on all 32 positions; the high half loses
// T-1 positions as above: 64-(T-1) valid positions per call instead of 2 x (32-(T-1)).
template <int L, int T>
MDBG_HD void filter_window64(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b0, uint32_t b1, uint32_t b2,
                             uint32_t& cand_lo, uint32_t& cand_hi) {
    uint32_t F0 = 0, R0 = 0, F1 = 0, R1 = 0;
#pragma unroll
    for (int j = 0; j < L; j++) {
        const uint32_t aj0 = fsr(a0, a1, j), bj0 = fsr(b0, b1, j), aj1 = fsr(a1, a2, j), bj1 = fsr(b1, b2, j);
        F0 = apply_tt(tt_fwd(63 - (L - 1) + j), F0, aj0, bj0);
        R0 = apply_tt(tt_rc((64 - T) - j), R0, aj0, bj0);
        F1 = apply_tt(tt_fwd(63 - (L - 1) + j), F1, aj1, bj1);
        R1 = apply_tt(tt_rc((64 - T) - j), R1, aj1, bj1);
    }
    const uint32_t aL0 = fsr(a0, a1, L), bL0 = fsr(b0, b1, L), aL1 = fsr(a1, a2, L), bL1 = fsr(b1, b2, L);
    uint32_t orF0 = F0, orR0 = R0, orF1 = F1, orR1 = R1;
#pragma unroll
    for (int t = 1; t < T; t++) {
        const int b = 64 - t, c = 64 - T + t;
        F0 = fsr(F0, F1, 1);
        F1 >>= 1;
        F0 = apply_tt(tt_fwd(b), apply_tt(tt_fwd(b - L), F0, a0, b0), aL0, bL0);
        F1 = apply_tt(tt_fwd(b), apply_tt(tt_fwd(b - L), F1, a1, b1), aL1, bL1);
        orF0 |= F0; orF1 |= F1;
        R0 = fsr(R0, R1, 1);
        R1 >>= 1;
        R0 = apply_tt(tt_rc(c - L), apply_tt(tt_rc(c), R0, a0, b0), aL0, bL0);
        R1 = apply_tt(tt_rc(c - L), apply_tt(tt_rc(c), R1, a1, b1), aL1, bL1);
        orR0 |= R0; orR1 |= R1;
    }
    cand_lo = ~(orF0 & orR0);
    cand_hi = ~(orF1 & orR1);
}

// ---- exact canonical hash of one window from its l codes (a = code bits 0, b = code bits 1) ----
// Four bases at a time: t4[(b4 << 4) | a4] = { XOR_j rol(H[c_j], 3-j), XOR_j rol(RC[c_j], j) }.
struct T4Entry { uint64_t f, r; };
MDBG_HD T4Entry t4_make(uint32_t idx) {
    T4Entry e{0, 0};
    for (uint32_t j = 0; j < 4; j++) {
        const uint32_t code = ((idx >> j) & 1u) | (((idx >> (4 + j)) & 1u) << 1);
        e.f ^= rol64(nt_fwd_code(code), 3 - j);
        e.r ^= rol64(nt_rc_code(code), j);
    }
    return e;
}
template <int L>
MDBG_HD uint64_t exact_hash(uint32_t av, uint32_t bv, const T4Entry* t4) {
    uint64_t fh = 0, rh = 0;
    constexpr int G = L / 4;
#pragma unroll
    for (int g = 0; g < G; g++) {
        const uint32_t idx = ((av >> (4 * g)) & 15u) | (((bv >> (4 * g)) & 15u) << 4);
        const T4Entry e = t4[idx];
        fh = rol64(fh, 4) ^ e.f;
        rh ^= rol64(e.r, 4 * g);
    }
#pragma unroll
    for (int j = 4 * G; j < L; j++) {
        const uint32_t code = ((av >> j) & 1u) | (((bv >> j) & 1u) << 1);
        fh = rol64(fh, 1) ^ nt_fwd_code(code);
        rh ^= rol64(nt_rc_code(code), j);
    }
    return fh < rh ? fh : rh;
}

// Which (l, bound) the bit-sliced kernel is built for: l in {10, 12, 14} (instantiated), and at
// least T = 8 leading zero bits in the bound (density < 2^-8).
inline bool supported(uint32_t l, uint64_t bound) {
    return (l == 10 || l == 12 || l == 14) && (bound >> 56) == 0;
}

}  // namespace bs
}  // namespace mdbg
