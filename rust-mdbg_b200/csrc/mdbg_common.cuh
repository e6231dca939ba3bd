// mdbg_common.cuh -- arithmetic shared by host and device code of libmdbg_b200.
//
// ntHash-1 as used by the reference (crate nthash, called at src/read.rs:196): canonical
// 64-bit hash with plain rotates.  fh(i) = XOR_j rol(H[s[i+j]], l-1-j),
// rh(i) = XOR_j rol(RC[s[i+j]], j), canonical = min(fh, rh); 'N' hashes as 0.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MDBG_HD __host__ __device__ __forceinline__
#else
#define MDBG_HD inline
#endif

namespace mdbg {

constexpr uint64_t NT_A = 0x3c8bfbb395c60474ULL;
constexpr uint64_t NT_C = 0x3193c18562a02b4cULL;
constexpr uint64_t NT_G = 0x20323ed082572324ULL;
constexpr uint64_t NT_T = 0x295549f54be24456ULL;

MDBG_HD uint64_t rol64(uint64_t x, uint32_t r) {
    r &= 63u;
    return r ? (x << r) | (x >> (64u - r)) : x;
}

// 3-bit class of an ASCII base: (c >> 1) & 7 maps A->0 C->1 T->2 G->3 N->7; the low two bits
// are the 2-bit code used by the rolling filter.
MDBG_HD uint32_t base_class(uint32_t c) { return (c >> 1) & 7u; }

// byte -> forward seed value; ok=false for anything but ACGTN (the crate panics there).
MDBG_HD uint64_t nt_fwd(uint32_t c, bool& ok) {
    ok = true;
    switch (c) {
        case 'A': return NT_A;
        case 'C': return NT_C;
        case 'G': return NT_G;
        case 'T': return NT_T;
        case 'N': return 0;
        default: ok = false; return 0;
    }
}
MDBG_HD uint64_t nt_rc(uint32_t c) {
    switch (c) {
        case 'A': return NT_T;
        case 'C': return NT_G;
        case 'G': return NT_C;
        case 'T': return NT_A;
        default: return 0;
    }
}
// 2-bit code (A0 C1 T2 G3) -> seed values
MDBG_HD uint64_t nt_fwd_code(uint32_t code) {
    return code == 0 ? NT_A : code == 1 ? NT_C : code == 2 ? NT_T : NT_G;
}
MDBG_HD uint64_t nt_rc_code(uint32_t code) {
    return code == 0 ? NT_T : code == 1 ? NT_G : code == 2 ? NT_A : NT_C;
}

// (density as f64 * u64::MAX as f64) as u64 -- src/read.rs:183.  `u64::MAX as f64` is 2^64;
// Rust's float->int cast saturates and maps NaN to 0.  Host only (never recomputed on device).
inline uint64_t hash_bound(double density) {
    double x = density * 18446744073709551616.0;
    if (!(x == x) || x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}

// Fingerprint of a canonical tuple: order-sensitive 64-bit hash, ONE multiply per element (the elements are
// ntHash values, already well mixed) and a murmur-style finish.  Used only to place a tuple (table slot, owner
// rank, sort key); identity is always decided by comparing the tuples themselves, and a collision only costs a
// retry with another seed (the seed enters before the first multiply, so collisions do not persist across seeds).
MDBG_HD uint64_t fp_mix(uint64_t h, uint64_t v) {
    h ^= v;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 32;
    return h;
}
MDBG_HD uint64_t fp_fin(uint64_t h) {
    h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 29;
    h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 32;
    return h;
}
MDBG_HD uint64_t fp_init(uint64_t seed, uint32_t k) { return seed ^ (0x9e3779b97f4a7c15ULL * (k + 1)); }

// ------------------------------------------------------------------------------------------
// Rolling 32-bit FILTER for the density threshold (K-A fast path).
//
// With h'(b) = H[b] >> 32 and rc'(b) = RC[b] >> 32 define, over the HPC string s,
//     F(i) = XOR_j  h'(s[i+j]) << (l-1-j)        (mod 2^32)
//     G(i) = XOR_j rc'(s[i+j]) >> (l-1-j)
// Shifts compose exactly, so both roll with one shift and one XOR per base:
//     F(i+1) = (F(i) << 1) ^ (h'(out) << l) ^ h'(in)
//     G(i+1) = (G(i) >> 1) ^ (rc'(out) >> l) ^ rc'(in)
// and  bits [31 : l-1] of F(i) == bits [63 : 32+l-1] of fh(i),
//      bits [32-l : 0] of G(i) == bits [63 : 32+l-1] of rh(i)          (E = 33-l exact bits).
// hash <= bound implies top_E(hash) <= top_E(bound), so
//     (F <= f_thresh) || ((G & g_zero) == 0)
// (g_zero = the exact bits of G above the bit length of top_E(bound): one LOP3 with a predicate
// result instead of mask + compare) is a superset test passing ~0.7 % of the windows at
// d = 0.003; survivors are re-evaluated exactly in 64 bits.  Valid for 4 <= l <= 14 (the 2-bit
// history register holds l+1 codes, pre-scaled by 4).
// The kernel keeps G BIT-REVERSED (so that it, too, rolls with a left shift, which the compiler
// issues on the otherwise idle FMA pipe as IMAD.SHL) and the 2-bit history pre-scaled by 4 (the
// table index (out<<2|in) then needs no scaling: tables are two arrays of 16 32-bit words).
struct FilterConsts {
    uint32_t tab_f[16];    // [out<<2 | in] -> TF
    uint32_t tab_g[16];    // [out<<2 | in] -> bitrev(TG)
    uint32_t f_init, g_init;            // g_init bit-reversed
    uint32_t f_thresh, g_mask, g_thresh, g_zero;   // g_zero bit-reversed; g_mask/g_thresh plain (tests)
    uint32_t hist_shift;   // 2l-2
    uint32_t usable;       // 0 => use the exact (dense) path
};

inline uint32_t brev32(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}

inline FilterConsts make_filter(uint32_t l, uint64_t bound) {
    FilterConsts fc{};
    fc.usable = 0;
    if (l < 4 || l > 14) return fc;   // history = l+1 codes, pre-scaled by 4: 2l+4 <= 32
    uint32_t bh = (uint32_t)(bound >> 32);
    // the filter only pays when few windows pass: require >= 5 leading zero bits (d < 1/32)
    if (bh >= (1u << 27)) return fc;
    for (uint32_t o = 0; o < 4; o++)
        for (uint32_t i = 0; i < 4; i++) {
            uint32_t ho = (uint32_t)(nt_fwd_code(o) >> 32), hi = (uint32_t)(nt_fwd_code(i) >> 32);
            uint32_t ro = (uint32_t)(nt_rc_code(o) >> 32), ri = (uint32_t)(nt_rc_code(i) >> 32);
            fc.tab_f[(o << 2) | i] = (ho << l) ^ hi;
            fc.tab_g[(o << 2) | i] = brev32((ro >> l) ^ ri);
        }
    uint32_t ha = (uint32_t)(nt_fwd_code(0) >> 32), ra = (uint32_t)(nt_rc_code(0) >> 32);
    fc.f_init = fc.g_init = 0;
    for (uint32_t j = 0; j < l; j++) {   // l phantom 'A's (history register starts at 0)
        fc.f_init ^= ha << (l - 1 - j);
        fc.g_init ^= ra >> (l - 1 - j);
    }
    uint32_t low = (1u << (l - 1)) - 1u;
    fc.f_thresh = bh | low;
    fc.g_mask = 0xffffffffu >> (l - 1);
    fc.g_thresh = bh >> (l - 1);
    uint32_t bl = 0;                         // bit length of top_E(bound)
    while (bl < 32 && (fc.g_thresh >> bl) != 0) bl++;
    fc.g_zero = brev32(fc.g_mask & ~((bl >= 32) ? 0xffffffffu : ((1u << bl) - 1u)));
    fc.g_init = brev32(fc.g_init);
    fc.hist_shift = 2 * l - 2;
    fc.usable = 1;
    return fc;
}

}  // namespace mdbg
