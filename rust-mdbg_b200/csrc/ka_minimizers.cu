// ka_minimizers.cu -- K-A: homopolymer compression + ntHash + density threshold, sm_100a.
//
// Replaces, for a whole batch of reads at once (reference file:line):
//   Read::encode_rle            src/read.rs:157-174   (run starts, raw position of each run)
//   nthash::NtHashIterator      crate, src/read.rs:196 (canonical ntHash of every HPC l-mer)
//   filter(<= hash_bound)+push  src/read.rs:183,196-208
//
// Layout.  The batch is ONE byte array (ASCII bases of all reads, concatenated) plus
// read_off[R+1].  The array is cut into fixed 4 KiB tiles on absolute offsets, so the grid
// does not depend on read lengths; every WARP of a persistent CTA claims tiles from an atomic
// counter and owns a private slice of shared memory, so warps never wait for each other (the
// only CTA-wide barrier is before the loop).  A tile is staged with coalesced 128-bit loads
// into rows of 128 bytes padded to 144 (conflict-free 128-bit reads with one row per thread).
// Each thread then walks ITS 128 bytes sequentially, entirely in registers:
//   * run start  = byte differs from its predecessor (no compaction pass, no position array:
//     the raw position of a run IS the loop counter);
//   * the density test uses the 32-bit rolling filter of mdbg_common.cuh (2 shifts, 2 xors,
//     one 8-byte table read from a 128-byte, conflict-free shared table per run) -- only
//     ~0.7 % of the windows survive it;
//   * a thread owns the l-mers whose FIRST run starts in its segment and keeps walking past
//     its end until l-1 more runs are consumed (data-dependent halo, no fixed bound);
//   * survivors go to a shared queue and are re-evaluated exactly (64-bit, N-aware, walking
//     the runs backwards/forwards) by the whole CTA afterwards; hits set a bit per tile byte.
// A block-wide popcount scan of that bitmap orders the tile's minimizers; the tile reserves a
// slice of a staging array with one atomicAdd (tiles never wait for each other) and a small
// finalize kernel scans the per-tile counts and moves every slice to its final position in the
// global, (read, position)-ordered output; the per-read offsets fall out of the same prefix.
// Anything the filter cannot represent (N or illegal bytes in the segment, l > 15, large
// densities, queue overflow on low-complexity sequence) takes the exact per-position path in
// the same kernel -- never the CPU.
//
// Roofline: HBM.  Algorithmic bytes = 1 B read per base + 12 B written per minimizer
// (8 B hash + 4 B position), ~1.04 B/base at d = 0.003.  The kernel is issue-bound well below
// the HBM line (~15 integer instructions per base); see DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdbg_common.cuh"
#include "mdbg_kernels.h"

namespace mdbg {

namespace {

constexpr int NT = 32;                 // threads per tile: one warp owns one tile, warps never wait
constexpr int NWARP = KA_THREADS / 32;  // for each other (no CTA barrier inside the tile loop)
constexpr int SEG = KA_SEG;
constexpr int TILE = KA_TILE;
constexpr int PRE = 128;
constexpr int HALO = 256;
constexpr int ROWS = (PRE + TILE + HALO) / 128;  // 35
constexpr int RSTRIDE = 144;
constexpr int WIN = ROWS * 128;
#ifndef MDBG_KA_QCAP
#define MDBG_KA_QCAP 256
#endif
#ifndef MDBG_KA_MIN_BLOCKS
#define MDBG_KA_MIN_BLOCKS 5
#endif
constexpr int QCAP = MDBG_KA_QCAP;
constexpr int WORDS = TILE / 32;  // 128

constexpr uint32_t Q_DROP = 0xFFFFFFFFu;
constexpr uint32_t Q_VERIFIED = 0x80000000u;
constexpr uint64_t VALID_BY_CLASS = 0x4E00000047544341ull;  // "ACTG\0\0\0N": the byte a class must equal

struct __align__(16) Smem {   // one per warp
    uint8_t raw[ROWS * RSTRIDE];
    uint64_t hq[QCAP];        // exact hash of a verified queue entry
    uint32_t queue[QCAP];     // candidates: window-relative position of the LAST run; after phase B:
                              // tile-relative start position of a verified minimizer, or Q_DROP
    uint32_t bitmap[WORDS];
    uint32_t prefix[WORDS];
    uint32_t qn;
    uint32_t total;
    uint32_t tile;
    unsigned long long base;
};
struct __align__(128) CtaSmem {
    __align__(128) uint32_t tab[32];   // [0,16): TF, [16,32): bit-reversed TG
    uint64_t hfw[8], hrc[8];           // ntHash seeds by base class (c>>1)&7: A0 C1 T2 G3 N7
    Smem w[NWARP];
};

struct Win {  // byte access through the staged window, falling back to global memory
    const uint8_t* raw;
    const uint8_t* g;
    int64_t w0;
    const uint64_t* hfw;
    const uint64_t* hrc;
    __device__ __forceinline__ uint32_t byte(int64_t p) const {
        int64_t x = p - w0;
        if (x >= 0 && x < WIN) return raw[(x >> 7) * RSTRIDE + (x & 127)];
        return g[p];
    }
};

__device__ __forceinline__ bool collapsible(uint32_t c) {  // "ACTGactgNn", read.rs:163
    uint32_t u = c & 0xDFu;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N';
}

// Exact canonical hash of the l-mer whose first run starts at p0 (read ends at re).
// fh is rolled left, rh right (one fixed rotate per run); rh gets its l-1 rotation at the end:
//   XOR_j rol(H_j, l-1-j)  and  rol(XOR_j ror(RC_j, l-1-j), l-1) = XOR_j rol(RC_j, j).
// 0: fewer than l runs remain; 1: ok; 2: ok but an illegal byte was hashed (inv = its position).
template <bool HPC>
__device__ int lmer_hash(const Win& W, int64_t p0, int64_t re, uint32_t l, uint64_t& h, int64_t& inv) {
    uint64_t fh = 0, rh = 0;
    int64_t p = p0;
    int res = 1;
    for (uint32_t j = 0; j < l; j++) {
        if (p >= re) return 0;
        uint32_t c = W.byte(p);
        uint32_t cls = (c >> 1) & 7u;
        bool valid = (uint32_t)((VALID_BY_CLASS >> (8 * cls)) & 0xFFu) == c;
        if (!valid && res == 1) { res = 2; inv = p; }
        uint64_t hv = valid ? W.hfw[cls] : 0, rv = valid ? W.hrc[cls] : 0;
        fh = ((fh << 1) | (fh >> 63)) ^ hv;
        rh = ((rh >> 1) | (rh << 63)) ^ rv;
        int64_t q = p + 1;
        if (HPC && (valid || collapsible(c)))
            while (q < re && W.byte(q) == c) q++;
        p = q;
    }
    rh = rol64(rh, l - 1);
    h = fh < rh ? fh : rh;
    return res;
}

// From the run starting at p_end walk l-1 runs back (not past rs).
template <bool HPC>
__device__ bool walk_back(const Win& W, int64_t p_end, int64_t rs, uint32_t l, int64_t& p0) {
    int64_t p = p_end;
    for (uint32_t j = 1; j < l; j++) {
        if (p <= rs) return false;
        int64_t q = p - 1;
        uint32_t c = W.byte(q);
        if (HPC && collapsible(c))
            while (q > rs && W.byte(q - 1) == c) q--;
        p = q;
    }
    p0 = p;
    return true;
}

// Fast exact evaluation of a filter survivor, entirely from the staged window: walk the runs
// BACKWARDS from the last run (at window-relative byte `rel_end`) and fold both strands on the way:
//   A <- ror(A,1) ^ H[c]   gives  fh = rol(A, l-1);      B <- rol(B,1) ^ RC[c]  gives  rh = B.
// Returns false when the window is not plain ACGT inside the staged bytes (N, illegal byte, read
// start or window edge within reach, homopolymer longer than the guard): the caller then takes the
// general path.
template <bool HPC>
__device__ __forceinline__ bool verify_fast(const uint8_t* raw, const uint64_t* hfw, const uint64_t* hrc,
                                            int rel_end, int64_t rel_rs, uint32_t l, int& rel_p0, uint64_t& h) {
    constexpr int GUARD = 96;
    if (rel_end < GUARD || (int64_t)rel_end - rel_rs < GUARD || rel_end >= WIN) return false;
    uint64_t A = 0, B = 0;
    int r = rel_end;
    uint32_t c = raw[r + (r >> 7) * 16];
    for (uint32_t need = l;;) {
        uint32_t cls = (c >> 1) & 7u;
        if (cls > 3u || (uint32_t)((VALID_BY_CLASS >> (8 * cls)) & 0xFFu) != c) return false;
        A = ((A >> 1) | (A << 63)) ^ hfw[cls];
        B = ((B << 1) | (B >> 63)) ^ hrc[cls];
        if (--need == 0) break;
        r--;
        c = raw[r + (r >> 7) * 16];
        if (HPC) {
            for (;;) {
                int q = r - 1;
                if (raw[q + (q >> 7) * 16] != c) break;
                r = q;
            }
        }
        if (rel_end - r > GUARD - 2) return false;
    }
    rel_p0 = r;
    uint64_t fh = rol64(A, l - 1);
    h = fh < B ? fh : B;
    return true;
}

// Alphabet check, word-parallel.  A byte is one of A C G T  iff  bits 7,5,3 are 0, bit 6 is 1,
// b0 != b4 and b4 == (b2 & ~b1).  The three terms are evaluated at bit 4 of every byte from LEFT
// shifted copies of the word (left shifts issue on the FMA pipe as IMAD.SHL; the ALU pipe is the
// bottleneck of this kernel) and OR-accumulated over a chunk; bad_of() folds them once at the end.
struct BadAcc { uint32_t k, y, x; };
__device__ __forceinline__ void bad_accumulate(BadAcc& b, uint32_t w) {
    uint32_t s4 = w << 4, s3 = w << 3, s2 = w << 2;
    b.k |= w ^ 0x40404040u;            // bits 7,5,3 (and 6) of any byte off the pattern
    b.y |= ~(s4 ^ w);                  // at bit 4: b0 == b4
    b.x |= (w ^ (s2 & ~s3));           // at bit 4: b4 != (b2 & ~b1)
}
__device__ __forceinline__ uint32_t bad_of(const BadAcc& b) {
    return (b.k & 0xE8E8E8E8u) | ((b.y | b.x) & 0x10101010u);
}
// per-byte form: byte j of the result is non-zero iff byte j of w is not A/C/G/T
__device__ __forceinline__ uint32_t badword(uint32_t w) {
    BadAcc b{0, 0, 0};
    bad_accumulate(b, w);
    return bad_of(b);
}

// last r in [lo, hi) with read_off[r] <= p   (read_off[lo] <= p guaranteed)
__device__ __forceinline__ uint64_t find_read(const uint64_t* __restrict__ read_off, uint64_t lo,
                                              uint64_t hi, int64_t p) {
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(read_off + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace

// tile_lb[i] = first r in [0, R] with read_off[r] >= i*TILE; tile_lb[n_tiles] = R+1.
// The same launch resets the batch's device counters (a kernel, not cudaMemsetAsync / an H2D copy:
// those may be served by the copy engine and would queue behind the bulk upload that K-A is
// supposed to overlap).
__global__ void ka_tile_lb_kernel(const uint64_t* __restrict__ read_off, uint64_t R, uint64_t n_tiles,
                                  uint64_t* __restrict__ tile_lb, KAInit I) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i == 0) {
        *I.total_out = 0; *I.err_pos = ~0ull; *I.stage_counter = 0; *I.dense_tiles = 0;
        for (uint32_t j = 0; j < I.n_counters; j++) I.counters[j] = 0;
        if (I.zero64) *I.zero64 = 0;
    }
    if (i > n_tiles) return;
    if (i == n_tiles) { tile_lb[i] = R + 1; return; }
    uint64_t target = i * (uint64_t)TILE;
    uint64_t lo = 0, hi = R + 1;  // lower_bound over R+1 entries
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (read_off[mid] < target) lo = mid + 1; else hi = mid;
    }
    tile_lb[i] = lo;
}

template <bool HPC>
__global__ void __launch_bounds__(KA_THREADS, MDBG_KA_MIN_BLOCKS) ka_minimizers_kernel(const KAArgs A) {
    __shared__ CtaSmem cs;
    const int ctid = threadIdx.x;
    const int tid = ctid & 31, lane = tid;      // position inside the warp's tile
    Smem& sm = cs.w[ctid >> 5];
    const uint32_t l = A.l;
    const uint64_t bound = A.bound;
    if (ctid < 16) { cs.tab[ctid] = A.fc.tab_f[ctid]; cs.tab[16 + ctid] = A.fc.tab_g[ctid]; }
    if (ctid < 8) {
        cs.hfw[ctid] = ctid < 4 ? nt_fwd_code(ctid) : 0;   // classes 4..7 (illegal bytes and N) hash as 0
        cs.hrc[ctid] = ctid < 4 ? nt_rc_code(ctid) : 0;
    }
    __syncthreads();   // the only CTA-wide barrier: tables are read-only from here on
    const bool use_filter = A.fc.usable && !A.force_dense;
    const uint32_t SH = A.fc.hist_shift, fth = A.fc.f_thresh, gz = A.fc.g_zero;
    const uint8_t* __restrict__ gb = A.bases;
    const int64_t B = (int64_t)A.n_bases;

    for (;;) {
        __syncwarp();  // previous tile fully done with this warp's shared memory
        if (tid == 0) {
            sm.tile = atomicAdd(A.tile_counter, 1u);
            sm.qn = 0;
        }
        for (int i = tid; i < WORDS; i += NT) sm.bitmap[i] = 0;
        __syncwarp();
        uint64_t tile;
        if (A.tile_list) {             // list mode: the tiles the bit-sliced variant handed over
            if (sm.tile >= *A.tile_list_n) break;
            tile = A.tile_list[sm.tile];
        } else {
            tile = A.tile_begin + sm.tile;
            if (tile >= A.tile_end) break;
        }
        const int64_t t0 = (int64_t)tile * TILE;
        const int64_t t1 = (t0 + TILE < B) ? t0 + TILE : B;
        const int64_t w0 = t0 - PRE;
        const uint64_t lb = __ldg(A.tile_lb + tile), lbn = __ldg(A.tile_lb + tile + 1);
        const uint64_t rlo = lb > 0 ? lb - 1 : 0;
        const uint64_t rhi = lbn < A.n_reads ? lbn : A.n_reads;  // exclusive

        // ---- stage the window -----------------------------------------------------------
        for (int q = tid; q < ROWS * 8; q += NT) {
            int64_t gp = w0 + (int64_t)q * 16;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gp >= 0 && gp + 16 <= B) {
                v = __ldg(reinterpret_cast<const uint4*>(gb + gp));
            } else if (gp + 16 > 0 && gp < B) {
                uint8_t tmp[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    int64_t x = gp + j;
                    tmp[j] = (x >= 0 && x < B) ? gb[x] : (uint8_t)0;
                }
                v.x = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((uint32_t)tmp[3] << 24);
                v.y = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((uint32_t)tmp[7] << 24);
                v.z = tmp[8] | (tmp[9] << 8) | (tmp[10] << 16) | ((uint32_t)tmp[11] << 24);
                v.w = tmp[12] | (tmp[13] << 8) | (tmp[14] << 16) | ((uint32_t)tmp[15] << 24);
            }
            *reinterpret_cast<uint4*>(sm.raw + (q >> 3) * RSTRIDE + (q & 7) * 16) = v;
        }
        __syncwarp();
        Win W{sm.raw, gb, w0, cs.hfw, cs.hrc};
        const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(cs.tab);
        const uint32_t raw_s = (uint32_t)__cvta_generic_to_shared(sm.raw);

        const int64_t a = t0 + (int64_t)tid * SEG;
        const int64_t seg_end = (a + SEG < t1) ? a + SEG : t1;

        // Exact evaluation of every run start of [pb, plim) (slow path).  QUEUE: also leave a
        // verified queue entry so that the emit phase need not recompute the hash.
        auto exact_portion_impl = [&](int64_t rs, int64_t re, int64_t pb, int64_t plim, bool to_queue) {
            uint32_t prev = (pb > rs) ? W.byte(pb - 1) : 0x100u;
            for (int64_t p = pb; p < plim; p++) {
                uint32_t c = W.byte(p);
                bool nr = !HPC || c != prev || !collapsible(c);
                prev = c;
                if (!nr) continue;
                uint64_t h;
                int64_t inv = 0;
                int res = lmer_hash<HPC>(W, p, re, l, h, inv);
                if (res == 0) break;  // no later start of this read can complete either
                if (res == 2) atomicMin(A.err_pos, (unsigned long long)inv);
                if (h <= bound) {
                    uint32_t rel = (uint32_t)(p - t0), bit = 1u << (rel & 31);
                    uint32_t old = atomicOr(&sm.bitmap[rel >> 5], bit);
                    if (to_queue && !(old & bit)) {
                        uint32_t qi = atomicAdd(&sm.qn, 1u);
                        if (qi < QCAP) { sm.queue[qi] = rel | Q_VERIFIED; sm.hq[qi] = h; }
                    }
                }
            }
        };
        auto exact_portion = [&](int64_t rs, int64_t re, int64_t pb, int64_t plim) {
            exact_portion_impl(rs, re, pb, plim, false);
        };

        // Filtered scan of one portion; falls back to exact_portion if a non-ACGT byte shows up.
        auto fast_portion = [&](int64_t rs, int64_t re, int64_t pb, int64_t plim) {
            uint32_t F = A.fc.f_init, G = A.fc.g_init, hist = 0, bad = 0;
            auto push = [&](int64_t pos) {
                uint32_t qi = atomicAdd(&sm.qn, 1u);
                if (qi < QCAP) sm.queue[qi] = (uint32_t)(pos - w0);
            };
            // one run (C++ form, used on ragged portions): roll the 2-bit history and the two filter
            // words; true if the window ending here may be a minimizer
            auto run_step = [&](uint32_t code) -> bool {   // hist pre-scaled by 4, G bit-reversed
                hist = (hist << 2) | (code << 2);
                uint32_t idx = ((hist >> SH) & 0x30u) | (code << 2);
                F = (F << 1) ^ cs.tab[idx >> 2];
                G = (G << 1) ^ cs.tab[16 + (idx >> 2)];
                return F <= fth || (G & gz) == 0;
            };
            // The same step for byte J of a word, hand-scheduled in PTX: every instruction is
            // predicated on "byte J starts a run" (p) -- no branch per byte, no select chains.
            //   x: word ^ (word shifted by one byte)  cw: 2-bit codes, one per byte
            // TAIL adds the data-dependent halo conditions (runs left to consume, bytes left in
            // the read) and only then looks at the per-byte "not ACGT" flags in bw.
            // ALU-pipe budget: the kernel is bound by the 16-lane ALU pipe (LOP3/SHF/PRMT/ISETP), so
            // whatever can run on the FMA pipe does: history update (IMAD), both rolls (IMAD.SHL, G is
            // kept bit-reversed for that).  A survivor is flagged whenever the CURRENT state passes,
            // also on the non-run bytes that follow a passing run: phase B drops those (one byte
            // compare) -- cheaper than and-ing the run predicate into every step.
#define MDBG_STEP(J, BITN)                                                                                  \
    asm("{\n\t.reg .pred p, q;\n\t.reg .b32 t, c, a, b, tx, ty;\n\t"                                        \
        "and.b32 t, %4, %10;\n\tsetp.ne.u32 p, t, 0;\n\t"                                                   \
        "prmt.b32 c, %5, 0, %11;\n\t"                                                                       \
        "@p mad.lo.u32 %0, %0, 4, c;\n\t"                                                                   \
        "shr.u32 t, %0, %6;\n\t"                                                                            \
        "lop3.b32 t, t, 48, c, 0xEA;\n\t"                                                                   \
        "add.u32 t, t, %7;\n\t"                                                                             \
        "ld.shared.u32 tx, [t];\n\tld.shared.u32 ty, [t+64];\n\t"                                           \
        "shl.b32 a, %1, 1;\n\t@p xor.b32 %1, a, tx;\n\t"                                                    \
        "shl.b32 b, %2, 1;\n\t@p xor.b32 %2, b, ty;\n\t"                                                    \
        "and.b32 a, %2, %9;\n\tsetp.eq.u32 q, a, 0;\n\t"                                                    \
        "setp.le.or.u32 q, %1, %8, q;\n\t"                                                                  \
        "@q or.b32 %3, %3, %12;\n\t}"                                                                       \
        : "+r"(hist), "+r"(F), "+r"(G), "+r"(cand)                                                          \
        : "r"(x), "r"(cw), "r"(SHr), "r"(tab_r), "r"(fthr), "r"(gzr), "n"(0xFFu << (8 * (J))),              \
          "n"(0x4440 | (J)), "n"(1u << (BITN)))
#define MDBG_TSTEP(J, BITN)                                                                                 \
    asm("{\n\t.reg .pred p, q, e;\n\t.reg .b32 t, c, a, b, tx, ty;\n\t"                                     \
        "setp.gt.u32 e, %4, 0;\n\tsetp.gt.and.u32 e, %15, %16, e;\n\t"                                      \
        "and.b32 t, %6, %12;\n\tsetp.ne.and.u32 p, t, 0, e;\n\t"                                            \
        "and.b32 t, %17, %12;\n\t@e or.b32 %5, %5, t;\n\t"                                                  \
        "prmt.b32 c, %7, 0, %13;\n\t"                                                                       \
        "@p mad.lo.u32 %0, %0, 4, c;\n\t"                                                                   \
        "@p add.u32 %4, %4, -1;\n\t"                                                                        \
        "shr.u32 t, %0, %8;\n\t"                                                                            \
        "lop3.b32 t, t, 48, c, 0xEA;\n\t"                                                                   \
        "add.u32 t, t, %9;\n\t"                                                                             \
        "ld.shared.u32 tx, [t];\n\tld.shared.u32 ty, [t+64];\n\t"                                           \
        "shl.b32 a, %1, 1;\n\t@p xor.b32 %1, a, tx;\n\t"                                                    \
        "shl.b32 b, %2, 1;\n\t@p xor.b32 %2, b, ty;\n\t"                                                    \
        "and.b32 a, %2, %11;\n\tsetp.eq.u32 q, a, 0;\n\t"                                                   \
        "setp.le.or.u32 q, %1, %10, q;\n\tand.pred q, q, p;\n\t"                                            \
        "@q or.b32 %3, %3, %14;\n\t}"                                                                       \
        : "+r"(hist), "+r"(F), "+r"(G), "+r"(cand), "+r"(rem), "+r"(bad)                                    \
        : "r"(x), "r"(cw), "r"(SHr), "r"(tab_r), "r"(fthr), "r"(gzr), "n"(0xFFu << (8 * (J))),              \
          "n"(0x4440 | (J)), "n"(1u << (BITN)), "r"(left), "n"(BITN), "r"(bw))
            uint32_t prevb;
            // A read that ends inside the NEXT thread's segment is walked to its end by this thread
            // (the next thread skips that head portion, see for_each_portion): the whole warp is
            // in its halo loop at that time, instead of one lane crawling through a ragged portion.
            const bool extend = (plim == a + SEG) && (re < plim + SEG) && (tid + 1 < NT);
            uint32_t rem = extend ? 0x7FFFFFFFu : l - 1;
            int64_t tail_from = plim;
            // chunked path: the portion ends with the segment and starts either with the segment or
            // with a read that begins inside it
            if (plim == a + SEG && (pb == a || pb == rs)) {
                // 8 x LDS.128, everything else in registers.  The loop constants are pinned in
                // ordinary registers (a volatile move cannot be re-materialised) so that nothing is
                // re-derived per step.
                uint32_t tab_r, SHr, fthr, gzr, row_r;
                asm volatile("mov.b32 %0, %1;" : "=r"(tab_r) : "r"(tab_s));
                asm volatile("mov.b32 %0, %1;" : "=r"(SHr) : "r"(SH));
                asm volatile("mov.b32 %0, %1;" : "=r"(fthr) : "r"(fth));
                asm volatile("mov.b32 %0, %1;" : "=r"(gzr) : "r"(gz));
                asm volatile("mov.b32 %0, %1;" : "=r"(row_r) : "r"(raw_s + (tid + 1) * RSTRIDE));
                uint32_t pw = 0;
                int c16 = 0;
                BadAcc bacc{0, 0, 0};
                if (pb == a) {
                    uint32_t first = sm.raw[(tid + 1) * RSTRIDE];
                    pw = ((pb > rs) ? (uint32_t)sm.raw[tid * RSTRIDE + 127] : (first ^ 0xFFu)) << 24;
                } else {
                    // the read starts at byte `off` of chunk c16: earlier bytes belong to the previous
                    // read (no runs, not inspected), byte `off` is a run start whatever precedes it
                    c16 = (int)(pb - a) >> 4;
                    const int off = (int)(pb - a) & 15;
                    uint32_t ws[4];
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ws[0]), "=r"(ws[1]), "=r"(ws[2]), "=r"(ws[3])
                        : "r"(row_r + c16 * 16));
                    uint32_t cand = 0;
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        uint32_t w = ws[wi];
                        uint32_t x = HPC ? (w ^ __funnelshift_l(pw, w, 8)) : 0xFFFFFFFFu;
                        uint32_t cw = (w << 1) & 0x0C0C0C0Cu;   // 2-bit codes, pre-scaled by 4
                        uint32_t bw = badword(w);
                        const int wlo = off - 4 * wi;
                        if (wlo >= 4) { x = 0; bw = 0; }
                        else if (wlo >= 0) {
                            uint32_t mk = 0xFFFFFFFFu << (8 * wlo);
                            x = (x & mk) | (0xFFu << (8 * wlo));
                            bw &= mk;
                        }
                        bad |= bw;
                        if (wi == 0) { MDBG_STEP(0, 0); MDBG_STEP(1, 1); MDBG_STEP(2, 2); MDBG_STEP(3, 3); }
                        if (wi == 1) { MDBG_STEP(0, 4); MDBG_STEP(1, 5); MDBG_STEP(2, 6); MDBG_STEP(3, 7); }
                        if (wi == 2) { MDBG_STEP(0, 8); MDBG_STEP(1, 9); MDBG_STEP(2, 10); MDBG_STEP(3, 11); }
                        if (wi == 3) { MDBG_STEP(0, 12); MDBG_STEP(1, 13); MDBG_STEP(2, 14); MDBG_STEP(3, 15); }
                        pw = w;
                    }
                    while (cand) {
                        int b = __ffs(cand) - 1;
                        cand &= cand - 1;
                        push(a + c16 * 16 + b);
                    }
                    c16++;
                }
#pragma unroll 1
                for (; c16 < 8; c16++) {
                    uint32_t ws[4];
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ws[0]), "=r"(ws[1]), "=r"(ws[2]), "=r"(ws[3])
                        : "r"(row_r + c16 * 16));
                    uint32_t cand = 0;
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        uint32_t w = ws[wi];
                        uint32_t x = HPC ? (w ^ __funnelshift_l(pw, w, 8)) : 0xFFFFFFFFu;
                        uint32_t cw = (w << 1) & 0x0C0C0C0Cu;   // 2-bit codes, pre-scaled by 4
                        bad_accumulate(bacc, w);
                        if (wi == 0) { MDBG_STEP(0, 0); MDBG_STEP(1, 1); MDBG_STEP(2, 2); MDBG_STEP(3, 3); }
                        if (wi == 1) { MDBG_STEP(0, 4); MDBG_STEP(1, 5); MDBG_STEP(2, 6); MDBG_STEP(3, 7); }
                        if (wi == 2) { MDBG_STEP(0, 8); MDBG_STEP(1, 9); MDBG_STEP(2, 10); MDBG_STEP(3, 11); }
                        if (wi == 3) { MDBG_STEP(0, 12); MDBG_STEP(1, 13); MDBG_STEP(2, 14); MDBG_STEP(3, 15); }
                        pw = w;
                    }
                    while (cand) {
                        int b = __ffs(cand) - 1;
                        cand &= cand - 1;
                        push(a + c16 * 16 + b);
                    }
                }
                bad |= bad_of(bacc);
                // data-dependent halo, same code chunk by chunk: the rows after this thread's row
                // (up to 256 more bytes) are in shared memory for every thread of the tile
#pragma unroll 1
                for (int tc = 0; tc < 16 && rem > 0; tc++) {
                    int64_t cpos = a + SEG + tc * 16;
                    if (cpos >= re) break;
                    uint32_t left = (re - cpos) > 16 ? 16u : (uint32_t)(re - cpos);
                    uint32_t ws[4];
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ws[0]), "=r"(ws[1]), "=r"(ws[2]), "=r"(ws[3])
                        : "r"(row_r + (1 + (tc >> 3)) * RSTRIDE + (tc & 7) * 16));
                    uint32_t cand = 0;
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        uint32_t w = ws[wi];
                        uint32_t x = HPC ? (w ^ __funnelshift_l(pw, w, 8)) : 0xFFFFFFFFu;
                        uint32_t cw = (w << 1) & 0x0C0C0C0Cu;   // 2-bit codes, pre-scaled by 4
                        uint32_t bw = badword(w);
                        if (wi == 0) { MDBG_TSTEP(0, 0); MDBG_TSTEP(1, 1); MDBG_TSTEP(2, 2); MDBG_TSTEP(3, 3); }
                        if (wi == 1) { MDBG_TSTEP(0, 4); MDBG_TSTEP(1, 5); MDBG_TSTEP(2, 6); MDBG_TSTEP(3, 7); }
                        if (wi == 2) { MDBG_TSTEP(0, 8); MDBG_TSTEP(1, 9); MDBG_TSTEP(2, 10); MDBG_TSTEP(3, 11); }
                        if (wi == 3) { MDBG_TSTEP(0, 12); MDBG_TSTEP(1, 13); MDBG_TSTEP(2, 14); MDBG_TSTEP(3, 15); }
                        pw = w;
                    }
                    while (cand) {
                        int b = __ffs(cand) - 1;
                        cand &= cand - 1;
                        push(cpos + b);
                    }
                    tail_from = cpos + 16;
                }
                prevb = pw >> 24;
            } else {
                prevb = (pb > rs) ? W.byte(pb - 1) : 0x100u;
                for (int64_t p = pb; p < plim; p++) {
                    uint32_t c = W.byte(p);
                    bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    if (!HPC || c != prevb) { if (run_step((c >> 1) & 3u)) push(p); }
                    prevb = c;
                }
            }
#undef MDBG_STEP
#undef MDBG_TSTEP
            // ragged portions, and halos longer than 256 bytes: byte-wise
            for (int64_t p = tail_from; p < re && rem > 0; p++) {
                uint32_t c = W.byte(p);
                bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                if (!HPC || c != prevb) { if (run_step((c >> 1) & 3u)) push(p); rem--; }
                prevb = c;
            }
            if (bad) exact_portion_impl(rs, re, pb, extend ? re : plim, true);
        };

        auto for_each_portion = [&](auto&& fn, bool delegate) {
            if (a >= t1) return;
            uint64_t r = find_read(A.read_off, rlo, rhi, a);
            int64_t p = a;
            while (p < seg_end) {
                int64_t rs = (int64_t)__ldg(A.read_off + r), re = (int64_t)__ldg(A.read_off + r + 1);
                if (re <= p) { r++; continue; }
                int64_t plim = seg_end < re ? seg_end : re;
                // the tail of a read that began before this segment and ends inside it was already
                // walked by the previous thread of the tile (`extend` in fast_portion)
                if (!(delegate && p == a && tid > 0 && rs < a && re < a + SEG)) fn(rs, re, p, plim);
                p = plim;
                if (p == re) r++;
            }
        };

        // ---- phase A: scan --------------------------------------------------------------
        bool queue_mode = use_filter;
        if (use_filter) for_each_portion(fast_portion, true);
        else for_each_portion(exact_portion, false);
        __syncwarp();
        if (use_filter) {
            const uint32_t qn = sm.qn;
            if (qn > QCAP) {  // low-complexity sequence flooded the queue: exact path for the tile
                queue_mode = false;
                for (int i = tid; i < WORDS; i += NT) sm.bitmap[i] = 0;
                __syncwarp();
                if (tid == 0) atomicAdd(A.dense_tiles, 1u);
                for_each_portion(exact_portion, false);
            } else {
                // ---- phase B: exact re-evaluation of the survivors -------------------------
                for (uint32_t qi = tid; qi < qn; qi += NT) {
                    uint32_t ent = sm.queue[qi];
                    if (ent & Q_VERIFIED) { sm.queue[qi] = ent & ~Q_VERIFIED; continue; }
                    sm.queue[qi] = Q_DROP;
                    int64_t pe = w0 + (int64_t)ent;
                    // the read containing pe (pe may lie past this tile: widen the range)
                    uint64_t hi2 = rhi;
                    if (pe >= t1) hi2 = A.n_reads;
                    uint64_t r = find_read(A.read_off, rlo, hi2, pe);
                    int64_t rs = (int64_t)__ldg(A.read_off + r);
                    // the scan flags every byte at which the filter state passes, also the non-run
                    // bytes after a passing run: only run starts are windows
                    if (HPC && pe > rs && W.byte(pe - 1) == W.byte(pe)) continue;
                    int64_t p0;
                    uint64_t h;
                    int rel_p0;
                    if (verify_fast<HPC>(sm.raw, cs.hfw, cs.hrc, (int)ent, rs - w0, l, rel_p0, h)) {
                        p0 = w0 + rel_p0;
                        if (p0 < t0 || p0 >= t1) continue;  // owned by another tile
                    } else {
                        int64_t re = (int64_t)__ldg(A.read_off + r + 1);
                        if (!walk_back<HPC>(W, pe, rs, l, p0)) continue;
                        if (p0 < t0 || p0 >= t1) continue;
                        int64_t inv = 0;
                        int res = lmer_hash<HPC>(W, p0, re, l, h, inv);
                        if (res == 0) continue;
                        if (res == 2) atomicMin(A.err_pos, (unsigned long long)inv);
                    }
                    if (h <= bound) {
                        uint32_t rel = (uint32_t)(p0 - t0), bit = 1u << (rel & 31);
                        uint32_t old = atomicOr(&sm.bitmap[rel >> 5], bit);
                        if (!(old & bit)) { sm.queue[qi] = rel; sm.hq[qi] = h; }  // first finder emits it
                    }
                }
            }
        } else if (tid == 0) {
            atomicAdd(A.dense_tiles, 1u);
        }
        __syncwarp();

        // ---- count + block scan of the bitmap -------------------------------------------
        uint32_t wv[4], cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { wv[i] = sm.bitmap[tid * 4 + i]; cnt += __popc(wv[i]); }
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += n;
        }
        const uint32_t woff = 0, total = __shfl_sync(0xffffffffu, inc, 31);
        uint32_t ex = woff + inc - cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) { sm.prefix[tid * 4 + i] = ex; ex += __popc(wv[i]); }

        // ---- reserve this tile's slice of the staging arrays (no ordering between tiles here: the
        //      finalize kernel scans the per-tile counts and moves the slices to their final place)
        if (tid == 0) {
            unsigned long long sb = atomicAdd(A.stage_counter, (unsigned long long)total);
            sm.base = sb;
            sm.total = total;
            A.tile_cnt[tile] = total;
            A.tile_soff[tile] = sb;
        }
        __syncwarp();
        const uint64_t obase = sm.base;

        // ---- emit: per-read offsets, then (hash, pos) at the final positions ---------------
        for (uint64_t r = lb + tid; r < lbn; r += NT) {
            int64_t x = (int64_t)__ldg(A.read_off + r) - t0;  // 0 <= x <= TILE
            uint32_t rank;
            if (x >= TILE) rank = sm.total;
            else rank = sm.prefix[x >> 5] + __popc(sm.bitmap[x >> 5] & ((1u << (x & 31)) - 1u));
            A.out_read_off[A.read_base + r] = (tile << 32) | rank;   // fixed up by ka_finalize_kernel
        }
        if (queue_mode) {
            // every verified queue entry knows its hash; its rank is a popcount prefix of the bitmap
            const uint32_t qn = sm.qn;
            for (uint32_t qi = tid; qi < qn; qi += NT) {
                uint32_t rel = sm.queue[qi];
                if (rel == Q_DROP) continue;
                uint32_t rank = sm.prefix[rel >> 5] + __popc(sm.bitmap[rel >> 5] & ((1u << (rel & 31)) - 1u));
                int64_t p0 = t0 + rel;
                uint64_t r = find_read(A.read_off, rlo, rhi, p0);
                uint64_t o = obase + rank;
                if (o < A.stage_cap) {
                    A.stage_hash[o] = sm.hq[qi];
                    A.stage_pos[o] = (uint32_t)(p0 - (int64_t)__ldg(A.read_off + r));
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < 4; i++) {
                uint32_t word = wv[i];
                uint32_t rank = sm.prefix[tid * 4 + i];
                while (word) {
                    int b = __ffs(word) - 1;
                    word &= word - 1;
                    int64_t p0 = t0 + (tid * 4 + i) * 32 + b;
                    uint64_t r = find_read(A.read_off, rlo, rhi, p0);
                    int64_t rs = (int64_t)__ldg(A.read_off + r), re = (int64_t)__ldg(A.read_off + r + 1);
                    uint64_t h = 0;
                    int64_t inv;
                    lmer_hash<HPC>(W, p0, re, l, h, inv);
                    uint64_t o = obase + rank;
                    if (o < A.stage_cap) {
                        A.stage_hash[o] = h;
                        A.stage_pos[o] = (uint32_t)(p0 - rs);
                    }
                    rank++;
                }
            }
        }
    }
}

// Eight lanes per tile (a tile holds ~4096 * 2 * density minimizers: a dozen): move the tile's staged slice
// to its final, globally ordered position; then fix up the per-read offsets (tile-relative rank -> global
// index).
__global__ void ka_finalize_kernel(const KAArgs A, const uint64_t* __restrict__ tile_excl) {
    const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t tile = gtid >> 3, sub = gtid & 7;
    if (tile < A.n_tiles) {
        const uint64_t cnt = A.tile_cnt[tile], src = A.tile_soff[tile], dst = A.out_base + tile_excl[tile];
        for (uint64_t i = sub; i < cnt; i += 8) {
            if (src + i < A.stage_cap && dst + i < A.out_cap) {
                A.out_hash[dst + i] = A.stage_hash[src + i];
                A.out_pos[dst + i] = A.stage_pos[src + i];
            }
        }
        if (tile + 1 == A.n_tiles && sub == 0) *A.total_out = A.out_base + tile_excl[tile] + cnt;
    }
    if (gtid == 0 && A.dirty_out) *A.dirty_out = A.dirty_n ? *A.dirty_n : 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = gtid; r <= A.n_reads; r += stride) {
        uint64_t v = A.out_read_off[A.read_base + r];
        A.out_read_off[A.read_base + r] = A.out_base + tile_excl[v >> 32] + (v & 0xFFFFFFFFull);
    }
}

// ---- host launchers ------------------------------------------------------------------------
cudaError_t ka_prepare(const KAArgs& A, const KAInit& I, cudaStream_t st, uint64_t* launches) {
    uint64_t nt = A.n_tiles;
    unsigned nb = (unsigned)((nt + 1 + 255) / 256);
    ka_tile_lb_kernel<<<nb, 256, 0, st>>>(A.read_off, A.n_reads, nt, A.tile_lb, I);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t ka_launch(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches) {
    if (A.tile_end <= A.tile_begin) return cudaSuccess;   // A.tile_counter: zeroed by ka_prepare, one per launch
    uint64_t warps = A.tile_end - A.tile_begin;
    uint64_t need = (warps + NWARP - 1) / NWARP;
    unsigned g = (unsigned)(need < (uint64_t)grid ? need : (uint64_t)grid);
    if (hpc) ka_minimizers_kernel<true><<<g, KA_THREADS, 0, st>>>(A);
    else ka_minimizers_kernel<false><<<g, KA_THREADS, 0, st>>>(A);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t ka_launch_list(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches) {
    if (!A.tile_list || !A.tile_list_n || grid < 1) return cudaErrorInvalidValue;
    if (hpc) ka_minimizers_kernel<true><<<grid, KA_THREADS, 0, st>>>(A);
    else ka_minimizers_kernel<false><<<grid, KA_THREADS, 0, st>>>(A);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t ka_finalize(const KAArgs& A, const uint64_t* tile_excl, cudaStream_t st, uint64_t* launches) {
    uint64_t threads = A.n_tiles * 8;
    ka_finalize_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(A, tile_excl);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

int ka_max_blocks_per_sm(int hpc) {
    int n = 0;
    if (hpc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka_minimizers_kernel<true>, KA_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka_minimizers_kernel<false>, KA_THREADS, 0);
    return n;
}

}  // namespace mdbg
