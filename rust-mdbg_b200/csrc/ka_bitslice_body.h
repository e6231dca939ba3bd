// ka_bitslice_body.h -- the per-warp tile loop of the bit-sliced K-A variant.
//
// Same contract as ka_minimizers.cu (reference: Read::encode_rle src/read.rs:157-174, the
// NtHashIterator loop src/read.rs:196, the `<= hash_bound` push src/read.rs:183,196-208) and same
// tile protocol (4 KiB tiles on absolute offsets, per-tile count + slice of the staging arrays,
// ka_finalize_kernel restores the global order), but the work per base is ~3.5x smaller:
//
//   P1  stage the tile (cp.async 16-byte chunks -> padded rows, one row of 128 bytes per lane); the
//       copy of tile n+1 is issued as soon as the rows of tile n are dead and lands under P4..P6;
//   P2  ASCII -> two bit planes, 8 bases per multiply (ka_bitslice_math.h), alphabet check;
//   P3  run starts = plane word XOR itself shifted by one (32 bases per instruction), read starts
//       forced; the planes are compacted by the run mask (parallel-suffix compress) and appended to
//       the warp's HPC bit streams in shared memory -- the HPC string is materialised at 2 bits/base;
//   P4  every lane filters windows of 64 HPC positions (stride 65-T) with filter_window64<L,T>;
//   P5  the ~0.8 % survivors are hashed exactly from 4-base tables; raw positions come from a
//       select in the run masks; windows that leave their read are dropped with a bitmap of read
//       starts in HPC space;
//   P6  ranks = popcount prefix of the accepted bitmap; (hash, pos) and per-read offsets are written
//       exactly like ka_minimizers_kernel does.
//
// A warp claims GROUPS of consecutive tiles and walks them from the top down: the first 32 HPC bases
// of tile t+1 are the look-ahead of tile t (kept in shared memory), so only the top tile of a group
// computes a look-ahead of its own (32 raw bytes, one lane).
//
// Whatever does not fit this scheme -- any byte outside ACGT (N included: it hashes as 0 and has
// no 2-bit code), a look-ahead with fewer than l-1 runs (long homopolymers), a candidate queue
// overflow (low-complexity sequence) -- marks the tile DIRTY: nothing is emitted for it here, the
// tile number goes to a list and ka_minimizers_kernel (exact for every input) processes the list
// afterwards.  Never the CPU.
//
// This file is compiled twice: by nvcc into ka_bitslice.cu, and by g++ into tests/model/ where 32
// host threads per warp execute it with emulated warp primitives (BS_* layer below) -- the kernel
// logic is checked on the CPU against the reference restatement, bit for bit, before it ever sees a GPU.
#pragma once
#include <stdint.h>

#include "ka_bitslice_math.h"
#include "mdbg_kernels.h"

namespace mdbg {
namespace bs {

constexpr int TILE = KA_TILE;          // 4096 raw bytes owned by a tile
constexpr int ROWS = 33;               // 32 rows of 128 bytes + the look-ahead row of a group's top tile
constexpr int RSTRIDE = 144;           // row pitch: conflict-free LDS.128 with one row per lane
constexpr int CW = 136;                // words of an HPC bit stream: 4096 + 32 bits, + reach of the filter
constexpr int NW = 128;                // run-mask words of a tile
constexpr int QCAP = 64;               // candidate queue (expected ~25 per tile at T = 8)
constexpr int HALO_RUNS = 16;          // look-ahead kept by a top tile (>= l-1 for every instantiated l)
constexpr uint32_t Q_DROP = 0xFFFFFFFFu;
#ifndef MDBG_BS_UNROLL
#define MDBG_BS_UNROLL 1               // words of a lane's row processed per iteration of the P2/P3 loop (A/B: 1, 2, 4)
#endif
constexpr int BS_UNROLL = MDBG_BS_UNROLL;

struct Carry {                         // the first HPC bases of tile t+1, seen from tile t
    uint32_t a, b, rs;                 // planes and read-start bits, cnt valid positions
    uint32_t cnt;
    uint32_t bad;                      // tile t+1 holds a byte outside ACGT
    uint32_t ok;                       // cnt >= l-1, or the data ends inside these cnt runs
};

struct __align__(16) WarpSmem {
    uint8_t raw[ROWS * RSTRIDE];       // the staged tile; refilled for the NEXT tile while P4..P6 run
    uint32_t CA[CW], CB[CW];           // HPC string of the tile (+ look-ahead), 1 bit per base and plane
    uint32_t RS[CW];                   // HPC positions that start a read; after P5: popcount prefix of ACC
    uint32_t ACC[CW];                  // P3: forced run starts (read starts), raw space; P5/P6: accepted windows
    uint32_t mraw[NW];                 // run starts, raw space
    uint32_t cpre[NW + 1];             // HPC position of the first run of every raw word
    union {
        struct { uint32_t ca[NW], cb[NW]; } t;                                   // P3: compacted plane words
        struct { uint64_t hq[QCAP]; uint32_t queue[QCAP], qpos[QCAP]; } q;      // P4..P6: candidates
    } u;
    Carry carry;
    uint32_t qn;
    unsigned long long base;
};
static_assert(sizeof(uint32_t) * 4 * CW % 16 == 0, "CA..ACC are cleared with 128-bit stores");

// ---- warp primitives: the real ones under nvcc, emulated ones (tests/model/) under g++ -------------
#if defined(__CUDACC__)
#define BS_DEV __device__ __forceinline__
BS_DEV uint32_t bs_shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
BS_DEV uint32_t bs_shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
BS_DEV void bs_syncwarp() { __syncwarp(); }
BS_DEV bool bs_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
BS_DEV uint32_t bs_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
BS_DEV uint32_t bs_atomic_or_s(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
BS_DEV uint32_t bs_atomic_add_s(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
BS_DEV uint32_t bs_atomic_add_g32(unsigned int* p, uint32_t v) { return atomicAdd(p, v); }
BS_DEV unsigned long long bs_atomic_add_g64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
BS_DEV uint64_t bs_ldg64(const uint64_t* p) { return __ldg(p); }
// 16 bytes global -> shared without a register round trip (LDGSTS); completion: bs_stage_wait()
BS_DEV void bs_cp_async16(uint8_t* dst_smem, const uint8_t* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
BS_DEV void bs_stage_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
BS_DEV void bs_stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
#define BS_DEV inline
uint32_t bs_shfl(uint32_t v, int src);
uint32_t bs_shfl_up(uint32_t v, int d);
void bs_syncwarp();
bool bs_any(bool p);
uint32_t bs_ballot(bool p);
uint32_t bs_atomic_or_s(uint32_t* p, uint32_t v);
uint32_t bs_atomic_add_s(uint32_t* p, uint32_t v);
uint32_t bs_atomic_add_g32(unsigned int* p, uint32_t v);
unsigned long long bs_atomic_add_g64(unsigned long long* p, unsigned long long v);
inline uint64_t bs_ldg64(const uint64_t* p) { return *p; }
inline void bs_cp_async16(uint8_t* dst, const uint8_t* src) { __builtin_memcpy(dst, src, 16); }
inline void bs_stage_commit() {}
inline void bs_stage_wait() {}
#endif

// last r in [lo, hi) with read_off[r] <= p   (read_off[lo] <= p guaranteed)
BS_DEV uint64_t find_read(const uint64_t* read_off, uint64_t lo, uint64_t hi, int64_t p) {
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if ((int64_t)bs_ldg64(read_off + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// OR the low bits of v into a bit stream at bit offset o (v has no bits above its length)
BS_DEV void put_bits(uint32_t* arr, uint32_t o, uint32_t v) {
    const uint32_t w = o >> 5, s = o & 31u;
    if (v << s) bs_atomic_or_s(arr + w, v << s);
    if (s && (v >> (32u - s))) bs_atomic_or_s(arr + w + 1, v >> (32u - s));
}
// the same for both plane streams at once, branch-free (the streams have a spare word at the end)
BS_DEV void put_bits2(uint32_t* arr_a, uint32_t* arr_b, uint32_t o, uint32_t va, uint32_t vb) {
    const uint32_t w = o >> 5, s = o & 31u;
    bs_atomic_or_s(arr_a + w, va << s);
    bs_atomic_or_s(arr_a + w + 1, shr_clamp(va, 32u - s));
    bs_atomic_or_s(arr_b + w, vb << s);
    bs_atomic_or_s(arr_b + w + 1, shr_clamp(vb, 32u - s));
}

// 32 raw bytes (8 words at p, 16-byte aligned) -> plane words; alphabet flags accumulate in acc
BS_DEV void gather32(const uint8_t* p, uint32_t& A, uint32_t& B, BadAcc& acc) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(p);
    const uint4 v1 = *reinterpret_cast<const uint4*>(p + 16);
    uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
    plane_pair(v0.x, v0.y, a0, b0); plane_pair(v0.z, v0.w, a1, b1);
    plane_pair(v1.x, v1.y, a2, b2); plane_pair(v1.z, v1.w, a3, b3);
    A = top_bytes(a0, a1, a2, a3);
    B = top_bytes(b0, b1, b2, b3);
    bad_accumulate(acc, v0.x); bad_accumulate(acc, v0.y); bad_accumulate(acc, v0.z); bad_accumulate(acc, v0.w);
    bad_accumulate(acc, v1.x); bad_accumulate(acc, v1.y); bad_accumulate(acc, v1.z); bad_accumulate(acc, v1.w);
}

// Queue the set bits of a 64-position candidate set (HPC positions s + k): ONE shared-memory atomic per lane
// and window (the lane reserves all its slots at once; lanes without a candidate -- most -- skip it.  A
// shuffle scan of the counts instead of the atomic was measured 2 % slower: it runs for every window).
BS_DEV void push_candidates(WarpSmem& sm, uint32_t c_lo, uint32_t c_hi, const uint32_t s) {
    if (!(c_lo | c_hi)) return;
    uint32_t qi = bs_atomic_add_s(&sm.qn, popc32(c_lo) + popc32(c_hi));
    uint32_t base = s;
    for (int half = 0; half < 2; half++) {
        uint32_t cand = half ? c_hi : c_lo;
        while (cand) {
#if defined(__CUDA_ARCH__)
            const uint32_t k = (uint32_t)__ffs((int)cand) - 1u;
#else
            const uint32_t k = (uint32_t)__builtin_ctz(cand);
#endif
            cand &= cand - 1u;
            if (qi < (uint32_t)QCAP) sm.u.q.queue[qi] = base + k;
            qi++;
        }
        base += 32u;
    }
}

// The tiles of one warp, in the order it walks them: groups of S consecutive tiles claimed from an
// atomic counter, each group from its top tile down.
struct TileIter {
    uint64_t tile, tlo;
    bool valid, top;
};
BS_DEV void iter_claim(const KAArgs& A, TileIter& it, const int lane) {
    const uint64_t S = A.bs_group ? A.bs_group : 1;
    const uint64_t ngroups = A.bs_ngroups;     // (a 64-bit division per claim was 40 instructions per tile)
    uint32_t g = 0;
    if (lane == 0) g = bs_atomic_add_g32(A.tile_counter, 1u);
    g = bs_shfl(g, 0);
    it.valid = (uint64_t)g < ngroups;
    if (it.valid) {
        it.tlo = A.tile_begin + (uint64_t)g * S;
        const uint64_t thi = (it.tlo + S < A.tile_end) ? it.tlo + S : A.tile_end;
        it.tile = thi - 1;
        it.top = true;
    }
}
BS_DEV void iter_next(const KAArgs& A, TileIter& it, const int lane) {
    if (it.tile > it.tlo) { it.tile--; it.top = false; }
    else iter_claim(A, it, lane);
}

// P1: start copying a tile (and the look-ahead row of a top tile) into the warp's rows.  Whole
// 16-byte chunks inside the batch go global -> shared asynchronously; the ragged end of the batch is
// written with ordinary stores, padded with 'A' (never a run start, see the limit mask).
BS_DEV void stage_issue(const KAArgs& A, WarpSmem& sm, const int lane, const uint64_t tile, const bool top) {
    const uint8_t* gb = A.bases;
    const int64_t B = (int64_t)A.n_bases;
    const int64_t t0 = (int64_t)tile * TILE;
    const int nchunks = (top ? ROWS : ROWS - 1) * 8;
    if (t0 + (int64_t)nchunks * 16 <= B) {
        const uint8_t* src = gb + t0 + lane * 16;
        uint8_t* dst = sm.raw + (lane >> 3) * RSTRIDE + (lane & 7) * 16;   // chunk q = lane + 32 i -> row 4 i + (lane >> 3)
#pragma unroll
        for (int i = 0; i < 8; i++) bs_cp_async16(dst + i * 4 * RSTRIDE, src + i * 512);
        if (top && lane < 8) bs_cp_async16(sm.raw + 32 * RSTRIDE + lane * 16, gb + t0 + TILE + lane * 16);
    } else {
        for (int q = lane; q < nchunks; q += 32) {
            const int64_t gp = t0 + (int64_t)q * 16;
            uint8_t* dst = sm.raw + (q >> 3) * RSTRIDE + (q & 7) * 16;
            if (gp + 16 <= B) {
                bs_cp_async16(dst, gb + gp);
            } else {
                uint32_t t[4];
                for (int wq = 0; wq < 4; wq++) {
                    uint32_t x = 0;
                    for (int j = 0; j < 4; j++) {
                        const int64_t pos = gp + wq * 4 + j;
                        x |= (uint32_t)(pos < B ? gb[pos] : (uint8_t)'A') << (8 * j);
                    }
                    t[wq] = x;
                }
                uint4 v;
                v.x = t[0]; v.y = t[1]; v.z = t[2]; v.w = t[3];
                *reinterpret_cast<uint4*>(dst) = v;
            }
        }
    }
    bs_stage_commit();
}

// One tile.  Its bytes were requested by stage_issue(); `next` is the tile this warp takes afterwards:
// its bytes are requested as soon as this tile's rows are no longer needed (they arrive while P4..P6 run).
template <int L, int T, bool HPC>
BS_DEV void process_tile(const KAArgs& A, WarpSmem& sm, const int lane, const uint64_t tile, const bool top,
                         const TileIter& next) {
    const uint8_t* gb = A.bases;
    const int64_t B = (int64_t)A.n_bases;
    const int64_t t0 = (int64_t)tile * TILE;
    const int vt = (B - t0 >= TILE) ? TILE : (B > t0 ? (int)(B - t0) : 0);   // valid bytes of the tile
    const uint64_t lb = bs_ldg64(A.tile_lb + tile), lbn = bs_ldg64(A.tile_lb + tile + 1);
    const uint64_t rlo = lb > 0 ? lb - 1 : 0;
    const uint64_t rhi = lbn < A.n_reads ? lbn : A.n_reads;   // exclusive
    constexpr uint32_t LM = (1u << L) - 1u;

    // ---- clear the bit streams; read starts -> forced run starts (raw space, in ACC) ---------------
    {
        uint4 z;
        z.x = 0; z.y = 0; z.z = 0; z.w = 0;
        uint4* p = reinterpret_cast<uint4*>(sm.CA);                  // CA, CB, RS, ACC are contiguous
        for (int i = lane; i < CW; i += 32) p[i] = z;
        if (lane == 0) sm.qn = 0;
    }
    const uint32_t preb = (t0 > 0 && vt > 0) ? (uint32_t)gb[t0 - 1] : 0u;   // the byte before the tile
    bs_syncwarp();
    // the first base of a read starts a run whatever precedes it (read.rs:157 works per read)
    for (uint64_t r = lb + lane; r < lbn; r += 32) {
        const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - t0;
        if (x < TILE) bs_atomic_or_s(sm.ACC + (x >> 5), 1u << (x & 31));
    }
    bs_stage_wait();                   // this lane's part of the tile has landed ...
    bs_syncwarp();                     // ... and everybody else's

    // ---- P2 + P3: planes, alphabet, run starts, compaction (one raw word of 32 bases at a time) ----
    BadAcc bacc{0, 0, 0};
    uint32_t tot = 0, tail;
    const bool has_starts = lbn > lb;                  // warp-uniform: most tiles hold no read start
    {
        const uint8_t* row = sm.raw + lane * RSTRIDE;
        const uint32_t prevb = lane ? (uint32_t)sm.raw[(lane - 1) * RSTRIDE + 127] : preb;
        const bool prev_ok = is_acgt(prevb);        // N / nothing before the tile: a run starts
        // planes of the previous word: only their top bits matter (the base before this lane's first)
        uint32_t qa = ((prevb >> 1) & 1u) << 31, qb = ((prevb >> 2) & 1u) << 31;
#pragma unroll BS_UNROLL
        for (int n = 0; n < 4; n++) {
            const int idx = lane * 4 + n;
            uint32_t PA, PB;
            gather32(row + 32 * n, PA, PB, bacc);
            uint32_t m = HPC ? ((PA ^ fsl(qa, PA, 1)) | (PB ^ fsl(qb, PB, 1))) : 0xFFFFFFFFu;
            if (n == 0 && !prev_ok) m |= 1u;
            if (has_starts) { m |= sm.ACC[idx]; sm.ACC[idx] = 0; }
            if (vt < TILE) {           // last tile of the batch: nothing starts at or after byte vt
                const int nv = vt - (lane * 128 + 32 * n);
                m &= nv >= 32 ? 0xFFFFFFFFu : (nv > 0 ? low_mask((uint32_t)nv) : 0u);
            }
            qa = PA; qb = PB;
            if (HPC) {                 // the last round of the network only when some lane of the warp needs it
                const PextState ps = pext_pair_rounds4(m, PA, PB);
                if (bs_any(ps.mk != 0)) pext_pair_round5(ps, PA, PB);
            } else { PA &= m; PB &= m; }
            sm.mraw[idx] = m;
            sm.u.t.ca[idx] = PA;
            sm.u.t.cb[idx] = PB;
            tot += popc32(m);
        }
        tail = (qa >> 31) | ((qb >> 31) << 1);       // code of this lane's last base
    }
    const bool tile_bad = bs_any(bad_of(bacc) != 0);
    const uint32_t tail31 = bs_shfl(tail, 31);
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t nb = bs_shfl_up(inc, d);
        if (lane >= d) inc += nb;
    }
    const uint32_t Ctile = bs_shfl(inc, 31);       // runs (HPC bases) owned by the tile
    {
        uint32_t o = inc - tot;
#pragma unroll 1
        for (int n = 0; n < 4; n++) {
            const int idx = lane * 4 + n;
            sm.cpre[idx] = o;
            put_bits2(sm.CA, sm.CB, o, sm.u.t.ca[idx], sm.u.t.cb[idx]);
            o += popc32(sm.mraw[idx]);
        }
        if (lane == 31) sm.cpre[NW] = o;
    }
    // look-ahead: the first HPC bases after the tile (uniform control flow: every lane takes part)
    uint32_t hcnt = 0, unusable = 0;
    if (t0 + TILE < B) {
        if (top) {                     // from the 32 bytes after the tile: one byte per lane
            const int64_t left = B - (t0 + TILE);
            const uint32_t c = sm.raw[32 * RSTRIDE + lane];
            const bool valid = (int64_t)lane < left;
            const uint32_t code = (c >> 1) & 3u;
            uint32_t prevc = bs_shfl_up(code, 1);
            if (lane == 0) prevc = tail31;
            uint32_t hs = 0;           // read starts inside the look-ahead (typically none or one)
            int guard = 0;
            for (uint64_t r = lbn; r <= A.n_reads && guard < 64; r++, guard++) {
                const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - (t0 + TILE);
                if (x >= 32) break;
                hs |= 1u << x;
            }
            const bool rstart = ((hs >> lane) & 1u) != 0;
            const bool isrun = valid && (!HPC || code != prevc || rstart);
            const uint32_t hm = bs_ballot(isrun);
            const bool badb = bs_any(valid && !is_acgt(c));
            const uint32_t rank = popc32(hm & low_mask((uint32_t)lane)), runs = popc32(hm);
            hcnt = runs < (uint32_t)HALO_RUNS ? runs : (uint32_t)HALO_RUNS;
            if (isrun && rank < (uint32_t)HALO_RUNS) {
                const uint32_t o = Ctile + rank, w = o >> 5, bit = 1u << (o & 31u);
                if (code & 1u) bs_atomic_or_s(sm.CA + w, bit);
                if (code & 2u) bs_atomic_or_s(sm.CB + w, bit);
                if (rstart) bs_atomic_or_s(sm.RS + w, bit);
            }
            if (guard >= 64 || badb || !(hcnt >= (uint32_t)(L - 1) || (left <= 32 && hcnt == runs))) unusable = 1;
        } else {                       // from the tile above, processed just before by this warp
            const Carry cy = sm.carry;
            hcnt = cy.cnt;
            if (cy.bad || !cy.ok) unusable = 1;
            if (lane == 0) {
                put_bits(sm.CA, Ctile, cy.a);
                put_bits(sm.CB, Ctile, cy.b);
                put_bits(sm.RS, Ctile, cy.rs);
            }
        }
    }
    bs_syncwarp();                     // the rows are dead from here on ...
    if (next.valid) stage_issue(A, sm, lane, next.tile, next.top);   // ... refill them for the next tile
    // read starts of this tile in HPC space
    for (uint64_t r = lb + lane; r < lbn; r += 32) {
        const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - t0;
        if (x < vt) {
            const uint32_t comp = sm.cpre[x >> 5] + popc32(sm.mraw[x >> 5] & low_mask((uint32_t)(x & 31)));
            bs_atomic_or_s(sm.RS + (comp >> 5), 1u << (comp & 31u));
        }
    }
    const uint32_t Ctotal = Ctile + hcnt;
    bool dirty = tile_bad || unusable;
    bs_syncwarp();
    const uint32_t rs0 = sm.RS[0];     // for the tile below (RS is recycled after P5)

    // ---- P4: filter -------------------------------------------------------------------------
    if (!dirty) {
        constexpr uint32_t STR = 65 - T;               // 64-position windows: the low half keeps all 32
        const uint32_t nwin = (Ctile + STR - 1) / STR;
        for (uint32_t idx = lane; idx < nwin; idx += 32) {
            const uint32_t s = idx * STR, w = s >> 5, sh = s & 31u;
            const uint32_t x0 = sm.CA[w], x1 = sm.CA[w + 1], x2 = sm.CA[w + 2], x3 = sm.CA[w + 3];
            const uint32_t y0 = sm.CB[w], y1 = sm.CB[w + 1], y2 = sm.CB[w + 2], y3 = sm.CB[w + 3];
            uint32_t c_lo, c_hi;
            filter_window64<L, T>(fsr(x0, x1, sh), fsr(x1, x2, sh), fsr(x2, x3, sh), fsr(y0, y1, sh), fsr(y1, y2, sh),
                                  fsr(y2, y3, sh), c_lo, c_hi);
            const uint32_t nv = Ctile - s;
            c_lo &= low_mask(nv);
            c_hi &= low_mask(nv > 32u ? (nv - 32u < STR - 32u ? nv - 32u : STR - 32u) : 0u);
            push_candidates(sm, c_lo, c_hi, s);
        }
        bs_syncwarp();
        if (sm.qn > (uint32_t)QCAP) dirty = true;   // low-complexity sequence: exact path
    }

    // ---- P5: exact evaluation of the survivors ----------------------------------------------
    if (!dirty) {
        const uint32_t qn = sm.qn;
        for (uint32_t qi = lane; qi < qn; qi += 32) {
            const uint32_t p = sm.u.q.queue[qi];
            sm.u.q.queue[qi] = Q_DROP;
            if (p + (uint32_t)L > Ctotal) continue;             // fewer than l runs left in the data
            const uint32_t w = p >> 5, sh = p & 31u;
            const uint32_t rsb = fsr(sm.RS[w], sm.RS[w + 1], sh);
            if ((rsb >> 1) & (LM >> 1)) continue;               // a read starts inside the window
            const uint32_t av = fsr(sm.CA[w], sm.CA[w + 1], sh) & LM;
            const uint32_t bv = fsr(sm.CB[w], sm.CB[w + 1], sh) & LM;
            const uint64_t h = exact_hash<L>(av, bv, A.bs_t4);
            if (h > A.bound) continue;
            // raw position of the window's first run: last raw word whose first run is at or before p
            uint32_t lo = 0, hi = NW;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (sm.cpre[mid] <= p) lo = mid; else hi = mid;
            }
            const uint32_t x = lo * 32u + select_bit(sm.mraw[lo], p - sm.cpre[lo]);
            const int64_t p0 = t0 + (int64_t)x;
            const uint64_t r = find_read(A.read_off, rlo, rhi, p0);
            bs_atomic_or_s(sm.ACC + w, 1u << sh);
            sm.u.q.queue[qi] = p;
            sm.u.q.hq[qi] = h;
            sm.u.q.qpos[qi] = (uint32_t)(p0 - (int64_t)bs_ldg64(A.read_off + r));
        }
        bs_syncwarp();
    }

    // ---- P6: ranks, reservation, emission ---------------------------------------------------
    uint32_t* accpre = sm.RS;          // RS is not read any more
    if (!dirty) {
        uint32_t wv[4], cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { wv[i] = sm.ACC[lane * 4 + i]; cnt += popc32(wv[i]); }
        uint32_t sc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t nb = bs_shfl_up(sc, d);
            if (lane >= d) sc += nb;
        }
        const uint32_t total = bs_shfl(sc, 31);
        uint32_t ex = sc - cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) { accpre[lane * 4 + i] = ex; ex += popc32(wv[i]); }
        if (lane == 31) accpre[NW] = total;                      // ACC[NW] stays 0
        if (lane == 0) {
            const unsigned long long sb = bs_atomic_add_g64(A.stage_counter, (unsigned long long)total);
            sm.base = sb;
            A.tile_cnt[tile] = total;
            A.tile_soff[tile] = sb;
        }
        bs_syncwarp();
        const uint64_t obase = sm.base;
        for (uint64_t r = lb + lane; r < lbn; r += 32) {
            const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - t0;
            uint32_t rank = total;
            if (x < vt) {
                const uint32_t comp = sm.cpre[x >> 5] + popc32(sm.mraw[x >> 5] & low_mask((uint32_t)(x & 31)));
                rank = accpre[comp >> 5] + popc32(sm.ACC[comp >> 5] & low_mask(comp & 31u));
            }
            A.out_read_off[A.read_base + r] = ((uint64_t)tile << 32) | rank;   // fixed up by ka_finalize_kernel
        }
        const uint32_t qn = sm.qn;
        for (uint32_t qi = lane; qi < qn; qi += 32) {
            const uint32_t p = sm.u.q.queue[qi];
            if (p == Q_DROP) continue;
            const uint32_t rank = accpre[p >> 5] + popc32(sm.ACC[p >> 5] & low_mask(p & 31u));
            const uint64_t o = obase + rank;
            if (o < A.stage_cap) {
                A.stage_hash[o] = sm.u.q.hq[qi];
                A.stage_pos[o] = sm.u.q.qpos[qi];
            }
        }
    } else if (lane == 0) {
        const uint32_t di = bs_atomic_add_g32(A.dirty_n, 1u);
        A.dirty_list[di] = (uint32_t)tile;
    }

    if (lane == 0) {
        if (A.dbg) {                   // per-tile state for the emulator-vs-GPU comparison
            uint32_t* d = A.dbg + tile * 8;
            d[0] = Ctile; d[1] = Ctotal; d[2] = unusable | (tile_bad ? 2u : 0u) | (dirty ? 4u : 0u);
            d[3] = sm.qn; d[4] = dirty ? 0u : accpre[NW]; d[5] = sm.CA[0]; d[6] = sm.CB[0]; d[7] = sm.mraw[0];
        }
        // what the tile below needs to know about this one
        Carry cy;
        cy.cnt = Ctile < 32u ? Ctile : 32u;
        cy.a = sm.CA[0] & low_mask(cy.cnt);
        cy.b = sm.CB[0] & low_mask(cy.cnt);
        cy.rs = rs0 & low_mask(cy.cnt);
        cy.bad = tile_bad ? 1u : 0u;
        cy.ok = (cy.cnt >= (uint32_t)(L - 1) || (Ctile == cy.cnt && t0 + TILE >= B)) ? 1u : 0u;
        sm.carry = cy;
    }
    bs_syncwarp();
}

// The persistent loop of one warp.
template <int L, int T, bool HPC>
BS_DEV void warp_loop(const KAArgs& A, WarpSmem& sm, const int lane) {
    TileIter it;
    iter_claim(A, it, lane);
    if (!it.valid) return;
    stage_issue(A, sm, lane, it.tile, it.top);
    for (;;) {
        const uint64_t tile = it.tile;
        const bool top = it.top;
        iter_next(A, it, lane);
        process_tile<L, T, HPC>(A, sm, lane, tile, top, it);
        if (!it.valid) break;
    }
}

}  // namespace bs
}  // namespace mdbg
