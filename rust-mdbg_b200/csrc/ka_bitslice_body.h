// ka_bitslice_body.h -- the per-warp tile loop of the bit-sliced K-A variant.
//
// Same contract as ka_minimizers.cu (reference: Read::encode_rle src/read.rs:157-174, the
// NtHashIterator loop src/read.rs:196, the `<= hash_bound` push src/read.rs:183,196-208) and same
// tile protocol (4 KiB tiles on absolute offsets, per-tile count + slice of the staging arrays,
// ka_finalize_kernel restores the global order), but the work per base is ~3.5x smaller:
//
//   P1  stage the tile (coalesced 128-bit loads -> padded rows, one row of 128 bytes per lane);
//   P2  ASCII -> two bit planes, 4 bases per multiply (ka_bitslice_math.h), alphabet check;
//   P3  run starts = plane word XOR itself shifted by one (32 bases per instruction), read starts
//       forced; the planes are compacted by the run mask (parallel-suffix compress) and appended to
//       the warp's HPC bit streams in shared memory -- the HPC string is materialised at 2 bits/base;
//   P4  every lane filters windows of 32 HPC positions (stride 33-T) with filter_window<L,T>;
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
constexpr int QCAP = 128;              // candidate queue (expected ~25 per tile at T = 8)
constexpr uint32_t Q_DROP = 0xFFFFFFFFu;

struct Carry {                         // the first HPC bases of tile t+1, seen from tile t
    uint32_t a, b, rs;                 // planes and read-start bits, cnt valid positions
    uint32_t cnt;
    uint32_t bad;                      // tile t+1 holds a byte outside ACGT
    uint32_t ok;                       // cnt >= l-1, or the data ends inside these cnt runs
};

struct Post {                          // lives in the raw rows once the planes are built
    uint64_t hq[QCAP];                 // exact hash of an accepted candidate
    uint32_t queue[QCAP];              // HPC position of a candidate; Q_DROP once rejected
    uint32_t qpos[QCAP];               // raw position inside its read
    uint32_t ACC[CW];                  // accepted windows, HPC space
    uint32_t accpre[CW];               // exclusive popcount prefix of ACC
};

struct __align__(16) WarpSmem {
    union {
        uint8_t raw[ROWS * RSTRIDE];
        Post post;
    } u;
    uint32_t CA[CW], CB[CW];           // HPC string of the tile (+ look-ahead), 1 bit per base and plane
    uint32_t RS[CW];                   // HPC positions that start a read
    uint32_t mraw[NW];                 // run starts, raw space
    uint32_t cpre[NW + 1];             // HPC position of the first run of every raw word
    Carry carry;
    uint32_t qn;
    uint32_t flags;                    // look-ahead verdict of the tile (bit 0: unusable)
    uint32_t hcnt;
    unsigned long long base;
};
static_assert(sizeof(Post) <= 32 * RSTRIDE, "post-phase arrays must not reach the look-ahead row");

struct CtaTables {
    T4Entry t4[256];
};

// ---- warp primitives: the real ones under nvcc, emulated ones (tests/model/warp_emu.h) under g++ ----
#if defined(__CUDACC__)
#define BS_DEV __device__ __forceinline__
BS_DEV uint32_t bs_shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
BS_DEV uint32_t bs_shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
BS_DEV void bs_syncwarp() { __syncwarp(); }
BS_DEV bool bs_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
BS_DEV uint32_t bs_atomic_or_s(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
BS_DEV uint32_t bs_atomic_add_s(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
BS_DEV uint32_t bs_atomic_add_g32(unsigned int* p, uint32_t v) { return atomicAdd(p, v); }
BS_DEV unsigned long long bs_atomic_add_g64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
BS_DEV uint64_t bs_ldg64(const uint64_t* p) { return __ldg(p); }
BS_DEV uint4 bs_ldg128(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
#else
#define BS_DEV inline
uint32_t bs_shfl(uint32_t v, int src);
uint32_t bs_shfl_up(uint32_t v, int d);
void bs_syncwarp();
bool bs_any(bool p);
uint32_t bs_atomic_or_s(uint32_t* p, uint32_t v);
uint32_t bs_atomic_add_s(uint32_t* p, uint32_t v);
uint32_t bs_atomic_add_g32(unsigned int* p, uint32_t v);
unsigned long long bs_atomic_add_g64(unsigned long long* p, unsigned long long v);
inline uint64_t bs_ldg64(const uint64_t* p) { return *p; }
inline uint4 bs_ldg128(const uint8_t* p) { uint4 v; __builtin_memcpy(&v, p, 16); return v; }
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

// 32 raw bytes (8 words at p, 16-byte aligned) -> plane words; alphabet flags accumulate in acc
BS_DEV void gather32(const uint8_t* p, uint32_t& A, uint32_t& B, BadAcc& acc) {
    const uint4 v0 = *reinterpret_cast<const uint4*>(p);
    const uint4 v1 = *reinterpret_cast<const uint4*>(p + 16);
    A = 0; B = 0;
    plane_push(v1.w, A, B); plane_push(v1.z, A, B); plane_push(v1.y, A, B); plane_push(v1.x, A, B);
    plane_push(v0.w, A, B); plane_push(v0.z, A, B); plane_push(v0.y, A, B); plane_push(v0.x, A, B);
    bad_accumulate(acc, v0.x); bad_accumulate(acc, v0.y); bad_accumulate(acc, v0.z); bad_accumulate(acc, v0.w);
    bad_accumulate(acc, v1.x); bad_accumulate(acc, v1.y); bad_accumulate(acc, v1.z); bad_accumulate(acc, v1.w);
}

// One tile.  `top`: the tile above it is not part of this warp's group (look-ahead from the halo row).
template <int L, int T, bool HPC>
BS_DEV void process_tile(const KAArgs& A, WarpSmem& sm, const CtaTables& ct, const int lane,
                         const uint64_t tile, const bool top) {
    const uint8_t* gb = A.bases;
    const int64_t B = (int64_t)A.n_bases;
    const int64_t t0 = (int64_t)tile * TILE;
    const int vt = (B - t0 >= TILE) ? TILE : (B > t0 ? (int)(B - t0) : 0);   // valid bytes of the tile
    const uint64_t lb = bs_ldg64(A.tile_lb + tile), lbn = bs_ldg64(A.tile_lb + tile + 1);
    const uint64_t rlo = lb > 0 ? lb - 1 : 0;
    const uint64_t rhi = lbn < A.n_reads ? lbn : A.n_reads;   // exclusive
    constexpr uint32_t LM = (1u << L) - 1u;

    // ---- P1: stage ------------------------------------------------------------------------
    for (int i = lane; i < CW; i += 32) { sm.CA[i] = 0; sm.CB[i] = 0; sm.RS[i] = 0; }
    const int nchunks = (top ? ROWS : ROWS - 1) * 8;
    for (int q = lane; q < nchunks; q += 32) {
        const int64_t gp = t0 + (int64_t)q * 16;
        uint4 v;
        if (gp + 16 <= B) {
            v = bs_ldg128(gb + gp);
        } else {                       // past the end of the batch: 'A' (never a run start, see the limit mask)
            uint32_t t[4];
            for (int wq = 0; wq < 4; wq++) {
                uint32_t x = 0;
                for (int j = 0; j < 4; j++) {
                    const int64_t pos = gp + wq * 4 + j;
                    x |= (uint32_t)(pos < B ? gb[pos] : (uint8_t)'A') << (8 * j);
                }
                t[wq] = x;
            }
            v.x = t[0]; v.y = t[1]; v.z = t[2]; v.w = t[3];
        }
        *reinterpret_cast<uint4*>(sm.u.raw + (q >> 3) * RSTRIDE + (q & 7) * 16) = v;
    }
    const uint32_t preb = (t0 > 0 && vt > 0) ? (uint32_t)gb[t0 - 1] : 0u;   // the byte before the tile
    bs_syncwarp();

    // ---- P2: planes + alphabet ------------------------------------------------------------
    uint32_t PA[4], PB[4], M[4];
    BadAcc bacc{0, 0, 0};
    {
        const uint8_t* row = sm.u.raw + lane * RSTRIDE;
#pragma unroll
        for (int n = 0; n < 4; n++) gather32(row + 32 * n, PA[n], PB[n], bacc);
    }
    const bool tile_bad = bs_any(bad_of(bacc) != 0);

    // ---- P3: run starts, compaction ---------------------------------------------------------
    const uint32_t tail = (PA[3] >> 31) | ((PB[3] >> 31) << 1);   // code of this lane's last base
    uint32_t up = bs_shfl_up(tail, 1);
    const uint32_t tail31 = bs_shfl(tail, 31);
    bool force_first = false;
    if (lane == 0) {
        if (is_acgt(preb)) up = (preb >> 1) & 3u; else force_first = true;   // N / nothing before: a run starts
    }
    if (HPC) {
        uint32_t pa = up & 1u, pb = up >> 1;
#pragma unroll
        for (int n = 0; n < 4; n++) {
            M[n] = (PA[n] ^ ((PA[n] << 1) | pa)) | (PB[n] ^ ((PB[n] << 1) | pb));
            pa = PA[n] >> 31; pb = PB[n] >> 31;
        }
        if (force_first) M[0] |= 1u;
    } else {
#pragma unroll
        for (int n = 0; n < 4; n++) M[n] = 0xFFFFFFFFu;
    }
    // the first base of a read starts a run whatever precedes it (read.rs:157 works per read)
    for (uint64_t r = lb; r < lbn; r++) {
        const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - t0;
        if (x < TILE && (x >> 7) == lane) {
            const uint32_t bit = 1u << (x & 31);
            const int n = (int)(x >> 5) & 3;
            if (n == 0) M[0] |= bit;
            if (n == 1) M[1] |= bit;
            if (n == 2) M[2] |= bit;
            if (n == 3) M[3] |= bit;
        }
    }
    if (vt < TILE) {                   // last tile of the batch: nothing starts at or after byte vt
#pragma unroll
        for (int n = 0; n < 4; n++) {
            const int nv = vt - (lane * 128 + 32 * n);
            M[n] &= nv >= 32 ? 0xFFFFFFFFu : (nv > 0 ? low_mask((uint32_t)nv) : 0u);
        }
    }
    uint32_t c[4], tot = 0;
#pragma unroll
    for (int n = 0; n < 4; n++) { c[n] = popc32(M[n]); tot += c[n]; }
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t nb = bs_shfl_up(inc, d);
        if (lane >= d) inc += nb;
    }
    const uint32_t Ctile = bs_shfl(inc, 31);       // runs (HPC bases) owned by the tile
    {
        uint32_t o = inc - tot;
#pragma unroll
        for (int n = 0; n < 4; n++) {
            uint32_t ca = PA[n], cb = PB[n];
            if (HPC) pext_pair(M[n], ca, cb); else { ca &= M[n]; cb &= M[n]; }
            sm.mraw[lane * 4 + n] = M[n];
            sm.cpre[lane * 4 + n] = o;
            if (c[n]) { put_bits(sm.CA, o, ca); put_bits(sm.CB, o, cb); }
            o += c[n];
        }
        if (lane == 31) sm.cpre[NW] = o;
    }
    // look-ahead: the first HPC bases after the tile
    if (lane == 0) {
        uint32_t ha = 0, hb = 0, hrs = 0, hcnt = 0, unusable = 0;
        if (t0 + TILE < B) {
            if (top) {
                BadAcc hacc{0, 0, 0};
                gather32(sm.u.raw + 32 * RSTRIDE, ha, hb, hacc);
                const int64_t left = B - (t0 + TILE);
                uint32_t hm = HPC ? ((ha ^ ((ha << 1) | (tail31 & 1u))) | (hb ^ ((hb << 1) | (tail31 >> 1)))) : 0xFFFFFFFFu;
                uint32_t hs = 0;
                uint64_t r = lbn;
                int guard = 0;
                for (; r <= A.n_reads && guard < 64; r++, guard++) {
                    const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - (t0 + TILE);
                    if (x >= 32) break;
                    hs |= 1u << x;
                }
                if (guard >= 64) unusable = 1;
                hm |= hs;
                if (left < 32) { hm &= low_mask((uint32_t)left); hs &= low_mask((uint32_t)left); }
                if (bad_of(hacc) != 0) unusable = 1;   // conservative: the filler past the batch end is 'A'
                hcnt = popc32(hm);
                uint32_t zero = 0;
                if (HPC) { pext_pair(hm, ha, hb); pext_pair(hm, hs, zero); } else { ha &= hm; hb &= hm; }
                hrs = hs;
                if (!(hcnt >= (uint32_t)(L - 1) || left <= 32)) unusable = 1;
            } else {
                const Carry cy = sm.carry;
                ha = cy.a; hb = cy.b; hrs = cy.rs; hcnt = cy.cnt;
                if (cy.bad || !cy.ok) unusable = 1;
            }
            put_bits(sm.CA, Ctile, ha);
            put_bits(sm.CB, Ctile, hb);
            put_bits(sm.RS, Ctile, hrs);
        }
        sm.hcnt = hcnt;
        sm.flags = unusable;
        sm.qn = 0;
    }
    bs_syncwarp();                     // raw rows are dead from here on: Post may be written
    for (int i = lane; i < CW; i += 32) sm.u.post.ACC[i] = 0;
    // read starts of this tile in HPC space
    for (uint64_t r = lb + lane; r < lbn; r += 32) {
        const int64_t x = (int64_t)bs_ldg64(A.read_off + r) - t0;
        if (x < vt) {
            const uint32_t comp = sm.cpre[x >> 5] + popc32(sm.mraw[x >> 5] & low_mask((uint32_t)(x & 31)));
            bs_atomic_or_s(sm.RS + (comp >> 5), 1u << (comp & 31u));
        }
    }
    const uint32_t Ctotal = Ctile + sm.hcnt;
    bool dirty = tile_bad || (sm.flags & 1u);
    bs_syncwarp();

    // ---- P4: filter -------------------------------------------------------------------------
    if (!dirty) {
        constexpr uint32_t STR = 33 - T;
        const uint32_t nwin = (Ctile + STR - 1) / STR;
        for (uint32_t idx = lane; idx < nwin; idx += 32) {
            const uint32_t s = idx * STR, w = s >> 5, sh = s & 31u;
            const uint32_t x0 = sm.CA[w], x1 = sm.CA[w + 1], x2 = sm.CA[w + 2];
            const uint32_t y0 = sm.CB[w], y1 = sm.CB[w + 1], y2 = sm.CB[w + 2];
            uint32_t cand = filter_window<L, T>(fsr(x0, x1, sh), fsr(x1, x2, sh), fsr(y0, y1, sh), fsr(y1, y2, sh));
            const uint32_t nv = Ctile - s;
            cand &= low_mask(nv < STR ? nv : STR);
            while (cand) {
#if defined(__CUDA_ARCH__)
                const uint32_t k = (uint32_t)__ffs((int)cand) - 1u;
#else
                const uint32_t k = (uint32_t)__builtin_ctz(cand);
#endif
                cand &= cand - 1u;
                const uint32_t qi = bs_atomic_add_s(&sm.qn, 1u);
                if (qi < (uint32_t)QCAP) sm.u.post.queue[qi] = s + k;
            }
        }
        bs_syncwarp();
        if (sm.qn > (uint32_t)QCAP) dirty = true;   // low-complexity sequence: exact path
    }

    // ---- P5: exact evaluation of the survivors ----------------------------------------------
    if (!dirty) {
        const uint32_t qn = sm.qn;
        for (uint32_t qi = lane; qi < qn; qi += 32) {
            const uint32_t p = sm.u.post.queue[qi];
            sm.u.post.queue[qi] = Q_DROP;
            if (p + (uint32_t)L > Ctotal) continue;             // fewer than l runs left in the data
            const uint32_t w = p >> 5, sh = p & 31u;
            const uint32_t rsb = fsr(sm.RS[w], sm.RS[w + 1], sh);
            if ((rsb >> 1) & (LM >> 1)) continue;               // a read starts inside the window
            const uint32_t av = fsr(sm.CA[w], sm.CA[w + 1], sh) & LM;
            const uint32_t bv = fsr(sm.CB[w], sm.CB[w + 1], sh) & LM;
            const uint64_t h = exact_hash<L>(av, bv, ct.t4);
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
            bs_atomic_or_s(sm.u.post.ACC + w, 1u << sh);
            sm.u.post.queue[qi] = p;
            sm.u.post.hq[qi] = h;
            sm.u.post.qpos[qi] = (uint32_t)(p0 - (int64_t)bs_ldg64(A.read_off + r));
        }
        bs_syncwarp();
    }

    // ---- P6: ranks, reservation, emission ---------------------------------------------------
    if (!dirty) {
        uint32_t wv[4], cnt = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { wv[i] = sm.u.post.ACC[lane * 4 + i]; cnt += popc32(wv[i]); }
        uint32_t sc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t nb = bs_shfl_up(sc, d);
            if (lane >= d) sc += nb;
        }
        const uint32_t total = bs_shfl(sc, 31);
        uint32_t ex = sc - cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) { sm.u.post.accpre[lane * 4 + i] = ex; ex += popc32(wv[i]); }
        if (lane == 31) sm.u.post.accpre[NW] = total;            // ACC[NW] stays 0
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
                rank = sm.u.post.accpre[comp >> 5] + popc32(sm.u.post.ACC[comp >> 5] & low_mask(comp & 31u));
            }
            A.out_read_off[A.read_base + r] = ((uint64_t)tile << 32) | rank;   // fixed up by ka_finalize_kernel
        }
        const uint32_t qn = sm.qn;
        for (uint32_t qi = lane; qi < qn; qi += 32) {
            const uint32_t p = sm.u.post.queue[qi];
            if (p == Q_DROP) continue;
            const uint32_t rank = sm.u.post.accpre[p >> 5] + popc32(sm.u.post.ACC[p >> 5] & low_mask(p & 31u));
            const uint64_t o = obase + rank;
            if (o < A.stage_cap) {
                A.stage_hash[o] = sm.u.post.hq[qi];
                A.stage_pos[o] = sm.u.post.qpos[qi];
            }
        }
    } else if (lane == 0) {
        const uint32_t di = bs_atomic_add_g32(A.dirty_n, 1u);
        A.dirty_list[di] = (uint32_t)tile;
    }

    if (A.dbg && lane == 0) {          // per-tile state for the emulator-vs-GPU comparison
        uint32_t* d = A.dbg + tile * 8;
        d[0] = Ctile; d[1] = Ctotal; d[2] = (sm.flags & 1u) | (tile_bad ? 2u : 0u) | (dirty ? 4u : 0u);
        d[3] = sm.qn; d[4] = dirty ? 0u : sm.u.post.accpre[NW]; d[5] = sm.CA[0]; d[6] = sm.CB[0]; d[7] = sm.mraw[0];
    }
    // ---- what the tile below needs to know about this one --------------------------------------
    if (lane == 0) {
        Carry cy;
        cy.cnt = Ctile < 32u ? Ctile : 32u;
        cy.a = sm.CA[0] & low_mask(cy.cnt);
        cy.b = sm.CB[0] & low_mask(cy.cnt);
        cy.rs = sm.RS[0] & low_mask(cy.cnt);
        cy.bad = tile_bad ? 1u : 0u;
        cy.ok = (cy.cnt >= (uint32_t)(L - 1) || (Ctile == cy.cnt && t0 + TILE >= B)) ? 1u : 0u;
        sm.carry = cy;
    }
    bs_syncwarp();
}

// The persistent loop of one warp: claim a group of A.bs_group consecutive tiles, walk it top down.
template <int L, int T, bool HPC>
BS_DEV void warp_loop(const KAArgs& A, WarpSmem& sm, const CtaTables& ct, const int lane) {
    const uint64_t S = A.bs_group ? A.bs_group : 1;
    const uint64_t ntl = A.tile_end - A.tile_begin;
    const uint64_t ngroups = (ntl + S - 1) / S;
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = bs_atomic_add_g32(A.tile_counter, 1u);
        g = bs_shfl(g, 0);
        if ((uint64_t)g >= ngroups) break;
        const uint64_t tlo = A.tile_begin + (uint64_t)g * S;
        const uint64_t thi = (tlo + S < A.tile_end) ? tlo + S : A.tile_end;
        for (uint64_t tile = thi; tile-- > tlo;) process_tile<L, T, HPC>(A, sm, ct, lane, tile, tile + 1 == thi);
    }
}

}  // namespace bs
}  // namespace mdbg
