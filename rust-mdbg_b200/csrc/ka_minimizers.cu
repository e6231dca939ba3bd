// ka_minimizers.cu -- K-A: homopolymer compression + ntHash + density threshold, sm_100a.
//
// Replaces, for a whole batch of reads at once (reference file:line):
//   Read::encode_rle            src/read.rs:157-174   (run starts, raw position of each run)
//   nthash::NtHashIterator      crate, src/read.rs:196 (canonical ntHash of every HPC l-mer)
//   filter(<= hash_bound)+push  src/read.rs:183,196-208
//
// Layout.  The batch is ONE byte array (ASCII bases of all reads, concatenated) plus
// read_off[R+1].  The array is cut into fixed 16 KiB tiles on absolute offsets, so the grid
// does not depend on read lengths; a persistent CTA (128 threads) claims tiles from an atomic
// counter.  A tile is staged in shared memory with coalesced 128-bit loads into rows of 128
// bytes padded to 144 (conflict-free 128-bit reads with one row per thread).  Each thread
// then walks ITS 128 bytes sequentially, entirely in registers:
//   * run start  = byte differs from its predecessor (no compaction pass, no position array:
//     the raw position of a run IS the loop counter);
//   * the density test uses the 32-bit rolling filter of mdbg_common.cuh (2 shifts, 2 xors,
//     one 8-byte table read from a 128-byte, conflict-free shared table per run) -- only
//     ~0.7 % of the windows survive it;
//   * a thread owns the l-mers whose FIRST run starts in its segment and keeps walking past
//     its end until l-1 more runs are consumed (data-dependent halo, no fixed bound);
//   * survivors go to a shared queue and are re-evaluated exactly (64-bit, N-aware, walking
//     the runs backwards/forwards) by the whole CTA afterwards; hits set a bit per tile byte.
// A block-wide popcount scan of that bitmap plus a decoupled look-back across tiles (single
// pass, tiles are claimed in order) gives every minimizer its final position in the global,
// (read, position)-ordered output, and the per-read offsets fall out of the same prefix.
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

constexpr int NT = KA_THREADS;
constexpr int SEG = KA_SEG;
constexpr int TILE = KA_TILE;
constexpr int PRE = 128;
constexpr int HALO = 256;
constexpr int ROWS = (PRE + TILE + HALO) / 128;  // 131
constexpr int RSTRIDE = 144;
constexpr int WIN = ROWS * 128;
constexpr int QCAP = 1024;
constexpr int WORDS = TILE / 32;  // 512

constexpr uint64_t ST_AGG = 1ull << 62;
constexpr uint64_t ST_PREFIX = 2ull << 62;
constexpr uint64_t ST_VAL = (1ull << 62) - 1;

struct __align__(16) Smem {
    uint8_t raw[ROWS * RSTRIDE];
    uint32_t queue[QCAP];
    uint32_t bitmap[WORDS];
    uint32_t prefix[WORDS];
    uint2 tab[16];
    uint32_t warp_sum[NT / 32];
    uint32_t qn;
    uint32_t total;
    uint32_t tile;
    unsigned long long base;
};

struct Win {  // byte access through the staged window, falling back to global memory
    const uint8_t* raw;
    const uint8_t* g;
    int64_t w0;
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
// 0: fewer than l runs remain; 1: ok; 2: ok but an illegal byte was hashed (inv = its position).
template <bool HPC>
__device__ int lmer_hash(const Win& W, int64_t p0, int64_t re, uint32_t l, uint64_t& h, int64_t& inv) {
    uint64_t fh = 0, rh = 0;
    int64_t p = p0;
    int res = 1;
    for (uint32_t j = 0; j < l; j++) {
        if (p >= re) return 0;
        uint32_t c = W.byte(p);
        bool ok;
        uint64_t hv = nt_fwd(c, ok);
        uint64_t rv = nt_rc(c);
        if (!ok && res == 1) { res = 2; inv = p; }
        fh ^= rol64(hv, l - 1 - j);
        rh ^= rol64(rv, j);
        int64_t q = p + 1;
        if (HPC && collapsible(c))
            while (q < re && W.byte(q) == c) q++;
        p = q;
    }
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

__device__ __forceinline__ uint32_t badword(uint32_t w) {  // nonzero iff some byte is not A/C/G/T
    uint32_t t = (w >> 2) & ~(w >> 1);
    uint32_t d = w >> 4;
    uint32_t b2 = (d & (~t | w)) | (~d & (~w | t));
    return ((w & 0xE8E8E8E8u) ^ 0x40404040u) | (b2 & 0x01010101u);
}

__device__ __forceinline__ uint64_t ld_state(const uint64_t* p) {
    return *reinterpret_cast<const volatile uint64_t*>(p);
}
__device__ __forceinline__ void st_state(uint64_t* p, uint64_t v) {
    *reinterpret_cast<volatile uint64_t*>(p) = v;
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
__global__ void ka_tile_lb_kernel(const uint64_t* __restrict__ read_off, uint64_t R, uint64_t n_tiles,
                                  uint64_t* __restrict__ tile_lb) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
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
__global__ void __launch_bounds__(NT, 6) ka_minimizers_kernel(const KAArgs A) {
    __shared__ Smem sm;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t l = A.l;
    const uint64_t bound = A.bound;
    if (tid < 16) sm.tab[tid] = make_uint2(A.fc.tab[tid][0], A.fc.tab[tid][1]);
    const bool use_filter = A.fc.usable && !A.force_dense;
    const uint32_t SH = A.fc.hist_shift, fth = A.fc.f_thresh, gm = A.fc.g_mask, gth = A.fc.g_thresh;
    const uint8_t* __restrict__ gb = A.bases;
    const int64_t B = (int64_t)A.n_bases;

    for (;;) {
        __syncthreads();  // previous tile fully done with shared memory
        if (tid == 0) {
            sm.tile = atomicAdd(A.tile_counter, 1u);
            sm.qn = 0;
        }
        for (int i = tid; i < WORDS; i += NT) sm.bitmap[i] = 0;
        __syncthreads();
        const uint64_t tile = sm.tile;
        if (tile >= A.n_tiles) break;
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
        __syncthreads();
        Win W{sm.raw, gb, w0};

        const int64_t a = t0 + (int64_t)tid * SEG;
        const int64_t seg_end = (a + SEG < t1) ? a + SEG : t1;

        // Exact evaluation of every run start of [pb, plim) (slow path).
        auto exact_portion = [&](int64_t rs, int64_t re, int64_t pb, int64_t plim) {
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
                if (h <= bound) atomicOr(&sm.bitmap[(p - t0) >> 5], 1u << ((p - t0) & 31));
            }
        };

        // Filtered scan of one portion; falls back to exact_portion if a non-ACGT byte shows up.
        auto fast_portion = [&](int64_t rs, int64_t re, int64_t pb, int64_t plim) {
            uint32_t F = A.fc.f_init, G = A.fc.g_init, hist = 0, bad = 0;
#define MDBG_RUN_STEP(code, pos)                                                 \
    {                                                                            \
        hist = (hist << 2) | (code);                                             \
        uint2 tt = sm.tab[((hist >> SH) & 0xCu) | (code)];                       \
        F = (F << 1) ^ tt.x;                                                     \
        G = (G >> 1) ^ tt.y;                                                     \
        if (F <= fth || (G & gm) <= gth) {                                       \
            uint32_t qi = atomicAdd(&sm.qn, 1u);                                 \
            if (qi < QCAP) sm.queue[qi] = (uint32_t)((pos) - w0);                \
        }                                                                        \
    }
            uint32_t prevb;
            if (pb == a && plim == a + SEG) {
                // full, 16-byte aligned segment: 8 x LDS.128, everything else in registers
                const uint8_t* row = sm.raw + (tid + 1) * RSTRIDE;
                uint32_t first = row[0];
                uint32_t pw = ((pb > rs) ? W.byte(pb - 1) : (first ^ 0xFFu)) << 24;
#pragma unroll 2
                for (int c16 = 0; c16 < 8; c16++) {
                    // the 8 chunks of thread t sit at chunk index (t+1)*9 + c16 of the padded rows
                    uint4 v = *reinterpret_cast<const uint4*>(row + c16 * 16);
                    uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int wi = 0; wi < 4; wi++) {
                        uint32_t w = ws[wi];
                        uint32_t x = w ^ __funnelshift_l(pw, w, 8);
                        uint32_t cw = (w >> 1) & 0x03030303u;
                        bad |= badword(w);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            bool nr = !HPC || ((x >> (8 * j)) & 0xFFu) != 0;
                            uint32_t code = (cw >> (8 * j)) & 3u;
                            if (nr) MDBG_RUN_STEP(code, a + c16 * 16 + wi * 4 + j)
                        }
                        pw = w;
                    }
                }
                prevb = pw >> 24;
            } else {
                prevb = (pb > rs) ? W.byte(pb - 1) : 0x100u;
                for (int64_t p = pb; p < plim; p++) {
                    uint32_t c = W.byte(p);
                    bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                    if (!HPC || c != prevb) MDBG_RUN_STEP((c >> 1) & 3u, p)
                    prevb = c;
                }
            }
            // data-dependent halo: l-1 more runs complete the l-mers that start in the portion
            uint32_t rem = l - 1;
            for (int64_t p = plim; p < re && rem > 0; p++) {
                uint32_t c = W.byte(p);
                bad |= !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
                if (!HPC || c != prevb) { MDBG_RUN_STEP((c >> 1) & 3u, p) rem--; }
                prevb = c;
            }
#undef MDBG_RUN_STEP
            if (bad) exact_portion(rs, re, pb, plim);
        };

        auto for_each_portion = [&](auto&& fn) {
            if (a >= t1) return;
            uint64_t r = find_read(A.read_off, rlo, rhi, a);
            int64_t p = a;
            while (p < seg_end) {
                int64_t rs = (int64_t)__ldg(A.read_off + r), re = (int64_t)__ldg(A.read_off + r + 1);
                if (re <= p) { r++; continue; }
                int64_t plim = seg_end < re ? seg_end : re;
                fn(rs, re, p, plim);
                p = plim;
                if (p == re) r++;
            }
        };

        // ---- phase A: scan --------------------------------------------------------------
        if (use_filter) for_each_portion(fast_portion);
        else for_each_portion(exact_portion);
        __syncthreads();
        if (use_filter) {
            const uint32_t qn = sm.qn;
            if (qn > QCAP) {  // low-complexity sequence flooded the queue: exact path for the tile
                if (tid == 0) atomicAdd(A.dense_tiles, 1u);
                for_each_portion(exact_portion);
            } else {
                // ---- phase B: exact re-evaluation of the survivors -------------------------
                for (uint32_t qi = tid; qi < qn; qi += NT) {
                    int64_t pe = w0 + (int64_t)sm.queue[qi];
                    // the read containing pe (pe may lie past this tile: widen the range)
                    uint64_t hi2 = rhi;
                    if (pe >= t1) hi2 = A.n_reads;
                    uint64_t r = find_read(A.read_off, rlo, hi2, pe);
                    int64_t rs = (int64_t)__ldg(A.read_off + r), re = (int64_t)__ldg(A.read_off + r + 1);
                    int64_t p0;
                    if (!walk_back<HPC>(W, pe, rs, l, p0)) continue;
                    if (p0 < t0 || p0 >= t1) continue;  // owned by another tile
                    uint64_t h;
                    int64_t inv = 0;
                    int res = lmer_hash<HPC>(W, p0, re, l, h, inv);
                    if (res == 0) continue;
                    if (res == 2) atomicMin(A.err_pos, (unsigned long long)inv);
                    if (h <= bound) atomicOr(&sm.bitmap[(p0 - t0) >> 5], 1u << ((p0 - t0) & 31));
                }
            }
        } else if (tid == 0) {
            atomicAdd(A.dense_tiles, 1u);
        }
        __syncthreads();

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
        if (lane == 31) sm.warp_sum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int i = 0; i < NT / 32; i++) {
            uint32_t s = sm.warp_sum[i];
            if (i < warp) woff += s;
            total += s;
        }
        uint32_t ex = woff + inc - cnt;
#pragma unroll
        for (int i = 0; i < 4; i++) { sm.prefix[tid * 4 + i] = ex; ex += __popc(wv[i]); }

        // ---- decoupled look-back across tiles (warp 0) ------------------------------------
        if (warp == 0) {
            if (lane == 0) st_state(A.tile_state + tile, (tile == 0 ? ST_PREFIX : ST_AGG) | total);
            uint64_t excl = 0;
            if (tile > 0) {
                int64_t look = (int64_t)tile - 1;
                for (;;) {
                    int64_t idx = look - lane;
                    uint64_t s = idx >= 0 ? ld_state(A.tile_state + idx) : ST_PREFIX;
                    uint32_t flag = (uint32_t)(s >> 62);
                    uint32_t inval = __ballot_sync(0xffffffffu, flag == 0);
                    uint32_t pref = __ballot_sync(0xffffffffu, flag == 2);
                    int fp = pref ? __ffs(pref) - 1 : 32;
                    uint32_t need = fp >= 31 ? 0xffffffffu : ((2u << fp) - 1u);
                    if (inval & need) continue;  // a predecessor has not published yet
                    uint64_t v = (lane <= fp) ? (s & ST_VAL) : 0;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                    excl += v;
                    if (fp < 32) break;
                    look -= 32;
                }
                if (lane == 0) st_state(A.tile_state + tile, ST_PREFIX | (excl + total));
            }
            if (lane == 0) {
                sm.base = excl;
                sm.total = total;
                if (tile + 1 == A.n_tiles) *A.total_out = A.out_base + excl + total;
            }
        }
        __syncthreads();
        const uint64_t obase = A.out_base + sm.base;

        // ---- emit: per-read offsets, then (hash, pos) at the final positions ---------------
        for (uint64_t r = lb + tid; r < lbn; r += NT) {
            int64_t x = (int64_t)__ldg(A.read_off + r) - t0;  // 0 <= x <= TILE
            uint32_t rank;
            if (x >= TILE) rank = sm.total;
            else rank = sm.prefix[x >> 5] + __popc(sm.bitmap[x >> 5] & ((1u << (x & 31)) - 1u));
            A.out_read_off[A.read_base + r] = obase + rank;
        }
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
                if (o < A.out_cap) {
                    A.out_hash[o] = h;
                    A.out_pos[o] = (uint32_t)(p0 - rs);
                }
                rank++;
            }
        }
    }
}

// ---- host launchers ------------------------------------------------------------------------
cudaError_t ka_launch(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches) {
    uint64_t nt = A.n_tiles;
    cudaError_t e = cudaMemsetAsync(A.tile_state, 0, sizeof(uint64_t) * nt, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(A.tile_counter, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    unsigned nb = (unsigned)((nt + 1 + 255) / 256);
    ka_tile_lb_kernel<<<nb, 256, 0, st>>>(A.read_off, A.n_reads, nt, A.tile_lb);
    if (hpc) ka_minimizers_kernel<true><<<grid, NT, 0, st>>>(A);
    else ka_minimizers_kernel<false><<<grid, NT, 0, st>>>(A);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

int ka_max_blocks_per_sm(int hpc) {
    int n = 0;
    if (hpc) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka_minimizers_kernel<true>, NT, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka_minimizers_kernel<false>, NT, 0);
    return n;
}

}  // namespace mdbg
