// graph_kernels.cuh -- K-B .. K-E device code (windowing, node table, segmented reduce, edges).
//
// Replaces (reference file:line):
//   K-B  windowing loop + KmerVec::normalize     src/main.rs:756-781, src/kmer_vec.rs:34-42
//   K-C  DashMap<Kmer,DbgEntry> insert/count     src/main.rs:595,657-686   (open-address table,
//        4-lane cooperative probing of one 32-byte sector, atomicCAS claim, atomicMin first sighting)
//   K-D  abundance / representative / index      src/main.rs:662,680-684,696,922-929 (radix sort by
//        slot + segmented reduce under serial-order semantics)
//   K-E  km_index + 4-orientation test + presimp src/main.rs:1015-1117
// All integer work; identity of tuples is always decided on the tuples themselves, fingerprints
// only place them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdbg_common.cuh"

namespace mdbg {

constexpr uint64_t KC_EMPTY = ~0ull;

struct MinArena {  // resident minimizers of all pushed reads (global read order)
    const uint64_t* hash;
    const uint32_t* pos;
    const uint64_t* off;   // [R+1]
    uint64_t R;
};

// last r in [0, R) with a[r] <= g < a[r+1]  (a = exclusive prefix array of R+1 entries, a[R] > g)
__device__ __forceinline__ uint64_t owner_read(const uint64_t* __restrict__ a, uint64_t R, uint64_t g) {
    uint64_t lo = 0, hi = R;  // invariant: a[lo] <= g, answer in [lo, hi)
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src, uint32_t mask = 0xffffffffu) {
    const uint32_t lo = __shfl_sync(mask, (uint32_t)v, src), hi = __shfl_sync(mask, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}

// ---- K-B ---------------------------------------------------------------------------------------
// cnt[r] = m > k ? m-k+1 : 0  (strict, main.rs:756); cnt[R] = 0 so the exclusive scan ends with K.
__global__ void kb_count_kernel(const uint64_t* __restrict__ m_off, uint64_t R, uint32_t k,
                                uint64_t* __restrict__ cnt) {
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > R) return;
    uint64_t c = 0;
    if (r < R) {
        uint64_t m = m_off[r + 1] - m_off[r];
        c = m > k ? m - k + 1 : 0;
    }
    cnt[r] = c;
}

// canonical orientation of the window h[0..k): reversed unless fwd < rev lexicographically
__device__ __forceinline__ bool window_reversed(const uint64_t* __restrict__ h, uint32_t k) {
    for (uint32_t j = 0; j < k / 2; j++) {
        uint64_t a = __ldg(h + j), b = __ldg(h + k - 1 - j);
        if (a != b) return !(a < b);
    }
    return true;  // palindrome => reversed (kmer_vec.rs:37-38)
}

// Record of one k-min-mer sighting, as exchanged between GPUs (SoA: fp / ord / wloc / info, 44 bytes).
//   fp    : table fingerprint of the canonical tuple (places it; identity is decided on the tuples)
//   ord   : global ordinal of the sighting (serial (read, i) order over the whole job), bit 63 =
//           the window was reversed
//   wloc  : first element of the window in the job-wide hash arena (the tuple itself)
//   info  : what add_kminmer needs from the sighting (main.rs:769-778)
struct RecInfo {
    uint32_t p0;      // raw position of the first minimizer           (read_offsets.0)
    uint32_t d01;     // p[i+1] - p[i]
    uint32_t dlast;   // p[i+k-1] - p[i+k-2]
    uint32_t span;    // p[i+k-1] - p[i]   (seqlen = span + 2, end = p0 + span + l)
    uint64_t read;    // global read index
};
static_assert(sizeof(RecInfo) == 24, "RecInfo is moved as three u64 words");
constexpr uint64_t ORD_REV = 1ull << 63;
constexpr uint64_t ORD_MASK = ORD_REV - 1;

// Where the canonical tuple of record j lives: nowhere of its own.  A record is a window of the
// minimizer hash arena (wloc[j], read backwards when ord[j] says reversed), so the K x k x 8 bytes of
// tuples are never materialised.  With N GPUs `hash` is the all-gathered arena of the whole job (8 B per
// minimizer), so this also holds for the owner of a tuple whose sightings came from other GPUs.
struct TupleSrc {
    const uint64_t* hash;   // arena hashes
    const uint32_t* wloc;   // [K] first arena element of the window
    const uint64_t* ord;    // [K] bit 63 = window is reversed
    uint32_t k;
    __device__ __forceinline__ void row(uint64_t j, const uint64_t*& p, int& step) const {
        bool rv = (__ldg(ord + j) & ORD_REV) != 0;
        p = hash + __ldg(wloc + j) + (rv ? k - 1 : 0);
        step = rv ? -1 : 1;
    }
};

// One thread per k-min-mer sighting of THIS GPU's reads (window g of the local arena, serial (read, i)
// order): orientation, job-wide ordinal (kbase + g: ranks hold contiguous read ranges, so ordinals
// of rank r start at the number of sightings of the ranks before it), RecInfo, the window's location
// in the job-wide arena (mbase + local index: with N GPUs the hash arenas are all-gathered at a fixed
// pitch, mbase = rank * pitch) and the table fingerprint of the first seed (masked, never KC_EMPTY).
// N > 1: owner[j] = GPU that counts this tuple = range partition of the UNMASKED fingerprint
// (mdbg_owner_of_fingerprint), so every copy of a tuple meets on one owner.
__global__ void kb_records_kernel(MinArena A, const uint64_t* __restrict__ kmer_off, uint64_t K, uint32_t k,
                                  uint64_t seed, uint64_t fp_mask, uint64_t read_base, uint64_t kbase,
                                  uint64_t mbase, uint32_t world,
                                  uint32_t* __restrict__ wloc, uint64_t* __restrict__ ord,
                                  RecInfo* __restrict__ info, uint64_t* __restrict__ fp,
                                  uint32_t* __restrict__ iota, uint8_t* __restrict__ owner) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= K) return;
    uint64_t r = owner_read(kmer_off, A.R, g);
    uint64_t i = g - __ldg(kmer_off + r);
    uint64_t lo = __ldg(A.off + r) + i;
    const uint64_t* h = A.hash + lo;
    const uint32_t* p = A.pos + lo;
    bool rv = window_reversed(h, k);
    uint64_t f = fp_init(seed, k);
    for (uint32_t q = 0; q < k; q++) f = fp_mix(f, rv ? __ldg(h + k - 1 - q) : __ldg(h + q));
    f = fp_fin(f);
    if (owner) owner[g] = (uint8_t)__umul64hi(f, (uint64_t)world);
    f &= fp_mask;
    if (f == KC_EMPTY) f = KC_EMPTY - 1;
    fp[g] = f;
    if (iota) iota[g] = (uint32_t)g;
    wloc[g] = (uint32_t)(mbase + lo);
    ord[g] = (kbase + g) | (rv ? ORD_REV : 0);
    RecInfo ri;
    ri.p0 = p[0];
    ri.d01 = p[1] - p[0];
    ri.dlast = p[k - 1] - p[k - 2];
    ri.span = p[k - 1] - p[0];
    ri.read = read_base + r;
    info[g] = ri;
}

// ---- the exchange (N > 1): records bucketed by owner, one all-to-all over NVLink ----------------
// {M, R, first read, K} of this rank, assembled on the device (K is the last entry of a device scan)
__global__ void kx_sizes_kernel(uint64_t* __restrict__ out, uint64_t M, uint64_t R, uint64_t read_base,
                                const uint64_t* __restrict__ k_total) {
    out[0] = M; out[1] = R; out[2] = read_base; out[3] = *k_total;
}
// cnt[w] = records whose owner is w, from the owner-sorted key array (W + 1 binary searches)
__global__ void kx_bounds_kernel(const uint8_t* __restrict__ sorted, uint64_t K, uint32_t W, uint64_t* __restrict__ cnt) {
    uint32_t w = threadIdx.x;
    if (w >= W) return;
    uint64_t b[2];
    for (int t = 0; t < 2; t++) {
        uint32_t key = w + t;
        uint64_t lo = 0, hi = K;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (__ldg(sorted + mid) < key) lo = mid + 1; else hi = mid; }
        b[t] = lo;
    }
    cnt[w] = b[1] - b[0];
}
// Bucketing fused with the exchange: every record is written STRAIGHT into its owner's inbox over NVLink (the
// inboxes of the peers are mapped into this process with CUDA IPC; `box[w]` is rank w's inbox as seen from here).
// Record i of the owner-sorted order goes to slot dst_off[w] + (i - bstart[w]) of owner w: dst_off[w] = records
// the lower ranks send to w, so an owner's inbox fills in rank order = ascending ordinal.  Consecutive threads
// write consecutive remote addresses (coalesced 8/4-byte stores through the NVSwitch).
struct InboxLayout {        // SoA inside one allocation of `cap` records
    uint64_t cap;
    __host__ __device__ uint64_t* fp(void* b) const { return (uint64_t*)b; }
    __host__ __device__ uint64_t* ord(void* b) const { return (uint64_t*)b + cap; }
    __host__ __device__ RecInfo* info(void* b) const { return (RecInfo*)((uint64_t*)b + 2 * cap); }
    __host__ __device__ uint32_t* wloc(void* b) const { return (uint32_t*)((uint64_t*)b + 5 * cap); }
    __host__ __device__ static uint64_t bytes(uint64_t cap) { return cap * 44 + 64; }
};
constexpr int KX_MAX_WORLD = 16;
struct ScatterPlan {
    void* box[KX_MAX_WORLD];
    uint64_t bstart[KX_MAX_WORLD + 1];   // first record of bucket w in the owner-sorted order
    uint64_t dst_off[KX_MAX_WORLD];
};
__global__ void kx_scatter_kernel(const uint32_t* __restrict__ perm, const uint8_t* __restrict__ owner_s, uint64_t K,
                                  const uint64_t* __restrict__ fp, const uint64_t* __restrict__ ord,
                                  const uint32_t* __restrict__ wloc, const RecInfo* __restrict__ info,
                                  const ScatterPlan P, const InboxLayout L, uint64_t rot) {
    // Thread t takes record (t + rot) mod K: rank r starts with the bucket of rank r + 1, so that at any moment the
    // GPUs of the job store to DIFFERENT owners (in plain bucket order everybody would write to owner 0 first, then
    // to owner 1, ...: an incast that serialises the exchange on one NVLink ingress at a time).
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool valid = t < K;
    const uint64_t i = t + rot < K ? t + rot : t + rot - K;
    uint64_t ibase = 0, v0 = 0, v1 = 0, v2 = 0;
    if (valid) {
        const uint32_t w = owner_s[i], j = __ldg(perm + i);
        const uint64_t d = P.dst_off[w] + (i - P.bstart[w]);
        void* b = P.box[w];
        L.fp(b)[d] = fp[j];
        L.ord(b)[d] = ord[j];
        L.wloc(b)[d] = wloc[j];
        const RecInfo ri = info[j];
        v0 = (uint64_t)ri.p0 | ((uint64_t)ri.d01 << 32);
        v1 = (uint64_t)ri.dlast | ((uint64_t)ri.span << 32);
        v2 = ri.read;
        ibase = (uint64_t)(uintptr_t)(L.info(b) + d);
    }
    // RecInfo is 24 bytes: stored per thread it would leave the SM as 8/16-byte pieces at stride 24 (three quarters of
    // every NVLink packet empty).  The warp's 32 records are 96 words: lane t stores words t, t + 32, t + 64, so the
    // records of one bucket leave as three contiguous 256-byte stores.
#pragma unroll
    for (int part = 0; part < 3; part++) {
        const uint32_t q = part * 32u + lane, r = q / 3u, cidx = q - 3u * r;
        const uint64_t a = shfl64(ibase, (int)r);
        const uint64_t x0 = shfl64(v0, (int)r), x1 = shfl64(v1, (int)r), x2 = shfl64(v2, (int)r);
        if (a) reinterpret_cast<uint64_t*>((uintptr_t)a)[cidx] = cidx == 0 ? x0 : (cidx == 1 ? x1 : x2);
    }
}

// send buffers: the records in owner order (stable, so every bucket ascends in ordinal)
__global__ void kx_pack_kernel(const uint32_t* __restrict__ perm, uint64_t K, const uint64_t* __restrict__ fp,
                               const uint64_t* __restrict__ ord, const uint32_t* __restrict__ wloc,
                               const RecInfo* __restrict__ info, uint64_t* __restrict__ s_fp,
                               uint64_t* __restrict__ s_ord, uint32_t* __restrict__ s_wloc, RecInfo* __restrict__ s_info) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= K) return;
    uint32_t j = __ldg(perm + i);
    s_fp[i] = fp[j]; s_ord[i] = ord[j]; s_wloc[i] = wloc[j]; s_info[i] = info[j];
}

// table fingerprint of the canonical tuples
__global__ void kc_fp_kernel(TupleSrc T, uint64_t K, uint64_t seed, uint64_t fp_mask, uint64_t* __restrict__ fp,
                             uint32_t* __restrict__ iota) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= K) return;
    const uint64_t* t; int step;
    T.row(j, t, step);
    uint64_t f = fp_init(seed, T.k);
    for (uint32_t q = 0; q < T.k; q++) f = fp_mix(f, __ldg(t + (int64_t)q * step));
    f = fp_fin(f);
    f &= fp_mask;
    if (f == KC_EMPTY) f = KC_EMPTY - 1;
    fp[j] = f;
    iota[j] = (uint32_t)j;
}

// export form for mdbg_window (Entry 2): canonical tuple, reversed, shift pair, read_offsets
__global__ void kb_export_kernel(MinArena A, const uint64_t* __restrict__ kmer_off, uint64_t K, uint32_t k,
                                 uint32_t l, uint64_t* __restrict__ out_tuple, uint8_t* __restrict__ out_rev,
                                 uint64_t* __restrict__ out_shift, uint64_t* __restrict__ out_offsets) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= K) return;
    uint64_t r = owner_read(kmer_off, A.R, g);
    uint64_t i = g - kmer_off[r];
    uint64_t lo = A.off[r] + i;
    const uint64_t* h = A.hash + lo;
    const uint32_t* p = A.pos + lo;
    bool rv = window_reversed(h, k);
    if (out_tuple)
        for (uint32_t j = 0; j < k; j++) out_tuple[g * k + j] = rv ? h[k - 1 - j] : h[j];
    if (out_rev) out_rev[g] = rv ? 1 : 0;
    uint64_t a = (uint64_t)p[1] - p[0], b = (uint64_t)p[k - 1] - p[k - 2];
    if (out_shift) { out_shift[2 * g] = rv ? b : a; out_shift[2 * g + 1] = rv ? a : b; }  // main.rs:769-777
    if (out_offsets) {  // main.rs:778
        out_offsets[3 * g] = p[0];
        out_offsets[3 * g + 1] = (uint64_t)p[k - 1] + l;
        out_offsets[3 * g + 2] = (uint64_t)p[k - 1] + 1 - p[0] + 1;
    }
}

// ---- --read-stats (main.rs:939-975): abundance of the k-min-mers of a second read set among the kept nodes
// fingerprint of every node's tuple (nodes are sorted by index, the lookup wants them by fingerprint)
__global__ void rs_node_fp_kernel(const uint64_t* __restrict__ tuple, uint32_t S, uint32_t k, uint64_t seed,
                                  uint64_t* __restrict__ fp, uint32_t* __restrict__ pos) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    uint64_t f = fp_init(seed, k);
    for (uint32_t q = 0; q < k; q++) f = fp_mix(f, __ldg(tuple + (uint64_t)j * k + q));
    f = fp_fin(f);
    fp[j] = f;
    pos[j] = j;
}
// One thread per window of the query reads: canonical orientation, fingerprint, binary search among the
// sorted node fingerprints, then the tuples decide (a fingerprint never does).  0 when the tuple is no node.
__global__ void rs_lookup_kernel(MinArena A, const uint64_t* __restrict__ kmer_off, uint64_t K, uint32_t k,
                                 uint64_t seed, const uint64_t* __restrict__ sfp, const uint32_t* __restrict__ spos,
                                 uint32_t S, const uint64_t* __restrict__ node_tuple,
                                 const uint16_t* __restrict__ node_ab, uint32_t* __restrict__ out) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= K) return;
    const uint64_t r = owner_read(kmer_off, A.R, g);
    const uint64_t* h = A.hash + A.off[r] + (g - kmer_off[r]);
    const bool rv = window_reversed(h, k);
    uint64_t f = fp_init(seed, k);
    for (uint32_t q = 0; q < k; q++) f = fp_mix(f, __ldg(rv ? h + (k - 1 - q) : h + q));
    f = fp_fin(f);
    uint32_t lo = 0, hi = S;           // lower bound of f in sfp
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(sfp + mid) < f) lo = mid + 1; else hi = mid;
    }
    uint32_t res = 0;
    for (uint32_t t = lo; t < S && __ldg(sfp + t) == f; t++) {
        const uint64_t* nt = node_tuple + (uint64_t)__ldg(spos + t) * k;
        bool eq = true;
        for (uint32_t q = 0; q < k && eq; q++) eq = __ldg(nt + q) == __ldg(rv ? h + (k - 1 - q) : h + q);
        if (eq) { res = __ldg(node_ab + __ldg(spos + t)); break; }
    }
    out[g] = res;
}

// ---- K-C ---------------------------------------------------------------------------------------
// Open-address table of 64-bit fingerprints.  Four lanes cooperate on one key: they read one
// aligned 32-byte sector (4 slots) per probe, vote with ballot, and the lane holding the first
// empty slot claims it with atomicCAS.  first[slot] = smallest ordinal that carries the key.
// A group of four lanes takes KC_U consecutive keys and works on them in phases -- all first probes, then all
// claims, then the few keys that need another look, one at a time -- so that KC_U sector reads and KC_U atomics
// of a group are in flight together (one key at a time, a group waits for a DRAM round trip twice per key).
constexpr int KC_U = 4;
__global__ void kc_insert_kernel(const uint64_t* __restrict__ fp, uint64_t K, uint64_t* keys,
                                 uint32_t* first, uint64_t cap_mask, uint32_t* __restrict__ slot_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sub = lane & 3, gbase = lane & ~3u;
    const uint64_t item0 = ((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 2) * KC_U;
    uint64_t key[KC_U], base[KC_U], cur[KC_U], slot[KC_U];
    bool done[KC_U];
#pragma unroll
    for (int u = 0; u < KC_U; u++) {
        const bool active = item0 + u < K;
        key[u] = active ? __ldg(fp + item0 + u) : 0;
        base[u] = ((key[u] * 0x9e3779b97f4a7c15ULL) >> 17) & cap_mask & ~3ull;
        done[u] = !active;
        slot[u] = 0;
    }
#pragma unroll
    for (int u = 0; u < KC_U; u++) cur[u] = done[u] ? 0 : *reinterpret_cast<volatile uint64_t*>(keys + base[u] + sub);
    uint32_t m_match[KC_U], m_empty[KC_U], leader[KC_U];
    uint64_t old[KC_U];
#pragma unroll
    for (int u = 0; u < KC_U; u++) {
        m_match[u] = (__ballot_sync(0xffffffffu, !done[u] && cur[u] == key[u]) >> gbase) & 0xFu;
        m_empty[u] = (__ballot_sync(0xffffffffu, !done[u] && cur[u] == KC_EMPTY) >> gbase) & 0xFu;
        leader[u] = m_empty[u] ? (uint32_t)__ffs(m_empty[u]) - 1 : 0;
        old[u] = 0;
        if (!done[u] && !m_match[u] && m_empty[u] && sub == leader[u])
            old[u] = atomicCAS(reinterpret_cast<unsigned long long*>(keys + base[u] + leader[u]), KC_EMPTY, key[u]);
    }
#pragma unroll
    for (int u = 0; u < KC_U; u++) {
        old[u] = shfl64(old[u], (int)(gbase + leader[u]));
        if (!done[u]) {
            if (m_match[u]) { slot[u] = base[u] + (uint32_t)__ffs(m_match[u]) - 1; done[u] = true; }
            else if (m_empty[u]) {
                if (old[u] == KC_EMPTY || old[u] == key[u]) { slot[u] = base[u] + leader[u]; done[u] = true; }
                // else: somebody else took that slot for another key; look at the window again
            } else base[u] = (base[u] + 4) & cap_mask;
        }
    }
    // what is left (a full window, a lost race): the same steps, one key at a time
#pragma unroll
    for (int u = 0; u < KC_U; u++) {
        while (__any_sync(0xffffffffu, !done[u])) {
            const uint64_t c = done[u] ? 0 : *reinterpret_cast<volatile uint64_t*>(keys + base[u] + sub);
            const uint32_t mm = (__ballot_sync(0xffffffffu, !done[u] && c == key[u]) >> gbase) & 0xFu;
            const uint32_t me = (__ballot_sync(0xffffffffu, !done[u] && c == KC_EMPTY) >> gbase) & 0xFu;
            const uint32_t ld = me ? (uint32_t)__ffs(me) - 1 : 0;
            uint64_t o = 0;
            if (!done[u] && !mm && me && sub == ld)
                o = atomicCAS(reinterpret_cast<unsigned long long*>(keys + base[u] + ld), KC_EMPTY, key[u]);
            o = shfl64(o, (int)(gbase + ld));
            if (!done[u]) {
                if (mm) { slot[u] = base[u] + (uint32_t)__ffs(mm) - 1; done[u] = true; }
                else if (me) {
                    if (o == KC_EMPTY || o == key[u]) { slot[u] = base[u] + ld; done[u] = true; }
                } else base[u] = (base[u] + 4) & cap_mask;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KC_U; u++)
        if (item0 + u < K && sub == (uint32_t)u % 4u) {
            atomicMin(first + slot[u], (uint32_t)(item0 + u));
            slot_out[item0 + u] = (uint32_t)slot[u];
        }
}

// exactness: every record must carry the same TUPLE as the first record of its slot
__global__ void kc_verify_kernel(TupleSrc T, uint64_t K, const uint32_t* __restrict__ slot,
                                 const uint32_t* __restrict__ first, unsigned long long* collisions) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= K) return;
    uint32_t f = __ldg(first + __ldg(slot + j));
    if (f == j) return;
    const uint64_t *a, *b; int sa, sb;
    T.row(j, a, sa);
    T.row(f, b, sb);
    // four elements per step without a branch in between: eight loads in flight instead of two (the windows are
    // equal in all but ~2^-40 of the cases, so leaving early buys nothing)
    uint64_t diff = 0;
    uint32_t q = 0;
    for (; q + 4 <= T.k; q += 4) {
        const uint64_t a0 = __ldg(a + (int64_t)q * sa), a1 = __ldg(a + (int64_t)(q + 1) * sa);
        const uint64_t a2 = __ldg(a + (int64_t)(q + 2) * sa), a3 = __ldg(a + (int64_t)(q + 3) * sa);
        const uint64_t b0 = __ldg(b + (int64_t)q * sb), b1 = __ldg(b + (int64_t)(q + 1) * sb);
        const uint64_t b2 = __ldg(b + (int64_t)(q + 2) * sb), b3 = __ldg(b + (int64_t)(q + 3) * sb);
        diff |= (a0 ^ b0) | (a1 ^ b1) | (a2 ^ b2) | (a3 ^ b3);
    }
    for (; q < T.k; q++) diff |= __ldg(a + (int64_t)q * sa) ^ __ldg(b + (int64_t)q * sb);
    if (diff) atomicAdd(collisions, 1ull);
}

// ---- K-D ---------------------------------------------------------------------------------------
__global__ void kd_heads_kernel(const uint32_t* __restrict__ sslot, uint64_t K, uint8_t* __restrict__ head) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= K) return;
    head[j] = (j == 0 || sslot[j] != sslot[j - 1]) ? 1 : 0;
}

// One thread per distinct tuple = segment of the slot-sorted records (ascending ordinal inside,
// because records arrive in ordinal order and the sort is stable).
//   first_ord : ordinal of the first sighting (it consumed a node index, main.rs:662)
//   solid     : abundance(u16) >= minabund (main.rs:922-929)
//   nseq      : sightings with previous_abundance == minabund-1 (u16 counter): main.rs:680,696
//   bf (main.rs:639-655, ideal filter): a tuple enters the table at its SECOND sighting, so the
//   index order is the order of second sightings and tuples seen once are not counted
// "How many tuples were first seen earlier" is a prefix sum over ordinal space.  ord_bits holds two
// bits per ordinal (16 ordinals per word): bit 0 = this ordinal consumed a node index, bit 1 = ... of
// a solid node.  N > 1: the bitmaps of the GPUs are summed (an ordinal belongs to one tuple, hence to
// one owner: the fields never collide, the sum is the union).
__global__ void kd_segments_kernel(const uint32_t* __restrict__ seg_start, uint32_t D, uint64_t K,
                                   const uint32_t* __restrict__ sj, const uint64_t* __restrict__ ord,
                                   uint32_t minab, uint32_t bf, uint64_t* __restrict__ first_ord,
                                   uint8_t* __restrict__ counted, uint8_t* __restrict__ solid,
                                   uint32_t* __restrict__ nseq, uint32_t* __restrict__ ord_bits) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D) return;
    uint32_t st = seg_start[s];
    uint32_t en = (s + 1 < D) ? seg_start[s + 1] : (uint32_t)K;
    uint32_t cnt = en - st;
    const bool in_table = !bf || cnt >= 2;
    const uint64_t fo = in_table ? (ord[sj[st + (bf ? 1 : 0)]] & ORD_MASK) : ORD_MASK;
    first_ord[s] = fo;
    counted[s] = in_table ? 1 : 0;
    uint32_t ab = cnt & 0xFFFFu;
    const bool sol = (minab == 1 || ab >= minab);
    solid[s] = sol ? 1 : 0;
    nseq[s] = cnt >= minab ? 1 + (cnt - minab) / 65536u : 0;
    if (in_table) atomicOr(ord_bits + (fo >> 4), (sol ? 3u : 1u) << (2 * (uint32_t)(fo & 15)));
}

// bitmap word -> packed u64 counters (low 32: index consumers, high 32: solid nodes) for ONE scan
struct FlagWordToU64 {
    __host__ __device__ __forceinline__ uint64_t operator()(uint32_t w) const {
#if defined(__CUDA_ARCH__)
        return (uint64_t)__popc(w & 0x55555555u) | ((uint64_t)__popc(w & 0xAAAAAAAAu) << 32);
#else
        return (uint64_t)__builtin_popcount(w & 0x55555555u) | ((uint64_t)__builtin_popcount(w & 0xAAAAAAAAu) << 32);
#endif
    }
};
// packed (node index | node position << 32) of the tuple first seen at ordinal fo: the exclusive scan of the
// bitmap words plus the fields below fo in its own word
__device__ __forceinline__ uint64_t ordinal_rank(const uint32_t* __restrict__ ord_bits,
                                                 const uint64_t* __restrict__ wscan, uint64_t fo) {
    const uint32_t w = __ldg(ord_bits + (fo >> 4)) & ((1u << (2 * (uint32_t)(fo & 15))) - 1u);
    return __ldg(wscan + (fo >> 4)) + ((uint64_t)__popc(w & 0x55555555u) | ((uint64_t)__popc(w & 0xAAAAAAAAu) << 32));
}

// What the other GPUs need to know about a node (20 bytes; the tuple is a window of the all-gathered hash
// arena, so it is not shipped): written by the owner at the node's final position n, zero elsewhere, and
// summed over the GPUs (N > 1).
struct NodeRec { uint32_t index, seqlen, wloc, ab_sh0, sh1_rev; };
static_assert(sizeof(NodeRec) == 20, "NodeRec is reduced as 5 u32 words");

// The exclusive scan of the (job-wide) ordinal bitmap gives, at a tuple's first sighting, its node
// index (low half) and -- because index order IS first-sighting order -- the position of a solid
// node in the ascending-index node list (high half): nodes are written in place, no sort, no
// binary searches.  One thread per distinct tuple of this GPU; also leaves every tuple's index in
// seg_index (for .sequences).
__global__ void kd_nodes_kernel(uint32_t D, uint32_t minab, uint64_t K,
                                const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ sj,
                                const uint64_t* __restrict__ first_ord, const uint8_t* __restrict__ counted,
                                const uint8_t* __restrict__ solid, const uint32_t* __restrict__ ord_bits,
                                const uint64_t* __restrict__ wscan, const uint32_t* __restrict__ wloc,
                                const uint64_t* __restrict__ ord, const RecInfo* __restrict__ info,
                                uint32_t* __restrict__ seg_index, NodeRec* __restrict__ nodes) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D) return;
    if (!counted[s]) { seg_index[s] = 0xFFFFFFFFu; return; }
    const uint64_t rk = ordinal_rank(ord_bits, wscan, first_ord[s]);
    const uint32_t index = (uint32_t)rk, n = (uint32_t)(rk >> 32);
    seg_index[s] = index;
    if (!solid[s]) return;
    uint32_t st = seg_start[s];
    uint32_t en = (s + 1 < D) ? seg_start[s + 1] : (uint32_t)K;
    uint32_t cnt = en - st;
    uint32_t rep_rank = (minab - 1) + 65536u * ((cnt - minab) / 65536u);  // last overwrite, main.rs:680-684
    uint32_t j = sj[st + rep_rank], j0 = sj[st];
    bool rv = (ord[j] & ORD_REV) != 0;
    RecInfo ri = info[j];
    NodeRec nr;
    nr.index = index;
    nr.seqlen = ri.span + 2;                                              // read_offsets.2, main.rs:778
    nr.wloc = wloc[j0];                                                   // the tuple: first sighting's window
    nr.ab_sh0 = (cnt & 0xFFFFu) | ((uint32_t)(uint16_t)(rv ? ri.dlast : ri.d01) << 16);   // lowprec_shift, main.rs:675
    nr.sh1_rev = (uint32_t)(uint16_t)(rv ? ri.d01 : ri.dlast) | ((ord[j0] & ORD_REV) ? 0x10000u : 0u);
    nodes[n] = nr;
}

// Node arrays of the whole job on this GPU: one thread per tuple element (coalesced), the tuple read from
// the hash arena in canonical orientation.
struct NodeOut { uint32_t* index; uint16_t* abundance; uint32_t* seqlen; uint16_t* shift; uint64_t* tuple; };
__global__ void kd_expand_kernel(const NodeRec* __restrict__ nodes, uint64_t S, uint32_t k,
                                 const uint64_t* __restrict__ hash, NodeOut O) {
    uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (e >= S * k) return;
    const uint64_t n = e / k;
    const uint32_t q = (uint32_t)(e - n * k);
    const NodeRec nr = nodes[n];
    const bool rv = (nr.sh1_rev & 0x10000u) != 0;
    O.tuple[e] = __ldg(hash + nr.wloc + (rv ? k - 1 - q : q));
    if (q == 0) {
        O.index[n] = nr.index;
        O.seqlen[n] = nr.seqlen;
        O.abundance[n] = (uint16_t)(nr.ab_sh0 & 0xFFFFu);
        O.shift[2 * n] = (uint16_t)(nr.ab_sh0 >> 16);
        O.shift[2 * n + 1] = (uint16_t)(nr.sh1_rev & 0xFFFFu);
    }
}

struct SeqRec { uint64_t ord; uint32_t index, pad; uint64_t read, start, end, shift0, shift1; };

// .sequences lines of this owner: one thread per segment, seq_off = exclusive scan of nseq
__global__ void kd_seqlines_kernel(uint32_t D, uint64_t K, uint32_t minab, uint32_t l,
                                   const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ sj,
                                   const uint32_t* __restrict__ nseq, const uint32_t* __restrict__ seq_off,
                                   const uint32_t* __restrict__ seg_index, const uint64_t* __restrict__ ord,
                                   const RecInfo* __restrict__ info, SeqRec* __restrict__ out) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D) return;
    uint32_t n = nseq[s];
    uint32_t st = seg_start[s];
    for (uint32_t t = 0; t < n; t++) {
        uint32_t j = sj[st + (minab - 1) + 65536u * t];
        uint64_t o = ord[j];
        bool rv = (o & ORD_REV) != 0;
        RecInfo ri = info[j];
        SeqRec q;
        q.ord = o;
        q.index = seg_index[s];
        q.pad = 0;
        q.read = ri.read;
        q.start = ri.p0;
        q.end = (uint64_t)ri.p0 + ri.span + l;
        q.shift0 = rv ? ri.dlast : ri.d01;      // untruncated (usize, usize), main.rs:702
        q.shift1 = rv ? ri.d01 : ri.dlast;
        out[seq_off[s] + t] = q;
    }
}

__global__ void kd_seq_keys_kernel(const SeqRec* __restrict__ rec, uint32_t Q, uint64_t* __restrict__ key,
                                   uint32_t* __restrict__ iota) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Q) return;
    key[n] = rec[n].ord & ORD_MASK;
    iota[n] = n;
}
__global__ void kd_seq_gather_kernel(const SeqRec* __restrict__ in, const uint32_t* __restrict__ perm, uint32_t n,
                                     SeqRec* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
struct SeqOut {
    uint32_t* index; uint64_t* read; uint64_t* start; uint64_t* end; uint8_t* reversed; uint64_t* shift;
};
__global__ void kd_unpack_seq_kernel(const SeqRec* __restrict__ rec, const uint32_t* __restrict__ perm, uint32_t Q,
                                     SeqOut O) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Q) return;
    SeqRec r = rec[perm[n]];
    O.index[n] = r.index; O.read[n] = r.read; O.start[n] = r.start; O.end[n] = r.end;
    O.reversed[n] = (r.ord & ORD_REV) ? 1 : 0;
    O.shift[2 * n] = r.shift0; O.shift[2 * n + 1] = r.shift1;
}

// ---- K-E ---------------------------------------------------------------------------------------
struct NodeView {
    const uint32_t* index; const uint16_t* abundance; const uint32_t* seqlen; const uint16_t* shift;
    const uint64_t* tuple; uint32_t S, k;
};

// entry 2n+w: w=0 prefix (k-1)-mer of node n, w=1 suffix; key = fingerprint of its normalised form
__global__ void ke_entries_kernel(NodeView N, uint64_t seed, uint64_t* __restrict__ ekey,
                                  uint32_t* __restrict__ eval, uint8_t* __restrict__ erev) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * N.S) return;
    uint32_t n = e >> 1, w = e & 1, k1 = N.k - 1;
    const uint64_t* t = N.tuple + (uint64_t)n * N.k + w;
    bool rv = true;
    for (uint32_t j = 0; j < k1 / 2; j++) {
        uint64_t a = t[j], b = t[k1 - 1 - j];
        if (a != b) { rv = !(a < b); break; }
    }
    uint64_t f = fp_init(seed, k1);
    for (uint32_t j = 0; j < k1; j++) f = fp_mix(f, rv ? t[k1 - 1 - j] : t[j]);
    f = fp_fin(f);
    ekey[e] = f >> 32;   // 32-bit bucket key (4 radix passes); equality is decided on the tuples
    eval[e] = e;
    erev[e] = rv ? 1 : 0;
}

struct EdgeRec { uint32_t n1, n2, ov; uint8_t o1, o2; };

__device__ __forceinline__ bool eq_range(const uint64_t* a, int sa, const uint64_t* b, int sb, uint32_t n) {
    // compares a[0], a[sa], a[2sa].. with b[0], b[sb], ..  (strides +1 / -1)
    for (uint32_t j = 0; j < n; j++)
        if (a[(int64_t)j * sa] != b[(int64_t)j * sb]) return false;
    return true;
}

// position of every entry in the key-sorted list (so a query finds its bucket without a search)
__global__ void ke_inverse_kernel(const uint32_t* __restrict__ sval, uint32_t n, uint32_t* __restrict__ inv) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n) inv[sval[m]] = m;
}

// EIGHT lanes per (n1, key) with key in [suffix_norm, prefix_norm] (main.rs:1050-1052): the lanes split the
// k-1 elements of every tuple comparison (a chain of dependent 8-byte loads when one thread walks it), vote
// with shuffles inside their group, and all of them keep the (identical) bookkeeping; lane 0 writes.
//   MODE 0 (capture): decide the edges once, park up to KE_CAP_E kept edges / KE_CAP_R presimp
//          removals of the query in fixed slots, write the packed counts (edges | removals << 32);
//          a query that needs more raises *overflow and the host falls back to MODE 1
//   MODE 1 (write): emit at the scanned packed offsets (exact, any bucket size)
constexpr uint32_t KE_CAP_E = 8, KE_CAP_R = 4;
constexpr uint32_t KE_LANES = 8;

// bit 0: the entry's normalised (k-1)-mer equals the query's; bits 1..4: the four orientation identities of
// main.rs:1062-1075 -- (+,+) n1.suffix == n2.prefix, (+,-) n1.suffix == rev_n2.prefix, (-,+) rev_n1.suffix ==
// n2.prefix, (-,-) rev_n1.suffix == rev_n2.prefix.  Every lane of the group returns the same word.
__device__ __forceinline__ uint32_t ke_compare(const uint64_t* __restrict__ t1, const uint64_t* __restrict__ t2,
                                               const uint64_t* __restrict__ qsub, bool qrev,
                                               const uint64_t* __restrict__ esub, bool er, uint32_t k,
                                               uint32_t sub, uint32_t gmask, bool own_entry) {
    const uint32_t k1 = k - 1;
    uint32_t ok = 31u;
    // six streams walked with a stride of +-1 element: base pointers and directions are fixed before the loop
    const uint64_t* qp = (qrev ? qsub + (k1 - 1) : qsub) + (qrev ? -(int64_t)sub : (int64_t)sub);
    const uint64_t* ep = (er ? esub + (k1 - 1) : esub) + (er ? -(int64_t)sub : (int64_t)sub);
    const int64_t qs = qrev ? -(int64_t)KE_LANES : (int64_t)KE_LANES, es = er ? -(int64_t)KE_LANES : (int64_t)KE_LANES;
    const uint64_t* a1p = t1 + 1 + sub;
    const uint64_t* a2p = t1 + k - 2 - sub;
    const uint64_t* c1p = t2 + sub;
    const uint64_t* c2p = t2 + k - 1 - sub;
    for (uint32_t base = 0; base < k1; base += KE_LANES) {   // same trip count for every lane of the group
        const uint32_t j = base + sub;
        if (j < k1) {
            const uint64_t qa = __ldg(qp), ea = __ldg(ep);
            const uint64_t a1 = __ldg(a1p), a2 = __ldg(a2p);
            const uint64_t c1 = __ldg(c1p), c2 = __ldg(c2p);
            if (qa != ea) ok &= ~1u;
            if (a1 != c1) ok &= ~2u;
            if (a1 != c2) ok &= ~4u;
            if (a2 != c1) ok &= ~8u;
            if (a2 != c2) ok &= ~16u;
        }
        qp += qs; ep += es; a1p += KE_LANES; a2p -= KE_LANES; c1p += KE_LANES; c2p -= KE_LANES;
        if (base == 0) {   // after the first 8 elements almost every foreign key is known: leave together
            uint32_t r = ok;
            r &= __shfl_xor_sync(gmask, r, 1);
            r &= __shfl_xor_sync(gmask, r, 2);
            r &= __shfl_xor_sync(gmask, r, 4);
            if ((r & 1u) == 0u) return 0u;
            // the query's own entry (every bucket holds it): the key is equal by construction, and a node is joined to
            // itself only if one of the four identities survives the first 8 elements (a periodic tuple)
            if (own_entry && (r & 30u) == 0u) return 1u;
        }
    }
    ok &= __shfl_xor_sync(gmask, ok, 1);
    ok &= __shfl_xor_sync(gmask, ok, 2);
    ok &= __shfl_xor_sync(gmask, ok, 4);
    return (ok & 1u) ? ok : 0u;
}

template <int MODE>
__global__ void ke_join_kernel(NodeView N, const uint64_t* __restrict__ ekey, const uint8_t* __restrict__ erev,
                               const uint64_t* __restrict__ skey, const uint32_t* __restrict__ sval,
                               const uint32_t* __restrict__ inv, float presimp, uint32_t q_lo, uint32_t q_n,
                               uint64_t* __restrict__ cnt, const uint64_t* __restrict__ off,
                               EdgeRec* __restrict__ edges, uint64_t* __restrict__ removed,
                               unsigned long long* __restrict__ overflow) {
    // queries [q_lo, q_lo + q_n) of the 2S (n1, key) pairs: the slice of nodes this GPU emits edges for
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t qq = (uint32_t)(tid / KE_LANES), sub = (uint32_t)(tid % KE_LANES);
    const uint32_t gmask = 0xFFu << ((threadIdx.x & 31u) & ~7u);
    const uint32_t S = N.S, k = N.k;
    if (qq >= q_n) return;      // whole groups leave together (blockDim is a multiple of 8)
    uint32_t q = q_lo + qq;
    uint32_t n1 = q >> 1, which = q & 1;
    uint32_t qe = 2 * n1 + (which == 0 ? 1 : 0);  // which 0: suffix entry, 1: prefix entry
    uint64_t kf = ekey[qe];
    // bucket [b0, b1) of equal fingerprints around the query's own entry in the sorted list
    const uint32_t own = inv[qe];
    uint32_t b0 = own, b1 = b0 + 1;
    while (b0 > 0 && skey[b0 - 1] == kf) b0--;
    while (b1 < 2 * S && skey[b1] == kf) b1++;
    const uint64_t* t1 = N.tuple + (uint64_t)n1 * k;
    const uint64_t* qsub = t1 + (qe & 1);
    bool qrev = erev[qe];
    uint32_t ab1 = N.abundance[n1];
    // pass 0: number of potential edges and max abundance; pass 1: decide each edge.  The comparison words
    // of the first 12 entries of the bucket are kept from pass 0 (5 bits each).
    uint32_t npot = 0, abmax = 0, ne = 0, nr = 0;
    uint64_t cache = 0;
    const uint64_t o64 = MODE == 1 ? off[qq] : 0;
    const uint32_t oe = MODE == 1 ? (uint32_t)o64 : qq * KE_CAP_E;
    const uint32_t orr = MODE == 1 ? (uint32_t)(o64 >> 32) : qq * KE_CAP_R;
    for (int pass = 0; pass < 2; pass++) {
        uint32_t abref = abmax < ab1 ? abmax : ab1;
        for (uint32_t e = b0; e < b1; e++) {
            uint32_t ent = sval[e];
            uint32_t n2 = ent >> 1;
            const uint64_t* t2 = N.tuple + (uint64_t)n2 * k;
            uint32_t t;
            if (pass == 1 && e - b0 < 12) t = (uint32_t)(cache >> (5 * (e - b0))) & 31u;
            else {
                t = ke_compare(t1, t2, qsub, qrev, t2 + (ent & 1), erev[ent] != 0, k, sub, gmask, e == own);
                if (pass == 0 && e - b0 < 12) cache |= (uint64_t)t << (5 * (e - b0));
            }
            if (!(t & 1u)) continue;   // another (k-1)-mer in the same fingerprint bucket
            uint32_t ab2 = N.abundance[n2];
            for (int o = 0; o < 4; o++) {
                if (!((t >> (1 + o)) & 1u)) continue;
                if (pass == 0) { npot++; abmax = ab2 > abmax ? ab2 : abmax; continue; }
                if (presimp > 0.0f && npot >= 2 && (float)ab2 < presimp * (float)abref) {  // main.rs:1086
                    if (sub == 0 && (MODE == 1 || nr < KE_CAP_R)) removed[orr + nr] = ((uint64_t)N.index[n1] << 32) | N.index[n2];
                    nr++;
                    continue;
                }
                if (sub == 0 && (MODE == 1 || ne < KE_CAP_E)) {
                    uint32_t sh = (o < 2) ? N.shift[2 * n1] : N.shift[2 * n1 + 1];
                    uint32_t a = N.seqlen[n1] - sh, b = N.seqlen[n2] - 1;  // main.rs:1091-1092
                    EdgeRec r;
                    r.n1 = N.index[n1]; r.n2 = N.index[n2]; r.ov = a < b ? a : b;
                    r.o1 = (o >> 1) & 1; r.o2 = o & 1;
                    edges[oe + ne] = r;
                }
                ne++;
            }
        }
        if (npot == 0) break;
    }
    if (MODE == 0 && sub == 0) {
        cnt[qq] = (uint64_t)ne | ((uint64_t)nr << 32);
        if (ne > KE_CAP_E || nr > KE_CAP_R) atomicAdd(overflow, 1ull);
    }
}

// The join emits the edges of node n1 contiguously and nodes in ascending index order, so the
// canonical order (n1, n2, o1, o2, overlap) only needs each node's few edges sorted among
// themselves: one thread per node, insertion sort in place (no device-wide radix sort).
__device__ __forceinline__ bool edge_less(const EdgeRec& a, const EdgeRec& b) {
    if (a.n2 != b.n2) return a.n2 < b.n2;
    if (a.o1 != b.o1) return a.o1 < b.o1;
    if (a.o2 != b.o2) return a.o2 < b.o2;
    return a.ov < b.ov;
}
__device__ __forceinline__ void edge_group_sort(EdgeRec* __restrict__ edges, uint32_t lo, uint32_t hi) {
    for (uint32_t a = lo + 1; a < hi; a++) {
        EdgeRec x = edges[a];
        uint32_t b = a;
        while (b > lo && edge_less(x, edges[b - 1])) { edges[b] = edges[b - 1]; b--; }
        edges[b] = x;
    }
}
// after MODE 1: off = packed exclusive offsets per query (2 per node)
__global__ void ke_group_sort_kernel(EdgeRec* __restrict__ edges, const uint64_t* __restrict__ off, uint32_t n_nodes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    edge_group_sort(edges, (uint32_t)off[2 * i], (uint32_t)off[2 * i + 2]);
}
// after MODE 0 without overflow: move the parked edges / removals of a node's two queries to their
// scanned places and sort the node's edges; one thread per node
__global__ void ke_compact_kernel(const EdgeRec* __restrict__ cap_e, const uint64_t* __restrict__ cap_r,
                                  const uint64_t* __restrict__ cnt, const uint64_t* __restrict__ off,
                                  uint32_t n_nodes, EdgeRec* __restrict__ edges, uint64_t* __restrict__ removed) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    for (uint32_t w = 0; w < 2; w++) {
        const uint32_t qq = 2 * i + w;
        const uint64_t c = cnt[qq], o = off[qq];
        const uint32_t ne = (uint32_t)c, nr = (uint32_t)(c >> 32), oe = (uint32_t)o, orr = (uint32_t)(o >> 32);
        for (uint32_t t = 0; t < ne; t++) edges[oe + t] = cap_e[qq * KE_CAP_E + t];
        for (uint32_t t = 0; t < nr; t++) removed[orr + t] = cap_r[qq * KE_CAP_R + t];
    }
    edge_group_sort(edges, (uint32_t)off[2 * i], (uint32_t)off[2 * i + 2]);
}

// keep[e] = 0 if (n1,n2) or (n2,n1) was presimp-removed (main.rs:1109)
__global__ void ke_filter_kernel(const EdgeRec* __restrict__ edges, uint32_t E, const uint64_t* __restrict__ rem,
                                 uint32_t NR, uint8_t* __restrict__ keep) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    EdgeRec r = edges[e];
    uint64_t k1 = ((uint64_t)r.n1 << 32) | r.n2, k2 = ((uint64_t)r.n2 << 32) | r.n1;
    bool found = false;
    for (int t = 0; t < 2 && !found; t++) {
        uint64_t key = t ? k2 : k1;
        uint32_t lo = 0, hi = NR;
        while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (rem[m] < key) lo = m + 1; else hi = m; }
        found = lo < NR && rem[lo] == key;
    }
    keep[e] = found ? 0 : 1;
}

struct EdgeOut { uint32_t* n1; uint8_t* o1; uint32_t* n2; uint8_t* o2; uint32_t* ov; };
__global__ void ke_gather_kernel(const EdgeRec* __restrict__ edges, const uint32_t* __restrict__ ids, uint32_t E,
                                 EdgeOut O) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    EdgeRec r = edges[ids ? ids[i] : i];
    O.n1[i] = r.n1; O.o1[i] = r.o1; O.n2[i] = r.n2; O.o2[i] = r.o2; O.ov[i] = r.ov;
}

__global__ void iota_kernel(uint32_t* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace mdbg
