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

// one thread per k-min-mer ordinal g (global (read, i) order)
__global__ void kb_window_kernel(MinArena A, const uint64_t* __restrict__ kmer_off, uint64_t K, uint32_t k,
                                 uint64_t seed, uint64_t fp_mask, uint64_t* __restrict__ fp,
                                 uint32_t* __restrict__ loc, uint8_t* __restrict__ rev,
                                 uint32_t* __restrict__ iota) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= K) return;
    uint64_t r = owner_read(kmer_off, A.R, g);
    uint64_t i = g - __ldg(kmer_off + r);
    uint64_t lo = __ldg(A.off + r) + i;
    const uint64_t* h = A.hash + lo;
    bool rv = window_reversed(h, k);
    uint64_t f = fp_init(seed, k);
    if (rv) for (uint32_t j = 0; j < k; j++) f = fp_mix(f, __ldg(h + k - 1 - j));
    else for (uint32_t j = 0; j < k; j++) f = fp_mix(f, __ldg(h + j));
    f &= fp_mask;
    if (f == KC_EMPTY) f = KC_EMPTY - 1;
    fp[g] = f;
    loc[g] = (uint32_t)lo;
    rev[g] = rv ? 1 : 0;
    iota[g] = (uint32_t)g;
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

// ---- K-C ---------------------------------------------------------------------------------------
// Open-address table of 64-bit fingerprints.  Four lanes cooperate on one key: they read one
// aligned 32-byte sector (4 slots) per probe, vote with ballot, and the lane holding the first
// empty slot claims it with atomicCAS.  first[slot] = smallest ordinal that carries the key.
__global__ void kc_insert_kernel(const uint64_t* __restrict__ fp, uint64_t K, uint64_t* keys,
                                 uint32_t* first, uint64_t cap_mask, uint32_t* __restrict__ slot_out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sub = lane & 3, gbase = lane & ~3u;
    uint64_t item = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 2;
    bool active = item < K;
    uint64_t key = active ? __ldg(fp + item) : 0;
    uint64_t base = ((key * 0x9e3779b97f4a7c15ULL) >> 17) & cap_mask & ~3ull;
    bool done = !active;
    uint64_t slot = 0;
    while (__any_sync(0xffffffffu, !done)) {
        uint64_t cur = done ? 0 : *reinterpret_cast<volatile uint64_t*>(keys + base + sub);
        uint32_t m_match = (__ballot_sync(0xffffffffu, !done && cur == key) >> gbase) & 0xFu;
        uint32_t m_empty = (__ballot_sync(0xffffffffu, !done && cur == KC_EMPTY) >> gbase) & 0xFu;
        uint32_t leader = m_empty ? (uint32_t)__ffs(m_empty) - 1 : 0;
        uint64_t old = 0;
        bool try_claim = !done && !m_match && m_empty;
        if (try_claim && sub == leader)
            old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + base + leader), KC_EMPTY, key);
        old = __shfl_sync(0xffffffffu, old, gbase + leader);
        if (!done) {
            if (m_match) { slot = base + (uint32_t)__ffs(m_match) - 1; done = true; }
            else if (m_empty) {
                if (old == KC_EMPTY || old == key) { slot = base + leader; done = true; }
                // else: somebody else took that slot for another key; look at the window again
            } else base = (base + 4) & cap_mask;
        }
    }
    if (active && sub == 0) {
        atomicMin(first + slot, (uint32_t)item);
        slot_out[item] = (uint32_t)slot;
    }
}

// exactness: every ordinal must carry the same TUPLE as the first ordinal of its slot
__global__ void kc_verify_kernel(MinArena A, uint64_t K, uint32_t k, const uint32_t* __restrict__ slot,
                                 const uint32_t* __restrict__ first, const uint32_t* __restrict__ loc,
                                 const uint8_t* __restrict__ rev, unsigned long long* collisions) {
    uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (g >= K) return;
    uint32_t f = __ldg(first + __ldg(slot + g));
    if (f == g) return;
    const uint64_t* a = A.hash + loc[g];
    const uint64_t* b = A.hash + loc[f];
    bool ra = rev[g], rb = rev[f];
    bool same = true;
    for (uint32_t j = 0; j < k && same; j++) {
        uint64_t x = ra ? __ldg(a + k - 1 - j) : __ldg(a + j);
        uint64_t y = rb ? __ldg(b + k - 1 - j) : __ldg(b + j);
        same = x == y;
    }
    if (!same) atomicAdd(collisions, 1ull);
}

// ---- K-D ---------------------------------------------------------------------------------------
__global__ void kd_heads_kernel(const uint32_t* __restrict__ sslot, uint64_t K, uint8_t* __restrict__ head) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= K) return;
    head[j] = (j == 0 || sslot[j] != sslot[j - 1]) ? 1 : 0;
}

// one thread per distinct tuple (segment of the slot-sorted ordinals, ascending ordinal inside)
__global__ void kd_segments_kernel(const uint32_t* __restrict__ seg_start, uint32_t D, uint64_t K,
                                   const uint32_t* __restrict__ sg, uint32_t minab,
                                   uint8_t* __restrict__ flag_first, uint8_t* __restrict__ flag_seq,
                                   uint8_t* __restrict__ solid, uint32_t* __restrict__ seg_first) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= D) return;
    uint32_t st = seg_start[s];
    uint32_t en = (s + 1 < D) ? seg_start[s + 1] : (uint32_t)K;
    uint32_t cnt = en - st;
    uint32_t fg = sg[st];
    seg_first[s] = fg;
    flag_first[fg] = 1;  // this sighting consumed a node index (main.rs:662)
    // sightings where previous_abundance == minabund-1 (u16 counter, wraps): main.rs:680,696
    for (uint32_t q = minab - 1; q < cnt; q += 65536u) flag_seq[sg[st + q]] = 1;
    uint32_t ab = cnt & 0xFFFFu;
    solid[s] = (minab == 1 || ab >= minab) ? 1 : 0;  // main.rs:922-929
}

__global__ void kd_node_keys_kernel(const uint32_t* __restrict__ solid_seg, uint32_t S,
                                    const uint32_t* __restrict__ seg_first,
                                    const uint32_t* __restrict__ first_rank, uint32_t* __restrict__ key) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S) return;
    key[n] = first_rank[seg_first[solid_seg[n]]];
}

struct NodeOut {
    uint32_t* index; uint16_t* abundance; uint32_t* seqlen; uint16_t* shift; uint64_t* tuple;
};

__global__ void kd_nodes_kernel(MinArena A, uint32_t S, uint32_t k, uint32_t minab, uint64_t K, uint32_t D,
                                const uint32_t* __restrict__ node_key, const uint32_t* __restrict__ node_seg,
                                const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ sg,
                                const uint32_t* __restrict__ loc, const uint8_t* __restrict__ rev,
                                uint32_t index_base, NodeOut O) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S) return;
    uint32_t s = node_seg[n];
    uint32_t st = seg_start[s];
    uint32_t en = (s + 1 < D) ? seg_start[s + 1] : (uint32_t)K;
    uint32_t cnt = en - st;
    uint32_t rep_rank = (minab - 1) + 65536u * ((cnt - minab) / 65536u);  // last overwrite, main.rs:680-684
    uint32_t g = sg[st + rep_rank];
    uint32_t lo = loc[g];
    bool rv = rev[g];
    const uint32_t* p = A.pos + lo;
    uint32_t a = p[1] - p[0], b = p[k - 1] - p[k - 2];
    O.index[n] = index_base + node_key[n];
    O.abundance[n] = (uint16_t)(cnt & 0xFFFFu);
    O.seqlen[n] = p[k - 1] + 1 - p[0] + 1;                    // read_offsets.2, main.rs:778
    O.shift[2 * n] = (uint16_t)(rv ? b : a);                  // lowprec_shift, main.rs:675
    O.shift[2 * n + 1] = (uint16_t)(rv ? a : b);
    const uint64_t* h = A.hash + lo;
    for (uint32_t j = 0; j < k; j++) O.tuple[(uint64_t)n * k + j] = rv ? h[k - 1 - j] : h[j];
}

struct SeqOut {
    uint32_t* index; uint64_t* read; uint64_t* start; uint64_t* end; uint8_t* reversed; uint64_t* shift;
};

__global__ void kd_seqlines_kernel(MinArena A, const uint64_t* __restrict__ kmer_off, uint32_t Q, uint32_t k,
                                   uint32_t l, const uint32_t* __restrict__ seq_g,
                                   const uint32_t* __restrict__ slot, const uint32_t* __restrict__ first,
                                   const uint32_t* __restrict__ first_rank, const uint32_t* __restrict__ loc,
                                   const uint8_t* __restrict__ rev, uint32_t index_base, uint64_t read_base,
                                   SeqOut O) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    uint32_t g = seq_g[q];
    uint32_t lo = loc[g];
    bool rv = rev[g];
    const uint32_t* p = A.pos + lo;
    uint64_t a = (uint64_t)p[1] - p[0], b = (uint64_t)p[k - 1] - p[k - 2];
    O.index[q] = index_base + first_rank[first[slot[g]]];
    O.read[q] = read_base + owner_read(kmer_off, A.R, g);
    O.start[q] = p[0];
    O.end[q] = (uint64_t)p[k - 1] + l;
    O.reversed[q] = rv ? 1 : 0;
    O.shift[2 * q] = rv ? b : a;
    O.shift[2 * q + 1] = rv ? a : b;
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
    ekey[e] = f;
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

// One thread per (n1, key) with key in [suffix_norm, prefix_norm] (main.rs:1050-1052).  WRITE=false
// counts kept edges / presimp removals, WRITE=true emits them at the scanned offsets.
template <bool WRITE>
__global__ void ke_join_kernel(NodeView N, const uint64_t* __restrict__ ekey, const uint8_t* __restrict__ erev,
                               const uint64_t* __restrict__ skey, const uint32_t* __restrict__ sval,
                               float presimp, uint32_t* __restrict__ cnt_edge, uint32_t* __restrict__ cnt_rem,
                               const uint32_t* __restrict__ off_edge, const uint32_t* __restrict__ off_rem,
                               EdgeRec* __restrict__ edges, uint64_t* __restrict__ removed) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t S = N.S, k = N.k, k1 = k - 1;
    if (q >= 2 * S) return;
    uint32_t n1 = q >> 1, which = q & 1;
    uint32_t qe = 2 * n1 + (which == 0 ? 1 : 0);  // which 0: suffix entry, 1: prefix entry
    uint64_t kf = ekey[qe];
    // bucket [lo, hi) of equal fingerprints in the sorted entry list
    uint32_t lo = 0, hi = 2 * S;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (skey[m] < kf) lo = m + 1; else hi = m; }
    uint32_t b0 = lo;
    hi = 2 * S;
    while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (skey[m] <= kf) lo = m + 1; else hi = m; }
    uint32_t b1 = lo;
    const uint64_t* t1 = N.tuple + (uint64_t)n1 * k;
    const uint64_t* qsub = t1 + (qe & 1);
    bool qrev = erev[qe];
    uint32_t ab1 = N.abundance[n1];
    // pass 0: number of potential edges and max abundance; pass 1: decide each edge
    uint32_t npot = 0, abmax = 0, ne = 0, nr = 0;
    uint32_t oe = WRITE ? off_edge[q] : 0, orr = WRITE ? off_rem[q] : 0;
    for (int pass = 0; pass < 2; pass++) {
        uint32_t abref = abmax < ab1 ? abmax : ab1;
        for (uint32_t e = b0; e < b1; e++) {
            uint32_t ent = sval[e];
            uint32_t n2 = ent >> 1;
            const uint64_t* t2 = N.tuple + (uint64_t)n2 * k;
            const uint64_t* esub = t2 + (ent & 1);
            bool er = erev[ent];
            // same (k-1)-mer key? (fingerprints only bucket; identity is on the tuples)
            bool same = eq_range(qrev ? qsub + k1 - 1 : qsub, qrev ? -1 : 1, er ? esub + k1 - 1 : esub, er ? -1 : 1, k1);
            if (!same) continue;
            // the four orientation identities of main.rs:1062-1075
            bool t[4];
            t[0] = eq_range(t1 + 1, 1, t2, 1, k1);              // n1.suffix == n2.prefix       (+,+)
            t[1] = eq_range(t1 + 1, 1, t2 + k - 1, -1, k1);     // n1.suffix == rev_n2.prefix   (+,-)
            t[2] = eq_range(t1 + k - 2, -1, t2, 1, k1);         // rev_n1.suffix == n2.prefix   (-,+)
            t[3] = eq_range(t1 + k - 2, -1, t2 + k - 1, -1, k1);  // rev_n1.suffix == rev_n2.prefix (-,-)
            uint32_t ab2 = N.abundance[n2];
            for (int o = 0; o < 4; o++) {
                if (!t[o]) continue;
                if (pass == 0) { npot++; abmax = ab2 > abmax ? ab2 : abmax; continue; }
                if (presimp > 0.0f && npot >= 2 && (float)ab2 < presimp * (float)abref) {  // main.rs:1086
                    if (WRITE) removed[orr + nr] = ((uint64_t)N.index[n1] << 32) | N.index[n2];
                    nr++;
                    continue;
                }
                if (WRITE) {
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
    if (!WRITE) { cnt_edge[q] = ne; cnt_rem[q] = nr; }
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

__global__ void ke_sortkeys_kernel(const EdgeRec* __restrict__ edges, const uint32_t* __restrict__ ids, uint32_t E,
                                   int major, uint64_t* __restrict__ key) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    EdgeRec r = edges[ids ? ids[i] : i];
    key[i] = major ? (((uint64_t)r.n1 << 32) | r.n2) : (((uint64_t)r.o1 << 33) | ((uint64_t)r.o2 << 32) | r.ov);
}

struct EdgeOut { uint32_t* n1; uint8_t* o1; uint32_t* n2; uint8_t* o2; uint32_t* ov; };
__global__ void ke_gather_kernel(const EdgeRec* __restrict__ edges, const uint32_t* __restrict__ ids, uint32_t E,
                                 EdgeOut O) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    EdgeRec r = edges[ids[i]];
    O.n1[i] = r.n1; O.o1[i] = r.o1; O.n2[i] = r.n2; O.o2[i] = r.o2; O.ov[i] = r.ov;
}

__global__ void iota_kernel(uint32_t* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace mdbg
