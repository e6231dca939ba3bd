// graph.cu -- mdbg_finish / mdbg_finish_device / mdbg_window: host orchestration of K-B .. K-E
// (kernels in graph_kernels.cuh; radix sort / scan / select are CUB device-wide primitives).
// Serial-order semantics: index = rank of a tuple's first sighting, abundance =
// sightings (u16), seqlen/shift/sequence from the minabund-th sighting.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "ctx.h"
#include "graph_kernels.cuh"

using namespace mdbg;

// Device-resident result of the last finish.
struct DeviceGraph {
    uint64_t n_kminmers = 0, n_distinct = 0, n_nodes = 0, n_edges = 0, presimp_removed = 0, n_seqlines = 0;
    uint32_t k = 0;
    Tmp<uint32_t> index; Tmp<uint16_t> abundance; Tmp<uint32_t> seqlen; Tmp<uint16_t> shift; Tmp<uint64_t> tuple;
    Tmp<uint32_t> e_n1, e_n2, e_ov; Tmp<uint8_t> e_o1, e_o2;
    Tmp<uint32_t> q_index; Tmp<uint64_t> q_read, q_start, q_end, q_shift; Tmp<uint8_t> q_rev;
};

extern "C" void mdbg_graph_device_free(mdbg_ctx* c) {
    if (c && c->dg) { delete c->dg; c->dg = nullptr; }
}

namespace {

inline unsigned nblk(uint64_t n, unsigned bs = 256) { return (unsigned)std::max<uint64_t>(1, (n + bs - 1) / bs); }

struct Runner {  // CUB call helper: size query, pooled temp storage, launch accounting
    mdbg_ctx* c;
    Tmp<uint8_t> temp;
    int err = 0;
    template <class F>
    int cub(F&& f) {
        size_t bytes = 0;
        cudaError_t e = f((void*)nullptr, bytes);
        if (e == cudaSuccess) {
            if (temp.cap < bytes || !temp.p) e = temp.get(c->pool, bytes + 256);
            if (e == cudaSuccess) e = f((void*)temp.p, bytes);
        }
        c->tm.launches_finish += 2;  // CUB device primitives launch >= 2 kernels each
        if (e != cudaSuccess) {
            c->err = std::string("CUB: ") + cudaGetErrorString(e);
            return MDBG_ERR_CUDA;
        }
        return MDBG_OK;
    }
};

#define RC(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
#define LAUNCHED(c) do { (c)->tm.launches_finish++; MDBG_CK(c, cudaGetLastError()); } while (0)

int read_scalars(mdbg_ctx* c) {
    MDBG_CK(c, cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}

int log2_ceil(uint64_t x) { int b = 0; while ((1ull << b) < x) b++; return b; }

// Everything from the resident minimizers to the device graph.
int build_device_graph(mdbg_ctx* c, bool want_seqlines) {
    mdbg_graph_device_free(c);
    DeviceGraph* G = new DeviceGraph();
    c->dg = G;
    const uint32_t k = c->p.k, l = c->p.l, minab = c->p.min_abundance;
    const float presimp = c->p.presimp;
    G->k = k;
    c->tm.launches_finish = 0;
    c->tm.table_attempts = 0;
    cudaStream_t st = c->st;
    Runner R{c};
    if (c->M >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers on one GPU"; return MDBG_ERR_RANGE; }
    MDBG_CK(c, cudaEventRecord(c->ev[5], st));
    MinArena A{c->m_hash, c->m_pos, c->m_off, c->R};
    const uint64_t nR = c->R;

    // ---- K-B: k-min-mer offsets per read ------------------------------------------------------
    Tmp<uint64_t> cnt, kmer_off;
    MDBG_CK(c, cnt.get(c->pool, nR + 1));
    MDBG_CK(c, kmer_off.get(c->pool, nR + 1));
    uint64_t K = 0;
    if (nR > 0 && c->M > 0) {
        kb_count_kernel<<<nblk(nR + 1), 256, 0, st>>>(c->m_off, nR, k, cnt);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, kmer_off.p, nR + 1, st); }));
        MDBG_CK(c, cudaMemcpyAsync(&c->h_sc->v[0], kmer_off.p + nR, 8, cudaMemcpyDeviceToHost, st));
        MDBG_CK(c, cudaStreamSynchronize(st));
        K = c->h_sc->v[0];
    }
    G->n_kminmers = K;
    if (K >= 0x7FFFFFF0ull) { c->err = "more than 2^31 k-min-mers on one GPU"; return MDBG_ERR_RANGE; }
    MDBG_CK(c, cudaEventRecord(c->ev[6], st));
    if (K == 0) {
        MDBG_CK(c, cudaEventRecord(c->ev[7], st));
        MDBG_CK(c, cudaEventRecord(c->ev[8], st));
        MDBG_CK(c, cudaEventRecord(c->ev[9], st));
        return MDBG_OK;
    }

    // ---- K-B window + K-C table (retry with a new seed on a fingerprint collision) --------------
    Tmp<uint64_t> fp; Tmp<uint32_t> loc, iota, slot, first; Tmp<uint8_t> rev; Tmp<uint64_t> keys;
    MDBG_CK(c, fp.get(c->pool, K));
    MDBG_CK(c, loc.get(c->pool, K));
    MDBG_CK(c, iota.get(c->pool, K));
    MDBG_CK(c, slot.get(c->pool, K));
    MDBG_CK(c, rev.get(c->pool, K));
    const int cap_bits = std::max(12, log2_ceil(2 * K));
    const uint64_t cap = 1ull << cap_bits;
    MDBG_CK(c, keys.get(c->pool, cap));
    MDBG_CK(c, first.get(c->pool, cap));
    float ms_kb = 0, ms_kc = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt >= 8) { c->err = "fingerprint collisions persisted over 8 seeds"; return MDBG_ERR_RANGE; }
        c->tm.table_attempts = attempt + 1;
        uint64_t seed = 0x6d64626700000000ull + 0x9e3779b97f4a7c15ull * (uint64_t)attempt;
        uint64_t mask = ~0ull;
        if (attempt == 0 && c->p.debug_fp_bits > 0 && c->p.debug_fp_bits < 64) mask = (1ull << c->p.debug_fp_bits) - 1;
        MDBG_CK(c, cudaEventRecord(c->ev[10], st));
        kb_window_kernel<<<nblk(K), 256, 0, st>>>(A, kmer_off, K, k, seed, mask, fp, loc, rev, iota);
        LAUNCHED(c);
        MDBG_CK(c, cudaEventRecord(c->ev[11], st));
        MDBG_CK(c, cudaMemsetAsync(keys, 0xFF, cap * 8, st));
        MDBG_CK(c, cudaMemsetAsync(first, 0xFF, cap * 4, st));
        MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[1], 0, 8, st));
        kc_insert_kernel<<<nblk(K * 4), 256, 0, st>>>(fp, K, keys, first, cap - 1, slot);
        LAUNCHED(c);
        kc_verify_kernel<<<nblk(K), 256, 0, st>>>(A, K, k, slot, first, loc, rev, &c->d_sc->v[1]);
        LAUNCHED(c);
        MDBG_CK(c, cudaEventRecord(c->ev[12], st));
        RC(read_scalars(c));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev[10], c->ev[11]);
        cudaEventElapsedTime(&b, c->ev[11], c->ev[12]);
        ms_kb += a; ms_kc += b;
        if (c->h_sc->v[1] == 0) break;  // every slot holds exactly one tuple
    }
    keys.reset();
    fp.reset();
    MDBG_CK(c, cudaEventRecord(c->ev[7], st));

    // ---- K-D: sort ordinals by slot (stable => ascending ordinal inside a slot), segment -------
    Tmp<uint32_t> sslot, sg;
    MDBG_CK(c, sslot.get(c->pool, K));
    MDBG_CK(c, sg.get(c->pool, K));
    RC(R.cub([&](void* t, size_t& b) {
        return cub::DeviceRadixSort::SortPairs(t, b, slot.p, sslot.p, iota.p, sg.p, (uint32_t)K, 0, cap_bits, st);
    }));
    iota.reset();
    Tmp<uint8_t> head;
    Tmp<uint32_t> seg_start;
    MDBG_CK(c, head.get(c->pool, K));
    MDBG_CK(c, seg_start.get(c->pool, K + 1));
    kd_heads_kernel<<<nblk(K), 256, 0, st>>>(sslot, K, head);
    LAUNCHED(c);
    RC(R.cub([&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<uint32_t>(0), head.p, seg_start.p,
                                          (uint32_t*)&c->d_sc->v[2], (uint32_t)K, st);
    }));
    RC(read_scalars(c));
    const uint32_t D = (uint32_t)(c->h_sc->v[2] & 0xFFFFFFFFu);
    G->n_distinct = D;
    sslot.reset();
    head.reset();
    Tmp<uint8_t> flag_first, flag_seq, solid;
    Tmp<uint32_t> seg_first, first_rank, solid_seg;
    MDBG_CK(c, flag_first.get(c->pool, K));
    MDBG_CK(c, flag_seq.get(c->pool, K));
    MDBG_CK(c, solid.get(c->pool, D));
    MDBG_CK(c, seg_first.get(c->pool, D));
    MDBG_CK(c, first_rank.get(c->pool, K));
    MDBG_CK(c, solid_seg.get(c->pool, D));
    MDBG_CK(c, cudaMemsetAsync(flag_first, 0, K, st));
    MDBG_CK(c, cudaMemsetAsync(flag_seq, 0, K, st));
    kd_segments_kernel<<<nblk(D), 256, 0, st>>>(seg_start, D, K, sg, minab, flag_first, flag_seq, solid, seg_first);
    LAUNCHED(c);
    // node index = number of earlier first sightings (NODE_INDEX order, main.rs:662)
    RC(R.cub([&](void* t, size_t& b) {
        return cub::DeviceScan::ExclusiveSum(t, b, flag_first.p, first_rank.p, (uint32_t)K, st);
    }));
    RC(R.cub([&](void* t, size_t& b) {
        return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<uint32_t>(0), solid.p, solid_seg.p,
                                          (uint32_t*)&c->d_sc->v[3], (uint32_t)D, st);
    }));
    Tmp<uint32_t> seq_g;
    if (want_seqlines) {
        MDBG_CK(c, seq_g.get(c->pool, K));
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<uint32_t>(0), flag_seq.p, seq_g.p,
                                              (uint32_t*)&c->d_sc->v[4], (uint32_t)K, st);
        }));
    }
    RC(read_scalars(c));
    const uint32_t S = (uint32_t)(c->h_sc->v[3] & 0xFFFFFFFFu);
    const uint32_t Q = want_seqlines ? (uint32_t)(c->h_sc->v[4] & 0xFFFFFFFFu) : 0;
    G->n_nodes = S;
    G->n_seqlines = Q;
    flag_first.reset();
    flag_seq.reset();
    solid.reset();

    // nodes in ascending index order
    MDBG_CK(c, G->index.get(c->pool, S));
    MDBG_CK(c, G->abundance.get(c->pool, S));
    MDBG_CK(c, G->seqlen.get(c->pool, S));
    MDBG_CK(c, G->shift.get(c->pool, 2 * (uint64_t)S));
    MDBG_CK(c, G->tuple.get(c->pool, (uint64_t)S * k));
    if (S > 0) {
        Tmp<uint32_t> nkey, nkey_s, nseg_s;
        MDBG_CK(c, nkey.get(c->pool, S));
        MDBG_CK(c, nkey_s.get(c->pool, S));
        MDBG_CK(c, nseg_s.get(c->pool, S));
        kd_node_keys_kernel<<<nblk(S), 256, 0, st>>>(solid_seg, S, seg_first, first_rank, nkey);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, nkey.p, nkey_s.p, solid_seg.p, nseg_s.p, S, 0, 32, st);
        }));
        NodeOut NO{G->index, G->abundance, G->seqlen, G->shift, G->tuple};
        kd_nodes_kernel<<<nblk(S), 256, 0, st>>>(A, S, k, minab, K, D, nkey_s, nseg_s, seg_start, sg, loc, rev, 0, NO);
        LAUNCHED(c);
    }
    if (want_seqlines) {
        MDBG_CK(c, G->q_index.get(c->pool, Q));
        MDBG_CK(c, G->q_read.get(c->pool, Q));
        MDBG_CK(c, G->q_start.get(c->pool, Q));
        MDBG_CK(c, G->q_end.get(c->pool, Q));
        MDBG_CK(c, G->q_rev.get(c->pool, Q));
        MDBG_CK(c, G->q_shift.get(c->pool, 2 * (uint64_t)Q));
        if (Q > 0) {
            SeqOut SO{G->q_index, G->q_read, G->q_start, G->q_end, G->q_rev, G->q_shift};
            kd_seqlines_kernel<<<nblk(Q), 256, 0, st>>>(A, kmer_off, Q, k, l, seq_g, slot, first, first_rank, loc, rev,
                                                        0, 0, SO);
            LAUNCHED(c);
        }
    }
    MDBG_CK(c, cudaEventRecord(c->ev[8], st));
    cudaEventElapsedTime(&c->tm.ms_kb, c->ev[5], c->ev[6]);
    c->tm.ms_kb += ms_kb;
    c->tm.ms_kc = ms_kc;
    // free table-stage scratch before the edge stage
    seq_g.reset(); slot.reset(); first.reset(); first_rank.reset(); solid_seg.reset(); seg_first.reset();
    seg_start.reset(); sg.reset(); loc.reset(); rev.reset(); cnt.reset(); kmer_off.reset();

    // ---- K-E: edges -----------------------------------------------------------------------------
    if (S > 0) {
        NodeView NV{G->index, G->abundance, G->seqlen, G->shift, G->tuple, S, k};
        const uint32_t E2 = 2 * S;
        Tmp<uint64_t> ekey, skey; Tmp<uint32_t> eval, sval; Tmp<uint8_t> erev;
        MDBG_CK(c, ekey.get(c->pool, E2));
        MDBG_CK(c, skey.get(c->pool, E2));
        MDBG_CK(c, eval.get(c->pool, E2));
        MDBG_CK(c, sval.get(c->pool, E2));
        MDBG_CK(c, erev.get(c->pool, E2));
        ke_entries_kernel<<<nblk(E2), 256, 0, st>>>(NV, 0x656467657300ull, ekey, eval, erev);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, ekey.p, skey.p, eval.p, sval.p, E2, 0, 64, st);
        }));
        Tmp<uint32_t> cnt_e, cnt_r, off_e, off_r;
        MDBG_CK(c, cnt_e.get(c->pool, E2 + 1));
        MDBG_CK(c, cnt_r.get(c->pool, E2 + 1));
        MDBG_CK(c, off_e.get(c->pool, E2 + 1));
        MDBG_CK(c, off_r.get(c->pool, E2 + 1));
        MDBG_CK(c, cudaMemsetAsync(cnt_e.p + E2, 0, 4, st));
        MDBG_CK(c, cudaMemsetAsync(cnt_r.p + E2, 0, 4, st));
        ke_join_kernel<false><<<nblk(E2, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, presimp, cnt_e, cnt_r, nullptr,
                                                              nullptr, nullptr, nullptr);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt_e.p, off_e.p, E2 + 1, st); }));
        RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt_r.p, off_r.p, E2 + 1, st); }));
        MDBG_CK(c, cudaMemcpyAsync(&c->h_sc->v[5], off_e.p + E2, 4, cudaMemcpyDeviceToHost, st));
        MDBG_CK(c, cudaMemcpyAsync(&c->h_sc->v[6], off_r.p + E2, 4, cudaMemcpyDeviceToHost, st));
        MDBG_CK(c, cudaStreamSynchronize(st));
        uint32_t EP = (uint32_t)(c->h_sc->v[5] & 0xFFFFFFFFu), NR = (uint32_t)(c->h_sc->v[6] & 0xFFFFFFFFu);
        G->presimp_removed = presimp > 0.0f ? NR : 0;
        Tmp<EdgeRec> pend, kept;
        Tmp<uint64_t> removed;
        MDBG_CK(c, pend.get(c->pool, EP));
        MDBG_CK(c, removed.get(c->pool, NR));
        if (EP > 0 || NR > 0) {
            ke_join_kernel<true><<<nblk(E2, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, presimp, nullptr, nullptr, off_e,
                                                                 off_r, pend, removed);
            LAUNCHED(c);
        }
        EdgeRec* edges = pend;
        uint32_t E = EP;
        if (NR > 0 && EP > 0) {  // drop (n1,n2) if it or its reverse was presimp-removed (main.rs:1107-1116)
            Tmp<uint64_t> rem_s;
            Tmp<uint8_t> keep;
            MDBG_CK(c, rem_s.get(c->pool, NR));
            MDBG_CK(c, keep.get(c->pool, EP));
            MDBG_CK(c, kept.get(c->pool, EP));
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortKeys(t, b, removed.p, rem_s.p, NR, 0, 64, st); }));
            ke_filter_kernel<<<nblk(EP), 256, 0, st>>>(pend, EP, rem_s, NR, keep);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) {
                return cub::DeviceSelect::Flagged(t, b, pend.p, keep.p, kept.p, (uint32_t*)&c->d_sc->v[7], EP, st);
            }));
            RC(read_scalars(c));
            E = (uint32_t)(c->h_sc->v[7] & 0xFFFFFFFFu);
            edges = kept;
        }
        G->n_edges = E;
        MDBG_CK(c, G->e_n1.get(c->pool, E));
        MDBG_CK(c, G->e_o1.get(c->pool, E));
        MDBG_CK(c, G->e_n2.get(c->pool, E));
        MDBG_CK(c, G->e_o2.get(c->pool, E));
        MDBG_CK(c, G->e_ov.get(c->pool, E));
        if (E > 0) {  // canonical order (n1, n2, o1, o2, overlap): two stable radix passes
            Tmp<uint64_t> k_a, k_b; Tmp<uint32_t> id_a, id_b, id_c;
            MDBG_CK(c, k_a.get(c->pool, E));
            MDBG_CK(c, k_b.get(c->pool, E));
            MDBG_CK(c, id_a.get(c->pool, E));
            MDBG_CK(c, id_b.get(c->pool, E));
            MDBG_CK(c, id_c.get(c->pool, E));
            iota_kernel<<<nblk(E), 256, 0, st>>>(id_a, E);
            LAUNCHED(c);
            ke_sortkeys_kernel<<<nblk(E), 256, 0, st>>>(edges, nullptr, E, 0, k_a);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, k_a.p, k_b.p, id_a.p, id_b.p, E, 0, 34, st); }));
            ke_sortkeys_kernel<<<nblk(E), 256, 0, st>>>(edges, id_b, E, 1, k_a);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, k_a.p, k_b.p, id_b.p, id_c.p, E, 0, 64, st); }));
            EdgeOut EO{G->e_n1, G->e_o1, G->e_n2, G->e_o2, G->e_ov};
            ke_gather_kernel<<<nblk(E), 256, 0, st>>>(edges, id_c, E, EO);
            LAUNCHED(c);
        }
    }
    MDBG_CK(c, cudaEventRecord(c->ev[9], st));
    return MDBG_OK;
}

void fill_counters(mdbg_ctx* c, mdbg_graph* out) {
    memset(out, 0, sizeof(*out));
    DeviceGraph* G = c->dg;
    out->n_reads = c->R; out->n_bases = c->n_bases; out->n_minimizers = c->M;
    out->k = c->p.k; out->l = c->p.l;
    if (!G) return;
    out->n_kminmers = G->n_kminmers; out->n_distinct = G->n_distinct; out->n_nodes = G->n_nodes;
    out->n_edges = G->n_edges; out->presimp_removed = G->presimp_removed; out->n_seqlines = G->n_seqlines;
}

int finish_timings(mdbg_ctx* c) {
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    cudaEventElapsedTime(&c->tm.ms_kd, c->ev[7], c->ev[8]);
    cudaEventElapsedTime(&c->tm.ms_ke, c->ev[8], c->ev[9]);
    cudaEventElapsedTime(&c->tm.ms_total_finish, c->ev[5], c->ev[9]);
    return MDBG_OK;
}

struct HostGraph {  // owner of the host arrays handed out through mdbg_graph
    std::vector<uint32_t> index, seqlen, e_n1, e_n2, e_ov, q_index;
    std::vector<uint16_t> abundance, shift;
    std::vector<uint64_t> tuple, q_read, q_start, q_end, q_shift;
    std::vector<uint8_t> e_o1, e_o2, q_rev;
};

template <class T>
int d2h(mdbg_ctx* c, std::vector<T>& v, const T* d, uint64_t n) {
    v.resize(n);
    if (n) MDBG_CK(c, cudaMemcpyAsync(v.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost, c->st));
    return MDBG_OK;
}

}  // namespace

extern "C" {

int mdbg_finish_device(mdbg_ctx* c, mdbg_graph* out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    int rc = build_device_graph(c, false);
    if (rc) return rc;
    RC(finish_timings(c));
    c->tm.ms_d2h = 0;
    fill_counters(c, out);
    return MDBG_OK;
}

int mdbg_finish(mdbg_ctx* c, int want_seqlines, mdbg_graph* out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    int rc = build_device_graph(c, want_seqlines != 0);
    if (rc) return rc;
    RC(finish_timings(c));
    fill_counters(c, out);
    DeviceGraph* G = c->dg;
    HostGraph* H = new HostGraph();
    out->_owner = H;
    const uint64_t S = G->n_nodes, E = G->n_edges, Q = G->n_seqlines, k = G->k;
    MDBG_CK(c, cudaEventRecord(c->ev[13], c->st));
    RC(d2h(c, H->index, G->index.p, S));
    RC(d2h(c, H->abundance, G->abundance.p, S));
    RC(d2h(c, H->seqlen, G->seqlen.p, S));
    RC(d2h(c, H->shift, G->shift.p, 2 * S));
    RC(d2h(c, H->tuple, G->tuple.p, S * k));
    RC(d2h(c, H->e_n1, G->e_n1.p, E));
    RC(d2h(c, H->e_o1, G->e_o1.p, E));
    RC(d2h(c, H->e_n2, G->e_n2.p, E));
    RC(d2h(c, H->e_o2, G->e_o2.p, E));
    RC(d2h(c, H->e_ov, G->e_ov.p, E));
    if (want_seqlines) {
        RC(d2h(c, H->q_index, G->q_index.p, Q));
        RC(d2h(c, H->q_read, G->q_read.p, Q));
        RC(d2h(c, H->q_start, G->q_start.p, Q));
        RC(d2h(c, H->q_end, G->q_end.p, Q));
        RC(d2h(c, H->q_rev, G->q_rev.p, Q));
        RC(d2h(c, H->q_shift, G->q_shift.p, 2 * Q));
    }
    MDBG_CK(c, cudaEventRecord(c->ev[14], c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    cudaEventElapsedTime(&c->tm.ms_d2h, c->ev[13], c->ev[14]);
    out->node_index = H->index.data(); out->abundance = H->abundance.data(); out->seqlen = H->seqlen.data();
    out->shift = H->shift.data(); out->tuple = H->tuple.data();
    out->e_n1 = H->e_n1.data(); out->e_o1 = H->e_o1.data(); out->e_n2 = H->e_n2.data();
    out->e_o2 = H->e_o2.data(); out->e_overlap = H->e_ov.data();
    if (want_seqlines) {
        out->q_index = H->q_index.data(); out->q_read = H->q_read.data(); out->q_start = H->q_start.data();
        out->q_end = H->q_end.data(); out->q_reversed = H->q_rev.data(); out->q_shift = H->q_shift.data();
    }
    return MDBG_OK;
}

void mdbg_graph_free(mdbg_graph* g) {
    if (!g) return;
    delete (HostGraph*)g->_owner;
    memset(g, 0, sizeof(*g));
}

// Entry 2 batch form: windows of caller-provided minimizers (main.rs:756-781).
int mdbg_window(mdbg_ctx* c, const uint64_t* hash, const uint64_t* pos, const uint64_t* min_read_off,
                uint64_t n_reads, uint64_t* out_tuple, uint8_t* out_reversed, uint64_t* out_shift,
                uint64_t* out_offsets, uint64_t* out_kmer_read_off, uint64_t cap, uint64_t* n_out) {
    if (!c || !min_read_off) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    cudaStream_t st = c->st;
    const uint32_t k = c->p.k, l = c->p.l;
    const uint64_t M = min_read_off[n_reads];
    Runner R{c};
    Tmp<uint64_t> d_hash, d_off, cnt, kmer_off;
    Tmp<uint32_t> d_pos;
    MDBG_CK(c, d_hash.get(c->pool, M));
    MDBG_CK(c, d_pos.get(c->pool, M));
    MDBG_CK(c, d_off.get(c->pool, n_reads + 1));
    MDBG_CK(c, cnt.get(c->pool, n_reads + 1));
    MDBG_CK(c, kmer_off.get(c->pool, n_reads + 1));
    std::vector<uint32_t> p32(M);
    for (uint64_t i = 0; i < M; i++) p32[i] = (uint32_t)pos[i];
    if (M) {
        MDBG_CK(c, cudaMemcpyAsync(d_hash, hash, M * 8, cudaMemcpyHostToDevice, st));
        MDBG_CK(c, cudaMemcpyAsync(d_pos, p32.data(), M * 4, cudaMemcpyHostToDevice, st));
    }
    MDBG_CK(c, cudaMemcpyAsync(d_off, min_read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    kb_count_kernel<<<nblk(n_reads + 1), 256, 0, st>>>(d_off, n_reads, k, cnt);
    MDBG_CK(c, cudaGetLastError());
    RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, kmer_off.p, n_reads + 1, st); }));
    std::vector<uint64_t> ko(n_reads + 1);
    MDBG_CK(c, cudaMemcpyAsync(ko.data(), kmer_off, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    const uint64_t K = ko[n_reads];
    if (n_out) *n_out = K;
    if (out_kmer_read_off) memcpy(out_kmer_read_off, ko.data(), (n_reads + 1) * 8);
    if (K > cap) { c->err = "output capacity too small"; return MDBG_ERR_CAPACITY; }
    if (K == 0) return MDBG_OK;
    Tmp<uint64_t> t_tuple, t_shift, t_offs; Tmp<uint8_t> t_rev;
    MDBG_CK(c, t_tuple.get(c->pool, K * k));
    MDBG_CK(c, t_shift.get(c->pool, K * 2));
    MDBG_CK(c, t_offs.get(c->pool, K * 3));
    MDBG_CK(c, t_rev.get(c->pool, K));
    MinArena A{d_hash, d_pos, d_off, n_reads};
    kb_export_kernel<<<nblk(K), 256, 0, st>>>(A, kmer_off, K, k, l, t_tuple, t_rev, t_shift, t_offs);
    MDBG_CK(c, cudaGetLastError());
    if (out_tuple) MDBG_CK(c, cudaMemcpyAsync(out_tuple, t_tuple, K * k * 8, cudaMemcpyDeviceToHost, st));
    if (out_reversed) MDBG_CK(c, cudaMemcpyAsync(out_reversed, t_rev, K, cudaMemcpyDeviceToHost, st));
    if (out_shift) MDBG_CK(c, cudaMemcpyAsync(out_shift, t_shift, K * 16, cudaMemcpyDeviceToHost, st));
    if (out_offsets) MDBG_CK(c, cudaMemcpyAsync(out_offsets, t_offs, K * 24, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    return MDBG_OK;
}

}  // extern "C"
