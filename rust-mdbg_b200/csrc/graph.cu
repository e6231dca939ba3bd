// graph.cu -- mdbg_finish / mdbg_finish_device / mdbg_window: host orchestration of K-B .. K-E
// (kernels in graph_kernels.cuh; radix sort / scan / select are CUB device-wide primitives).
//
// Serial-order semantics: index = rank of a tuple's first sighting, abundance =
// sightings (u16), seqlen/shift/sequence from the minabund-th sighting.
//
// One code path for 1 and N GPUs (one process per GPU, NCCL over NVLink):
//   1. N > 1: the minimizer arenas (~12 B x 2d per base) are all-gathered, so every GPU sees every
//      window of the job in serial order; nothing larger is exchanged
//   2. every k-min-mer sighting this GPU owns (N > 1: tuple fingerprint in this rank's range, so
//      every copy of a tuple meets on one owner) becomes a RECORD {window location, ordinal,
//      RecInfo}; the canonical tuple is read from the arena, never materialised
//   3. owner: open-address table (fingerprint placed, tuple verified), stable radix sort by slot,
//      segmented reduce -> abundance / first sighting / representative sighting
//   4. node index and node placement = ONE exclusive scan over ordinal space of the
//      first-sighting flags (summed over the GPUs with one all-reduce)
//   5. every GPU writes its nodes at their final places, the node arrays are all-reduced; every
//      GPU builds the (k-1)-mer entry index and emits the edges of its slice of the nodes; presimp
//      removals are all-gathered before the final filter
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "ctx.h"
#include "graph_kernels.cuh"
#include "nccl_dl.h"

using namespace mdbg;

// Device-resident result of the last finish.
struct DeviceGraph {
    uint64_t n_kminmers = 0, n_distinct = 0, n_nodes = 0, n_edges = 0, presimp_removed = 0, n_seqlines = 0;
    uint64_t n_edges_local = 0, n_seq_local = 0;
    uint32_t k = 0;
    Tmp<uint32_t> index; Tmp<uint16_t> abundance; Tmp<uint32_t> seqlen; Tmp<uint16_t> shift; Tmp<uint64_t> tuple;
    Tmp<uint32_t> e_n1, e_n2, e_ov; Tmp<uint8_t> e_o1, e_o2;   // this GPU's slice (sorted)
    Tmp<SeqRec> seq;                                           // this GPU's lines (ordinal order)
};

extern "C" void mdbg_graph_device_free(mdbg_ctx* c) {
    if (c && c->dg) { delete c->dg; c->dg = nullptr; }
}

namespace {

inline unsigned nblk(uint64_t n, unsigned bs = 256) { return (unsigned)std::max<uint64_t>(1, (n + bs - 1) / bs); }

struct Runner {  // CUB call helper: size query, pooled temp storage, launch accounting
    mdbg_ctx* c;
    Tmp<uint8_t> temp;
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
#define NCK(c, call)                                                                        \
    do {                                                                                    \
        ncclResult_t _r = (call);                                                           \
        if (_r != ncclSuccess) {                                                            \
            (c)->err = std::string(#call) + ": " + nccl().GetErrorString(_r);               \
            return MDBG_ERR_NCCL;                                                           \
        }                                                                                   \
    } while (0)

int read_scalars(mdbg_ctx* c) {
    MDBG_CK(c, cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}

int log2_ceil(uint64_t x) { int b = 0; while ((1ull << b) < x) b++; return b; }

// ---- NCCL helpers (world > 1) -----------------------------------------------------------------
// all[s * n + i] = value i of rank s
int allgather_u64(mdbg_ctx* c, const uint64_t* mine, int n, std::vector<uint64_t>& all) {
    const int W = c->world;
    all.assign((size_t)W * n, 0);
    if (W == 1) { for (int i = 0; i < n; i++) all[i] = mine[i]; return MDBG_OK; }
    Tmp<uint64_t> d_in, d_out;
    MDBG_CK(c, d_in.get(c->pool, n));
    MDBG_CK(c, d_out.get(c->pool, (size_t)W * n));
    MDBG_CK(c, cudaMemcpyAsync(d_in, mine, n * 8, cudaMemcpyHostToDevice, c->st));
    NCK(c, nccl().AllGather(d_in, d_out, n, ncclUint64, (ncclComm_t)c->comm, c->st));
    MDBG_CK(c, cudaMemcpyAsync(all.data(), d_out, (size_t)W * n * 8, cudaMemcpyDeviceToHost, c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}

// all-gather of variable-length arrays: every rank ends with the concatenation in rank order
int allgatherv(mdbg_ctx* c, const void* mine, uint64_t my_cnt, const std::vector<uint64_t>& cnt, void* out,
               size_t elem) {
    const int W = c->world;
    if (W == 1) {
        if (my_cnt && out != mine) MDBG_CK(c, cudaMemcpyAsync(out, mine, my_cnt * elem, cudaMemcpyDeviceToDevice, c->st));
        return MDBG_OK;
    }
    NcclApi& N = nccl();
    size_t ro = 0;
    NCK(c, N.GroupStart());
    for (int p = 0; p < W; p++) {
        if (my_cnt) NCK(c, N.Send(mine, my_cnt * elem, ncclChar, p, (ncclComm_t)c->comm, c->st));
        if (cnt[p]) NCK(c, N.Recv((char*)out + ro, cnt[p] * elem, ncclChar, p, (ncclComm_t)c->comm, c->st));
        ro += cnt[p] * elem;
    }
    NCK(c, N.GroupEnd());
    return MDBG_OK;
}

// gather variable-length arrays on rank 0 (concatenated in rank order)
int gatherv_root(mdbg_ctx* c, const void* mine, uint64_t my_cnt, const std::vector<uint64_t>& cnt, void* out,
                 size_t elem) {
    const int W = c->world;
    if (W == 1) {
        if (my_cnt && out != mine) MDBG_CK(c, cudaMemcpyAsync(out, mine, my_cnt * elem, cudaMemcpyDeviceToDevice, c->st));
        return MDBG_OK;
    }
    NcclApi& N = nccl();
    NCK(c, N.GroupStart());
    if (my_cnt) NCK(c, N.Send(mine, my_cnt * elem, ncclChar, 0, (ncclComm_t)c->comm, c->st));
    if (c->rank == 0) {
        size_t ro = 0;
        for (int p = 0; p < W; p++) {
            if (cnt[p]) NCK(c, N.Recv((char*)out + ro, cnt[p] * elem, ncclChar, p, (ncclComm_t)c->comm, c->st));
            ro += cnt[p] * elem;
        }
    }
    NCK(c, N.GroupEnd());
    return MDBG_OK;
}

// Everything from the resident minimizers to the device graph.
int build_device_graph(mdbg_ctx* c, bool want_seqlines) {
    mdbg_graph_device_free(c);
    DeviceGraph* G = new DeviceGraph();
    c->dg = G;
    const uint32_t k = c->p.k, l = c->p.l, minab = c->p.min_abundance;
    const float presimp = c->p.presimp;
    const uint32_t bf = (c->p.bf && minab > 1) ? 1 : 0;   // main.rs:639
    const int W = c->world, rank = c->rank;
    G->k = k;
    c->tm.launches_finish = 0;
    c->tm.table_attempts = 0;
    c->tm.ms_kb = c->tm.ms_kc = c->tm.ms_kd = c->tm.ms_ke = 0;
    cudaStream_t st = c->st;
    Runner R{c};
    if (W > 1 && !c->comm) { c->err = "world > 1 but mdbg_comm_init was not called"; return MDBG_ERR_BAD_ARG; }
    if (c->M >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers on one GPU"; return MDBG_ERR_RANGE; }
    MDBG_CK(c, cudaEventRecord(c->ev[5], st));
    // ---- the minimizer arena of the whole job ---------------------------------------------------
    // N > 1: the per-GPU arenas (~2d x 12 bytes per base) are all-gathered, so every GPU sees every
    // window of the job in serial order; nothing larger is ever exchanged.
    MinArena A{c->m_hash, c->m_pos, c->m_off, c->R};
    Tmp<uint64_t> g_hash, g_off, d_rmap;
    Tmp<uint32_t> g_pos;
    const uint64_t* d_rpre = nullptr; const uint64_t* d_rbase = nullptr;
    if (W > 1) {
        std::vector<uint64_t> all;
        uint64_t mine[3] = {c->M, c->R, c->read_base_set ? c->read_base : ~0ull};
        RC(allgather_u64(c, mine, 3, all));
        std::vector<uint64_t> mcnt(W), rcnt(W), rmap(3 * (size_t)W + 2, 0);   // rpre[W+1] | mpre[W+1] | rbase[W]
        uint64_t* rpre = rmap.data(); uint64_t* mpre = rpre + W + 1; uint64_t* rbase = mpre + W + 1;
        for (int r = 0; r < W; r++) {
            mcnt[r] = all[3 * r]; rcnt[r] = all[3 * r + 1];
            mpre[r + 1] = mpre[r] + mcnt[r]; rpre[r + 1] = rpre[r] + rcnt[r];
            rbase[r] = all[3 * r + 2] != ~0ull ? all[3 * r + 2] : rpre[r];
        }
        const uint64_t Mtot = mpre[W], Rtot = rpre[W];
        if (Mtot >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers in the job"; return MDBG_ERR_RANGE; }
        MDBG_CK(c, g_hash.get(c->pool, Mtot)); MDBG_CK(c, g_pos.get(c->pool, Mtot)); MDBG_CK(c, g_off.get(c->pool, Rtot + 1));
        MDBG_CK(c, d_rmap.get(c->pool, rmap.size()));
        MDBG_CK(c, cudaMemcpyAsync(d_rmap, rmap.data(), rmap.size() * 8, cudaMemcpyHostToDevice, st));
        NCK(c, nccl().GroupStart());
        RC(allgatherv(c, c->m_hash, c->M, mcnt, g_hash, 8));
        RC(allgatherv(c, c->m_pos, c->M, mcnt, g_pos, 4));
        RC(allgatherv(c, c->m_off, c->R, rcnt, g_off, 8));
        NCK(c, nccl().GroupEnd());
        d_rpre = d_rmap.p; d_rbase = d_rmap.p + 2 * (W + 1);
        kb_rebase_off_kernel<<<nblk(Rtot + 1), 256, 0, st>>>(g_off, Rtot, d_rpre, d_rmap.p + (W + 1), (uint32_t)W);
        LAUNCHED(c);
        A = MinArena{g_hash, g_pos, g_off, Rtot};
    }
    const uint64_t nR = A.R;

    // ---- K-B: records -------------------------------------------------------------------------
    Tmp<uint64_t> cnt, kmer_off;
    MDBG_CK(c, cnt.get(c->pool, nR + 1));
    MDBG_CK(c, kmer_off.get(c->pool, nR + 1));
    uint64_t Ktot = 0;  // sightings of the whole job (serial ordinals are 0 .. Ktot-1)
    if (nR > 0) {
        kb_count_kernel<<<nblk(nR + 1), 256, 0, st>>>(A.off, nR, k, cnt);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, kmer_off.p, nR + 1, st); }));
        MDBG_CK(c, cudaMemcpyAsync(&c->h_sc->v[0], kmer_off.p + nR, 8, cudaMemcpyDeviceToHost, st));
        MDBG_CK(c, cudaStreamSynchronize(st));
        Ktot = c->h_sc->v[0];
    }
    if (Ktot >= 0x7FFFFFF0ull) { c->err = "more than 2^31 k-min-mers in the job"; return MDBG_ERR_RANGE; }
    G->n_kminmers = Ktot;
    const int ord_bits = std::max(1, log2_ceil(Ktot + 1));   // serial ordinals are < Ktot
    // records this GPU owns: all of them on one GPU; with N > 1 the windows whose tuple fingerprint
    // falls into this rank's range (every copy of a tuple meets on one owner), in ordinal order
    uint64_t K = Ktot;
    Tmp<uint32_t> own_g;
    if (W > 1 && Ktot > 0) {
        Tmp<uint8_t> own;
        MDBG_CK(c, own.get(c->pool, Ktot)); MDBG_CK(c, own_g.get(c->pool, Ktot));
        kb_own_kernel<<<nblk(Ktot), 256, 0, st>>>(A, kmer_off, Ktot, k, 0x6d64626700000000ull, (uint32_t)W, (uint32_t)rank, own);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<uint32_t>(0), own.p, own_g.p,
                                              (uint32_t*)&c->d_sc->v[0], (uint32_t)Ktot, st);
        }));
        RC(read_scalars(c));
        K = c->h_sc->v[0] & 0xFFFFFFFFu;
    }
    Tmp<uint64_t> r_ord_t, fp;
    Tmp<RecInfo> r_info_t;
    Tmp<uint32_t> wloc, iota;
    MDBG_CK(c, r_ord_t.get(c->pool, K)); MDBG_CK(c, r_info_t.get(c->pool, K)); MDBG_CK(c, fp.get(c->pool, K));
    MDBG_CK(c, wloc.get(c->pool, K)); MDBG_CK(c, iota.get(c->pool, K));
    // The canonical tuples are never materialised: a record is a window of the (L2-resident)
    // arena, read through TupleSrc.  The same pass computes the table fingerprint of the first seed.
    const uint64_t table_seed0 = 0x7461626c65000000ull;
    uint64_t fp_mask0 = ~0ull;
    if (c->p.debug_fp_bits > 0 && c->p.debug_fp_bits < 64) fp_mask0 = (1ull << c->p.debug_fp_bits) - 1;
    if (K) {
        kb_records_kernel<<<nblk(K), 256, 0, st>>>(A, kmer_off, K, k, W > 1 ? own_g.p : nullptr, table_seed0, fp_mask0, d_rpre,
                                                   d_rbase, (uint32_t)W, c->read_base, wloc, r_ord_t, r_info_t, fp, iota);
        LAUNCHED(c);
    }
    const uint64_t* r_ord = r_ord_t; const RecInfo* r_info = r_info_t;
    const TupleSrc T{A.hash, wloc.p, r_ord_t.p, k};
    cnt.reset(); kmer_off.reset(); own_g.reset();
    MDBG_CK(c, cudaEventRecord(c->ev[6], st));

    // ---- K-C table + K-D sort by slot (retry with a new seed on a fingerprint collision) ---------
    uint32_t D = 0, Q_local = 0;
    Tmp<uint32_t> slot, first, sslot, sj, seg_start, seg_index, nseq, seq_off;
    Tmp<uint64_t> first_ord;
    Tmp<uint8_t> solid;
    const int cap_bits = std::max(12, log2_ceil(2 * std::max<uint64_t>(K, 1)));
    if (K > 0) {
        Tmp<uint64_t> keys;
        Tmp<uint8_t> head;
        const uint64_t cap = 1ull << cap_bits;
        MDBG_CK(c, slot.get(c->pool, K));
        MDBG_CK(c, keys.get(c->pool, cap)); MDBG_CK(c, first.get(c->pool, cap));
        MDBG_CK(c, sslot.get(c->pool, K)); MDBG_CK(c, sj.get(c->pool, K));
        MDBG_CK(c, head.get(c->pool, K)); MDBG_CK(c, seg_start.get(c->pool, K + 1));
        for (int attempt = 0;; attempt++) {
            if (attempt >= 8) { c->err = "fingerprint collisions persisted over 8 seeds"; return MDBG_ERR_RANGE; }
            c->tm.table_attempts = attempt + 1;
            if (attempt > 0) {   // attempt 0 came out of kb_records
                uint64_t seed = table_seed0 + 0x9e3779b97f4a7c15ull * (uint64_t)attempt;
                kc_fp_kernel<<<nblk(K), 256, 0, st>>>(T, K, seed, ~0ull, fp, iota);
                LAUNCHED(c);
            }
            MDBG_CK(c, cudaMemsetAsync(keys, 0xFF, cap * 8, st));
            MDBG_CK(c, cudaMemsetAsync(first, 0xFF, cap * 4, st));
            MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[1], 0, 8, st));
            kc_insert_kernel<<<nblk(K * 4), 256, 0, st>>>(fp, K, keys, first, cap - 1, slot);
            LAUNCHED(c);
            kc_verify_kernel<<<nblk(K), 256, 0, st>>>(T, K, slot, first, &c->d_sc->v[1]);
            LAUNCHED(c);
            // K-D: stable sort by slot, segment heads (speculatively: the collision flag is read
            // together with the segment count, one host round trip for both)
            RC(R.cub([&](void* t, size_t& b) {
                return cub::DeviceRadixSort::SortPairs(t, b, slot.p, sslot.p, iota.p, sj.p, (uint32_t)K, 0, cap_bits, st);
            }));
            kd_heads_kernel<<<nblk(K), 256, 0, st>>>(sslot, K, head);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) {
                return cub::DeviceSelect::Flagged(t, b, cub::CountingInputIterator<uint32_t>(0), head.p, seg_start.p,
                                                  (uint32_t*)&c->d_sc->v[2], (uint32_t)K, st);
            }));
            RC(read_scalars(c));
            if (c->h_sc->v[1] == 0) break;  // every slot holds exactly one tuple
        }
        D = (uint32_t)(c->h_sc->v[2] & 0xFFFFFFFFu);
    }
    MDBG_CK(c, cudaEventRecord(c->ev[7], st));   // ms_kc = table + sort by slot, ms_kd = reduce + nodes
    slot.reset(); first.reset(); iota.reset(); sslot.reset(); fp.reset();
    MDBG_CK(c, first_ord.get(c->pool, D)); MDBG_CK(c, solid.get(c->pool, D)); MDBG_CK(c, nseq.get(c->pool, (uint64_t)D + 1));
    MDBG_CK(c, seq_off.get(c->pool, (uint64_t)D + 1)); MDBG_CK(c, seg_index.get(c->pool, D));
    Tmp<uint8_t> counted;
    MDBG_CK(c, counted.get(c->pool, D));
    uint64_t Dtot = 0, Stot = 0, Qtot = 0;
    {
        // ---- node index and node placement from ONE prefix sum over ordinal space -----------------
        // Every distinct tuple marks the ordinal of its first sighting (bit 0: it consumed a node
        // index, bit 1: it is a solid node).  N > 1: the flag arrays are summed over the GPUs (an
        // ordinal belongs to one tuple, hence one owner).  The exclusive scan then holds, at that
        // ordinal, the tuple's node index and the place of the node in the ascending-index list:
        // no sort of first sightings, no binary searches, no sort of nodes.
        Tmp<uint8_t> ord_flags;
        Tmp<uint64_t> rank64;
        MDBG_CK(c, ord_flags.get(c->pool, Ktot + 1)); MDBG_CK(c, rank64.get(c->pool, Ktot + 1));
        MDBG_CK(c, cudaMemsetAsync(ord_flags.p, 0, Ktot + 1, st));
        MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[3], 0, 16, st));
        if (D > 0) {
            kd_segments_kernel<<<nblk(D), 256, 0, st>>>(seg_start, D, K, sj, r_ord, minab, bf, first_ord, counted, solid, nseq,
                                                        ord_flags);
            LAUNCHED(c);
        }
        if (W > 1) NCK(c, nccl().AllReduce(ord_flags.p, ord_flags.p, Ktot + 1, ncclUint8, ncclSum, (ncclComm_t)c->comm, st));
        RC(R.cub([&](void* t, size_t& b) {
            cub::TransformInputIterator<uint64_t, FlagPairToU64, const uint8_t*> in(ord_flags.p, FlagPairToU64());
            return cub::DeviceScan::ExclusiveSum(t, b, in, rank64.p, (uint32_t)(Ktot + 1), st);
        }));
        MDBG_CK(c, cudaMemcpyAsync(&c->d_sc->v[3], rank64.p + Ktot, 8, cudaMemcpyDeviceToDevice, st));
        if (want_seqlines && D > 0) {
            MDBG_CK(c, cudaMemsetAsync(nseq.p + D, 0, 4, st));
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, nseq.p, seq_off.p, D + 1, st); }));
            MDBG_CK(c, cudaMemcpyAsync(&c->d_sc->v[4], seq_off.p + D, 4, cudaMemcpyDeviceToDevice, st));
        }
        RC(read_scalars(c));   // v[3] = tuples in the table | solid nodes << 32 (job-wide), v[4] = own .sequences lines
        Dtot = c->h_sc->v[3] & 0xFFFFFFFFu;
        Stot = c->h_sc->v[3] >> 32;
        Q_local = (uint32_t)(c->h_sc->v[4] & 0xFFFFFFFFu);
        Qtot = Q_local;
        if (W > 1 && want_seqlines) {
            std::vector<uint64_t> allQ;
            uint64_t mineq[1] = {Q_local};
            RC(allgather_u64(c, mineq, 1, allQ));
            Qtot = 0;
            for (int r = 0; r < W; r++) Qtot += allQ[r];
        }
        if (Stot >= 0x7FFFFFF0ull) { c->err = "more than 2^31 nodes"; return MDBG_ERR_RANGE; }
        G->n_distinct = Dtot; G->n_nodes = Stot; G->n_seqlines = want_seqlines ? Qtot : 0;
        // node arrays: every GPU writes the nodes it owns at their final places; N > 1: the arrays
        // (zero elsewhere) are summed so that every GPU holds all nodes for the edge stage
        const uint64_t Sp = Stot + (Stot & 1);   // u16 arrays are reduced as u32 words
        MDBG_CK(c, G->index.get(c->pool, Stot)); MDBG_CK(c, G->abundance.get(c->pool, Sp));
        MDBG_CK(c, G->seqlen.get(c->pool, Stot)); MDBG_CK(c, G->shift.get(c->pool, 2 * Stot));
        MDBG_CK(c, G->tuple.get(c->pool, Stot * k));
        if (W > 1 && Stot > 0) {
            MDBG_CK(c, cudaMemsetAsync(G->index.p, 0, Stot * 4, st)); MDBG_CK(c, cudaMemsetAsync(G->abundance.p, 0, Sp * 2, st));
            MDBG_CK(c, cudaMemsetAsync(G->seqlen.p, 0, Stot * 4, st)); MDBG_CK(c, cudaMemsetAsync(G->shift.p, 0, Stot * 4, st));
            MDBG_CK(c, cudaMemsetAsync(G->tuple.p, 0, Stot * k * 8, st));
        }
        if (D > 0) {
            NodeOut NO{G->index, G->abundance, G->seqlen, G->shift, G->tuple};
            kd_nodes_direct_kernel<<<nblk(D), 256, 0, st>>>(D, minab, K, seg_start, sj, first_ord, counted, solid, rank64, T,
                                                            r_ord, r_info, seg_index, NO);
            LAUNCHED(c);
        }
        if (W > 1 && Stot > 0) {
            ncclComm_t cm = (ncclComm_t)c->comm;
            NCK(c, nccl().GroupStart());
            NCK(c, nccl().AllReduce(G->index.p, G->index.p, Stot, ncclUint32, ncclSum, cm, st));
            NCK(c, nccl().AllReduce(G->abundance.p, G->abundance.p, Sp / 2, ncclUint32, ncclSum, cm, st));
            NCK(c, nccl().AllReduce(G->seqlen.p, G->seqlen.p, Stot, ncclUint32, ncclSum, cm, st));
            NCK(c, nccl().AllReduce(G->shift.p, G->shift.p, Stot, ncclUint32, ncclSum, cm, st));
            NCK(c, nccl().AllReduce(G->tuple.p, G->tuple.p, Stot * k, ncclUint64, ncclSum, cm, st));
            NCK(c, nccl().GroupEnd());
        }
        first_ord.reset(); solid.reset();
    }
    // .sequences lines of this owner, in ordinal (= emission) order
    G->n_seq_local = 0;
    if (want_seqlines) {
        MDBG_CK(c, G->seq.get(c->pool, Q_local));
        G->n_seq_local = Q_local;
        if (Q_local > 0) {
            Tmp<SeqRec> raw;
            Tmp<uint64_t> qk, qk_s; Tmp<uint32_t> qi, qi_s;
            MDBG_CK(c, raw.get(c->pool, Q_local));
            MDBG_CK(c, qk.get(c->pool, Q_local)); MDBG_CK(c, qk_s.get(c->pool, Q_local));
            MDBG_CK(c, qi.get(c->pool, Q_local)); MDBG_CK(c, qi_s.get(c->pool, Q_local));
            kd_seqlines_kernel<<<nblk(D), 256, 0, st>>>(D, K, minab, l, seg_start, sj, nseq, seq_off, seg_index, r_ord,
                                                        r_info, raw);
            LAUNCHED(c);
            kd_seq_keys_kernel<<<nblk(Q_local), 256, 0, st>>>(raw, Q_local, qk, qi);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, qk.p, qk_s.p, qi.p, qi_s.p, Q_local, 0, ord_bits, st); }));
            kd_seq_gather_kernel<<<nblk(Q_local), 256, 0, st>>>(raw, qi_s, Q_local, G->seq);
            LAUNCHED(c);
        }
    }
    MDBG_CK(c, cudaEventRecord(c->ev[8], st));
    // free table-stage scratch before the edge stage
    nseq.reset(); seq_off.reset(); seg_index.reset(); seg_start.reset(); sj.reset();
    wloc.reset(); r_ord_t.reset(); r_info_t.reset(); g_hash.reset(); g_pos.reset(); g_off.reset();

    // ---- K-E: edges of this GPU's slice of the nodes ----------------------------------------------
    uint32_t E = 0;
    uint64_t rem_local = 0;
    if (Stot > 0) {
        const uint32_t S = (uint32_t)Stot;
        NodeView NV{G->index, G->abundance, G->seqlen, G->shift, G->tuple, S, k};
        const uint32_t E2 = 2 * S;
        uint64_t n_lo, n_hi;
        mdbg_shard_reads(S, W, rank, &n_lo, &n_hi);
        const uint32_t q_lo = 2 * (uint32_t)n_lo, q_n = 2 * (uint32_t)(n_hi - n_lo);
        Tmp<uint64_t> ekey, skey; Tmp<uint32_t> eval, sval; Tmp<uint8_t> erev;
        MDBG_CK(c, ekey.get(c->pool, E2)); MDBG_CK(c, skey.get(c->pool, E2));
        MDBG_CK(c, eval.get(c->pool, E2)); MDBG_CK(c, sval.get(c->pool, E2)); MDBG_CK(c, erev.get(c->pool, E2));
        ke_entries_kernel<<<nblk(E2), 256, 0, st>>>(NV, 0x656467657300ull, ekey, eval, erev);
        LAUNCHED(c);
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, ekey.p, skey.p, eval.p, sval.p, E2, 0, 32, st);
        }));
        Tmp<uint32_t> inv;
        MDBG_CK(c, inv.get(c->pool, E2));
        ke_inverse_kernel<<<nblk(E2), 256, 0, st>>>(sval, E2, inv);
        LAUNCHED(c);
        // The join decides every edge ONCE: kept edges / presimp removals are parked in fixed
        // per-query slots together with their counts; one packed scan places them.  Only if some
        // query outgrows its slots (flagged by the kernel) is the exact two-pass form run.
        Tmp<uint64_t> cnt_q, off_q, cap_r;
        Tmp<EdgeRec> cap_e;
        MDBG_CK(c, cnt_q.get(c->pool, (uint64_t)q_n + 1)); MDBG_CK(c, off_q.get(c->pool, (uint64_t)q_n + 1));
        MDBG_CK(c, cap_e.get(c->pool, (uint64_t)q_n * KE_CAP_E)); MDBG_CK(c, cap_r.get(c->pool, (uint64_t)q_n * KE_CAP_R));
        MDBG_CK(c, cudaMemsetAsync(cnt_q.p + q_n, 0, 8, st));
        MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[5], 0, 24, st));   // v[5] packed totals, v[6] overflow, v[7] kept edges
        uint32_t EP = 0, NR = 0;
        bool overflow = false;
        if (q_n > 0) {
            ke_join_kernel<0><<<nblk(q_n, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, inv, presimp, q_lo, q_n, cnt_q, nullptr,
                                                              cap_e, cap_r, &c->d_sc->v[6]);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt_q.p, off_q.p, q_n + 1, st); }));
            MDBG_CK(c, cudaMemcpyAsync(&c->d_sc->v[5], off_q.p + q_n, 8, cudaMemcpyDeviceToDevice, st));
            RC(read_scalars(c));
            EP = (uint32_t)(c->h_sc->v[5] & 0xFFFFFFFFu);
            NR = (uint32_t)(c->h_sc->v[5] >> 32);
            overflow = c->h_sc->v[6] != 0;
        }
        rem_local = presimp > 0.0f ? NR : 0;
        Tmp<EdgeRec> pend, kept;
        Tmp<uint64_t> removed;
        MDBG_CK(c, pend.get(c->pool, EP)); MDBG_CK(c, removed.get(c->pool, NR));
        if (EP > 0 || NR > 0) {
            if (overflow) {
                ke_join_kernel<1><<<nblk(q_n, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, inv, presimp, q_lo, q_n, nullptr,
                                                                  off_q, pend, removed, nullptr);
                LAUNCHED(c);
                ke_group_sort_kernel<<<nblk(q_n / 2), 256, 0, st>>>(pend, off_q, q_n / 2);
                LAUNCHED(c);
            } else {   // canonical order: nodes ascend already, each node's edges are sorted while they move
                ke_compact_kernel<<<nblk(q_n / 2), 256, 0, st>>>(cap_e, cap_r, cnt_q, off_q, q_n / 2, pend, removed);
                LAUNCHED(c);
            }
        }
        cap_e.reset(); cap_r.reset();
        // presimp removals of every GPU (an edge is dropped if it or its reverse was removed anywhere)
        std::vector<uint64_t> allNR;
        { uint64_t mine[1] = {NR}; RC(allgather_u64(c, mine, 1, allNR)); }
        uint64_t NRtot = 0;
        for (int r = 0; r < W; r++) NRtot += allNR[r];
        EdgeRec* edges = pend;
        E = EP;
        if (NRtot > 0) {
            Tmp<uint64_t> rem_all, rem_s;
            Tmp<uint8_t> keep;
            MDBG_CK(c, rem_all.get(c->pool, NRtot)); MDBG_CK(c, rem_s.get(c->pool, NRtot));
            RC(allgatherv(c, removed, NR, allNR, rem_all, 8));
            RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortKeys(t, b, rem_all.p, rem_s.p, (uint32_t)NRtot, 0, 64, st); }));
            if (EP > 0) {
                MDBG_CK(c, keep.get(c->pool, EP)); MDBG_CK(c, kept.get(c->pool, EP));
                ke_filter_kernel<<<nblk(EP), 256, 0, st>>>(pend, EP, rem_s, (uint32_t)NRtot, keep);
                LAUNCHED(c);
                RC(R.cub([&](void* t, size_t& b) {
                    return cub::DeviceSelect::Flagged(t, b, pend.p, keep.p, kept.p, (uint32_t*)&c->d_sc->v[7], EP, st);
                }));
                RC(read_scalars(c));
                E = (uint32_t)(c->h_sc->v[7] & 0xFFFFFFFFu);
                edges = kept;
            }
        }
        MDBG_CK(c, G->e_n1.get(c->pool, E)); MDBG_CK(c, G->e_o1.get(c->pool, E)); MDBG_CK(c, G->e_n2.get(c->pool, E));
        MDBG_CK(c, G->e_o2.get(c->pool, E)); MDBG_CK(c, G->e_ov.get(c->pool, E));
        if (E > 0) {   // (the stable compaction above kept the canonical order)
            EdgeOut EO{G->e_n1, G->e_o1, G->e_n2, G->e_o2, G->e_ov};
            ke_gather_kernel<<<nblk(E), 256, 0, st>>>(edges, nullptr, E, EO);
            LAUNCHED(c);
        }
    }
    G->n_edges_local = E;
    {   // job-wide counters
        std::vector<uint64_t> all;
        uint64_t mine[2] = {E, rem_local};
        RC(allgather_u64(c, mine, 2, all));
        G->n_edges = 0; G->presimp_removed = 0;
        for (int r = 0; r < W; r++) { G->n_edges += all[2 * r]; G->presimp_removed += all[2 * r + 1]; }
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
    cudaEventElapsedTime(&c->tm.ms_kb, c->ev[5], c->ev[6]);
    cudaEventElapsedTime(&c->tm.ms_kc, c->ev[6], c->ev[7]);
    cudaEventElapsedTime(&c->tm.ms_kd, c->ev[7], c->ev[8]);
    cudaEventElapsedTime(&c->tm.ms_ke, c->ev[8], c->ev[9]);
    cudaEventElapsedTime(&c->tm.ms_total_finish, c->ev[5], c->ev[9]);
    return MDBG_OK;
}

// Owner of the host arrays handed out through mdbg_graph: ONE pinned block (so the D2H copies run
// at PCIe speed), carved into the arrays; blocks of freed graphs are recycled through the context.
struct HostGraph {
    mdbg_ctx* c = nullptr;
    void* block = nullptr;
    size_t bytes = 0, used = 0;
    template <class T>
    T* take(uint64_t n) {
        used = (used + 63) & ~(size_t)63;
        T* p = (T*)((char*)block + used);
        used += n * sizeof(T);
        return p;
    }
};

int host_block(mdbg_ctx* c, HostGraph* H, size_t need) {
    need = (need + 4095) & ~(size_t)4095;
    for (size_t i = 0; i < c->pinned_cache.size(); i++) {
        if (c->pinned_cache[i].first >= need && c->pinned_cache[i].first <= 2 * need + (1u << 20)) {
            H->block = c->pinned_cache[i].second;
            H->bytes = c->pinned_cache[i].first;
            c->pinned_cache.erase(c->pinned_cache.begin() + i);
            return MDBG_OK;
        }
    }
    MDBG_CK(c, cudaMallocHost(&H->block, need));
    H->bytes = need;
    return MDBG_OK;
}

template <class T>
int d2h(mdbg_ctx* c, T* dst, const T* d, uint64_t n) {
    if (n) MDBG_CK(c, cudaMemcpyAsync(dst, d, n * sizeof(T), cudaMemcpyDeviceToHost, c->st));
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

// Host copy.  With N GPUs every rank gets all nodes; rank 0 additionally gets ALL edges and
// .sequences lines (gathered over NCCL), the other ranks their own slices.
int mdbg_finish(mdbg_ctx* c, int want_seqlines, mdbg_graph* out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    int rc = build_device_graph(c, want_seqlines != 0);
    if (rc) return rc;
    RC(finish_timings(c));
    fill_counters(c, out);
    DeviceGraph* G = c->dg;
    HostGraph* H = new HostGraph();
    H->c = c;
    out->_owner = H;
    const int W = c->world;
    const uint64_t S = G->n_nodes, k = G->k;
    MDBG_CK(c, cudaEventRecord(c->ev[13], c->st));
    // edges / seqlines: gather on rank 0
    uint64_t E = G->n_edges_local, Q = want_seqlines ? G->n_seq_local : 0;
    Tmp<uint32_t> g_n1, g_n2, g_ov; Tmp<uint8_t> g_o1, g_o2; Tmp<SeqRec> g_seq, g_seq_s;
    const uint32_t* p_n1 = G->e_n1; const uint32_t* p_n2 = G->e_n2; const uint32_t* p_ov = G->e_ov;
    const uint8_t* p_o1 = G->e_o1; const uint8_t* p_o2 = G->e_o2; const SeqRec* p_seq = G->seq;
    if (W > 1) {
        std::vector<uint64_t> all;
        uint64_t mine[2] = {E, Q};
        RC(allgather_u64(c, mine, 2, all));
        std::vector<uint64_t> ec(W), qc(W);
        uint64_t Et = 0, Qt = 0;
        for (int r = 0; r < W; r++) { ec[r] = all[2 * r]; qc[r] = all[2 * r + 1]; Et += ec[r]; Qt += qc[r]; }
        uint64_t En = c->rank == 0 ? Et : 0, Qn = c->rank == 0 ? Qt : 0;
        MDBG_CK(c, g_n1.get(c->pool, En)); MDBG_CK(c, g_n2.get(c->pool, En)); MDBG_CK(c, g_ov.get(c->pool, En));
        MDBG_CK(c, g_o1.get(c->pool, En)); MDBG_CK(c, g_o2.get(c->pool, En)); MDBG_CK(c, g_seq.get(c->pool, Qn));
        NCK(c, nccl().GroupStart());
        RC(gatherv_root(c, G->e_n1, E, ec, g_n1, 4)); RC(gatherv_root(c, G->e_n2, E, ec, g_n2, 4));
        RC(gatherv_root(c, G->e_ov, E, ec, g_ov, 4)); RC(gatherv_root(c, G->e_o1, E, ec, g_o1, 1));
        RC(gatherv_root(c, G->e_o2, E, ec, g_o2, 1));
        if (want_seqlines) RC(gatherv_root(c, G->seq, Q, qc, g_seq, sizeof(SeqRec)));
        NCK(c, nccl().GroupEnd());
        if (c->rank == 0) {   // slices are contiguous node ranges: the concatenation is already sorted
            E = Et; p_n1 = g_n1; p_n2 = g_n2; p_ov = g_ov; p_o1 = g_o1; p_o2 = g_o2;
            if (want_seqlines && Qt > 0) {   // emission order = ordinal order over all owners
                Runner R{c};
                Tmp<uint64_t> qk, qk_s; Tmp<uint32_t> qi, qi_s;
                MDBG_CK(c, qk.get(c->pool, Qt)); MDBG_CK(c, qk_s.get(c->pool, Qt));
                MDBG_CK(c, qi.get(c->pool, Qt)); MDBG_CK(c, qi_s.get(c->pool, Qt)); MDBG_CK(c, g_seq_s.get(c->pool, Qt));
                kd_seq_keys_kernel<<<nblk(Qt), 256, 0, c->st>>>(g_seq, (uint32_t)Qt, qk, qi);
                RC(R.cub([&](void* t, size_t& b) { return cub::DeviceRadixSort::SortPairs(t, b, qk.p, qk_s.p, qi.p, qi_s.p, (uint32_t)Qt, 0, 63, c->st); }));
                kd_seq_gather_kernel<<<nblk(Qt), 256, 0, c->st>>>(g_seq, qi_s, (uint32_t)Qt, g_seq_s);
                MDBG_CK(c, cudaGetLastError());
                p_seq = g_seq_s;
            }
            Q = Qt;
        }
    }
    Tmp<uint32_t> q_index; Tmp<uint64_t> q_read, q_start, q_end, q_shift; Tmp<uint8_t> q_rev;
    if (want_seqlines) {
        MDBG_CK(c, q_index.get(c->pool, Q)); MDBG_CK(c, q_read.get(c->pool, Q)); MDBG_CK(c, q_start.get(c->pool, Q));
        MDBG_CK(c, q_end.get(c->pool, Q)); MDBG_CK(c, q_rev.get(c->pool, Q)); MDBG_CK(c, q_shift.get(c->pool, 2 * Q));
        if (Q > 0) {
            Tmp<uint32_t> idn;
            MDBG_CK(c, idn.get(c->pool, Q));
            iota_kernel<<<nblk(Q), 256, 0, c->st>>>(idn, (uint32_t)Q);
            SeqOut SO{q_index, q_read, q_start, q_end, q_rev, q_shift};
            kd_unpack_seq_kernel<<<nblk(Q), 256, 0, c->st>>>(p_seq, idn, (uint32_t)Q, SO);
            MDBG_CK(c, cudaGetLastError());
        }
    }
    size_t need = S * (4 + 2 + 4 + 4 + 8 * k) + E * 14 + Q * (4 + 8 * 3 + 1 + 16) + 64 * 20;
    RC(host_block(c, H, need));
    out->node_index = H->take<uint32_t>(S); out->abundance = H->take<uint16_t>(S); out->seqlen = H->take<uint32_t>(S);
    out->shift = H->take<uint16_t>(2 * S); out->tuple = H->take<uint64_t>(S * k);
    out->e_n1 = H->take<uint32_t>(E); out->e_o1 = H->take<uint8_t>(E); out->e_n2 = H->take<uint32_t>(E);
    out->e_o2 = H->take<uint8_t>(E); out->e_overlap = H->take<uint32_t>(E);
    RC(d2h(c, out->node_index, G->index.p, S)); RC(d2h(c, out->abundance, G->abundance.p, S));
    RC(d2h(c, out->seqlen, G->seqlen.p, S)); RC(d2h(c, out->shift, G->shift.p, 2 * S));
    RC(d2h(c, out->tuple, G->tuple.p, S * k));
    RC(d2h(c, out->e_n1, p_n1, E)); RC(d2h(c, out->e_o1, p_o1, E)); RC(d2h(c, out->e_n2, p_n2, E));
    RC(d2h(c, out->e_o2, p_o2, E)); RC(d2h(c, out->e_overlap, p_ov, E));
    out->n_edges = (W > 1 && c->rank != 0) ? E : out->n_edges;
    if (want_seqlines) {
        out->q_index = H->take<uint32_t>(Q); out->q_read = H->take<uint64_t>(Q); out->q_start = H->take<uint64_t>(Q);
        out->q_end = H->take<uint64_t>(Q); out->q_reversed = H->take<uint8_t>(Q); out->q_shift = H->take<uint64_t>(2 * Q);
        RC(d2h(c, out->q_index, q_index.p, Q)); RC(d2h(c, out->q_read, q_read.p, Q)); RC(d2h(c, out->q_start, q_start.p, Q));
        RC(d2h(c, out->q_end, q_end.p, Q)); RC(d2h(c, out->q_reversed, q_rev.p, Q)); RC(d2h(c, out->q_shift, q_shift.p, 2 * Q));
        out->n_seqlines = (W > 1 && c->rank != 0) ? Q : out->n_seqlines;
    }
    MDBG_CK(c, cudaEventRecord(c->ev[14], c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    cudaEventElapsedTime(&c->tm.ms_d2h, c->ev[13], c->ev[14]);
    return MDBG_OK;
}

// Graphs must be freed before their context is destroyed (the pinned block returns to its cache).
void mdbg_graph_free(mdbg_graph* g) {
    if (!g) return;
    HostGraph* H = (HostGraph*)g->_owner;
    if (H) {
        if (H->block) {
            if (H->c && H->c->pinned_cache.size() < 8) H->c->pinned_cache.emplace_back(H->bytes, H->block);
            else cudaFreeHost(H->block);
        }
        delete H;
    }
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

// --read-stats (main.rs:939-975): for every k-min-mer of every read of the batch, the abundance of its
// canonical tuple among the nodes of the last mdbg_finish (`dbg_nodes` after the abundance filter), 0 when
// it is none of them.  The batch goes through the same K-A as pushed reads (appended to the minimizer arena
// and rolled back afterwards); the nodes are indexed by a sort of their tuple fingerprints.
int mdbg_read_stats(mdbg_ctx* c, const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads,
                    uint32_t* out_counts, uint64_t* out_read_off, uint64_t cap, uint64_t* n_out) {
    if (!c || !read_off) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    DeviceGraph* G = c->dg;
    if (!G) { c->err = "mdbg_read_stats needs the graph of a previous mdbg_finish"; return MDBG_ERR_BAD_ARG; }
    if (G->k != c->p.k) { c->err = "k changed since the last mdbg_finish"; return MDBG_ERR_BAD_ARG; }
    cudaStream_t st = c->st;
    const uint32_t k = c->p.k;
    const uint64_t M0 = c->M, R0 = c->R, B0 = c->n_bases;
    int rc = mdbg_push_reads(c, bases, read_off, n_reads);      // minimizers of the batch: arena [M0, M), reads [R0, R)
    if (rc) { c->M = M0; c->R = R0; c->n_bases = B0; return rc; }
    struct Rollback { mdbg_ctx* c; uint64_t M, R, B; ~Rollback() { c->M = M; c->R = R; c->n_bases = B; } } rollback{c, M0, R0, B0};
    Runner R{c};
    Tmp<uint64_t> cnt, kmer_off;
    MDBG_CK(c, cnt.get(c->pool, n_reads + 1));
    MDBG_CK(c, kmer_off.get(c->pool, n_reads + 1));
    const uint64_t* q_off = c->m_off + R0;                      // absolute arena offsets of the batch's reads
    kb_count_kernel<<<nblk(n_reads + 1), 256, 0, st>>>(q_off, n_reads, k, cnt);
    MDBG_CK(c, cudaGetLastError());
    RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, kmer_off.p, n_reads + 1, st); }));
    std::vector<uint64_t> ko(n_reads + 1);
    MDBG_CK(c, cudaMemcpyAsync(ko.data(), kmer_off, (n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    const uint64_t K = ko[n_reads];
    if (n_out) *n_out = K;
    if (out_read_off) memcpy(out_read_off, ko.data(), (n_reads + 1) * 8);
    if (K > cap) { c->err = "output capacity too small"; return MDBG_ERR_CAPACITY; }
    if (K == 0) return MDBG_OK;
    if (K >= 0xFFFFFFF0ull) { c->err = "more than 2^32 k-min-mers in one read-stats batch"; return MDBG_ERR_RANGE; }
    Tmp<uint32_t> d_out;
    MDBG_CK(c, d_out.get(c->pool, K));
    const uint32_t S = (uint32_t)G->n_nodes;
    const uint64_t seed = 0x7273746174730000ull;
    Tmp<uint64_t> fp, sfp;
    Tmp<uint32_t> pos, spos;
    MDBG_CK(c, fp.get(c->pool, S)); MDBG_CK(c, sfp.get(c->pool, S));
    MDBG_CK(c, pos.get(c->pool, S)); MDBG_CK(c, spos.get(c->pool, S));
    if (S) {
        rs_node_fp_kernel<<<nblk(S), 256, 0, st>>>(G->tuple, S, k, seed, fp, pos);
        MDBG_CK(c, cudaGetLastError());
        RC(R.cub([&](void* t, size_t& b) {
            return cub::DeviceRadixSort::SortPairs(t, b, fp.p, sfp.p, pos.p, spos.p, (int)S, 0, 64, st);
        }));
    }
    MinArena A{c->m_hash, c->m_pos, q_off, n_reads};
    rs_lookup_kernel<<<nblk(K), 256, 0, st>>>(A, kmer_off, K, k, seed, sfp, spos, S, G->tuple, G->abundance, d_out);
    MDBG_CK(c, cudaGetLastError());
    if (out_counts) MDBG_CK(c, cudaMemcpyAsync(out_counts, d_out, K * 4, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    return MDBG_OK;
}

}  // extern "C"
