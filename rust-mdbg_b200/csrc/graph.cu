// graph.cu -- mdbg_finish / mdbg_finish_device / mdbg_window: host orchestration of K-B .. K-E
// (kernels in graph_kernels.cuh; radix sort / scan / select are CUB device-wide primitives).
//
// Serial-order semantics: index = rank of a tuple's first sighting, abundance =
// sightings (u16), seqlen/shift/sequence from the minabund-th sighting.
//
// One code path for 1 and N GPUs (one process per GPU, NCCL over NVLink).  Per-GPU work is
// that GPU's share of the job (its reads, the tuples it owns, its slice of the nodes):
//   1. K-B: every GPU turns the k-min-mer sightings of ITS reads into records {fingerprint, ordinal,
//      window location, RecInfo} (44 bytes; the canonical tuple is a window of the hash arena, never
//      materialised)
//   2. N > 1: the records are bucketed by the fingerprint's prefix (range partition, stable) and cross
//      NVLink in ONE all-to-all (grouped ncclSend/ncclRecv after an N x N count exchange), so every copy
//      of a tuple meets on one owner; the hash arenas (8 B per minimizer) are all-gathered at a fixed
//      pitch so that an owner can read the tuples of the records it received
//   3. owner: open-address table (fingerprint placed, tuple verified), stable radix sort by slot,
//      segmented reduce -> abundance / first sighting / representative sighting
//   4. node index and node placement = ONE exclusive scan over a 2-bit-per-ordinal bitmap of the
//      first sightings (N > 1: the bitmaps are summed with one all-reduce: Ktot / 4 bytes)
//   5. every owner writes a 20-byte NodeRec at the final place of each of its nodes (N > 1: summed with
//      one all-reduce; tuples are not shipped, they are windows of the gathered arena); every GPU expands
//      the node arrays, builds the (k-1)-mer entry index and emits the edges of ITS slice of the nodes;
//      presimp removals are all-gathered before the final filter
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cub/cub.cuh>
#include <memory>
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

void mdbg_inbox_release(mdbg_ctx* c);   // comm.cu
void mdbg_inbox_unmap_peers(mdbg_ctx* c);

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

// A group of sends / receives that is always closed, whatever fails inside it (an open group would leave
// the peers hanging): the first error is kept and reported after ncclGroupEnd.
struct NcclGroup {
    mdbg_ctx* c;
    ncclResult_t bad = ncclSuccess;
    bool open = false;
    explicit NcclGroup(mdbg_ctx* ctx) : c(ctx) {
        bad = nccl().GroupStart();
        open = bad == ncclSuccess;
    }
    void send(const void* p, size_t bytes, int peer) {
        if (bytes && bad == ncclSuccess) bad = nccl().Send(p, bytes, ncclChar, peer, (ncclComm_t)c->comm, c->st);
    }
    void recv(void* p, size_t bytes, int peer) {
        if (bytes && bad == ncclSuccess) bad = nccl().Recv(p, bytes, ncclChar, peer, (ncclComm_t)c->comm, c->st);
    }
    int close() {
        if (open) { ncclResult_t r = nccl().GroupEnd(); open = false; if (bad == ncclSuccess) bad = r; }
        if (bad != ncclSuccess) { c->err = std::string("NCCL send/recv group: ") + nccl().GetErrorString(bad); return MDBG_ERR_NCCL; }
        return MDBG_OK;
    }
    ~NcclGroup() { if (open) nccl().GroupEnd(); }
};

// all-gather of variable-length arrays: every rank ends with the concatenation in rank order
int allgatherv(mdbg_ctx* c, const void* mine, uint64_t my_cnt, const std::vector<uint64_t>& cnt, void* out,
               size_t elem) {
    const int W = c->world;
    if (W == 1) {
        if (my_cnt && out != mine) MDBG_CK(c, cudaMemcpyAsync(out, mine, my_cnt * elem, cudaMemcpyDeviceToDevice, c->st));
        return MDBG_OK;
    }
    NcclGroup g(c);
    size_t ro = 0;
    for (int p = 0; p < W; p++) {
        g.send(mine, my_cnt * elem, p);
        g.recv((char*)out + ro, cnt[p] * elem, p);
        ro += cnt[p] * elem;
    }
    return g.close();
}

// gather variable-length arrays on rank 0 (concatenated in rank order); `g` is the caller's open group
void gatherv_root(mdbg_ctx* c, NcclGroup& g, const void* mine, uint64_t my_cnt, const std::vector<uint64_t>& cnt,
                  void* out, size_t elem) {
    g.send(mine, my_cnt * elem, 0);
    if (c->rank == 0) {
        size_t ro = 0;
        for (int p = 0; p < c->world; p++) {
            g.recv((char*)out + ro, cnt[p] * elem, p);
            ro += cnt[p] * elem;
        }
    }
}

// all-to-all of `narr` SoA arrays: send[a] holds scnt[0] items for rank 0, then scnt[1] for rank 1, ...;
// recv[a] receives rcnt[0] items from rank 0, then rcnt[1] from rank 1, ...  One NCCL group; the group is
// always closed, whatever fails inside it.
int alltoallv(mdbg_ctx* c, int narr, const void* const* send, void* const* recv, const size_t* elem,
              const uint64_t* scnt, const uint64_t* rcnt) {
    const int W = c->world;
    NcclApi& N = nccl();
    ncclResult_t bad = ncclSuccess;
    cudaError_t cbad = cudaSuccess;
    ncclResult_t r = N.GroupStart();
    if (r != ncclSuccess) { c->err = std::string("ncclGroupStart: ") + N.GetErrorString(r); return MDBG_ERR_NCCL; }
    for (int a = 0; a < narr; a++) {
        size_t so = 0, ro = 0;
        for (int p = 0; p < W; p++) {
            const size_t sb = scnt[p] * elem[a], rb = rcnt[p] * elem[a];
            if (p == c->rank) {
                if (sb && cbad == cudaSuccess)
                    cbad = cudaMemcpyAsync((char*)recv[a] + ro, (const char*)send[a] + so, sb, cudaMemcpyDeviceToDevice, c->st);
            } else {
                if (sb && bad == ncclSuccess) bad = N.Send((const char*)send[a] + so, sb, ncclChar, p, (ncclComm_t)c->comm, c->st);
                if (rb && bad == ncclSuccess) bad = N.Recv((char*)recv[a] + ro, rb, ncclChar, p, (ncclComm_t)c->comm, c->st);
            }
            so += sb; ro += rb;
        }
    }
    r = N.GroupEnd();
    if (bad == ncclSuccess) bad = r;
    if (bad != ncclSuccess) { c->err = std::string("all-to-all: ") + N.GetErrorString(bad); return MDBG_ERR_NCCL; }
    if (cbad != cudaSuccess) { c->err = std::string("all-to-all (local part): ") + cudaGetErrorString(cbad); return MDBG_ERR_CUDA; }
    return MDBG_OK;
}

// Collective: make every rank's record inbox hold `cap_records` and map the peers' inboxes (CUDA IPC).  Every rank
// calls it with the same value.  Sets c->p2p_state = 1 on success, -1 when any rank could not map a peer (then the
// records travel with ncclSend/ncclRecv).
int ensure_inbox(mdbg_ctx* c, uint64_t cap_records) {
    if (c->p2p_state < 0) return MDBG_OK;
    if (c->p2p_state == 1 && c->inbox_cap >= cap_records) return MDBG_OK;
    const int W = c->world;
    cudaStream_t st = c->st;
    // nobody may still be writing into / reading from the old inboxes: the previous finish ended with a collective
    // and a stream synchronisation on every rank, and this call follows an all-gather of this finish.  Growing:
    // every rank unmaps its peers, a barrier, then every rank frees its own inbox.
    if (c->inbox) {
        mdbg_inbox_unmap_peers(c);
        NCK(c, nccl().AllReduce(c->d_mail, c->d_mail, 1, ncclUint64, ncclSum, (ncclComm_t)c->comm, st));
        MDBG_CK(c, cudaStreamSynchronize(st));
    }
    mdbg_inbox_release(c);
    const uint64_t cap = cap_records + cap_records / 4 + 4096;
    int ok = cudaMalloc(&c->inbox, InboxLayout::bytes(cap)) == cudaSuccess ? 1 : 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && cudaIpcGetMemHandle(&mine, c->inbox) != cudaSuccess) ok = 0;
    (void)cudaGetLastError();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // handles (64 B) + ok flag (8 B) of every rank: 9 words each, through the mailbox
    uint64_t h_words[9];
    memcpy(h_words, &mine, 64);
    h_words[8] = (uint64_t)ok;
    uint64_t* d_in = c->d_mail + 64 + MAX_WORLD * MAX_WORLD;            // scratch behind the count matrix
    uint64_t* d_all = d_in + 16;
    MDBG_CK(c, cudaMemcpyAsync(d_in, h_words, 72, cudaMemcpyHostToDevice, st));
    NCK(c, nccl().AllGather(d_in, d_all, 9, ncclUint64, (ncclComm_t)c->comm, st));
    std::vector<uint64_t> all(9 * (size_t)W);
    MDBG_CK(c, cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    bool all_ok = true;
    for (int p = 0; p < W; p++) all_ok = all_ok && all[9 * p + 8] == 1;
    int mapped = all_ok ? 1 : 0;
    if (all_ok) {
        for (int p = 0; p < W && mapped; p++) {
            if (p == c->rank) { c->peer_inbox[p] = c->inbox; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, &all[9 * p], 64);
            if (cudaIpcOpenMemHandle(&c->peer_inbox[p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                (void)cudaGetLastError();
                c->peer_inbox[p] = nullptr;
                mapped = 0;
            }
        }
    }
    // every rank must take the same path: agree on the outcome
    uint64_t flag = (uint64_t)mapped;
    MDBG_CK(c, cudaMemcpyAsync(d_in, &flag, 8, cudaMemcpyHostToDevice, st));
    NCK(c, nccl().AllGather(d_in, d_all, 1, ncclUint64, (ncclComm_t)c->comm, st));
    MDBG_CK(c, cudaMemcpyAsync(all.data(), d_all, (size_t)W * 8, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    bool everyone = true;
    for (int p = 0; p < W; p++) everyone = everyone && all[p] == 1;
    if (everyone) { c->inbox_cap = cap; c->p2p_state = 1; }
    else { mdbg_inbox_release(c); c->p2p_state = -1; }
    return MDBG_OK;
}

// Everything from the resident minimizers to the device graph (installed in the context only on success).
int build_device_graph(mdbg_ctx* c, bool want_seqlines) {
    mdbg_graph_device_free(c);
    std::unique_ptr<DeviceGraph> Gp(new DeviceGraph());
    DeviceGraph* G = Gp.get();
    const uint32_t k = c->p.k, l = c->p.l, minab = c->p.min_abundance;
    const float presimp = c->p.presimp;
    const uint32_t bf = (c->p.bf && minab > 1) ? 1 : 0;   // main.rs:639
    const int W = c->world, rank = c->rank;
    G->k = k;
    c->tm.launches_finish = 0;
    c->tm.table_attempts = 0;
    c->tm.ms_kb = c->tm.ms_kc = c->tm.ms_kd = c->tm.ms_ke = 0;
    c->tm.ms_exchange = 0; c->tm.exchange_bytes = 0;
    cudaStream_t st = c->st;
    Runner R{c};
    if (W > 1 && !c->comm) { c->err = "world > 1 but mdbg_comm_init was not called"; return MDBG_ERR_BAD_ARG; }
    if (c->M >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers on one GPU"; return MDBG_ERR_RANGE; }
    MDBG_CK(c, cudaEventRecord(c->ev[5], st));
    const MinArena A{c->m_hash, c->m_pos, c->m_off, c->R};   // this GPU's reads
    const uint64_t nR = A.R;

    // ---- K-B: the sightings of this GPU's reads ---------------------------------------------------
    Tmp<uint64_t> cnt, kmer_off;
    MDBG_CK(c, cnt.get(c->pool, nR + 1));
    MDBG_CK(c, kmer_off.get(c->pool, nR + 1));
    kb_count_kernel<<<nblk(nR + 1), 256, 0, st>>>(A.off, nR, k, cnt);   // nR == 0: a single zero
    LAUNCHED(c);
    RC(R.cub([&](void* t, size_t& b) { return cub::DeviceScan::ExclusiveSum(t, b, cnt.p, kmer_off.p, nR + 1, st); }));
    // sizes of every rank: {M, R, first read, K}; one host round trip (also on one GPU: K sizes the table)
    kx_sizes_kernel<<<1, 1, 0, st>>>(c->d_mail, c->M, c->R, c->read_base_set ? c->read_base : ~0ull, kmer_off.p + nR);
    LAUNCHED(c);
    if (W > 1) NCK(c, nccl().AllGather(c->d_mail, c->d_mail + 4, 4, ncclUint64, (ncclComm_t)c->comm, st));
    else MDBG_CK(c, cudaMemcpyAsync(c->d_mail + 4, c->d_mail, 32, cudaMemcpyDeviceToDevice, st));
    MDBG_CK(c, cudaMemcpyAsync(c->h_mail, c->d_mail + 4, (size_t)W * 32, cudaMemcpyDeviceToHost, st));
    MDBG_CK(c, cudaStreamSynchronize(st));
    uint64_t Ktot = 0, kbase = 0, Mpitch = 0, rbase = 0, rcount = 0;
    for (int r = 0; r < W; r++) {
        const uint64_t* v = c->h_mail + 4 * r;
        if (r == rank) { kbase = Ktot; rbase = v[2] != ~0ull ? v[2] : rcount; }
        Ktot += v[3]; rcount += v[1];
        Mpitch = std::max(Mpitch, v[0]);
    }
    const uint64_t K_local = c->h_mail[4 * rank + 3];
    if (W == 1) rbase = c->read_base;   // one GPU: reads are numbered in push order from read_base (0 unless set)
    Mpitch = (Mpitch + 3) & ~3ull;      // 32-byte aligned rows of the gathered arena
    if (K_local >= 0x7FFFFFF0ull) { c->err = "more than 2^31 k-min-mers on one GPU"; return MDBG_ERR_RANGE; }
    if (W > 1 && Mpitch * (uint64_t)W >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers in the job"; return MDBG_ERR_RANGE; }
    G->n_kminmers = Ktot;
    const int ord_bits = std::max(1, log2_ceil(Ktot + 1));   // serial ordinals are < Ktot

    // N > 1: the hash arenas of all GPUs at a fixed pitch (8 B per minimizer; positions stay at home: what a
    // sighting needs from them travels inside its record)
    Tmp<uint64_t> g_hash;
    const uint64_t* arena = c->m_hash;
    bool arena_pending = false;   // the all-gather runs on the copy stream (own communicator) under K-B / the exchange
    struct PendingGuard {         // an error return must not release g_hash while the copy stream still fills it
        bool& pending; cudaStream_t s;
        ~PendingGuard() { if (pending) cudaStreamSynchronize(s); }
    } pending_guard{arena_pending, c->st_copy};
    if (W > 1) {
        MDBG_CK(c, g_hash.get(c->pool, Mpitch * (uint64_t)W));
        cudaStream_t sa = c->comm2 ? c->st_copy : st;
        ncclComm_t ca = (ncclComm_t)(c->comm2 ? c->comm2 : c->comm);
        if (c->comm2) {
            MDBG_CK(c, cudaEventRecord(c->ev[19], st));
            MDBG_CK(c, cudaStreamWaitEvent(sa, c->ev[19], 0));
        }
        if (c->M) MDBG_CK(c, cudaMemcpyAsync(g_hash.p + Mpitch * rank, c->m_hash, c->M * 8, cudaMemcpyDeviceToDevice, sa));
        if (Mpitch) NCK(c, nccl().AllGather(g_hash.p + Mpitch * rank, g_hash.p, Mpitch, ncclUint64, ca, sa));
        if (c->comm2) { MDBG_CK(c, cudaEventRecord(c->ev[20], sa)); arena_pending = true; }
        arena = g_hash.p;
        c->tm.exchange_bytes += Mpitch * 8 * (uint64_t)(W - 1);
    }

    // records of this GPU's sightings.  The canonical tuples are never materialised: a record is a window
    // of the arena, read through TupleSrc.  The same pass computes the table fingerprint of the first seed
    // and (N > 1) the owner of the tuple.
    Tmp<uint64_t> l_ord, l_fp;
    Tmp<RecInfo> l_info;
    Tmp<uint32_t> l_wloc, iota;
    Tmp<uint8_t> l_owner;
    MDBG_CK(c, l_ord.get(c->pool, K_local)); MDBG_CK(c, l_info.get(c->pool, K_local)); MDBG_CK(c, l_fp.get(c->pool, K_local));
    MDBG_CK(c, l_wloc.get(c->pool, K_local));
    if (W > 1) MDBG_CK(c, l_owner.get(c->pool, K_local));
    else MDBG_CK(c, iota.get(c->pool, K_local));
    const uint64_t table_seed0 = 0x7461626c65000000ull;
    uint64_t fp_mask0 = ~0ull;
    if (c->p.debug_fp_bits > 0 && c->p.debug_fp_bits < 64) fp_mask0 = (1ull << c->p.debug_fp_bits) - 1;
    MDBG_CK(c, cudaEventRecord(c->evk[0], st));
    if (K_local) {
        kb_records_kernel<<<nblk(K_local), 256, 0, st>>>(A, kmer_off, K_local, k, table_seed0, fp_mask0, rbase, kbase,
                                                         W > 1 ? Mpitch * rank : 0, (uint32_t)W, l_wloc, l_ord, l_info, l_fp,
                                                         W > 1 ? nullptr : iota.p, W > 1 ? l_owner.p : nullptr);
        LAUNCHED(c);
    }
    MDBG_CK(c, cudaEventRecord(c->evk[1], st));
    cnt.reset(); kmer_off.reset();
    MDBG_CK(c, cudaEventRecord(c->ev[6], st));

    // ---- N > 1: bucket by owner, all-to-all --------------------------------------------------------
    uint64_t K = K_local;   // records this GPU owns
    Tmp<uint64_t> x_ord, x_fp;
    Tmp<RecInfo> x_info;
    Tmp<uint32_t> x_wloc;
    const uint64_t* r_ord = l_ord; const RecInfo* r_info = l_info; const uint32_t* r_wloc = l_wloc;
    uint64_t* fp = l_fp.p;   // the owned records' table fingerprints (rewritten in place when a collision forces a new seed)
    if (W > 1) {
        Tmp<uint8_t> owner_s;
        Tmp<uint32_t> perm_in, perm;
        Tmp<uint64_t> s_ord, s_fp; Tmp<RecInfo> s_info; Tmp<uint32_t> s_wloc;
        MDBG_CK(c, owner_s.get(c->pool, K_local)); MDBG_CK(c, perm_in.get(c->pool, K_local)); MDBG_CK(c, perm.get(c->pool, K_local));
        if (K_local) {
            iota_kernel<<<nblk(K_local), 256, 0, st>>>(perm_in, (uint32_t)K_local);
            LAUNCHED(c);
            RC(R.cub([&](void* t, size_t& b) {   // stable: every bucket keeps the ordinal order
                return cub::DeviceRadixSort::SortPairs(t, b, l_owner.p, owner_s.p, perm_in.p, perm.p, (uint32_t)K_local, 0,
                                                       std::max(1, log2_ceil((uint64_t)W)), st);
            }));
        }
        kx_bounds_kernel<<<1, 32, 0, st>>>(owner_s, K_local, (uint32_t)W, c->d_mail);
        LAUNCHED(c);
        uint64_t* d_cnt = c->d_mail + 64;   // [W][W]: row s = what rank s sends to each rank
        NCK(c, nccl().AllGather(c->d_mail, d_cnt, W, ncclUint64, (ncclComm_t)c->comm, st));
        MDBG_CK(c, cudaMemcpyAsync(c->h_mail, d_cnt, (size_t)W * W * 8, cudaMemcpyDeviceToHost, st));
        MDBG_CK(c, cudaStreamSynchronize(st));
        uint64_t scnt[MAX_WORLD], rcnt[MAX_WORLD], dst_off_of[MAX_WORLD];
        K = 0;
        for (int p = 0; p < W; p++) {
            scnt[p] = c->h_mail[(size_t)rank * W + p]; rcnt[p] = c->h_mail[(size_t)p * W + rank]; K += rcnt[p];
            dst_off_of[p] = 0;
            for (int r = 0; r < rank; r++) dst_off_of[p] += c->h_mail[(size_t)r * W + p];   // the lower ranks' records come first
        }
        if (K >= 0x7FFFFFF0ull) { c->err = "more than 2^31 k-min-mers owned by one GPU"; return MDBG_ERR_RANGE; }
        MDBG_CK(c, iota.get(c->pool, K));
        // The records reach their owners either through the owners' inboxes, written over NVLink by kx_scatter_kernel
        // (bucketing and exchange in one kernel; needs CUDA IPC between the ranks), or packed and sent with NCCL.
        uint64_t kown_max = 0;
        for (int p = 0; p < W; p++) {
            uint64_t kp = 0;
            for (int r = 0; r < W; r++) kp += c->h_mail[(size_t)r * W + p];
            kown_max = std::max(kown_max, kp);
        }
        const char* p2p_env = getenv("MDBG_P2P");
        if (p2p_env && p2p_env[0] == '0') c->p2p_state = -1;
        // (the count matrix in h_mail is consumed above: ensure_inbox reuses the mailbox)
        RC(ensure_inbox(c, kown_max));
        MDBG_CK(c, cudaEventRecord(c->ev[10], st));
        if (c->p2p_state == 1) {
            ScatterPlan SP{};
            uint64_t acc = 0;
            for (int p = 0; p < W; p++) {
                SP.box[p] = c->peer_inbox[p];
                SP.bstart[p] = acc; acc += scnt[p];
                SP.dst_off[p] = 0;
            }
            SP.bstart[W] = acc;
            for (int p = 0; p < W; p++) SP.dst_off[p] = dst_off_of[p];
            const InboxLayout IL{c->inbox_cap};
            if (K_local) {
                kx_scatter_kernel<<<nblk(K_local), 256, 0, st>>>(perm, owner_s, K_local, l_fp, l_ord, l_wloc, l_info, SP, IL,
                                                                 SP.bstart[(rank + 1) % W]);
                LAUNCHED(c);
            }
            // every rank's scatter must have landed before any owner reads its inbox: a collective after the kernel
            // (stream order on every rank) is that barrier
            NCK(c, nccl().AllReduce(c->d_mail, c->d_mail, 1, ncclUint64, ncclSum, (ncclComm_t)c->comm, st));
            MDBG_CK(c, cudaEventRecord(c->ev[11], st));
            r_ord = IL.ord(c->inbox); r_info = IL.info(c->inbox); r_wloc = IL.wloc(c->inbox);
            fp = IL.fp(c->inbox);
        } else {
            MDBG_CK(c, x_ord.get(c->pool, K)); MDBG_CK(c, x_fp.get(c->pool, K)); MDBG_CK(c, x_info.get(c->pool, K));
            MDBG_CK(c, x_wloc.get(c->pool, K));
            MDBG_CK(c, s_ord.get(c->pool, K_local)); MDBG_CK(c, s_fp.get(c->pool, K_local));
            MDBG_CK(c, s_info.get(c->pool, K_local)); MDBG_CK(c, s_wloc.get(c->pool, K_local));
            if (K_local) {
                kx_pack_kernel<<<nblk(K_local), 256, 0, st>>>(perm, K_local, l_fp, l_ord, l_wloc, l_info, s_fp, s_ord, s_wloc, s_info);
                LAUNCHED(c);
            }
            const void* sv[4] = {s_fp.p, s_ord.p, s_wloc.p, s_info.p};
            void* rv[4] = {x_fp.p, x_ord.p, x_wloc.p, x_info.p};
            const size_t el[4] = {8, 8, 4, sizeof(RecInfo)};
            RC(alltoallv(c, 4, sv, rv, el, scnt, rcnt));
            MDBG_CK(c, cudaEventRecord(c->ev[11], st));
            r_ord = x_ord; r_info = x_info; r_wloc = x_wloc;
            fp = x_fp.p;
        }
        c->tm.exchange_bytes += (K_local - scnt[rank]) * (8 + 8 + 4 + sizeof(RecInfo));
        c->tm.exchange_p2p = c->p2p_state == 1 ? 1 : 0;
        if (K) { iota_kernel<<<nblk(K), 256, 0, st>>>(iota, (uint32_t)K); LAUNCHED(c); }
        l_owner.reset();
        // (the local arrays stay alive until the sends / the scatter have run: freed with the other table scratch)
    }
    const TupleSrc T{arena, r_wloc, r_ord, k};

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
            MDBG_CK(c, cudaEventRecord(c->evk[2], st));
            kc_insert_kernel<<<nblk((K + KC_U - 1) / KC_U * 4), 256, 0, st>>>(fp, K, keys, first, cap - 1, slot);
            LAUNCHED(c);
            MDBG_CK(c, cudaEventRecord(c->evk[3], st));
            if (arena_pending) { MDBG_CK(c, cudaStreamWaitEvent(st, c->ev[20], 0)); arena_pending = false; }   // tuples from here on
            MDBG_CK(c, cudaEventRecord(c->evk[4], st));
            kc_verify_kernel<<<nblk(K), 256, 0, st>>>(T, K, slot, first, &c->d_sc->v[1]);
            LAUNCHED(c);
            MDBG_CK(c, cudaEventRecord(c->evk[5], st));
            // K-D: stable sort by slot, segment heads (speculatively: the collision flag is read
            // together with the segment count, one host round trip for both)
            RC(R.cub([&](void* t, size_t& b) {
                return cub::DeviceRadixSort::SortPairs(t, b, slot.p, sslot.p, iota.p, sj.p, (uint32_t)K, 0, cap_bits, st);
            }));
            MDBG_CK(c, cudaEventRecord(c->evk[6], st));
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
    if (arena_pending) { MDBG_CK(c, cudaStreamWaitEvent(st, c->ev[20], 0)); arena_pending = false; }
    MDBG_CK(c, cudaEventRecord(c->ev[7], st));   // ms_kc = table + sort by slot, ms_kd = reduce + nodes
    slot.reset(); first.reset(); iota.reset(); sslot.reset();
    MDBG_CK(c, first_ord.get(c->pool, D)); MDBG_CK(c, solid.get(c->pool, D)); MDBG_CK(c, nseq.get(c->pool, (uint64_t)D + 1));
    MDBG_CK(c, seq_off.get(c->pool, (uint64_t)D + 1)); MDBG_CK(c, seg_index.get(c->pool, D));
    Tmp<uint8_t> counted;
    MDBG_CK(c, counted.get(c->pool, D));
    uint64_t Dtot = 0, Stot = 0, Qtot = 0;
    {
        // ---- node index and node placement from ONE prefix sum over ordinal space -----------------
        // Every distinct tuple marks the ordinal of its first sighting in a bitmap of two bits per ordinal
        // (it consumed a node index / it is a solid node).  N > 1: the bitmaps are summed over the GPUs (an
        // ordinal belongs to one tuple, hence one owner).  The exclusive scan of the words' popcounts then
        // holds, at that ordinal, the tuple's node index and the place of the node in the ascending-index
        // list: no sort of first sightings, no binary searches, no sort of nodes.
        const uint64_t nwords = (Ktot + 1 + 15) / 16;
        if (nwords >= 0x7FFFFFF0ull) { c->err = "more than 2^35 k-min-mers in the job"; return MDBG_ERR_RANGE; }
        Tmp<uint32_t> ord_bits_map;
        Tmp<uint64_t> wscan;
        MDBG_CK(c, ord_bits_map.get(c->pool, nwords + 1)); MDBG_CK(c, wscan.get(c->pool, nwords + 1));
        MDBG_CK(c, cudaMemsetAsync(ord_bits_map.p, 0, (nwords + 1) * 4, st));
        MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[3], 0, 16, st));
        if (D > 0) {
            kd_segments_kernel<<<nblk(D), 256, 0, st>>>(seg_start, D, K, sj, r_ord, minab, bf, first_ord, counted, solid, nseq,
                                                        ord_bits_map);
            LAUNCHED(c);
        }
        if (W > 1) {
            NCK(c, nccl().AllReduce(ord_bits_map.p, ord_bits_map.p, nwords, ncclUint32, ncclSum, (ncclComm_t)c->comm, st));
            c->tm.exchange_bytes += nwords * 4;
        }
        RC(R.cub([&](void* t, size_t& b) {
            cub::TransformInputIterator<uint64_t, FlagWordToU64, const uint32_t*> in(ord_bits_map.p, FlagWordToU64());
            return cub::DeviceScan::ExclusiveSum(t, b, in, wscan.p, (uint32_t)(nwords + 1), st);
        }));
        MDBG_CK(c, cudaMemcpyAsync(&c->d_sc->v[3], wscan.p + nwords, 8, cudaMemcpyDeviceToDevice, st));
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
        // nodes: every owner writes a NodeRec at the final place of each node it owns; N > 1: the records
        // (zero elsewhere) are summed, then every GPU expands the node arrays (tuples from the arena)
        Tmp<NodeRec> nrec;
        MDBG_CK(c, nrec.get(c->pool, Stot));
        MDBG_CK(c, G->index.get(c->pool, Stot)); MDBG_CK(c, G->abundance.get(c->pool, Stot));
        MDBG_CK(c, G->seqlen.get(c->pool, Stot)); MDBG_CK(c, G->shift.get(c->pool, 2 * Stot));
        MDBG_CK(c, G->tuple.get(c->pool, Stot * k));
        if (W > 1 && Stot > 0) MDBG_CK(c, cudaMemsetAsync(nrec.p, 0, Stot * sizeof(NodeRec), st));
        MDBG_CK(c, cudaEventRecord(c->evk[12], st));
        if (D > 0) {
            kd_nodes_kernel<<<nblk(D), 256, 0, st>>>(D, minab, K, seg_start, sj, first_ord, counted, solid, ord_bits_map, wscan,
                                                     r_wloc, r_ord, r_info, seg_index, nrec);
            LAUNCHED(c);
        }
        if (W > 1 && Stot > 0) {
            NCK(c, nccl().AllReduce(nrec.p, nrec.p, Stot * (sizeof(NodeRec) / 4), ncclUint32, ncclSum, (ncclComm_t)c->comm, st));
            c->tm.exchange_bytes += Stot * sizeof(NodeRec);
        }
        if (Stot > 0) {
            NodeOut NO{G->index, G->abundance, G->seqlen, G->shift, G->tuple};
            kd_expand_kernel<<<nblk(Stot * k), 256, 0, st>>>(nrec, Stot, k, arena, NO);
            LAUNCHED(c);
        }
        MDBG_CK(c, cudaEventRecord(c->evk[13], st));
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
    l_wloc.reset(); l_ord.reset(); l_info.reset(); l_fp.reset(); x_wloc.reset(); x_ord.reset(); x_info.reset(); x_fp.reset();
    g_hash.reset();

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
            MDBG_CK(c, cudaEventRecord(c->evk[8], st));
            ke_join_kernel<0><<<nblk((uint64_t)q_n * KE_LANES, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, inv, presimp, q_lo, q_n, cnt_q, nullptr,
                                                              cap_e, cap_r, &c->d_sc->v[6]);
            LAUNCHED(c);
            MDBG_CK(c, cudaEventRecord(c->evk[9], st));
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
                ke_join_kernel<1><<<nblk((uint64_t)q_n * KE_LANES, 128), 128, 0, st>>>(NV, ekey, erev, skey, sval, inv, presimp, q_lo, q_n, nullptr,
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
    c->dg = Gp.release();
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
    if (c->world > 1) cudaEventElapsedTime(&c->tm.ms_exchange, c->ev[10], c->ev[11]);
    // single kernels (an event pair that was not recorded in this finish leaves 0)
    static const int pairs[][3] = {{0, 0, 1}, {1, 2, 3}, {2, 4, 5}, {3, 5, 6}, {4, 8, 9}, {6, 12, 13}};
    for (auto& p : pairs) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->evk[p[1]], c->evk[p[2]]) != cudaSuccess) { cudaGetLastError(); ms = 0; }
        c->tm.ms_kernels[p[0]] = ms > 0 ? ms : 0;
    }
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

// Host copy.  With N GPUs rank 0 gets the whole graph (edges and .sequences lines are gathered over
// NCCL); the other ranks get the job-wide counters only (array members NULL).
int mdbg_finish(mdbg_ctx* c, int want_seqlines, mdbg_graph* out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    int rc = build_device_graph(c, want_seqlines != 0);
    if (rc) return rc;
    RC(finish_timings(c));
    fill_counters(c, out);
    DeviceGraph* G = c->dg;
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
        {
            NcclGroup g(c);
            gatherv_root(c, g, G->e_n1, E, ec, g_n1, 4); gatherv_root(c, g, G->e_n2, E, ec, g_n2, 4);
            gatherv_root(c, g, G->e_ov, E, ec, g_ov, 4); gatherv_root(c, g, G->e_o1, E, ec, g_o1, 1);
            gatherv_root(c, g, G->e_o2, E, ec, g_o2, 1);
            if (want_seqlines) gatherv_root(c, g, G->seq, Q, qc, g_seq, sizeof(SeqRec));
            RC(g.close());
        }
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
    // N > 1: the host copy lives on rank 0 (all nodes, all edges, all .sequences lines); the other ranks
    // return the job-wide counters with NULL arrays -- nobody downloads a graph nobody reads
    if (W > 1 && c->rank != 0) {
        MDBG_CK(c, cudaEventRecord(c->ev[14], c->st));
        MDBG_CK(c, cudaStreamSynchronize(c->st));
        cudaEventElapsedTime(&c->tm.ms_d2h, c->ev[13], c->ev[14]);
        return MDBG_OK;
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
    HostGraph* H = new HostGraph();
    H->c = c;
    out->_owner = H;   // from here on mdbg_graph_free releases whatever was allocated
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
    if (want_seqlines) {
        out->q_index = H->take<uint32_t>(Q); out->q_read = H->take<uint64_t>(Q); out->q_start = H->take<uint64_t>(Q);
        out->q_end = H->take<uint64_t>(Q); out->q_reversed = H->take<uint8_t>(Q); out->q_shift = H->take<uint64_t>(2 * Q);
        RC(d2h(c, out->q_index, q_index.p, Q)); RC(d2h(c, out->q_read, q_read.p, Q)); RC(d2h(c, out->q_start, q_start.p, Q));
        RC(d2h(c, out->q_end, q_end.p, Q)); RC(d2h(c, out->q_reversed, q_rev.p, Q)); RC(d2h(c, out->q_shift, q_shift.p, 2 * Q));
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
    if (M > 0 && (!hash || !pos)) { c->err = "null hash / pos with minimizers present"; return MDBG_ERR_BAD_ARG; }
    if (M >= 0xFFFFFFF0ull) { c->err = "more than 2^32 minimizers in one mdbg_window call"; return MDBG_ERR_RANGE; }
    for (uint64_t r = 0; r < n_reads; r++)
        if (min_read_off[r + 1] < min_read_off[r]) { c->err = "min_read_off is not non-decreasing"; return MDBG_ERR_BAD_ARG; }
    for (uint64_t i = 0; i < M; i++)
        if (pos[i] > 0xFFFFFFFFull) { c->err = "a minimizer position of 4 G or more (positions inside a read are u32)"; return MDBG_ERR_RANGE; }
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
