// comm.cu -- NCCL plumbing for the multi-GPU path (one process per GPU).  The unique id is
// created on rank 0 by mdbg_nccl_unique_id and shipped by the host (torch.distributed
// broadcast in bench.py / tests); the exchange itself is in graph.cu (build_device_graph).
#include <cstdlib>
#include <cstring>
#include <string>

#include "ctx.h"
#include "nccl_dl.h"

using namespace mdbg;

// called by mdbg_ctx_destroy and before a re-init
// unmap the peers' inboxes (an exporter must not free memory a peer still has mapped: growing the inboxes is
// unmap everywhere -> barrier -> free, see ensure_inbox in graph.cu)
void mdbg_inbox_unmap_peers(mdbg_ctx* c) {
    for (int p = 0; p < MAX_WORLD; p++) {
        if (c->peer_inbox[p] && c->peer_inbox[p] != c->inbox) cudaIpcCloseMemHandle(c->peer_inbox[p]);
        c->peer_inbox[p] = nullptr;
    }
    (void)cudaGetLastError();
}
// ... and free the own one
void mdbg_inbox_release(mdbg_ctx* c) {
    mdbg_inbox_unmap_peers(c);
    if (c->inbox) cudaFree(c->inbox);
    c->inbox = nullptr;
    c->inbox_cap = 0;
    (void)cudaGetLastError();
}

extern "C" void mdbg_comm_release(mdbg_ctx* c) {
    if (c) { cudaSetDevice(c->device); mdbg_inbox_release(c); c->p2p_state = 0; }
    if (c && c->comm) {
        NcclApi& N = nccl();
        if (N.ok && c->comm2) N.CommDestroy((ncclComm_t)c->comm2);
        c->comm2 = nullptr;
        if (N.ok) N.CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
        c->rank = 0;
        c->world = 1;
    }
}

extern "C" {

int mdbg_nccl_unique_id(uint8_t id[MDBG_NCCL_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= MDBG_NCCL_ID_BYTES, "ncclUniqueId larger than MDBG_NCCL_ID_BYTES");
    NcclApi& N = nccl();
    if (!N.ok) return MDBG_ERR_NCCL;
    ncclUniqueId u;
    if (N.GetUniqueId(&u) != ncclSuccess) return MDBG_ERR_NCCL;
    memset(id, 0, MDBG_NCCL_ID_BYTES);
    memcpy(id, &u, sizeof(u));
    return MDBG_OK;
}

int mdbg_comm_init(mdbg_ctx* c, const uint8_t id[MDBG_NCCL_ID_BYTES], int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return MDBG_ERR_BAD_ARG;
    if (world > MAX_WORLD) { c->err = "more GPUs than MAX_WORLD (16) in one job"; return MDBG_ERR_BAD_ARG; }
    NcclApi& N = nccl();
    if (!N.ok) { c->err = "libnccl.so.2 could not be loaded"; return MDBG_ERR_NCCL; }
    MDBG_CK(c, cudaSetDevice(c->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm;
    ncclResult_t r = N.CommInitRank(&comm, world, u, rank);
    if (r != ncclSuccess) {
        c->err = std::string("ncclCommInitRank: ") + N.GetErrorString(r);
        return MDBG_ERR_NCCL;
    }
    mdbg_comm_release(c);   // a second init replaces the first communicator
    c->comm = (void*)comm;
    // second communicator over the same ranks: lets the arena all-gather overlap the record exchange (operations
    // of ONE communicator run in issue order).  MDBG_COMM2=0 turns it off: everything then runs on `comm`, in order.
    c->comm2 = nullptr;
    const char* want2 = getenv("MDBG_COMM2");
    if (N.CommSplit && world > 1 && !(want2 && want2[0] == '0')) {
        ncclComm_t c2 = nullptr;
        if (N.CommSplit(comm, 0, rank, &c2, nullptr) == ncclSuccess) c->comm2 = (void*)c2;
    }
    c->rank = rank;
    c->world = world;
    return MDBG_OK;
}

// Global index of the first read this rank pushes (reads are sharded by record in contiguous
// ranges).  Without it the ranks number their reads in rank order of what they pushed.
int mdbg_comm_set_read_base(mdbg_ctx* c, uint64_t first_read) {
    if (!c) return MDBG_ERR_BAD_ARG;
    c->read_base = first_read;
    c->read_base_set = true;
    return MDBG_OK;
}

}  // extern "C"
