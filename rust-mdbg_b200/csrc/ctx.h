// ctx.h -- the per-GPU context behind the C ABI (include/mdbg.h): stream, caching device
// allocator, the resident minimizer arena, scalar mailboxes, timing events.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mdbg.h"
#include "mdbg_kernels.h"
#include "pack_host.h"

namespace mdbg {

// Caching allocator: after warm-up a push/finish cycle performs no cudaMalloc/cudaFree.
class Pool {
public:
    ~Pool() { trim(); }
    cudaError_t alloc(size_t bytes, void** out, size_t* cap_out);
    void release(void* p, size_t cap);
    void trim();
private:
    std::multimap<size_t, void*> free_;
};

// RAII scratch buffer from the pool.
template <class T>
struct Tmp {
    Pool* pool = nullptr;
    T* p = nullptr;
    size_t cap = 0, n = 0;
    Tmp() {}
    Tmp(const Tmp&) = delete;
    Tmp& operator=(const Tmp&) = delete;
    Tmp(Tmp&& o) noexcept { *this = std::move(o); }
    Tmp& operator=(Tmp&& o) noexcept {
        if (this != &o) { reset(); pool = o.pool; p = o.p; cap = o.cap; n = o.n; o.p = nullptr; o.cap = 0; o.n = 0; }
        return *this;
    }
    ~Tmp() { reset(); }
    cudaError_t get(Pool& pl, size_t count) {
        reset();
        pool = &pl;
        n = count;
        void* q = nullptr;
        cudaError_t e = pl.alloc((count ? count : 1) * sizeof(T), &q, &cap);
        p = (T*)q;
        return e;
    }
    void reset() {
        if (p && pool) pool->release(p, cap);
        p = nullptr; cap = 0; n = 0;
    }
    operator T*() const { return p; }
};

constexpr int MAX_WORLD = 16;                 // GPUs of one job (count matrices are MAX_WORLD^2 words)
constexpr int MAIL_WORDS = 2 * MAX_WORLD * MAX_WORLD + 64;

struct Scalars {  // device mailbox mirrored into pinned host memory
    unsigned long long total_out;
    unsigned long long err_pos;
    unsigned int dense_tiles;
    unsigned int tile_counter;
    unsigned long long stage_counter;
    unsigned long long v[12];   // stage-specific counters (see api.cu)
};

}  // namespace mdbg

struct mdbg_ctx {
    mdbg_params p{};
    uint64_t bound = 0;
    mdbg::FilterConsts fc{};
    int device = 0, num_sms = 0, ka_grid = 0;
    bool ka_bs = false;                             // bit-sliced K-A variant selected and applicable
    int ka_bs_grid = 0;
    bool upload_hybrid = false;                     // ... and chunks go as ASCII whenever the copy engine runs dry
    bool upload_packed = false;                     // mdbg_push_reads: 2-bit planes over PCIe, expanded on the device
    mdbg::PackPool* pack_pool = nullptr;            // host worker threads of the packer
    void* h_planes = nullptr; size_t h_planes_cap = 0;   // pinned staging of the bit planes
    void* ka_bs_t4 = nullptr;                       // device copy of the 4-base ntHash tables
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;                 // uploads overlapped with K-A
    std::vector<cudaEvent_t> copy_ev;
    std::vector<std::pair<size_t, void*>> pinned_cache;   // host blocks of freed graphs
    std::string err;
    mdbg::Pool pool;
    // resident minimizer arena (global read order)
    uint64_t* m_hash = nullptr; size_t m_hash_cap = 0;   // bytes
    uint32_t* m_pos = nullptr;  size_t m_pos_cap = 0;
    uint64_t* m_off = nullptr;  size_t m_off_cap = 0;
    uint64_t M = 0, R = 0, n_bases = 0;
    uint64_t m_cap_items = 0, r_cap_items = 0;
    mdbg::Scalars* d_sc = nullptr;
    mdbg::Scalars* h_sc = nullptr;   // pinned
    uint64_t* d_mail = nullptr;      // MAIL_WORDS u64: sizes / counts exchanged between the GPUs ...
    uint64_t* h_mail = nullptr;      // ... and their pinned host mirror
    void* l2_flush = nullptr; size_t l2_flush_bytes = 0;
    cudaEvent_t ev[24]{};
    cudaEvent_t evk[16]{};   // begin/end pairs around single kernels (mdbg_timings.ms_kernels)
    mdbg_timings tm{};
    // NCCL (multi-GPU)
    void* comm = nullptr; int rank = 0, world = 1;
    void* comm2 = nullptr;   // a copy of the communicator for the arena all-gather, which runs on st_copy under K-B / K-C
    // record inbox (N > 1): the peers' kx_scatter_kernel writes their records for this owner straight into this
    // buffer over NVLink (CUDA IPC mappings of each other's inboxes); capacity in records, equal on every rank
    void* inbox = nullptr; uint64_t inbox_cap = 0;
    void* peer_inbox[mdbg::MAX_WORLD] = {};          // [p] = rank p's inbox as mapped into this process ([rank] = inbox)
    int p2p_state = 0;                               // 0 untried, 1 mapped, -1 unavailable (NCCL send/recv instead)
    uint64_t read_base = 0; bool read_base_set = false;   // global index of this rank's first read
    // device-resident result of the last finish (kept until the next finish/reset)
    struct DeviceGraph* dg = nullptr;
};

#define MDBG_CK(ctx, call)                                                                   \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                 \
            return MDBG_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)
