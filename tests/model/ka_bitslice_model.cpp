// ka_bitslice_model.cpp -- TEST INFRASTRUCTURE: runs the bit-sliced K-A kernel body
// (rust-mdbg_b200/csrc/ka_bitslice_body.h, the very source nvcc compiles for sm_100a) on the CPU.
// One warp = 32 host threads in lock step at every warp primitive (shuffles, votes, __syncwarp are
// pthread barriers; shared/global atomics are GCC atomics), shared memory = a heap struct.  The
// CPU tests compare its per-tile output with the oracle, so index arithmetic, look-ahead hand-over,
// read boundaries and the dirty-tile rules are checked without a GPU.  Not part of the product.
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../rust-mdbg_b200/csrc/expand_math.h"
#include "../../rust-mdbg_b200/csrc/ka_bitslice_body.h"

namespace {
struct WarpEmu {
    pthread_barrier_t bar;
    uint32_t xbuf[32];
    uint32_t vote[32];
};
thread_local WarpEmu* g_warp = nullptr;
thread_local int g_lane = 0;
void sync() { pthread_barrier_wait(&g_warp->bar); }
}  // namespace

namespace mdbg {
namespace bs {
uint32_t bs_shfl(uint32_t v, int src) {
    g_warp->xbuf[g_lane] = v;
    sync();
    uint32_t r = g_warp->xbuf[src & 31];
    sync();
    return r;
}
uint32_t bs_shfl_up(uint32_t v, int d) {
    g_warp->xbuf[g_lane] = v;
    sync();
    uint32_t r = g_lane >= d ? g_warp->xbuf[g_lane - d] : v;
    sync();
    return r;
}
void bs_syncwarp() { sync(); }
bool bs_any(bool p) {
    g_warp->vote[g_lane] = p ? 1u : 0u;
    sync();
    bool r = false;
    for (int i = 0; i < 32; i++) r = r || g_warp->vote[i];
    sync();
    return r;
}
uint32_t bs_ballot(bool p) {
    g_warp->vote[g_lane] = p ? 1u : 0u;
    sync();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= g_warp->vote[i] << i;
    sync();
    return r;
}
uint32_t bs_atomic_or_s(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
uint32_t bs_atomic_add_s(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
uint32_t bs_atomic_add_g32(unsigned int* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
unsigned long long bs_atomic_add_g64(unsigned long long* p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
}  // namespace bs
}  // namespace mdbg

using namespace mdbg;

template <int L, bool HPC>
static void run_warps(const KAArgs& A, int n_warps) {
    std::vector<std::thread> th;
    std::vector<WarpEmu*> warps;
    std::vector<bs::WarpSmem*> smems;
    for (int w = 0; w < n_warps; w++) {
        WarpEmu* we = new WarpEmu();
        pthread_barrier_init(&we->bar, nullptr, 32);
        bs::WarpSmem* sm = (bs::WarpSmem*)aligned_alloc(16, (sizeof(bs::WarpSmem) + 15) / 16 * 16);
        memset(sm, 0xAB, sizeof(bs::WarpSmem));   // shared memory is not zero on a GPU either
        warps.push_back(we);
        smems.push_back(sm);
        for (int lane = 0; lane < 32; lane++)
            th.emplace_back([=, &A]() {
                g_warp = we;
                g_lane = lane;
                bs::warp_loop<L, 8, HPC>(A, *sm, lane);
            });
    }
    for (auto& t : th) t.join();
    for (auto* we : warps) { pthread_barrier_destroy(&we->bar); delete we; }
    for (auto* sm : smems) free(sm);
}

extern "C" {

// Mirrors run_ka() of api.cu for the bit-sliced launch: tile_lb as ka_tile_lb_kernel computes it,
// then the warp loop.  Outputs are the raw per-tile protocol (before ka_finalize_kernel).
// Returns 0, or -1 when (l, bound) is not supported by the variant.
int bs_model_run(const uint8_t* bases, const uint64_t* read_off, uint64_t R, uint64_t B, uint32_t l, uint64_t bound,
                 int hpc, uint32_t group, int n_warps, uint64_t tile_begin, uint64_t tile_end_or_0,
                 uint64_t* tile_cnt, uint64_t* tile_soff, uint64_t* stage_hash, uint32_t* stage_pos,
                 uint64_t stage_cap, uint64_t* out_read_off, uint32_t* dirty_list, uint32_t* dirty_n,
                 uint64_t* stage_total, uint32_t* dbg) {
    if (!bs::supported(l, bound)) return -1;
    const uint64_t n_tiles = std::max<uint64_t>(1, (B + KA_TILE - 1) / KA_TILE);
    std::vector<uint64_t> tile_lb(n_tiles + 1);
    for (uint64_t i = 0; i < n_tiles; i++)
        tile_lb[i] = std::lower_bound(read_off, read_off + R + 1, i * (uint64_t)KA_TILE) - read_off;
    tile_lb[n_tiles] = R + 1;
    // 16-byte aligned copy of the bases, exactly B bytes readable (the kernel must not read past B)
    uint8_t* gb = (uint8_t*)aligned_alloc(16, (B + 15) / 16 * 16 + 16);
    memcpy(gb, bases, B);
    memset(gb + B, 0xEE, (B + 15) / 16 * 16 + 16 - B);   // poison: an illegal byte if ever looked at
    unsigned int tile_counter = 0, dn = *dirty_n;              // in/out: several launches share the dirty list
    unsigned long long stage_counter = *stage_total;          // ... and the staging arrays (like run_ka's chunks)
    KAArgs A{};
    A.bases = gb; A.read_off = read_off; A.n_reads = R; A.n_bases = B;
    A.l = l; A.bound = bound;
    A.out_read_off = out_read_off; A.read_base = 0;
    A.stage_hash = stage_hash; A.stage_pos = stage_pos; A.stage_cap = stage_cap;
    A.stage_counter = &stage_counter; A.tile_cnt = tile_cnt; A.tile_soff = tile_soff;
    A.tile_lb = tile_lb.data(); A.tile_counter = &tile_counter; A.n_tiles = n_tiles;
    A.tile_begin = tile_begin; A.tile_end = tile_end_or_0 ? tile_end_or_0 : n_tiles;
    A.dirty_list = dirty_list; A.dirty_n = &dn; A.bs_group = group; A.dbg = dbg;
    { const uint64_t S = group ? group : 1; A.bs_ngroups = (uint32_t)((A.tile_end - A.tile_begin + S - 1) / S); }   // as ka_bs_launch does
    static bs::T4Entry t4[256];
    for (uint32_t i = 0; i < 256; i++) t4[i] = bs::t4_make(i);
    A.bs_t4 = t4;
    switch (l) {
        case 10: hpc ? run_warps<10, true>(A, n_warps) : run_warps<10, false>(A, n_warps); break;
        case 12: hpc ? run_warps<12, true>(A, n_warps) : run_warps<12, false>(A, n_warps); break;
        case 14: hpc ? run_warps<14, true>(A, n_warps) : run_warps<14, false>(A, n_warps); break;
        default: free(gb); return -1;
    }
    *dirty_n = dn;
    *stage_total = stage_counter;
    free(gb);
    return 0;
}

// unit hooks for the arithmetic
uint32_t bs_model_filter(uint32_t l, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    switch (l) {
        case 10: return bs::filter_window<10, 8>(a0, a1, b0, b1);
        case 12: return bs::filter_window<12, 8>(a0, a1, b0, b1);
        case 14: return bs::filter_window<14, 8>(a0, a1, b0, b1);
    }
    return 0;
}
uint64_t bs_model_exact(uint32_t l, uint32_t av, uint32_t bv) {
    static bs::T4Entry t4[256];
    static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) t4[i] = bs::t4_make(i); init = true; }
    switch (l) {
        case 10: return bs::exact_hash<10>(av, bv, t4);
        case 12: return bs::exact_hash<12>(av, bv, t4);
        case 14: return bs::exact_hash<14>(av, bv, t4);
    }
    return 0;
}
void bs_model_planes(const uint8_t* p32, uint32_t* a, uint32_t* b, uint32_t* bad) {
    alignas(16) uint8_t buf[32];
    memcpy(buf, p32, 32);
    bs::BadAcc acc{0, 0, 0};
    bs::gather32(buf, *a, *b, acc);
    *bad = bs::bad_of(acc);
}
void bs_model_pext(uint32_t m, uint32_t* x, uint32_t* y) { bs::pext_pair(m, *x, *y); }
uint32_t bs_model_select(uint32_t m, uint32_t k) { return bs::select_bit(m, k); }
uint32_t bs_model_expand4(uint32_t a, uint32_t b) { return mdbg::expand4(a, b); }

}  // extern "C"
