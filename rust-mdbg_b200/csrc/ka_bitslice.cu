// ka_bitslice.cu -- K-A, bit-sliced variant: the kernel wrapper and its launcher.
//
// The algorithm is in ka_bitslice_body.h (one warp per tile, 32 HPC positions per instruction) and
// ka_bitslice_math.h; this file only provides the shared memory, the 4-base hash tables (host side) and
// the dispatch over the instantiated (l, hpc) pairs.  Roofline: HBM; algorithmic bytes as for
// ka_minimizers_kernel (1 B read per base + 12 B written per minimizer).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ka_bitslice_body.h"

namespace mdbg {

namespace {

constexpr int BS_WARPS = KA_THREADS / 32;
constexpr int BS_T = 8;                 // top bits tested by the filter: bound < 2^56

// MDBG_BS_SMEM_PAD=<bytes>: extra dynamic shared memory per CTA (unused) -- lowers the number of resident CTAs per
// SM for occupancy-sensitivity measurements; 0 in production.
size_t smem_pad() {
    static const size_t v = [] {
        const char* e = getenv("MDBG_BS_SMEM_PAD");
        long x = e ? atol(e) : 0;
        return (size_t)((x > 0 && x <= 160 * 1024) ? x : 0);
    }();
    return v;
}

struct __align__(16) BsCta {
    bs::WarpSmem w[BS_WARPS];
};

#ifndef MDBG_BS_MINBLOCKS
#define MDBG_BS_MINBLOCKS 6            // resident CTAs per SM (shared memory allows 6): up to 85 registers per thread
#endif
template <int L, bool HPC>
__global__ void __launch_bounds__(KA_THREADS, MDBG_BS_MINBLOCKS) ka_bitslice_kernel(const KAArgs A) {
    extern __shared__ __align__(16) unsigned char bs_smem_raw[];
    BsCta& cs = *reinterpret_cast<BsCta*>(bs_smem_raw);
    // warps are independent: no CTA-wide barrier anywhere
    bs::warp_loop<L, BS_T, HPC>(A, cs.w[threadIdx.x >> 5], (int)(threadIdx.x & 31));
}

template <int L, bool HPC>
cudaError_t launch_one(const KAArgs& A, unsigned g, cudaStream_t st) {
    static bool attr_set = false;       // per instantiation; contexts of one process share the device code
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(ka_bitslice_kernel<L, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(sizeof(BsCta) + smem_pad()));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    ka_bitslice_kernel<L, HPC><<<g, KA_THREADS, sizeof(BsCta) + smem_pad(), st>>>(A);
    return cudaGetLastError();
}

template <int L, bool HPC>
int occupancy_one() {
    int n = 0;
    cudaFuncSetAttribute(ka_bitslice_kernel<L, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(BsCta) + smem_pad()));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ka_bitslice_kernel<L, HPC>, KA_THREADS, sizeof(BsCta) + smem_pad());
    return n;
}

}  // namespace

bool ka_bs_supported(uint32_t l, uint64_t bound) { return bs::supported(l, bound); }

void ka_bs_tables(void* host_out) {
    static_assert(sizeof(bs::T4Entry) == 16, "table entry layout");
    bs::T4Entry* t = reinterpret_cast<bs::T4Entry*>(host_out);
    for (uint32_t i = 0; i < 256; i++) t[i] = bs::t4_make(i);
}

// Tiles [A.tile_begin, A.tile_end) in groups of A.bs_group; dirty tiles are appended to A.dirty_list.
cudaError_t ka_bs_launch(const KAArgs& A0, int hpc, int grid, cudaStream_t st, uint64_t* launches) {
    if (A0.tile_end <= A0.tile_begin) return cudaSuccess;
    KAArgs A = A0;
    const uint64_t S = A.bs_group ? A.bs_group : 1;
    const uint64_t groups = (A.tile_end - A.tile_begin + S - 1) / S;
    if (groups > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    A.bs_ngroups = (uint32_t)groups;
    const uint64_t need = (groups + BS_WARPS - 1) / BS_WARPS;
    const unsigned g = (unsigned)(need < (uint64_t)grid ? need : (uint64_t)grid);
    cudaError_t e = cudaErrorInvalidValue;
    switch (A.l) {
        case 10: e = hpc ? launch_one<10, true>(A, g, st) : launch_one<10, false>(A, g, st); break;
        case 12: e = hpc ? launch_one<12, true>(A, g, st) : launch_one<12, false>(A, g, st); break;
        case 14: e = hpc ? launch_one<14, true>(A, g, st) : launch_one<14, false>(A, g, st); break;
        default: break;
    }
    if (e == cudaSuccess && launches) *launches += 1;
    return e;
}

int ka_bs_max_blocks_per_sm(uint32_t l, int hpc) {
    switch (l) {
        case 10: return hpc ? occupancy_one<10, true>() : occupancy_one<10, false>();
        case 12: return hpc ? occupancy_one<12, true>() : occupancy_one<12, false>();
        case 14: return hpc ? occupancy_one<14, true>() : occupancy_one<14, false>();
        default: return 0;
    }
}

}  // namespace mdbg
