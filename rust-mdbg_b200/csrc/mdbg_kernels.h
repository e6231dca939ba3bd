// mdbg_kernels.h -- internal launch interfaces between the context (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mdbg_common.cuh"

namespace mdbg {

namespace bs { struct T4Entry; }

// ---- K-A (ka_minimizers.cu) ----------------------------------------------------------------
constexpr int KA_THREADS = 128;                // CTA size: 4 independent warps
constexpr int KA_SEG = 128;                    // bytes walked by one thread
constexpr int KA_TILE = 32 * KA_SEG;           // 4 KiB of bases per WARP iteration (one tile per warp)

struct KAArgs {
    const uint8_t* bases;        // concatenated ASCII reads (16-byte aligned)
    const uint64_t* read_off;    // [n_reads + 1], offsets into bases
    uint64_t n_reads, n_bases;
    uint32_t l;
    uint64_t bound;              // hash_bound(density), computed on the host in f64
    FilterConsts fc;
    int force_dense;
    // outputs (global, (read, position)-ordered)
    uint64_t* out_hash;
    uint32_t* out_pos;           // raw position inside the read
    uint64_t* out_read_off;      // [read_base + r] = index of read r's first minimizer
    uint64_t out_base, out_cap, read_base;
    unsigned long long* total_out;   // out_base + minimizers of this batch
    unsigned long long* err_pos;     // min byte offset of an illegal base that was hashed
    unsigned int* dense_tiles;
    // staging (tile order is restored by ka_finalize_kernel)
    uint64_t* stage_hash;
    uint32_t* stage_pos;
    uint64_t stage_cap;
    unsigned long long* stage_counter;
    uint64_t* tile_cnt;          // [n_tiles] minimizers found in the tile
    uint64_t* tile_soff;         // [n_tiles] offset of the tile's slice in the staging arrays
    // scratch
    uint64_t* tile_lb;           // [n_tiles + 1]
    unsigned int* tile_counter;
    uint64_t n_tiles;
    uint64_t tile_begin, tile_end;   // tiles handled by this launch (chunked, overlapped uploads)
    // list mode of ka_minimizers_kernel: process tile_list[0 .. *tile_list_n) instead of a range
    const uint32_t* tile_list;
    const unsigned int* tile_list_n;
    // bit-sliced variant (ka_bitslice.cu): tiles it cannot take are appended here
    uint32_t* dirty_list;
    unsigned int* dirty_n;
    unsigned long long* dirty_out;   // ka_finalize_kernel copies *dirty_n here (host mailbox)
    uint32_t bs_group;           // consecutive tiles claimed by a warp at a time
    uint32_t bs_ngroups;         // groups of this launch = ceil((tile_end - tile_begin) / bs_group), set by ka_bs_launch
    const bs::T4Entry* bs_t4;    // 256 x 16 B: 4-base ntHash tables (ka_bs_tables), device memory
    uint32_t* dbg;               // optional: 8 words of per-tile state (tests / MDBG_BS_DEBUG_DUMP)
};
// prepare: per-tile read lookup + counters; launch: tiles [A.tile_begin, A.tile_end); finalize: order
// device counters reset by ka_prepare: the batch scalars, one tile counter per K-A launch of the
// batch, and optionally one u64 (the first read offset of an empty arena)
struct KAInit {
    unsigned long long* total_out; unsigned long long* err_pos; unsigned long long* stage_counter;
    unsigned int* dense_tiles; unsigned int* counters; uint32_t n_counters; uint64_t* zero64;
};
cudaError_t ka_prepare(const KAArgs& A, const KAInit& I, cudaStream_t st, uint64_t* launches);
cudaError_t ka_launch(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches);
cudaError_t ka_finalize(const KAArgs& A, const uint64_t* tile_excl, cudaStream_t st, uint64_t* launches);
int ka_max_blocks_per_sm(int hpc);
// list mode: A.tile_list / A.tile_list_n set; the grid is sized by the caller (the list length lives
// on the device)
cudaError_t ka_launch_list(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches);

// ---- 4:1 upload (expand.cu): bit planes -> ASCII bases in HBM -----------------------------------
cudaError_t expand_planes(const uint32_t* planes, uint8_t* bases, uint64_t w_begin, uint64_t w_end, int num_sms,
                          cudaStream_t st, uint64_t* launches);

// ---- K-A, bit-sliced variant (ka_bitslice.cu) -----------------------------------------------
bool ka_bs_supported(uint32_t l, uint64_t bound);
constexpr size_t KA_BS_TABLE_BYTES = 256 * 16;
void ka_bs_tables(void* host_out);           // fills KA_BS_TABLE_BYTES (host); the caller uploads them
cudaError_t ka_bs_launch(const KAArgs& A, int hpc, int grid, cudaStream_t st, uint64_t* launches);
int ka_bs_max_blocks_per_sm(uint32_t l, int hpc);

}  // namespace mdbg
