// pack_host.h -- host-side 2-bit packing of a read batch for the 4:1 upload (pack_host.cc) and the
// device-side expansion back to ASCII (expand.cu).
#pragma once
#include <stdint.h>

#include <functional>

namespace mdbg {

constexpr uint64_t PACK_TILE_WORDS = 128;      // 32-base words per 4 KiB tile (= KA_TILE / 32)

// planes[2w] / planes[2w+1] = code bit 0 / bit 1 of bases [32w, 32w+32); bases past n_bases read as 'A'.
// bad_tiles[t] is set to 1 when tile t (4096 bases) holds a byte outside ACGT (never cleared here).
void pack_words(const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end, uint32_t* planes,
                uint8_t* bad_tiles);

class PackPool {
public:
    explicit PackPool(int threads);
    ~PackPool();
    PackPool(const PackPool&) = delete;
    PackPool& operator=(const PackPool&) = delete;
    int threads() const;
    void parallel_for(uint64_t n_items, const std::function<void(uint64_t)>& fn);   // fn(i) for i in [0, n_items)
    // The same in two halves: begin() hands the items to the workers and returns, finish() lets the caller work
    // on what is left and waits for the rest -- the caller is free in between (mdbg_push_reads enqueues the copies
    // of the previous chunk there).  One burst at a time; finish() without a burst in flight does nothing.
    void begin(uint64_t n_items, std::function<void(uint64_t)> fn);
    void finish();
private:
    struct Impl;
    Impl* impl_;
};

void pack_parallel(PackPool& pool, const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end,
                   uint32_t* planes, uint8_t* bad_tiles);
// begin()/finish() form of pack_parallel: the packing of words [w_begin, w_end) starts on the workers at once
void pack_parallel_begin(PackPool& pool, const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end,
                         uint32_t* planes, uint8_t* bad_tiles);

}  // namespace mdbg
