// pack_host.cc -- host side of the 4:1 upload of mdbg_push_reads ("pack on host to
// cut PCIe 4x").  ASCII bases -> two bit planes per 32 bases (plane a = bit 1, plane b = bit 2 of the
// byte: A 00, C 01, T 10, G 11 -- the 2-bit code of the kernels), validated on the way: a 4 KiB tile
// holding any byte outside ACGT is flagged and travels as ASCII instead.  The device expands the
// planes back to ASCII in HBM (expand.cu) and the kernels run unchanged, so what crosses PCIe is 2 bits
// per base while every result stays bit-identical.  Plain g++ (SSSE3 when available), worker threads.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSSE3__)
#include <immintrin.h>
#define MDBG_PACK_X86 1
#endif

#include "pack_host.h"

namespace mdbg {

namespace {

inline bool acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// one word of 32 bases at p (all 32 readable) -> planes; returns true if every byte is ACGT
inline bool pack_word(const uint8_t* p, uint32_t& a, uint32_t& b) {
#if defined(__SSSE3__)
    const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p));
    const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(p + 16));
    // movemask takes bit 7 of every byte: shifting the 16-bit lanes left by 6 (5) brings bit 1 (2) there
    a = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v0, 6)) | ((uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v1, 6)) << 16);
    b = (uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v0, 5)) | ((uint32_t)_mm_movemask_epi8(_mm_slli_epi16(v1, 5)) << 16);
    const __m128i lut = _mm_setr_epi8('A', 'C', 'T', 'G', 'A', 'C', 'T', 'G', 'A', 'C', 'T', 'G', 'A', 'C', 'T', 'G');
    const __m128i three = _mm_set1_epi8(3);
    const __m128i c0 = _mm_and_si128(_mm_srli_epi16(v0, 1), three), c1 = _mm_and_si128(_mm_srli_epi16(v1, 1), three);
    const __m128i ok = _mm_and_si128(_mm_cmpeq_epi8(_mm_shuffle_epi8(lut, c0), v0), _mm_cmpeq_epi8(_mm_shuffle_epi8(lut, c1), v1));
    return _mm_movemask_epi8(ok) == 0xFFFF;
#else
    uint32_t aa = 0, bb = 0;
    bool ok = true;
    for (int k = 0; k < 32; k++) {
        const uint8_t c = p[k];
        aa |= (uint32_t)((c >> 1) & 1u) << k;
        bb |= (uint32_t)((c >> 2) & 1u) << k;
        ok = ok && acgt(c);
    }
    a = aa; b = bb;
    return ok;
#endif
}

#if defined(MDBG_PACK_X86)
// the same with one 256-bit load per word; taken when the CPU has AVX2 (runtime dispatch)
__attribute__((target("avx2"))) inline bool pack_word_avx2(const uint8_t* p, uint32_t& a, uint32_t& b) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
    a = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
    b = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
    // table indexed by the low nibble of the byte itself; the unused entries hold a byte with a different low nibble
    // (1 at index 0, 0 elsewhere) and a byte with bit 7 set selects 0: equal to its entry iff A, C, G or T
    const __m256i lut = _mm256_setr_epi8(1, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0,
                                         1, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i ok = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, v), v);
    return (uint32_t)_mm256_movemask_epi8(ok) == 0xFFFFFFFFu;
}
__attribute__((target("avx2"))) void pack_full_words_avx2(const uint8_t* bases, uint64_t w_begin, uint64_t w_end,
                                                          uint32_t* planes, uint8_t* bad_tiles) {
    for (uint64_t w = w_begin; w < w_end; w++) {
        uint32_t a, b;
        const bool ok = pack_word_avx2(bases + w * 32, a, b);
        planes[2 * w] = a;
        planes[2 * w + 1] = b;
        if (!ok && bad_tiles) bad_tiles[w / PACK_TILE_WORDS] = 1;
    }
}
bool cpu_has_avx2() {
    static const bool v = __builtin_cpu_supports("avx2") && !getenv("MDBG_PACK_NO_AVX2");
    return v;
}
// AVX-512BW: two words (64 bases) per load; the mask registers ARE the bit planes (vptestmb against the plane's
// bit), the alphabet check is one vpshufb + one compare-to-mask: the table is indexed by the LOW NIBBLE of the byte
// itself (A 1, C 3, T 4, G 7 -> the letter; every other entry holds a byte with a DIFFERENT low nibble -- 1 at index
// 0, 0 elsewhere -- and a byte with bit 7 set selects 0), so a byte equals its table entry iff it is one of the four
// letters.  ~11 uops per 64 bases.
__attribute__((target("avx512f,avx512bw"))) void pack_full_words_avx512(const uint8_t* bases, uint64_t w_begin,
                                                                        uint64_t w_end, uint32_t* planes,
                                                                        uint8_t* bad_tiles) {
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(1, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i bit1 = _mm512_set1_epi8(2), bit2 = _mm512_set1_epi8(4);
    uint64_t w = w_begin;
    for (; w + 2 <= w_end; w += 2) {
        const __m512i v = _mm512_loadu_si512(reinterpret_cast<const void*>(bases + w * 32));
        const uint64_t a = (uint64_t)_mm512_test_epi8_mask(v, bit1);
        const uint64_t b = (uint64_t)_mm512_test_epi8_mask(v, bit2);
        const uint64_t ok = (uint64_t)_mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, v), v);
        uint64_t* out = reinterpret_cast<uint64_t*>(planes + 2 * w);       // {a lo, b lo}, {a hi, b hi}
        out[0] = (a & 0xFFFFFFFFull) | (b << 32);
        out[1] = (a >> 32) | (b & 0xFFFFFFFF00000000ull);
        if (ok != ~0ull && bad_tiles) {
            if ((uint32_t)ok != 0xFFFFFFFFu) bad_tiles[w / PACK_TILE_WORDS] = 1;
            if ((uint32_t)(ok >> 32) != 0xFFFFFFFFu) bad_tiles[(w + 1) / PACK_TILE_WORDS] = 1;
        }
    }
    if (w < w_end) pack_full_words_avx2(bases, w, w_end, planes, bad_tiles);
}
bool cpu_has_avx512bw() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                          !getenv("MDBG_PACK_NO_AVX512") && !getenv("MDBG_PACK_NO_AVX2");
    return v;
}
#endif

}  // namespace

void pack_words(const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end, uint32_t* planes,
                uint8_t* bad_tiles) {
#if defined(MDBG_PACK_X86)
    if (cpu_has_avx2()) {              // whole words with AVX2, the ragged last word below
        const uint64_t full = std::min<uint64_t>(w_end, n_bases / 32);
        if (w_begin < full) {
            if (cpu_has_avx512bw()) pack_full_words_avx512(bases, w_begin, full, planes, bad_tiles);
            else pack_full_words_avx2(bases, w_begin, full, planes, bad_tiles);
        }
        w_begin = std::max(w_begin, full);
    }
#endif
    for (uint64_t w = w_begin; w < w_end; w++) {
        const uint64_t off = w * 32;
        uint32_t a, b;
        bool ok;
        if (off + 32 <= n_bases) {
            ok = pack_word(bases + off, a, b);
        } else {                       // ragged end of the batch: pad with 'A' (code 0)
            uint8_t buf[32];
            memset(buf, 'A', 32);
            if (off < n_bases) memcpy(buf, bases + off, n_bases - off);
            ok = pack_word(buf, a, b);
        }
        planes[2 * w] = a;
        planes[2 * w + 1] = b;
        if (!ok && bad_tiles) bad_tiles[w / PACK_TILE_WORDS] = 1;   // racing writers store the same value
    }
}

// ---- a small persistent worker pool --------------------------------------------------------------
// Work arrives in bursts (one parallel_for per upload chunk, every ~100 us): workers poll for the next
// burst for a short while before they go to sleep on the condition variable, and the caller polls for
// completion -- a futex wake-up per chunk and worker would cost as much as the packing itself.
namespace {
inline void cpu_relax() {
#if defined(MDBG_PACK_X86)
    _mm_pause();
#endif
}
constexpr int SPIN_LIMIT = 1 << 15;    // ~0.2-0.5 ms of polling
}  // namespace

struct PackPool::Impl {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::function<void(uint64_t)> fn_store;        // the burst's function (begin() keeps a copy)
    const std::function<void(uint64_t)>* fn = nullptr;
    bool pending = false;                          // a burst was begun and not finished yet (caller's thread only)
    uint64_t n_items = 0;
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> generation{0};
    std::atomic<int> active{0};
    std::atomic<bool> stop{false};

    void work() {
        for (;;) {
            const uint64_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= n_items) break;
            (*fn)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_acquire)) {
                if (++spins < SPIN_LIMIT) { cpu_relax(); continue; }
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop.load() || generation.load() != seen; });
            }
            if (stop.load(std::memory_order_acquire)) return;
            seen = generation.load(std::memory_order_acquire);
            work();
            if (active.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                std::lock_guard<std::mutex> lk(mu);
                cv_done.notify_all();
            }
        }
    }
};

PackPool::PackPool(int threads) : impl_(new Impl()) {
    if (threads < 1) threads = 1;
    for (int i = 0; i + 1 < threads; i++) impl_->workers.emplace_back([this] { impl_->loop(); });
}

PackPool::~PackPool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop.store(true, std::memory_order_release);
    }
    impl_->cv_work.notify_all();
    for (auto& t : impl_->workers) t.join();
    delete impl_;
}

int PackPool::threads() const { return (int)impl_->workers.size() + 1; }

void PackPool::begin(uint64_t n_items, std::function<void(uint64_t)> fn) {
    finish();                          // one burst at a time
    if (n_items == 0) return;
    impl_->fn_store = std::move(fn);
    impl_->pending = true;
    if (impl_->workers.empty()) {      // no workers: everything happens in finish()
        impl_->fn = &impl_->fn_store;
        impl_->n_items = n_items;
        impl_->next.store(0, std::memory_order_relaxed);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(impl_->mu);   // a worker about to sleep sees the new generation
        impl_->fn = &impl_->fn_store;
        impl_->n_items = n_items;
        impl_->next.store(0, std::memory_order_relaxed);
        impl_->active.store((int)impl_->workers.size(), std::memory_order_relaxed);
        impl_->generation.fetch_add(1, std::memory_order_release);
    }
    impl_->cv_work.notify_all();
}

void PackPool::finish() {
    if (!impl_->pending) return;
    impl_->work();                     // the caller works too
    if (!impl_->workers.empty()) {
        int spins = 0;
        while (impl_->active.load(std::memory_order_acquire) != 0) {
            if (++spins < SPIN_LIMIT) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> lk(impl_->mu);
            impl_->cv_done.wait(lk, [&] { return impl_->active.load() == 0; });
        }
    }
    impl_->fn = nullptr;
    impl_->fn_store = nullptr;
    impl_->pending = false;
}

void PackPool::parallel_for(uint64_t n_items, const std::function<void(uint64_t)>& fn) {
    begin(n_items, fn);
    finish();
}

static std::function<void(uint64_t)> pack_items(const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end,
                                                uint32_t* planes, uint8_t* bad_tiles, uint64_t& n_items) {
    const uint64_t BLK = 16 * PACK_TILE_WORDS;     // 16 tiles (64 KiB of bases) per work item
    const uint64_t first = w_begin / BLK, last = (w_end + BLK - 1) / BLK;
    n_items = last - first;
    return [=](uint64_t i) {
        const uint64_t lo = std::max(w_begin, (first + i) * BLK), hi = std::min(w_end, (first + i + 1) * BLK);
        if (lo < hi) pack_words(bases, n_bases, lo, hi, planes, bad_tiles);
    };
}

void pack_parallel_begin(PackPool& pool, const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end,
                         uint32_t* planes, uint8_t* bad_tiles) {
    uint64_t n = 0;
    auto fn = pack_items(bases, n_bases, w_begin, w_end, planes, bad_tiles, n);
    pool.begin(n, std::move(fn));
}

void pack_parallel(PackPool& pool, const uint8_t* bases, uint64_t n_bases, uint64_t w_begin, uint64_t w_end,
                   uint32_t* planes, uint8_t* bad_tiles) {
    pack_parallel_begin(pool, bases, n_bases, w_begin, w_end, planes, bad_tiles);
    pool.finish();
}

}  // namespace mdbg
