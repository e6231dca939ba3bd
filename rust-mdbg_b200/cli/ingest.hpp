// ingest.hpp -- parallel FASTA/FASTQ ingest for the front end (reference side: the
// seq_io reader + `threads` workers of src/main.rs:163-178,830-839).
//
//   plain files : mmap; a batch is a window of the file cut at record starts, split into one slice per
//                 worker thread (memchr for the line ends, one memcpy per sequence line straight into the
//                 pinned batch buffer) -- count pass, prefix sum, copy pass, no intermediate strings;
//   .gz files   : one thread inflates (zlib) into 64 MB blocks cut at record starts while the workers
//                 parse the previous block and the GPU runs the batch before that (a gzip stream has no
//                 random access: inflate is the floor, everything else hides behind it).
//
// A Batch holds the bases of its reads back to back (what mdbg_push_reads takes) + offsets (+ ids when
// the caller wants them).  Records are delivered in file order.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include "gz_inflate.hpp"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ingest {

struct Batch {
    uint8_t* bases = nullptr;   // caller-provided (pinned) buffer of `cap` bytes
    size_t cap = 0, fill = 0;
    std::vector<uint64_t> off;  // [n_reads + 1]
    std::vector<std::string> ids;
    uint64_t n_reads() const { return off.empty() ? 0 : off.size() - 1; }
};

// One slice of text that starts at a record start and ends at a record start / end of data.
struct SliceStat { uint64_t reads = 0, bases = 0; };

inline const char* line_end(const char* p, const char* end) {
    const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
    return e ? e : end;
}

// FASTA: '>' at the start of a line.  FASTQ: a line starting with '@' whose next-but-one line starts with '+'
// (a quality line may start with '@', a sequence line never starts with '+').
inline const char* next_record_start(const char* p, const char* begin, const char* end, bool fasta) {
    if (p <= begin) return begin;
    // move to the start of the next line
    const char* q = (const char*)memchr(p - 1, '\n', (size_t)(end - (p - 1)));
    if (!q) return end;
    q++;
    while (q < end) {
        if (fasta) {
            if (*q == '>') return q;
        } else if (*q == '@') {
            const char* l1 = line_end(q, end);
            const char* l2 = l1 < end ? line_end(l1 + 1, end) : end;
            if (l2 < end && l2 + 1 < end && l2[1] == '+') return q;   // (a candidate cut off by `end` is not accepted)
        }
        const char* e = line_end(q, end);
        if (e >= end) return end;
        q = e + 1;
    }
    return end;
}

// Parses [p, end) (starts at a record start).  dst == nullptr: count only.
inline SliceStat parse_slice(const char* p, const char* end, bool fasta, uint8_t* dst, uint64_t* off_out,
                             uint64_t base_off, std::string* ids) {
    SliceStat st;
    uint64_t w = 0;
    while (p < end) {
        // header line
        const char* he = line_end(p, end);
        if (fasta ? *p != '>' : *p != '@') {   // stray line before the first record: skip it
            p = he < end ? he + 1 : end;
            continue;
        }
        if (ids) {
            const char* s = p + 1;
            const char* sp = (const char*)memchr(s, ' ', (size_t)(he - s));
            size_t n = (size_t)((sp ? sp : he) - s);
            while (n && (s[n - 1] == '\r')) n--;
            ids[st.reads].assign(s, n);
        }
        p = he < end ? he + 1 : end;
        if (fasta) {
            while (p < end && *p != '>') {
                const char* e = line_end(p, end);
                size_t n = (size_t)(e - p);
                if (n && p[n - 1] == '\r') n--;
                if (dst) memcpy(dst + w, p, n);
                w += n;
                p = e < end ? e + 1 : end;
            }
        } else {
            const char* e = line_end(p, end);
            size_t n = (size_t)(e - p);
            if (n && p[n - 1] == '\r') n--;
            if (dst) memcpy(dst + w, p, n);
            w += n;
            p = e < end ? e + 1 : end;
            for (int skip = 0; skip < 2 && p < end; skip++) {   // '+' line, quality line
                const char* e2 = line_end(p, end);
                p = e2 < end ? e2 + 1 : end;
            }
        }
        st.reads++;
        if (off_out) off_out[st.reads] = base_off + w;
    }
    st.bases = w;
    return st;
}

class Reader {
public:
    Reader(int threads) : nthreads_(std::max(1, threads)) {}
    ~Reader() { close(); }

    bool open(const char* path, bool fasta, bool want_ids) {
        close();
        fasta_ = fasta; want_ids_ = want_ids;
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) return false;
        struct stat sb;
        if (fstat(fd_, &sb) != 0) return false;
        size_ = (size_t)sb.st_size;
        unsigned char magic[2] = {0, 0};
        if (size_ >= 2 && pread(fd_, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            // the compressed file is mapped and decoded by gz_inflate.hpp (MDBG_GZ_ZLIB=1: zlib's gzread instead)
            gz_on_ = true; gz_err_.clear();
            if (!getenv("MDBG_GZ_ZLIB")) {
                gzmap_ = (const uint8_t*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
                if (gzmap_ == MAP_FAILED) { gzmap_ = nullptr; return false; }
                madvise((void*)gzmap_, size_, MADV_SEQUENTIAL);
                zfast_.reset(gzmap_, size_);
                // block gzip (bgzip / BGZF: members of <= 64 KiB that carry their own compressed size): members are
                // independent, so a block's worth of them is inflated by all threads at once
                gzpos_ = 0;
                bgzf_ = bgzf_member(0, nullptr, nullptr) && !getenv("MDBG_GZ_SERIAL");
                // one plain gzip stream of some size: spans of it are inflated by all threads (GzParallel); below
                // three threads or a few MB the serial decoder is as fast
                size_t par_min = 16u << 20, span = 2u << 20;
                if (const char* e = getenv("MDBG_GZ_PAR_MIN")) par_min = (size_t)atoll(e);
                if (const char* e = getenv("MDBG_GZ_SPAN")) span = (size_t)std::max<long long>(1 << 16, atoll(e));
                gzpar_ = !bgzf_ && nthreads_ >= 3 && size_ >= par_min && !getenv("MDBG_GZ_SERIAL");
                if (gzpar_) zpar_.reset(gzmap_, size_, nthreads_, span);
            } else {
                gz_ = gzdopen(dup(fd_), "rb");
                if (!gz_) return false;
                gzbuffer(gz_, 1 << 20);
            }
            blk_[0].resize(GZ_BLOCK + GZ_SLACK); blk_[1].resize(GZ_BLOCK + GZ_SLACK);
            gz_eof_ = false; carry_.clear();
            inflater_ = std::thread([this] { inflate_loop(); });
            return true;
        }
        if (size_ == 0) { map_ = nullptr; pos_ = 0; return true; }
        map_ = (const char*)mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (map_ == MAP_FAILED) { map_ = nullptr; return false; }
        madvise((void*)map_, size_, MADV_SEQUENTIAL);
        pos_ = 0;
        return true;
    }

    void close() {
        if (inflater_.joinable()) {
            { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
            cv_.notify_all();
            inflater_.join();
        }
        stop_ = false; ready_[0] = ready_[1] = false; blk_len_[0] = blk_len_[1] = 0; take_ = 0; put_ = 0;
        gz_eof_ = false; gz_eof_done_ = false; carry_.clear();
        if (gz_) { gzclose(gz_); gz_ = nullptr; }
        if (gzmap_) { munmap((void*)gzmap_, size_); gzmap_ = nullptr; }
        gz_on_ = false; gzpar_ = false; bgzf_ = false;
        if (map_) { munmap((void*)map_, size_); map_ = nullptr; }
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
    }

    // Fills `b` (b.bases / b.cap set by the caller) with the next reads of the file, about `target` bytes of
    // text; false when the file is exhausted and nothing was added.  A record larger than the buffer fails.
    bool next_batch(Batch& b, size_t target, std::string& err) {
        b.fill = 0; b.off.assign(1, 0); b.ids.clear();
        if (gz_on_) {
            for (;;) {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return ready_[take_] || gz_eof_done_; });
                if (!ready_[take_]) {
                    if (!gz_err_.empty()) err = gz_err_;     // a damaged stream is an error, not a short file
                    return false;
                }
                const char* text = blk_[take_].data();
                const size_t len = blk_len_[take_];
                lk.unlock();
                const bool ok = parse_text(text, len, b, err);
                lk.lock();
                ready_[take_] = false; take_ ^= 1;
                lk.unlock();
                cv_.notify_all();
                if (!ok) return false;
                if (b.n_reads() > 0) return true;   // (a block of stray lines only: take the next one)
            }
        }
        if (pos_ >= size_) return false;
        const char* begin = map_ + pos_;
        const char* fend = map_ + size_;
        const char* end = (size_t)(fend - begin) <= target ? fend : next_record_start(begin + target, begin, fend, fasta_);
        // text never yields more bases than bytes: make sure the window fits the buffer
        while ((size_t)(end - begin) > b.cap && end > begin) {
            const char* e2 = next_record_start(begin + (end - begin) / 2, begin, fend, fasta_);
            if (e2 >= end) { err = "a record longer than the staging buffer"; return false; }
            end = e2;
        }
        pos_ = (size_t)(end - map_);
        return parse_text(begin, (size_t)(end - begin), b, err);
    }

private:
    static constexpr size_t GZ_BLOCK = 64u << 20, GZ_SLACK = 64u << 20;
    static constexpr long GZ_FULL = -2;    // gz_read: nothing read because the buffer has no room for the next unit

    bool parse_text(const char* text, size_t len, Batch& b, std::string& err) {
        if (len > b.cap) { err = "a block of reads longer than the staging buffer"; return false; }
        const char* end = text + len;
        const int T = (int)std::min<size_t>((size_t)nthreads_, std::max<size_t>(1, len >> 20));
        std::vector<const char*> cut(T + 1);
        cut[0] = text; cut[T] = end;
        for (int t = 1; t < T; t++) cut[t] = std::max(cut[t - 1], next_record_start(text + len / T * t, text, end, fasta_));
        std::vector<SliceStat> st(T);
        run_parallel(T, [&](int t) { st[t] = parse_slice(cut[t], cut[t + 1], fasta_, nullptr, nullptr, 0, nullptr); });
        std::vector<uint64_t> rbase(T + 1, 0), bbase(T + 1, 0);
        for (int t = 0; t < T; t++) { rbase[t + 1] = rbase[t] + st[t].reads; bbase[t + 1] = bbase[t] + st[t].bases; }
        b.off.resize(rbase[T] + 1);
        b.off[0] = 0;
        if (want_ids_) b.ids.resize(rbase[T]);
        run_parallel(T, [&](int t) {
            parse_slice(cut[t], cut[t + 1], fasta_, b.bases + bbase[t], b.off.data() + rbase[t], bbase[t],
                        want_ids_ ? b.ids.data() + rbase[t] : nullptr);
        });
        b.fill = bbase[T];
        return true;
    }

    template <class F>
    void run_parallel(int T, F&& f) {
        if (T <= 1) { f(0); return; }
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back([&f, t] { f(t); });
        f(0);
        for (auto& x : th) x.join();
    }

    // inflate thread: fills blk_[put_] with whole records (the tail of a block that ends inside a record is
    // carried to the next one)
    // next decompressed bytes; <= 0 at the end of the stream or on a damaged one (gz_err_ says which)
    // BGZF member at byte `at` of the mapped file: its total size and uncompressed size (SAM spec 4.1: gzip member whose
    // extra field holds the subfield 'B' 'C' 2 0 BSIZE-1, ISIZE in its last four bytes)
    bool bgzf_member(size_t at, size_t* msize, size_t* isize) const {
        if (size_ < at + 28) return false;
        const uint8_t* h = gzmap_ + at;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
        const size_t xlen = h[10] | (h[11] << 8);
        if (size_ < at + 12 + xlen) return false;
        for (size_t x = 0; x + 4 <= xlen;) {
            const uint8_t* f = h + 12 + x;
            const size_t slen = f[2] | (f[3] << 8);
            if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) {
                const size_t total = (size_t)(f[4] | (f[5] << 8)) + 1;
                if (total < 12 + xlen + 8 || at + total > size_) return false;
                if (msize) *msize = total;
                const uint8_t* t = h + total - 4;
                if (isize) *isize = t[0] | (t[1] << 8) | (t[2] << 16) | ((size_t)t[3] << 24);
                return true;
            }
            x += 4 + slen;
        }
        return false;
    }
    // as many whole BGZF members as fit `cap`, inflated by all threads; 0 when the next member is not BGZF any more
    // (the serial decoder takes over from there) or the file has ended
    size_t bgzf_read(char* dst, size_t cap) {
        struct M { size_t at, msize, isize, out; };
        std::vector<M> ms;
        size_t out = 0, at = gzpos_;
        while (at < size_) {
            size_t msz, isz;
            if (!bgzf_member(at, &msz, &isz)) break;
            if (isz > (1u << 16) || out + isz > cap) break;          // (BGZF caps a member at 64 KiB; a fuller block next time)
            ms.push_back(M{at, msz, isz, out});
            out += isz; at += msz;
        }
        if (ms.empty()) return 0;
        const int T = (int)std::min<size_t>((size_t)nthreads_, std::max<size_t>(1, ms.size() / 8));
        if ((int)zpool_.size() < T) zpool_.resize(T);
        std::vector<std::string> errs(T);
        run_parallel(T, [&](int t) {
            GzInflate& z = zpool_[t];
            const size_t a = ms.size() * t / T, b = ms.size() * (t + 1) / T;
            for (size_t i = a; i < b && errs[t].empty(); i++) {
                z.reset(gzmap_ + ms[i].at, ms[i].msize);
                const size_t got = z.read(reinterpret_cast<uint8_t*>(dst) + ms[i].out, ms[i].isize);
                uint8_t extra;
                if (!z.error().empty()) errs[t] = z.error();
                else if (got != ms[i].isize || z.read(&extra, 1) != 0 || !z.error().empty()) errs[t] = "gzip: a BGZF member does not hold what its header says";
            }
        });
        for (const std::string& e : errs) if (!e.empty()) { gz_err_ = e; return 0; }
        gzpos_ = at;
        return out;
    }
    long gz_read(char* dst, size_t cap) {
        if (gzmap_ && bgzf_) {
            const size_t got = bgzf_read(dst, cap);
            if (got || !gz_err_.empty()) return (long)got;
            if (gzpos_ >= size_) return 0;
            // not BGZF from here on (or a member too large for what is left of the block): the serial decoder goes on.
            // A block with room for less than one member must not switch: ask again with a fresh block.
            size_t msz, isz;
            if (bgzf_member(gzpos_, &msz, &isz) && isz <= (1u << 16) && cap < (1u << 16)) return GZ_FULL;
            bgzf_ = false;
            zfast_.reset(gzmap_ + gzpos_, size_ - gzpos_);
        }
        if (gzmap_ && gzpar_) {
            const size_t got = zpar_.read(reinterpret_cast<uint8_t*>(dst), cap);
            if (got == 0 && !zpar_.error().empty()) gz_err_ = zpar_.error();
            return (long)got;
        }
        if (gzmap_) {
            const size_t got = zfast_.read(reinterpret_cast<uint8_t*>(dst), cap);
            if (got == 0 && !zfast_.error().empty()) gz_err_ = zfast_.error();
            return (long)got;
        }
        const int got = gzread(gz_, dst, (unsigned)std::min<size_t>(cap, 1u << 30));
        if (got < 0) { int en = 0; const char* m = gzerror(gz_, &en); gz_err_ = std::string("gzip: ") + (m ? m : "read error"); }
        return got;
    }
    void inflate_loop() {
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || !ready_[put_]; });
                if (stop_) return;
            }
            RawBuf& B = blk_[put_];
            size_t n = carry_.size();
            if (n) memcpy(B.data(), carry_.data(), n);
            carry_.clear();
            while (!gz_eof_ && n < GZ_BLOCK) {
                const long got = gz_read(B.data() + n, GZ_BLOCK - n);
                if (got == GZ_FULL) break;                       // block gzip: no room for one more member
                if (got <= 0) { gz_eof_ = true; break; }
                n += (size_t)got;
            }
            size_t use = n;
            if (!gz_eof_) {   // cut at the last record start; a single record longer than a block keeps growing
                const char* cutp = last_record_start(B.data(), B.data() + n);
                while (cutp == B.data() && !gz_eof_ && n < B.size()) {
                    const long got = gz_read(B.data() + n, std::min<size_t>(B.size() - n, 16u << 20));
                    if (got == GZ_FULL) break;
                    if (got <= 0) { gz_eof_ = true; break; }
                    n += (size_t)got;
                    cutp = gz_eof_ ? B.data() + n : last_record_start(B.data(), B.data() + n);
                }
                use = gz_eof_ ? n : (size_t)(cutp - B.data());
                if (use < n) carry_.assign(B.data() + use, B.data() + n);
            }
            {
                std::lock_guard<std::mutex> g(mu_);
                blk_len_[put_] = use;
                if (use) { ready_[put_] = true; put_ ^= 1; }
                if (gz_eof_ && carry_.empty()) gz_eof_done_ = true;
            }
            cv_.notify_all();
            if (gz_eof_ && carry_.empty()) return;
        }
    }
    const char* last_record_start(const char* begin, const char* end) const {
        // the last position in (begin, end) that next_record_start accepts; begin if there is none
        const char* best = begin;
        size_t back = 1u << 16;
        for (;;) {
            const char* from = (size_t)(end - begin) > back ? end - back : begin;
            const char* p = from;
            const char* found = nullptr;
            for (;;) {
                const char* q = next_record_start(p == begin ? begin + 1 : p, begin, end, fasta_);
                if (q >= end) break;
                found = q;
                p = q + 1;
            }
            if (found) return found;
            if (from == begin) return best;
            back <<= 2;
        }
    }

    int nthreads_;
    bool fasta_ = true, want_ids_ = false;
    int fd_ = -1;
    size_t size_ = 0, pos_ = 0;
    const char* map_ = nullptr;
    gzFile gz_ = nullptr;            // MDBG_GZ_ZLIB=1 only
    bool gz_on_ = false;             // the file is gzip: blocks come from the inflate thread
    const uint8_t* gzmap_ = nullptr; // the compressed file, mapped
    GzInflate zfast_;
    std::vector<GzInflate> zpool_;   // one decoder per thread for block gzip
    bool bgzf_ = false, gzpar_ = false;
    GzParallel zpar_;                // one plain gzip stream inflated by all threads
    size_t gzpos_ = 0;               // block gzip: next member's offset in the file
    std::string gz_err_;
    // gz pipeline
    std::thread inflater_;
    std::mutex mu_;
    std::condition_variable cv_;
    struct RawBuf {                  // a block of text, NOT zero-filled (2 x 128 MB of page touching at open otherwise)
        char* p = nullptr; size_t n = 0;
        ~RawBuf() { free(p); }
        void resize(size_t m) { if (m != n) { free(p); p = (char*)malloc(m); n = p ? m : 0; } }
        char* data() { return p; }
        size_t size() const { return n; }
    };
    RawBuf blk_[2];
    size_t blk_len_[2] = {0, 0};
    bool ready_[2] = {false, false};
    int take_ = 0, put_ = 0;
    bool stop_ = false, gz_eof_ = false, gz_eof_done_ = false;
    std::vector<char> carry_;
};

}  // namespace ingest
