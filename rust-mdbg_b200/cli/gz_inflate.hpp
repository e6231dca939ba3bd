// gz_inflate.hpp -- the front end's gzip reader: a DEFLATE (RFC 1951) decoder for (multi-member) gzip files
// (RFC 1952) that are mapped into memory, written for the one thing the ingest does with them -- turn a .fa.gz /
// .fq.gz into text as fast as one thread can, because a gzip stream has no random access and its inflate is the
// floor of the whole file -> .gfa path (reference side: `seq_io` over flate2, main.rs:163-178: one thread too).
//
// What makes it faster than zlib's inflate on this kind of input (short literal codes, many mid-length matches):
//   * a 64-bit bit buffer refilled with one unaligned 8-byte load, once per symbol (a length/distance pair needs at
//     most 48 bits);
//   * one 11-bit table lookup resolves almost every literal/length code (sub-tables for the rare longer codes),
//     entries carry base value, extra-bit count and code length, so there is no second table for bases/extras;
//   * up to three literals are decoded per refill;
//   * matches are copied eight bytes at a time into a buffer with slack (no per-byte bounds checks inside);
//   * the output of a round (1 MiB) is CRC-checked with zlib's crc32 and handed over with one memcpy.
// Every member's CRC-32 and length are verified like gzread does; any malformed stream is an error, never a read
// or write outside the buffers (lengths/distances are validated, tables are built only from complete codes).
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>      // crc32 only
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace ingest {

class GzInflate {
public:
    // `in` must stay mapped while the object is used
    void reset(const uint8_t* in, size_t n) {
        in_ = in; ip_ = in; iend_ = in + n;
        bitbuf_ = 0; bitcnt_ = 0;
        state_ = ST_HEADER; final_block_ = false; eof_ = false; err_.clear();
        obuf_.resize(WIN + ROUND + SLACK);
        opos_ = WIN; pend_begin_ = pend_end_ = WIN; hist_valid_ = 0;
        crc_ = 0; member_out_ = 0; stored_left_ = 0; member_start_ = WIN; crc_from_ = WIN;
    }
    // Goes on in the MIDDLE of a member: at bit `bit` of in[0, n) a block starts, window[0, wn) is the text before it
    // (at most 32 KiB matter), crc / member_out are the member's CRC-32 and length so far.
    void resume(const uint8_t* in, size_t n, uint64_t bit, const uint8_t* window, size_t wn, uint32_t crc, uint64_t member_out) {
        reset(in, n);
        ip_ = in + (size_t)(bit >> 3);
        refill();
        if (bitcnt_ >= (int)(bit & 7)) drop((int)(bit & 7));
        if (wn > WIN) { window += wn - WIN; wn = WIN; }
        if (wn) memcpy(obuf_.data() + WIN - wn, window, wn);
        hist_valid_ = wn;
        state_ = ST_BLOCK_HEADER;
        crc_ = crc; member_out_ = member_out;
    }
    bool eof() const { return eof_ && pend_begin_ == pend_end_; }
    const std::string& error() const { return err_; }

    // Up to `cap` bytes of decompressed text into `out`; 0 at the end of the file or on an error (see error()).
    size_t read(uint8_t* out, size_t cap) {
        size_t done = 0;
        while (done < cap) {
            if (pend_begin_ == pend_end_) {
                if (eof_ || !err_.empty()) break;
                if (!round()) break;
                continue;
            }
            const size_t n = std::min(cap - done, pend_end_ - pend_begin_);
            memcpy(out + done, obuf_.data() + pend_begin_, n);
            pend_begin_ += n;
            done += n;
        }
        return done;
    }

protected:                                   // (GzSpan below decodes spans of a stream with the same tables and bit reader)
    static constexpr size_t WIN = 32768, ROUND = 1u << 20, SLACK = 258 + 16;
    static constexpr int LBITS = 11, DBITS = 8;                      // primary table bits
    enum { ST_HEADER, ST_BLOCK_HEADER, ST_STORED, ST_CODES, ST_TRAILER };
    // table entry: bits 0-4 code length (bits to drop), 5-9 extra bits (sub-table pointer: its index bits),
    // 10-11 type (0 literal, 1 base value of a length / distance, 2 end of block, 3 sub-table pointer),
    // 16-31 value (literal, base, or first entry of the sub-table); 0 = no such code.  (Entries holding TWO literals
    // whose codes fit the primary index together were tried: no gain on sequence text, the loop got longer.)
    static constexpr uint32_t T_LIT = 0u << 10, T_BASE = 1u << 10, T_EOB = 2u << 10, T_SUB = 3u << 10, T_MASK = 3u << 10;

    const uint8_t *in_ = nullptr, *ip_ = nullptr, *iend_ = nullptr;
    uint64_t bitbuf_ = 0;
    int bitcnt_ = 0;
    int state_ = ST_HEADER;
    bool final_block_ = false, eof_ = false;
    std::string err_;
    std::vector<uint8_t> obuf_;
    size_t opos_ = 0, pend_begin_ = 0, pend_end_ = 0, hist_valid_ = 0;
    uint32_t crc_ = 0;
    uint64_t member_out_ = 0;
    size_t member_start_ = WIN;                    // where the current member's output begins in obuf_ (this round's view)
    uint32_t stored_left_ = 0;
    std::vector<uint32_t> ltab_, dtab_;

    bool fail(const char* m) { err_ = std::string("gzip: ") + m; return false; }

    // ---- bits -------------------------------------------------------------------------------------------------
    inline void refill() {
        if (ip_ + 8 <= iend_) {
            uint64_t w;
            memcpy(&w, ip_, 8);
            bitbuf_ |= w << bitcnt_;
            ip_ += (63 - bitcnt_) >> 3;
            bitcnt_ |= 56;
        } else {
            while (bitcnt_ <= 56 && ip_ < iend_) { bitbuf_ |= (uint64_t)*ip_++ << bitcnt_; bitcnt_ += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(bitbuf_ & ((1ull << n) - 1)); }
    inline void drop(int n) { bitbuf_ >>= n; bitcnt_ -= n; }
    // bits for the slow paths (headers): false when the input ends
    bool need(int n) { if (bitcnt_ < n) refill(); return bitcnt_ >= n; }
    bool getbits(int n, uint32_t& v) { if (!need(n)) return false; v = peek(n); drop(n); return true; }
    void align_byte() { drop(bitcnt_ & 7); }
    // byte-aligned access: give the whole bytes of the bit buffer back to the input first
    void unget_bytes() { align_byte(); ip_ -= bitcnt_ >> 3; bitbuf_ = 0; bitcnt_ = 0; }

    // ---- canonical Huffman code -> lookup table -----------------------------------------------------------------
    static uint32_t rev(uint32_t c, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) { r = (r << 1) | (c & 1); c >>= 1; } return r; }
    // lens[0..n): code lengths (0 = unused).  entry_of(sym) = type | extra bits << 5 | value << 16 (without the length).
    template <class F>
    bool build(const uint8_t* lens, int n, int pbits, bool allow_incomplete, F entry_of, std::vector<uint32_t>& tab) {
        int count[16] = {0};
        for (int i = 0; i < n; i++) count[lens[i]]++;
        count[0] = 0;
        int used = 0;
        for (int l = 1; l < 16; l++) used += count[l];
        uint32_t kraft = 0;                        // in units of 2^-15
        for (int l = 1; l < 16; l++) kraft += (uint32_t)count[l] << (15 - l);
        if (kraft > (1u << 15)) return false;      // over-subscribed
        // incomplete: only no code at all (a block without matches has no distance code) or ONE code of length 1
        if (kraft < (1u << 15) && !(allow_incomplete && (used == 0 || (used == 1 && count[1] == 1)))) return false;
        uint32_t next[16];
        { uint32_t c = 0; for (int l = 1; l < 16; l++) { c = (c + (uint32_t)count[l - 1]) << 1; next[l] = c; } }
        // sub-table sizes: per primary prefix, the longest code that starts with it (only when some code is longer
        // than the primary index: rare on sequence text, so the common case touches nothing but the table itself)
        bool any_long = false;
        for (int l = pbits + 1; l < 16; l++) any_long = any_long || count[l] != 0;
        if (!any_long) {
            tab.assign((size_t)1 << pbits, 0);
            for (int s = 0; s < n; s++) {
                const int l = lens[s];
                if (!l) continue;
                const uint32_t r = rev(next[l]++, l);
                const uint32_t e = entry_of(s) | (uint32_t)l;
                for (uint32_t i = r; i < (1u << pbits); i += 1u << l) tab[i] = e;
            }
            return true;
        }
        std::vector<uint8_t> submax((size_t)1 << pbits, 0);
        {
            uint32_t nx[16];
            memcpy(nx, next, sizeof nx);
            for (int s = 0; s < n; s++) {
                const int l = lens[s];
                if (l > pbits) { const uint32_t r = rev(nx[l], l) & ((1u << pbits) - 1); if (submax[r] < l) submax[r] = (uint8_t)l; }
                if (l) nx[l]++;
            }
        }
        size_t total = (size_t)1 << pbits;
        std::vector<uint32_t> substart((size_t)1 << pbits, 0);
        for (size_t p = 0; p < ((size_t)1 << pbits); p++)
            if (submax[p]) { substart[p] = (uint32_t)total; total += (size_t)1 << (submax[p] - pbits); }
        tab.assign(total, 0);
        for (size_t p = 0; p < ((size_t)1 << pbits); p++)
            if (submax[p]) tab[p] = T_SUB | ((uint32_t)(submax[p] - pbits) << 5) | (uint32_t)pbits | (substart[p] << 16);
        if (total > 65535) return false;           // (cannot happen with <= 288 symbols of <= 15 bits)
        for (int s = 0; s < n; s++) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t r = rev(next[l]++, l);
            const uint32_t e = entry_of(s);
            if (l <= pbits) {
                for (uint32_t i = r; i < (1u << pbits); i += 1u << l) tab[i] = e | (uint32_t)l;
            } else {
                const uint32_t p = r & ((1u << pbits) - 1);
                const int sb = submax[p] - pbits;
                for (uint32_t i = r >> pbits; i < (1u << sb); i += 1u << (l - pbits)) tab[substart[p] + i] = e | (uint32_t)(l - pbits);
            }
        }
        return true;
    }
    static uint32_t litlen_entry(int s) {
        static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        if (s < 256) return T_LIT | ((uint32_t)s << 16);
        if (s == 256) return T_EOB;
        if (s > 285) return T_BASE | (0xFFFFu << 16);      // 286, 287: never valid in a stream
        return T_BASE | ((uint32_t)extra[s - 257] << 5) | ((uint32_t)base[s - 257] << 16);
    }
    static uint32_t dist_entry(int s) {
        static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                          4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        if (s >= 30) return T_BASE | (0xFFFFu << 16);      // 30, 31: never valid
        return T_BASE | ((uint32_t)extra[s] << 5) | ((uint32_t)base[s] << 16);
    }

    // ---- stream structure -----------------------------------------------------------------------------------------
    bool member_header() {
        unget_bytes();
        if (ip_ == iend_) { eof_ = true; return true; }
        // trailing zero padding after the last member is tolerated (as gzip -d does)
        if (*ip_ == 0) { const uint8_t* q = ip_; while (q < iend_ && *q == 0) q++; if (q == iend_) { ip_ = q; eof_ = true; return true; } }
        if (iend_ - ip_ < 18) return fail("truncated header");
        if (ip_[0] != 0x1f || ip_[1] != 0x8b) return fail("not a gzip member");
        if (ip_[2] != 8) return fail("unknown compression method");
        const uint8_t flg = ip_[3];
        if (flg & 0xE0) return fail("reserved header flags set");
        const uint8_t* p = ip_ + 10;
        if (flg & 4) { if (iend_ - p < 2) return fail("truncated header"); const size_t xl = p[0] | (p[1] << 8); p += 2; if ((size_t)(iend_ - p) < xl) return fail("truncated header"); p += xl; }
        for (int f = 3; f <= 4; f++)
            if (flg & (1 << f)) { while (p < iend_ && *p) p++; if (p == iend_) return fail("truncated header"); p++; }
        if (flg & 2) { if (iend_ - p < 2) return fail("truncated header"); p += 2; }
        ip_ = p;
        crc_ = 0; member_out_ = 0;
        member_start_ = opos_;
        state_ = ST_BLOCK_HEADER;
        return true;
    }
    bool member_trailer() {
        unget_bytes();
        if (iend_ - ip_ < 8) return fail("truncated trailer");
        const uint32_t crc = ip_[0] | (ip_[1] << 8) | (ip_[2] << 16) | ((uint32_t)ip_[3] << 24);
        const uint32_t isize = ip_[4] | (ip_[5] << 8) | (ip_[6] << 16) | ((uint32_t)ip_[7] << 24);
        ip_ += 8;
        flush_crc();
        if (crc != crc_) return fail("CRC-32 mismatch");
        if (isize != (uint32_t)member_out_) return fail("length mismatch");
        state_ = ST_HEADER;
        return true;
    }
    bool block_header() {
        uint32_t fin, type;
        if (!getbits(1, fin) || !getbits(2, type)) return fail("truncated stream");
        final_block_ = fin != 0;
        if (type == 0) {
            unget_bytes();
            if (iend_ - ip_ < 4) return fail("truncated stored block");
            const uint32_t len = ip_[0] | (ip_[1] << 8), nlen = ip_[2] | (ip_[3] << 8);
            if ((len ^ nlen) != 0xFFFFu) return fail("stored block length check failed");
            ip_ += 4;
            stored_left_ = len;
            state_ = ST_STORED;
            return true;
        }
        uint8_t lens[320];
        int nl, nd;
        if (type == 1) {
            nl = 288; nd = 32;
            for (int i = 0; i < 144; i++) lens[i] = 8;
            for (int i = 144; i < 256; i++) lens[i] = 9;
            for (int i = 256; i < 280; i++) lens[i] = 7;
            for (int i = 280; i < 288; i++) lens[i] = 8;
            for (int i = 0; i < 32; i++) lens[288 + i] = 5;
        } else if (type == 2) {
            uint32_t hlit, hdist, hclen;
            if (!getbits(5, hlit) || !getbits(5, hdist) || !getbits(4, hclen)) return fail("truncated stream");
            nl = (int)hlit + 257; nd = (int)hdist + 1;
            if (nl > 286 || nd > 30) return fail("too many length or distance codes");
            static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            uint8_t cl[19] = {0};
            for (uint32_t i = 0; i < hclen + 4; i++) { uint32_t v; if (!getbits(3, v)) return fail("truncated stream"); cl[order[i]] = (uint8_t)v; }
            std::vector<uint32_t> ctab;
            if (!build(cl, 19, 7, false, [](int s) { return T_LIT | ((uint32_t)s << 16); }, ctab)) return fail("invalid code-length code");
            int i = 0;
            while (i < nl + nd) {
                if (!need(7 + 7)) { if (bitcnt_ < 1) return fail("truncated stream"); }
                const uint32_t e = ctab[peek(7)];
                if (!e) return fail("invalid code-length symbol");
                if ((int)(e & 31) > bitcnt_) return fail("truncated stream");
                drop(e & 31);
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                int rep; uint8_t val = 0; uint32_t x;
                if (sym == 16) { if (i == 0) return fail("repeat without a previous length"); val = lens[i - 1]; if (!getbits(2, x)) return fail("truncated stream"); rep = 3 + (int)x; }
                else if (sym == 17) { if (!getbits(3, x)) return fail("truncated stream"); rep = 3 + (int)x; }
                else { if (!getbits(7, x)) return fail("truncated stream"); rep = 11 + (int)x; }
                if (i + rep > nl + nd) return fail("code lengths overrun");
                while (rep--) lens[i++] = val;
            }
            if (lens[256] == 0) return fail("no end-of-block code");
            memmove(lens + 288, lens + nl, (size_t)nd);          // distance lengths behind a fixed offset
            for (int k = nl; k < 288; k++) lens[k] = 0;
        } else {
            return fail("invalid block type");
        }
        if (!build(lens, type == 1 ? 288 : nl, LBITS, true, litlen_entry, ltab_)) return fail("invalid literal/length code");
        if (!build(lens + 288, nd, DBITS, true, dist_entry, dtab_)) return fail("invalid distance code");
        state_ = ST_CODES;
        return true;
    }

    void flush_crc() {                             // CRC of what this round has produced for the current member so far
        if (opos_ > crc_from_) { crc_ = crc32_fast(crc_, obuf_.data() + crc_from_, opos_ - crc_from_); crc_from_ = opos_; }
    }

    // CRC-32 (the gzip polynomial) by carry-less multiplication where the CPU has PCLMULQDQ: four 16-byte lanes folded
    // per step, then Barrett reduction (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ");
    // ~10x zlib's table-driven crc32_z, which takes the head/tail bytes and every CPU without the instruction.  Any
    // mistake here shows as a CRC error on a valid file (tests/test_cpu_gunzip.py), never as wrong output.
#if defined(__x86_64__)
    __attribute__((target("pclmul,sse4.1"))) static uint32_t crc32_clmul(uint32_t crc, const uint8_t* buf, size_t len) {   // len >= 64, multiple of 16
        const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll), k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
        const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124ll), poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
        __m128i x1 = _mm_loadu_si128((const __m128i*)(buf + 0)), x2 = _mm_loadu_si128((const __m128i*)(buf + 16));
        __m128i x3 = _mm_loadu_si128((const __m128i*)(buf + 32)), x4 = _mm_loadu_si128((const __m128i*)(buf + 48));
        x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
        buf += 64; len -= 64;
        while (len >= 64) {
            const __m128i y1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), y2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
            const __m128i y3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), y4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
            x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11); x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
            x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11); x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
            x1 = _mm_xor_si128(_mm_xor_si128(x1, y1), _mm_loadu_si128((const __m128i*)(buf + 0)));
            x2 = _mm_xor_si128(_mm_xor_si128(x2, y2), _mm_loadu_si128((const __m128i*)(buf + 16)));
            x3 = _mm_xor_si128(_mm_xor_si128(x3, y3), _mm_loadu_si128((const __m128i*)(buf + 32)));
            x4 = _mm_xor_si128(_mm_xor_si128(x4, y4), _mm_loadu_si128((const __m128i*)(buf + 48)));
            buf += 64; len -= 64;
        }
        for (const __m128i* nx : {&x2, &x3, &x4}) {        // four lanes -> one
            const __m128i y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
            x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), *nx), y);
        }
        while (len >= 16) {
            const __m128i y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
            x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), _mm_loadu_si128((const __m128i*)buf)), y);
            buf += 16; len -= 16;
        }
        // 128 -> 64 -> 32 bits
        const __m128i m32 = _mm_setr_epi32(~0, 0, ~0, 0);
        __m128i x0 = _mm_clmulepi64_si128(x1, k3k4, 0x10);
        x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x0);
        x0 = _mm_srli_si128(x1, 4);
        x1 = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(x1, m32), k5k0, 0x00), x0);
        x0 = _mm_clmulepi64_si128(_mm_and_si128(x1, m32), poly, 0x10);
        x0 = _mm_clmulepi64_si128(_mm_and_si128(x0, m32), poly, 0x00);
        return (uint32_t)_mm_extract_epi32(_mm_xor_si128(x1, x0), 1);
    }
#endif
public:
    static uint32_t crc32_of(const uint8_t* buf, size_t len) { return crc32_fast(0, buf, len); }
protected:
    static uint32_t crc32_fast(uint32_t crc, const uint8_t* buf, size_t len) {
#if defined(__x86_64__)
        static const bool have = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
        if (have && len >= 64) {
            const size_t n = len & ~(size_t)15;
            crc = ~crc32_clmul(~crc, buf, n);
            buf += n; len -= n;
        }
#endif
        return len ? (uint32_t)crc32_z(crc, buf, len) : crc;
    }
    size_t crc_from_ = WIN;

    // Decodes symbols of the current block until the round's output is full or the block ends.  The entry of the NEXT
    // symbol is looked up right after a refill and before the bytes of a match are copied, so the table load runs under
    // the copy (sequence text is mostly matches: ~7 bases on average at gzip -6, ~4 at -1).
    inline uint32_t litlen_entry_at_bits() const {
        uint32_t e = ltab_[peek(LBITS)];
        if ((e & T_MASK) == T_SUB) {
            const int pb = (int)(e & 31), sb = (int)((e >> 5) & 31);
            e = ltab_[(e >> 16) + ((uint32_t)(bitbuf_ >> pb) & ((1u << sb) - 1))];
            if (e) e += (uint32_t)pb;
        }
        return e;
    }
    bool codes() {
        uint8_t* const ob = obuf_.data();
        const size_t limit = WIN + ROUND;
        const uint32_t* const dt = dtab_.data();
        size_t op = opos_;
        const size_t floor = member_start_;        // lowest valid position for a back reference: the member's first byte
        refill();
        uint32_t e = litlen_entry_at_bits();
        while (op < limit) {                       // (e is only looked at, not consumed, when the round is full)
            if (!e) { opos_ = op; return fail("invalid literal/length code in the stream"); }
            if ((int)(e & 31) > bitcnt_) { opos_ = op; return fail("truncated stream"); }
            if ((e & T_MASK) == T_LIT) {           // up to three literals on one refill (3 x 15 bits at most)
                drop(e & 31);
                ob[op++] = (uint8_t)(e >> 16);
                for (int rep = 0; rep < 2; rep++) {
                    const uint32_t e2 = ltab_[peek(LBITS)];
                    if ((e2 & T_MASK) != T_LIT || !e2 || (int)(e2 & 31) > bitcnt_) break;
                    drop(e2 & 31);
                    ob[op++] = (uint8_t)(e2 >> 16);
                }
                refill();
                e = litlen_entry_at_bits();
                continue;
            }
            drop(e & 31);
            if ((e & T_MASK) == T_EOB) {
                opos_ = op;
                state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
                return true;
            }
            // length + distance (at most 5 + 15 + 13 more bits: they are there, a refill leaves 56)
            const int lx = (int)((e >> 5) & 31);
            uint32_t len = (e >> 16);
            if (len == 0xFFFFu) { opos_ = op; return fail("invalid length symbol"); }
            len += peek(lx);
            drop(lx);
            uint32_t d = dt[peek(DBITS)];
            if ((d & T_MASK) == T_SUB) { const int pb = (int)(d & 31), sb = (int)((d >> 5) & 31); d = dt[(d >> 16) + ((uint32_t)(bitbuf_ >> pb) & ((1u << sb) - 1))]; if (d) d += (uint32_t)pb; }
            if (!d || (d >> 16) == 0xFFFFu) { opos_ = op; return fail("invalid distance code in the stream"); }
            const int dx = (int)((d >> 5) & 31);
            if ((int)(d & 31) + dx > bitcnt_) { opos_ = op; return fail("truncated stream"); }
            drop(d & 31);
            const size_t dist = (size_t)(d >> 16) + peek(dx);
            drop(dx);
            if (dist > op - floor) { opos_ = op; return fail("distance reaches before the start of the data"); }
            refill();
            e = litlen_entry_at_bits();            // the next symbol's entry: its load overlaps the copy
            uint8_t* dst = ob + op;
            const uint8_t* src = dst - dist;
            op += len;
            if (dist >= 8) {                       // eight bytes at a time (may write up to 7 bytes into the slack)
                uint8_t* const end = dst + len;
                do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < end);
            } else if (dist == 1) {
                memset(dst, *src, len);
            } else {
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
            }
        }
        opos_ = op;
        return true;
    }

    bool stored() {
        while (stored_left_ && opos_ < WIN + ROUND) {
            const size_t n = std::min<size_t>({stored_left_, WIN + ROUND - opos_, (size_t)(iend_ - ip_)});
            if (n == 0) return fail("truncated stored block");
            memcpy(obuf_.data() + opos_, ip_, n);
            ip_ += n; opos_ += n; stored_left_ -= (uint32_t)n;
        }
        if (!stored_left_) state_ = final_block_ ? ST_TRAILER : ST_BLOCK_HEADER;
        return true;
    }

    // One round: slide the window, decode until ROUND bytes are there (or the file ends), publish them.
    bool round() {
        // keep the last WIN bytes of what was produced so far in front of the new output
        const size_t produced = opos_ - WIN;
        if (produced) {
            const size_t keep = std::min(WIN, hist_valid_ + produced);
            memmove(obuf_.data() + WIN - keep, obuf_.data() + opos_ - keep, keep);
            hist_valid_ = keep;
        }
        opos_ = WIN; crc_from_ = WIN;
        // a member that goes on from the last round: its history is what it has produced so far, at most the window
        if (state_ != ST_HEADER) member_start_ = WIN - (size_t)std::min<uint64_t>(hist_valid_, member_out_);
        bool ok = true;
        while (ok && opos_ < WIN + ROUND && !eof_) {
            switch (state_) {
                case ST_HEADER: ok = member_header(); break;
                case ST_BLOCK_HEADER: ok = block_header(); break;
                case ST_STORED: ok = stored(); break;
                case ST_CODES: ok = codes(); break;
                case ST_TRAILER: member_out_ += opos_ - crc_from_; ok = member_trailer(); break;
            }
        }
        if (ok && state_ != ST_HEADER) { member_out_ += opos_ - crc_from_; flush_crc(); }   // the member goes on in the next round
        pend_begin_ = WIN;
        pend_end_ = ok ? opos_ : WIN;              // nothing of a failed round is handed out
        return ok && pend_end_ > pend_begin_;
    }
};


// ---- one gzip stream, many threads ---------------------------------------------------------------------------------
// A deflate stream has no index, but its blocks can be FOUND: a thread that starts in the middle of the file tries bit
// offsets until a dynamic block header parses and the data behind it decodes (Kerbiriou & Chikhi, "Parallel decompression
// of gzip-compressed files and random access to DNA sequences", 2019; rapidgzip).  What it cannot know is the 32 KiB of
// text before its start, so it decodes into 16-bit symbols: a value >= 256 stands for "byte (value - 256) of the unknown
// window" and is copied by matches like any literal.  When the span before it is done its last 32 KiB are known and the
// markers are replaced.  Spans must join exactly (the span before ends on the bit the next one started at), else the later
// span is thrown away and decoded again from the true position -- correctness never rests on the block finder, and the
// member's CRC-32 (combined from the spans' CRCs) and length are verified at its end.

class GzSpan : public GzInflate {
public:
    static constexpr uint16_t UNKNOWN = 0xFFFF;    // before the start of the data: a match that reaches it is invalid
    std::vector<uint16_t> sym;                     // [0, WIN): image of the window before the span, then the span's text
    uint64_t end_bit = 0;                          // where decoding stopped: a block boundary (or the end of the final block)
    bool hit_final = false;
    size_t used = 0;                               // symbols of `sym` in use (the vector itself is kept larger between blocks)

    // Decodes blocks from bit `start_bit` of in[0, n) until the first block boundary at or after `stop_bit` or the end
    // of the final block.  window[0, wn): the text before the span if known, else markers.  False on any error.
    bool run(const uint8_t* in, size_t n, uint64_t start_bit, uint64_t stop_bit, const uint8_t* window, size_t wn,
             bool window_known, size_t max_out) {
        in_ = in; iend_ = in + n; err_.clear();
        seek_bit(start_bit);
        hit_final = false;
        bool imaged = false;                       // (the block finder calls this for one bit offset in eight: the 64 KiB
        for (;;) {                                 //  image is written only once a header has parsed)
            const uint64_t at = bit_pos();
            if (at >= stop_bit && imaged) { end_bit = at; sym.resize(used); return true; }
            if (!block_header()) return false;
            if (!imaged) {
                if (sym.size() < WIN + (1u << 16)) sym.resize(WIN + (1u << 16));
                used = WIN;
                if (window_known) {
                    if (wn > WIN) { window += wn - WIN; wn = WIN; }
                    for (size_t j = 0; j < WIN - wn; j++) sym[j] = UNKNOWN;
                    for (size_t j = 0; j < wn; j++) sym[WIN - wn + j] = window[j];
                } else {
                    for (size_t j = 0; j < WIN; j++) sym[j] = (uint16_t)(256 + j);
                }
                imaged = true;
            }
            if (state_ == ST_STORED) {
                if ((size_t)(iend_ - ip_) < stored_left_) return fail("truncated stored block");
                const size_t o = used;
                if (sym.size() < o + stored_left_) sym.resize(std::max(sym.size() * 2, o + stored_left_));
                for (uint32_t j = 0; j < stored_left_; j++) sym[o + j] = ip_[j];
                used = o + stored_left_;
                ip_ += stored_left_;
                stored_left_ = 0;
            } else if (!codes16(max_out)) {
                return false;
            }
            if (final_block_) { hit_final = true; end_bit = bit_pos(); sym.resize(used); return true; }
        }
    }
    // quick look: do the three header bits at `bit` say "not final, dynamic"?
    static bool plausible(const uint8_t* in, size_t n, uint64_t bit) {
        const size_t by = (size_t)(bit >> 3);
        if (by + 2 > n) return false;
        const uint32_t v = ((uint32_t)in[by] | ((uint32_t)in[by + 1] << 8)) >> (bit & 7);
        return (v & 7u) == 4u;
    }
    size_t n_out() const { return used - WIN; }

private:
    uint64_t bit_pos() const { return 8ull * (uint64_t)(ip_ - in_) - (uint64_t)bitcnt_; }
    void seek_bit(uint64_t bit) {
        ip_ = in_ + (size_t)(bit >> 3);
        bitbuf_ = 0; bitcnt_ = 0;
        refill();
        if (bitcnt_ >= (int)(bit & 7)) drop((int)(bit & 7));
    }
    // the symbol loop of GzInflate::codes with 16-bit output and no round limit (one block)
    bool codes16(size_t max_out) {
        const uint32_t* const dt = dtab_.data();
        size_t op = used;
        refill();
        uint32_t e = litlen_entry_at_bits();
        for (;;) {
            if (op + 600 > sym.size()) {
                if (op > max_out + WIN) { used = op; return fail("span grows beyond its bound"); }
                sym.resize(sym.size() * 2);
            }
            uint16_t* const ob = sym.data();
            if (!e) { used = op; return fail("invalid literal/length code in the stream"); }
            if ((int)(e & 31) > bitcnt_) { used = op; return fail("truncated stream"); }
            if ((e & T_MASK) == T_LIT) {
                drop(e & 31);
                ob[op++] = (uint16_t)(e >> 16);
                for (int rep = 0; rep < 2; rep++) {
                    const uint32_t e2 = ltab_[peek(LBITS)];
                    if ((e2 & T_MASK) != T_LIT || !e2 || (int)(e2 & 31) > bitcnt_) break;
                    drop(e2 & 31);
                    ob[op++] = (uint16_t)(e2 >> 16);
                }
                refill();
                e = litlen_entry_at_bits();
                continue;
            }
            drop(e & 31);
            if ((e & T_MASK) == T_EOB) { used = op; return true; }
            const int lx = (int)((e >> 5) & 31);
            uint32_t len = (e >> 16);
            if (len == 0xFFFFu) { used = op; return fail("invalid length symbol"); }
            len += peek(lx);
            drop(lx);
            uint32_t d = dt[peek(DBITS)];
            if ((d & T_MASK) == T_SUB) { const int pb = (int)(d & 31), sb = (int)((d >> 5) & 31); d = dt[(d >> 16) + ((uint32_t)(bitbuf_ >> pb) & ((1u << sb) - 1))]; if (d) d += (uint32_t)pb; }
            if (!d || (d >> 16) == 0xFFFFu) { used = op; return fail("invalid distance code in the stream"); }
            const int dx = (int)((d >> 5) & 31);
            if ((int)(d & 31) + dx > bitcnt_) { used = op; return fail("truncated stream"); }
            drop(d & 31);
            const size_t dist = (size_t)(d >> 16) + peek(dx);
            drop(dx);
            if (dist > op) { used = op; return fail("distance reaches before the window"); }
            refill();
            e = litlen_entry_at_bits();
            const uint16_t* src = ob + op - dist;
            uint16_t* dst = ob + op;
            if (src[0] == UNKNOWN) { used = op; return fail("distance reaches before the start of the data"); }
            op += len;
            if (dist >= 8) {                       // sixteen bytes = eight symbols at a time (the slack absorbs the overrun)
                uint16_t* const end = dst + len;
                do { uint64_t w0, w1; memcpy(&w0, src, 8); memcpy(&w1, src + 4, 8); memcpy(dst, &w0, 8); memcpy(dst + 4, &w1, 8); src += 8; dst += 8; } while (dst < end);
            } else if (dist >= 4) {
                uint16_t* const end = dst + len;
                do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 4; dst += 4; } while (dst < end);
            } else {
                for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
            }
        }
    }
};

// Drives spans over one gzip member: text comes out in order through read().  Anything it is not made for (a second
// member, a span that outgrows its bound, a file that is too small to be worth it) goes to the serial decoder.
class GzParallel {
public:
    void reset(const uint8_t* in, size_t n, int threads, size_t span_bytes) {
        in_ = in; n_ = n; T_ = std::max(1, threads); span_ = std::max<size_t>(span_bytes, 1u << 16);
        err_.clear(); text_.clear(); tpos_ = 0; done_ = false; serial_ = false; started_ = false;
        window_.clear(); crc_ = 0; total_ = 0; next_bit_ = 0;
        spans_.resize((size_t)T_);
    }
    const std::string& error() const { return err_; }
    size_t read(uint8_t* out, size_t cap) {
        size_t got = 0;
        while (got < cap) {
            if (tpos_ < text_.size()) {            // what the last batch produced
                const size_t m = std::min(cap - got, text_.size() - tpos_);
                memcpy(out + got, text_.data() + tpos_, m);
                tpos_ += m; got += m;
                continue;
            }
            if (serial_) {
                const size_t g = z_.read(out + got, cap - got);
                if (g == 0) { if (!z_.error().empty()) err_ = z_.error(); break; }
                got += g;
                continue;
            }
            if (done_ || !err_.empty()) break;
            if (!batch() && !serial_) break;
        }
        return got;
    }

private:
    const uint8_t* in_ = nullptr;
    size_t n_ = 0, span_ = 0;
    int T_ = 1;
    std::string err_;
    std::vector<uint8_t> text_;                    // resolved text of the last batch
    size_t tpos_ = 0;
    bool done_ = false, serial_ = false, started_ = false;
    std::vector<uint8_t> window_;                  // last <= 32 KiB of text before next_bit_
    uint32_t crc_ = 0;
    uint64_t total_ = 0, next_bit_ = 0;            // member so far: CRC, length; where its next block starts
    std::vector<GzSpan> spans_;
    GzInflate z_;

    bool fail(const std::string& m) { err_ = m; return false; }
    // the rest of the file through the serial decoder, as a fresh stream from byte `at`
    void go_serial(size_t at) { z_.reset(in_ + at, n_ - at); serial_ = true; }

    template <class F>
    static void parallel(int T, F&& f) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; t++) th.emplace_back([&f, t] { f(t); });
        f(0);
        for (auto& x : th) x.join();
    }

    bool batch() {
        if (!started_) {                           // the member header: the serial decoder's parser knows it
            struct H : GzInflate { bool parse(const uint8_t* in, size_t n, size_t& deflate_at, std::string& e) { reset(in, n); if (!member_header()) { e = err_; return false; } if (eof_) { deflate_at = n; return true; } deflate_at = (size_t)(ip_ - in_); return true; } } h;
            size_t at = 0;
            if (!h.parse(in_, n_, at, err_)) return false;
            if (at >= n_) { done_ = true; return false; }
            next_bit_ = 8ull * at;
            started_ = true;
        }
        const uint64_t nbits = 8ull * n_;
        // span t nominally covers compressed bytes [base + t span_, base + (t + 1) span_)
        const size_t base = (size_t)(next_bit_ >> 3);
        std::vector<uint64_t> start((size_t)T_, 0);
        std::vector<char> ok((size_t)T_, 0);
        const size_t max_out = span_ * 24;         // a span that expands more than this is not sequence text (3.5-5x): serial
                                                   // decoder; also bounds the memory of a batch (2 B per symbol and span)
        parallel(T_, [&](int t) {
            GzSpan& S = spans_[(size_t)t];
            const uint64_t nominal = 8ull * (base + (size_t)t * span_), stop = std::min<uint64_t>(nbits, 8ull * (base + (size_t)(t + 1) * span_));
            if (t == 0) {
                start[0] = next_bit_;
                ok[0] = S.run(in_, n_, next_bit_, stop, window_.data(), window_.size(), true, max_out) ? 1 : 0;
                return;
            }
            if (nominal + 64 >= nbits) return;
            for (uint64_t b = nominal; b < stop; b++) {          // the first offset from which everything decodes
                if (!GzSpan::plausible(in_, n_, b)) continue;
                if (S.run(in_, n_, b, stop, nullptr, 0, false, max_out)) { start[(size_t)t] = b; ok[(size_t)t] = 1; return; }
            }
        });
        if (!ok[0]) {
            // span 0 starts at a true block boundary with its true window: its failure is the stream's (or the bound's)
            const std::string& e = spans_[0].error();
            if (e.find("beyond its bound") == std::string::npos) return fail(e.empty() ? "gzip: damaged stream" : e);
            // text that expands more than 24x is not what the marker scheme is for: the serial decoder takes the
            // member over where it stands
            z_.resume(in_, n_, next_bit_, window_.data(), window_.size(), crc_, total_);
            serial_ = true;
            text_.clear(); tpos_ = 0;
            return true;
        }
        // join the spans: span t is kept only if the text before it ended exactly where it started
        int used = 1;
        uint64_t end = spans_[0].end_bit;
        bool final = spans_[0].hit_final;
        while (!final && used < T_ && ok[(size_t)used] && start[(size_t)used] == end) {
            end = spans_[(size_t)used].end_bit;
            final = spans_[(size_t)used].hit_final;
            used++;
        }
        // windows in order (each needs the resolved tail of the span before it), then every span's text in parallel
        std::vector<std::vector<uint8_t>> win((size_t)used + 1);
        win[0] = window_;
        std::vector<size_t> off((size_t)used + 1, 0);
        for (int t = 0; t < used; t++) off[(size_t)t + 1] = off[(size_t)t] + spans_[(size_t)t].n_out();
        for (int t = 0; t < used; t++) {
            const GzSpan& S = spans_[(size_t)t];
            const std::vector<uint8_t>& w = win[(size_t)t];
            const size_t have = S.used;            // window image + text
            const size_t take = std::min<size_t>(32768, S.n_out() + w.size());
            std::vector<uint8_t>& nw = win[(size_t)t + 1];
            nw.resize(take);
            for (size_t j = 0; j < take; j++) {
                const uint16_t v = S.sym[have - take + j];
                if (v < 256) nw[j] = (uint8_t)v;
                else {
                    const size_t k = (size_t)v - 256;            // byte k of the 32 KiB image before the span
                    if (v == GzSpan::UNKNOWN || k + w.size() < 32768) return fail("gzip: distance reaches before the start of the data");
                    nw[j] = w[k - (32768 - w.size())];
                }
            }
        }
        text_.resize(off[(size_t)used]);
        tpos_ = 0;
        std::vector<uint32_t> crcs((size_t)used, 0);
        std::vector<char> bad((size_t)used, 0);
        parallel(used, [&](int t) {
            const GzSpan& S = spans_[(size_t)t];
            const std::vector<uint8_t>& w = win[(size_t)t];
            uint8_t* dst = text_.data() + off[(size_t)t];
            const uint16_t* src = S.sym.data() + 32768;
            const size_t m = S.n_out(), wpad = 32768 - w.size();
            size_t j = 0;
            while (j < m) {
                // eight symbols at once while none of them is a marker (all of them, a window's reach into the span)
                while (j + 8 <= m) {
                    uint64_t a, b;
                    memcpy(&a, src + j, 8); memcpy(&b, src + j + 4, 8);
                    if ((a | b) & 0xFF00FF00FF00FF00ull) break;
                    const uint64_t lo = (a & 0xFF) | ((a >> 8) & 0xFF00) | ((a >> 16) & 0xFF0000) | ((a >> 24) & 0xFF000000ull);
                    const uint64_t hi = (b & 0xFF) | ((b >> 8) & 0xFF00) | ((b >> 16) & 0xFF0000) | ((b >> 24) & 0xFF000000ull);
                    const uint64_t o = lo | (hi << 32);
                    memcpy(dst + j, &o, 8);
                    j += 8;
                }
                if (j >= m) break;
                const uint16_t v = src[j];
                if (v < 256) dst[j] = (uint8_t)v;
                else if (v == GzSpan::UNKNOWN || (size_t)v - 256 < wpad) { bad[(size_t)t] = 1; dst[j] = 0; }
                else dst[j] = w[(size_t)v - 256 - wpad];
                j++;
            }
            crcs[(size_t)t] = GzInflate::crc32_of(dst, m);
        });
        for (int t = 0; t < used; t++) {
            if (bad[(size_t)t]) return fail("gzip: distance reaches before the start of the data");
            crc_ = (uint32_t)crc32_combine(crc_, crcs[(size_t)t], (z_off_t)spans_[(size_t)t].n_out());
            total_ += spans_[(size_t)t].n_out();
        }
        window_ = win[(size_t)used];
        next_bit_ = end;
        if (final) {                               // trailer, then whatever follows goes to the serial decoder
            const size_t tr = (size_t)((end + 7) >> 3);
            if (n_ - tr < 8) return fail("gzip: truncated trailer");
            const uint32_t crc = in_[tr] | (in_[tr + 1] << 8) | (in_[tr + 2] << 16) | ((uint32_t)in_[tr + 3] << 24);
            const uint32_t isz = in_[tr + 4] | (in_[tr + 5] << 8) | (in_[tr + 6] << 16) | ((uint32_t)in_[tr + 7] << 24);
            if (crc != crc_) return fail("gzip: CRC-32 mismatch");
            if (isz != (uint32_t)total_) return fail("gzip: length mismatch");
            go_serial(tr + 8);                     // further members, zero padding or the end of the file
            if (text_.empty()) return true;        // (read() goes on with the serial decoder)
        }
        return true;
    }
};

}  // namespace ingest
