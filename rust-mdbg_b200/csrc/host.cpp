// host.cpp -- host-side value helpers of the C ABI: KmerVec ops (src/kmer_vec.rs:16-47,73-77),
// read sharding and fingerprint->owner planning for the multi-GPU path (pure functions).
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../include/mdbg.h"
#include "pack_host.h"
#include "mdbg_common.cuh"

extern "C" {

// kmer_vec.rs:73-77 Ord = lexicographic on the u64 vector
int mdbg_kminmer_cmp(const uint64_t* a, const uint64_t* b, uint32_t k) {
    for (uint32_t i = 0; i < k; i++) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
// kmer_vec.rs:28-32
void mdbg_kminmer_reverse(const uint64_t* in, uint32_t k, uint64_t* out) {
    std::vector<uint64_t> t(in, in + k);
    for (uint32_t i = 0; i < k; i++) out[i] = t[k - 1 - i];
}
// kmer_vec.rs:34-39: (self,false) if self < rev else (rev,true); palindrome => reversed
void mdbg_kminmer_normalize(const uint64_t* in, uint32_t k, uint64_t* out, int* reversed) {
    bool fwd_less = false;
    for (uint32_t i = 0; i < k; i++) {
        uint64_t a = in[i], b = in[k - 1 - i];
        if (a != b) { fwd_less = a < b; break; }
    }
    if (fwd_less) {
        if (out != in) memmove(out, in, (size_t)k * 8);
        if (reversed) *reversed = 0;
    } else {
        mdbg_kminmer_reverse(in, k, out);
        if (reversed) *reversed = 1;
    }
}
// kmer_vec.rs:22-26 / 16-20
void mdbg_kminmer_prefix(const uint64_t* in, uint32_t k, uint64_t* out) { memmove(out, in, (size_t)(k - 1) * 8); }
void mdbg_kminmer_suffix(const uint64_t* in, uint32_t k, uint64_t* out) { memmove(out, in + 1, (size_t)(k - 1) * 8); }

// Reads are sharded by record in contiguous global-index ranges.
void mdbg_shard_reads(uint64_t n, int world, int rank, uint64_t* lo, uint64_t* hi) {
    uint64_t q = n / (uint64_t)world, r = n % (uint64_t)world;
    uint64_t a = q * (uint64_t)rank + std::min<uint64_t>((uint64_t)rank, r);
    *lo = a;
    *hi = a + q + ((uint64_t)rank < r ? 1 : 0);
}
uint64_t mdbg_tuple_fingerprint(const uint64_t* t, uint32_t k, uint64_t seed) {
    uint64_t h = mdbg::fp_init(seed, k);
    for (uint32_t i = 0; i < k; i++) h = mdbg::fp_mix(h, t[i]);
    return mdbg::fp_fin(h);
}
// owner = top bits of the fingerprint, scaled to the world size (works for any world)
uint32_t mdbg_owner_of_fingerprint(uint64_t fp, int world) {
    return (uint32_t)(((unsigned __int128)fp * (unsigned __int128)(uint64_t)world) >> 64);
}

}  // extern "C"

// ---- file formats ---------------------------------------------------------
#include <cstdio>
#include <string>

namespace {

// utils::revcomp, src/utils.rs:3-24 (anything but ACGTU/acgtu becomes 'N')
void revcomp_into(const uint8_t* in, uint64_t len, std::string& out) {
    out.resize(len);
    for (uint64_t i = 0; i < len; i++) {
        char o;
        switch (in[len - 1 - i]) {
            case 'a': o = 't'; break; case 'c': o = 'g'; break; case 't': o = 'a'; break;
            case 'g': o = 'c'; break; case 'u': o = 'a'; break; case 'A': o = 'T'; break;
            case 'C': o = 'G'; break; case 'T': o = 'A'; break; case 'G': o = 'C'; break;
            case 'U': o = 'A'; break; default: o = 'N';
        }
        out[i] = o;
    }
}

uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
uint32_t xxh32(const uint8_t* p, size_t len, uint32_t seed) {  // XXH32, needed for the LZ4 frame header check byte
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t* end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        while (p + 16 <= end) {
            uint32_t w;
            memcpy(&w, p, 4); v1 = rotl32(v1 + w * P2, 13) * P1; p += 4;
            memcpy(&w, p, 4); v2 = rotl32(v2 + w * P2, 13) * P1; p += 4;
            memcpy(&w, p, 4); v3 = rotl32(v3 + w * P2, 13) * P1; p += 4;
            memcpy(&w, p, 4); v4 = rotl32(v4 + w * P2, 13) * P1; p += 4;
        }
        h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
    } else h = seed + P5;
    h += (uint32_t)len;
    while (p + 4 <= end) { uint32_t w; memcpy(&w, p, 4); h = rotl32(h + w * P3, 17) * P4; p += 4; }
    while (p < end) { h = rotl32(h + (*p) * P5, 11) * P1; p++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

// LZ4 frame writer using only STORED blocks: a valid frame any LZ4 reader (lzzzz in
// to_basespace.rs:233, python lz4.frame) decodes, without needing liblz4 headers here.
struct Lz4StoredWriter {
    FILE* f;
    bool frame;
    std::string buf;
    static constexpr size_t BLOCK = 4u << 20;
    bool begin() {
        if (!frame) return true;
        uint8_t hdr[7] = {0x04, 0x22, 0x4D, 0x18, 0x60, 0x70, 0};  // magic, FLG(v1, indep), BD(4 MiB)
        hdr[6] = (uint8_t)((xxh32(hdr + 4, 2, 0) >> 8) & 0xFF);
        return fwrite(hdr, 1, 7, f) == 7;
    }
    bool flush_block() {
        if (buf.empty()) return true;
        if (frame) {
            uint32_t sz = (uint32_t)buf.size() | 0x80000000u;  // high bit: uncompressed block
            uint8_t b[4] = {(uint8_t)sz, (uint8_t)(sz >> 8), (uint8_t)(sz >> 16), (uint8_t)(sz >> 24)};
            if (fwrite(b, 1, 4, f) != 4) return false;
        }
        bool ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        buf.clear();
        return ok;
    }
    bool write(const char* p, size_t n) {
        while (n) {
            size_t take = std::min(n, BLOCK - buf.size());
            buf.append(p, take);
            p += take; n -= take;
            if (buf.size() == BLOCK && !flush_block()) return false;
        }
        return true;
    }
    bool end() {
        if (!flush_block()) return false;
        if (frame) { uint8_t z[4] = {0, 0, 0, 0}; if (fwrite(z, 1, 4, f) != 4) return false; }
        return true;
    }
};

}  // namespace

extern "C" {

// {prefix}.gfa: header main.rs:1011, S lines main.rs:1021 (all before any L), L lines main.rs:1095/1113.
// Lines come out in ascending node index / sorted edge order (the reference's order is
// DashMap-iteration dependent; the line SET is what is specified).
int mdbg_write_gfa(const mdbg_graph* g, const char* path) {
    if (!g || !path) return MDBG_ERR_BAD_ARG;
    FILE* f = fopen(path, "w");
    if (!f) return MDBG_ERR_IO;
    std::vector<char> big(8u << 20);
    setvbuf(f, big.data(), _IOFBF, big.size());
    fputs("H\tVN:Z:1.0\n", f);
    for (uint64_t i = 0; i < g->n_nodes; i++)
        fprintf(f, "S\t%u\t*\tLN:i:%u\tKC:i:%u\n", g->node_index[i], g->seqlen[i], (unsigned)g->abundance[i]);
    for (uint64_t i = 0; i < g->n_edges; i++)
        fprintf(f, "L\t%u\t%c\t%u\t%c\t%uM\n", g->e_n1[i], g->e_o1[i] ? '-' : '+', g->e_n2[i],
                g->e_o2[i] ? '-' : '+', g->e_overlap[i]);
    return fclose(f) == 0 ? MDBG_OK : MDBG_ERR_IO;
}

// {prefix}.{tid}.sequences: 4 '#' header lines (main.rs:625-628) then one line per q-entry
// "{index}\t{:?Vec<u64>}\t{seq}\t*\t*\t({s0}, {s1})" (main.rs:702), seq = raw[start..end),
// reverse-complemented when the node was reversed (main.rs:700-701).
// Streaming form: the q-entries are in serial (read, window) order, so a host that walks its reads once
// more in input order hands every read the writer asks for to mdbg_seq_writer_read -- nobody has to keep
// the whole read set in memory (the reference writes each line while it still holds the read, main.rs:696).
struct mdbg_seq_writer {
    const mdbg_graph* g = nullptr;
    FILE* f = nullptr;
    Lz4StoredWriter w{nullptr, false, {}};
    uint64_t q = 0;
    bool ok = true;
    std::string line, rc;
    // one of n_parts writers that share the lines of a graph (the reference writes one file per worker thread,
    // main.rs:614-630; its readers glob them): this one owns the runs of SEQ_RUN lines with run % n_parts == part
    uint32_t part = 0, n_parts = 1;
    static constexpr uint64_t SEQ_RUN = 32;
    bool owns(uint64_t qq) const { return n_parts <= 1 || (qq / SEQ_RUN) % n_parts == part; }
    void skip_foreign() { while (q < g->n_seqlines && !owns(q)) q++; }
};

int mdbg_seq_writer_open(const mdbg_graph* g, const char* path, int lz4_frame, mdbg_seq_writer** out) {
    return mdbg_seq_writer_open_part(g, path, lz4_frame, 0, 1, out);
}

int mdbg_seq_writer_open_part(const mdbg_graph* g, const char* path, int lz4_frame, uint32_t part, uint32_t n_parts,
                              mdbg_seq_writer** out) {
    if (!g || !path || !out || (g->n_seqlines && !g->q_index) || n_parts == 0 || part >= n_parts) return MDBG_ERR_BAD_ARG;
    *out = nullptr;
    FILE* f = fopen(path, "wb");
    if (!f) return MDBG_ERR_IO;
    mdbg_seq_writer* W = new mdbg_seq_writer();
    W->g = g; W->f = f;
    W->w = Lz4StoredWriter{f, lz4_frame != 0, {}};
    W->ok = W->w.begin();
    std::string hdr = "# k = " + std::to_string(g->k) + "\n# l = " + std::to_string(g->l) +
                      "\n# Structure of remaining of the file:\n"
                      "# [node name]\t[list of minimizers]\t[sequence of node]\t[abundance]\t[origin]\t[shift]\n";
    W->ok = W->ok && W->w.write(hdr.data(), hdr.size());
    W->part = part; W->n_parts = n_parts;
    W->skip_foreign();
    *out = W;
    return MDBG_OK;
}

// global index of the read the next line is cut from; UINT64_MAX when every line has been written
uint64_t mdbg_seq_writer_next_read(const mdbg_seq_writer* W) {
    return (W && W->q < W->g->n_seqlines) ? W->g->q_read[W->q] : UINT64_MAX;
}

// Writes every pending line cut from read `read_index` (bases of that read: read_bases[0 .. read_len)).
// Reads must come in ascending order; reads no line needs may be skipped.
int mdbg_seq_writer_read(mdbg_seq_writer* W, uint64_t read_index, const uint8_t* read_bases, uint64_t read_len) {
    if (!W) return MDBG_ERR_BAD_ARG;
    const mdbg_graph* g = W->g;
    while (W->ok && W->q < g->n_seqlines && g->q_read[W->q] == read_index) {
        const uint64_t q = W->q++;
        if (!W->owns(q)) continue;                 // another part's line of the same read
        if (g->q_end[q] > read_len || g->q_start[q] > g->q_end[q] || !read_bases) return MDBG_ERR_BAD_ARG;
        const uint32_t idx = g->q_index[q];
        const uint32_t* it = std::lower_bound(g->node_index, g->node_index + g->n_nodes, idx);   // nodes ascend in index
        std::string& line = W->line;
        line = std::to_string(idx) + "\t[";
        if (it != g->node_index + g->n_nodes && *it == idx) {
            const uint64_t* t = g->tuple + (uint64_t)(it - g->node_index) * g->k;
            for (uint32_t j = 0; j < g->k; j++) { if (j) line += ", "; line += std::to_string(t[j]); }
        }
        line += "]\t";
        const uint8_t* s = read_bases + g->q_start[q];
        const uint64_t len = g->q_end[q] - g->q_start[q];
        if (g->q_reversed[q]) { revcomp_into(s, len, W->rc); line += W->rc; }
        else line.append((const char*)s, len);
        line += "\t*\t*\t(" + std::to_string(g->q_shift[2 * q]) + ", " + std::to_string(g->q_shift[2 * q + 1]) + ")\n";
        W->ok = W->w.write(line.data(), line.size());
    }
    W->skip_foreign();
    return W->ok ? MDBG_OK : MDBG_ERR_IO;
}

int mdbg_seq_writer_close(mdbg_seq_writer* W) {
    if (!W) return MDBG_ERR_BAD_ARG;
    bool ok = W->ok && W->q == W->g->n_seqlines;   // a line whose read never came is an error
    ok = W->w.end() && ok;
    ok = (fclose(W->f) == 0) && ok;
    delete W;
    return ok ? MDBG_OK : MDBG_ERR_IO;
}

// The same in one call when the host still holds all reads (bases / read_off in global read order).
int mdbg_write_sequences(const mdbg_graph* g, const uint8_t* bases, const uint64_t* read_off, const char* path,
                         int lz4_frame) {
    if (!g || !path || (g->n_seqlines && (!g->q_index || !bases || !read_off))) return MDBG_ERR_BAD_ARG;
    mdbg_seq_writer* W = nullptr;
    int rc = mdbg_seq_writer_open(g, path, lz4_frame, &W);
    if (rc) return rc;
    for (uint64_t r; rc == MDBG_OK && (r = mdbg_seq_writer_next_read(W)) != UINT64_MAX;)
        rc = mdbg_seq_writer_read(W, r, bases + read_off[r], read_off[r + 1] - read_off[r]);
    const int rc2 = mdbg_seq_writer_close(W);
    return rc ? rc : rc2;
}

int mdbg_pack_bases_host(const uint8_t* bases, uint64_t n_bases, uint32_t* planes, uint8_t* bad_tiles, int threads) {
    if ((!bases && n_bases) || !planes) return MDBG_ERR_BAD_ARG;
    const uint64_t n_words = (n_bases + 31) / 32;
    if (bad_tiles) memset(bad_tiles, 0, (size_t)((n_bases + 4095) / 4096));
    if (threads <= 1) {
        mdbg::pack_words(bases, n_bases, 0, n_words, planes, bad_tiles);
    } else {
        mdbg::PackPool pool(threads);
        mdbg::pack_parallel(pool, bases, n_bases, 0, n_words, planes, bad_tiles);
    }
    return MDBG_OK;
}

}  // extern "C"
