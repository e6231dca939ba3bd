// rust-mdbg (B200) -- command line front end with the reference's flags (src/main.rs:228-423) for
// the hot path: reads (FASTA/FASTQ, plain or .gz) -> {prefix}.gfa + {prefix}.0.sequences.
// Host ingest only: parse, batch, mdbg_push_reads; all compute is in
// libmdbg_b200.so.  Modes outside the hot path are refused, never silently run elsewhere.
// stdout follows the reference line by line.
#include <sys/resource.h>
#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <glob.h>
#include <string>
#include <vector>

#include "../../include/mdbg.h"

namespace {

struct Fastx {  // FASTA (multi-line) / FASTQ (4-line) records from a plain or gzip file (zlib reads both)
    gzFile f = nullptr;
    bool fasta = true;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    std::string pending;  // FASTA: header line already consumed
    std::string id;       // of the record next() returned: the header up to the first space (seq_io's id())
    void set_id(const std::string& header) {
        size_t e = header.find(' ');
        id = header.substr(1, e == std::string::npos ? std::string::npos : e - 1);
    }
    bool open(const char* path, bool is_fasta) {
        f = gzopen(path, "rb");
        if (!f) return false;
        gzbuffer(f, 1 << 20);
        buf.resize(1 << 22);
        fasta = is_fasta;
        return true;
    }
    bool getline(std::string& out) {
        out.clear();
        for (;;) {
            if (pos == len) {
                int n = gzread(f, buf.data(), (unsigned)buf.size());
                if (n <= 0) return !out.empty();
                pos = 0; len = (size_t)n;
            }
            char* s = buf.data() + pos;
            char* e = (char*)memchr(s, '\n', len - pos);
            if (e) { out.append(s, e - s); pos = (e - buf.data()) + 1; break; }
            out.append(s, len - pos);
            pos = len;
        }
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return true;
    }
    // appends the sequence to `seq`; false at end of file
    bool next(std::string& seq) {
        std::string line;
        seq.clear();
        if (fasta) {
            if (pending.empty()) {
                do { if (!getline(line)) return false; } while (line.empty() || line[0] != '>');
                pending = line;
            }
            set_id(pending);
            pending.clear();
            while (getline(line)) {
                if (!line.empty() && line[0] == '>') { pending = line; return true; }
                seq += line;
            }
            return true;
        }
        if (!getline(line)) return false;      // @id
        if (line.empty()) return false;
        set_id(line);
        if (!getline(seq)) return false;       // sequence
        getline(line);                         // +
        getline(line);                         // qualities
        return true;
    }
    void close() { if (f) gzclose(f); f = nullptr; }
};

[[noreturn]] void die(const std::string& m) { fprintf(stderr, "error: %s\n", m.c_str()); exit(1); }

std::string rust_f64(double v) {  // Rust `{}` of an f64: shortest representation that round-trips
    char b[64];
    for (int p = 1; p <= 17; p++) { snprintf(b, sizeof b, "%.*g", p, v); if (strtod(b, nullptr) == v) break; }
    std::string s(b);
    if (s.find('e') != std::string::npos) { snprintf(b, sizeof b, "%.17f", v); s = b; while (s.back() == '0') s.pop_back(); }
    return s;
}

std::string rust_f32(float v) {  // Rust `{}` of an f32
    char b[64];
    for (int p = 1; p <= 9; p++) { snprintf(b, sizeof b, "%.*g", p, (double)v); if (strtof(b, nullptr) == v) break; }
    return std::string(b);
}

}  // namespace

int main(int argc, char** argv) {
    auto t_start = std::chrono::steady_clock::now();
    std::string reads, prefix, read_stats;
    long k = -1, l = -1, minabund = -1, threads = -1;
    double density = -1;
    float presimp = -1;
    bool skiphpc = false, no_basespace = false, use_bf = false;
    int device = 0;
    auto need = [&](int& i) -> const char* { if (i + 1 >= argc) die(std::string("missing value for ") + argv[i]); return argv[++i]; };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-k") k = atol(need(i));
        else if (a == "-l") l = atol(need(i));
        else if (a == "-d" || a == "--density") density = atof(need(i));
        else if (a == "--minabund") minabund = atol(need(i));
        else if (a == "--presimp") presimp = (float)atof(need(i));
        else if (a == "--threads") threads = atol(need(i));
        else if (a == "-p" || a == "--prefix") prefix = need(i);
        else if (a == "--skiphpc") skiphpc = true;
        else if (a == "--no-basespace") no_basespace = true;
        else if (a == "--debug") {}
        else if (a == "--device") device = atoi(need(i));   // extension: CUDA device ordinal
        else if (a == "--bf") use_bf = true;   // ideal-filter numbering, main.rs:639-655
        else if (a == "--read-stats") read_stats = need(i);   // main.rs:939-1004
        else if (a == "--syncmers" || a == "--uhs" || a == "--lcp" || a == "--error-correct" ||
                 a == "--restart-from-postcor" || a == "--reference" || a == "--lmer-counts" ||
                 a == "--lmer-counts-min" || a == "--lmer-counts-max" || a == "-n" || a == "-t" || a == "-s" ||
                 a == "--distance" || a == "--correction-threshold")
            die("`" + a + "` selects a mode outside the reads->mdBG hot path; this build refuses it rather than "
                "falling back (SURVEY.md 2.1)");
        else if (!a.empty() && a[0] == '-') die("unknown flag " + a);
        else reads = a;
    }
    if (reads.empty()) die("Please specify an input file.");
    bool fasta = reads.find(".fasta.") != std::string::npos || reads.find(".fa.") != std::string::npos ||
                 (reads.size() >= 3 && reads.compare(reads.size() - 3, 3, ".fa") == 0) ||
                 (reads.size() >= 6 && reads.compare(reads.size() - 6, 6, ".fasta") == 0);   // main.rs:463
    if (fasta) { printf("Input file: %s\n", reads.c_str()); printf("Format: FASTA\n"); }
    long kk = 10, ll = 12, mab = 2;
    double dd = 0.10;
    float ps = 0.01f;
    if (k < 0 && l < 0 && density < 0) {   // main.rs:468-472, 214-226
        printf("Autodetecting values for k, l, and density.\n");
        printf("Parsing input sequences to estimate mean read length...\n");
        Fastx fx;
        if (!fx.open(reads.c_str(), fasta)) die("Error opening compressed file: " + reads);
        std::string s;
        unsigned long long tot = 0, n = 0;
        while (n < 100 && fx.next(s)) { tot += s.size(); n++; }
        fx.close();
        unsigned long long mean = n ? tot / n : 0;
        printf("Detected mean read length of %llu bp.\n", mean);
        dd = 0.003; kk = (long)(dd * (double)mean); ll = 12;
        printf("Setting k = %ld l = %ld density = %s.\n", kk, ll, rust_f64(dd).c_str());
    } else {
        if (k >= 0) kk = k; else printf("Warning: Using default k value (%ld).\n", kk);
        if (l >= 0) ll = l; else printf("Warning: Using default l value (%ld).\n", ll);
        if (density >= 0) dd = density; else printf("Warning: Using default density value (%s%%).\n", rust_f64(dd * 100.0).c_str());
    }
    if (minabund >= 0) mab = minabund; else printf("Warning: Using default minimum k-mer abundance value (%ld).\n", mab);
    if (presimp >= 0) ps = presimp; else printf("Warning: Using default pre-simp value (0.01).\n");
    if (threads < 0) printf("Warning: Using default number of threads (8).\n");
    if (prefix.empty()) {
        prefix = "graph-k" + std::to_string(kk) + "-d" + rust_f64(dd) + "-l" + std::to_string(ll);
        printf("Warning: Using default output prefix (%s).\n", prefix.c_str());
    }
    {   // main.rs:608-613
        glob_t g;
        if (glob((prefix + "*.sequences").c_str(), 0, nullptr, &g) == 0) {
            for (size_t i = 0; i < g.gl_pathc; i++) { printf("Removing old sequences file: %s.\n", g.gl_pathv[i]); remove(g.gl_pathv[i]); }
            globfree(&g);
        }
    }
    mdbg_params P;
    memset(&P, 0, sizeof P);
    P.k = (uint32_t)kk; P.l = (uint32_t)ll; P.density = dd; P.min_abundance = (uint32_t)mab; P.presimp = ps;
    P.hpc = skiphpc ? 0 : 1; P.device = device; P.bf = use_bf ? 1 : 0;
    mdbg_ctx* ctx = nullptr;
    if (mdbg_ctx_create(&P, &ctx) != MDBG_OK) die(mdbg_last_error(nullptr));

    printf("Parsing input sequences...\n");
    Fastx fx;
    if (!fx.open(reads.c_str(), fasta)) die("Error opening compressed file: " + reads);
    const size_t BATCH = 256u << 20;
    std::vector<uint8_t> all_bases;        // kept only for the .sequences slices
    std::vector<uint64_t> all_off{0};
    uint8_t* pin = nullptr;
    if (mdbg_host_alloc_pinned(BATCH + (64u << 20), (void**)&pin) != MDBG_OK) die("cudaMallocHost failed");
    std::vector<uint64_t> off{0};
    size_t fill = 0;
    unsigned long long nb_reads = 0;
    std::string s;
    auto flush = [&]() {
        if (off.size() == 1) return;
        if (mdbg_push_reads(ctx, pin, off.data(), off.size() - 1) != MDBG_OK) die(mdbg_last_error(ctx));
        if (!no_basespace) {
            size_t base = all_bases.size();
            all_bases.insert(all_bases.end(), pin, pin + fill);
            for (size_t i = 1; i < off.size(); i++) all_off.push_back(base + off[i]);
        }
        off.assign(1, 0);
        fill = 0;
    };
    while (fx.next(s)) {
        if (s.size() > BATCH + (64u << 20)) die("a record longer than the staging buffer (use a larger batch)");
        if (fill + s.size() > BATCH + (64u << 20) || fill >= BATCH) flush();
        memcpy(pin + fill, s.data(), s.size());
        fill += s.size();
        off.push_back(fill);
        nb_reads++;
    }
    flush();
    fx.close();
    fprintf(stderr, "Converted reads to k-min-mers.\n");
    printf("Number of reads: %llu\n", nb_reads);
    mdbg_graph g;
    if (mdbg_finish(ctx, no_basespace ? 0 : 1, &g) != MDBG_OK) die(mdbg_last_error(ctx));
    if (mab > 1) {
        printf("Number of nodes before abundance filter: %llu\n", (unsigned long long)g.n_distinct);
        printf("Number of nodes after abundance filter: %llu\n", (unsigned long long)g.n_nodes);
    } else printf("Number of mdBG nodes: %llu\n", (unsigned long long)g.n_nodes);
    if (!read_stats.empty()) {   // main.rs:939-1004: abundance of every k-min-mer of a second read set, then exit
        const std::string stats_path = read_stats + ".read_stats";               // read_stats.rs:26
        FILE* sf = fopen(stats_path.c_str(), "w");
        if (!sf) die("Couldn't create " + stats_path);
        printf("Stats module initialized.\n");
        printf("Parsing sequences from \"%s\"...\n", read_stats.c_str());
        const bool sfa = read_stats.find(".fasta.") != std::string::npos || read_stats.find(".fa.") != std::string::npos ||
                         (read_stats.size() >= 3 && read_stats.compare(read_stats.size() - 3, 3, ".fa") == 0) ||
                         (read_stats.size() >= 6 && read_stats.compare(read_stats.size() - 6, 6, ".fasta") == 0);
        printf(sfa ? "Format: FASTA\n" : "Format: FASTQ\n");
        Fastx sx;
        if (!sx.open(read_stats.c_str(), sfa)) die("Error opening compressed file: " + read_stats);
        std::vector<std::string> ids;
        std::vector<uint32_t> counts;
        std::vector<uint64_t> coff;
        auto flush_stats = [&]() {
            if (off.size() == 1) return;
            const uint64_t n = off.size() - 1;
            coff.assign(n + 1, 0);
            uint64_t need_n = 0;
            int rc = mdbg_read_stats(ctx, pin, off.data(), n, counts.data(), coff.data(), counts.size(), &need_n);
            if (rc == MDBG_ERR_CAPACITY) {
                counts.resize(need_n + need_n / 8 + 1024);
                rc = mdbg_read_stats(ctx, pin, off.data(), n, counts.data(), coff.data(), counts.size(), &need_n);
            }
            if (rc != MDBG_OK) die(mdbg_last_error(ctx));
            std::string line;
            for (uint64_t r = 0; r < n; r++) {                                    // read_stats.rs:53-64
                line = ids[r] + ": ";
                for (uint64_t j = coff[r]; j < coff[r + 1]; j++) { line += std::to_string(counts[j]); line += ' '; }
                line += '\n';
                fwrite(line.data(), 1, line.size(), sf);
            }
            off.assign(1, 0);
            fill = 0;
            ids.clear();
        };
        while (sx.next(s)) {
            if (s.size() > BATCH + (64u << 20)) die("a record longer than the staging buffer (use a larger batch)");
            if (fill + s.size() > BATCH + (64u << 20) || fill >= BATCH) flush_stats();
            memcpy(pin + fill, s.data(), s.size());
            fill += s.size();
            off.push_back(fill);
            ids.push_back(sx.id);
        }
        flush_stats();
        sx.close();
        fclose(sf);
        printf("Read stats written, exiting.\n");
        mdbg_graph_free(&g);
        mdbg_host_free_pinned(pin);
        mdbg_ctx_destroy(ctx);
        return 0;
    }
    if (mdbg_write_gfa(&g, (prefix + ".gfa").c_str()) != MDBG_OK) die("Couldn't create " + prefix + ".gfa");
    if (!no_basespace &&
        mdbg_write_sequences(&g, all_bases.data(), all_off.data(), (prefix + ".0.sequences").c_str(), 1) != MDBG_OK)
        die("Couldn't create file: " + prefix + ".0.sequences");
    printf("Number of mdBG edges: %llu\n", (unsigned long long)g.n_edges);
    if (ps > 0.0f) printf("Pre-simp = %s: %llu edges removed.\n", rust_f32(ps).c_str(), (unsigned long long)g.presimp_removed);
    mdbg_graph_free(&g);
    mdbg_host_free_pinned(pin);
    mdbg_ctx_destroy(ctx);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf("Total execution time: %.9gs\n", sec);
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    printf("Maximum RSS: %.7gGB\n", (double)ru.ru_maxrss * 1024.0 / 1e9);
    return 0;
}
