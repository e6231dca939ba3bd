// rust-mdbg (B200) -- command line front end with the reference's flags (src/main.rs:228-423) for
// the hot path: reads (FASTA/FASTQ, plain or .gz) -> {prefix}.gfa + {prefix}.0.sequences.
// Host ingest only: parallel parse (ingest.hpp) into two pinned batch buffers, one
// being filled while the other is inside mdbg_push_reads; all compute is in libmdbg_b200.so.  The reads are
// NOT kept in memory: the .sequences lines are cut in a second pass over the input (mdbg_seq_writer_*).
// Modes outside the hot path are refused, never silently run elsewhere.
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

#include <future>
#include <thread>

#include "../../include/mdbg.h"
#include "ingest.hpp"

namespace {

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

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Runs `consume(batch, first_read_index)` over every batch of `path`, in file order, while the next batch is
// being parsed into the other pinned buffer.  Returns the number of reads.
template <class F>
unsigned long long for_each_batch(const std::string& path, bool fasta, bool want_ids, int threads, uint8_t* pin[2],
                                  size_t cap, size_t target, F&& consume) {
    ingest::Reader rd(threads);
    if (!rd.open(path.c_str(), fasta, want_ids)) die("Error opening compressed file: " + path);
    ingest::Batch bt[2];
    for (int i = 0; i < 2; i++) { bt[i].bases = pin[i]; bt[i].cap = cap; }
    std::string err;
    unsigned long long n = 0;
    auto fill = [&](int i) { return rd.next_batch(bt[i], target, err); };
    std::future<bool> nxt = std::async(std::launch::async, fill, 0);
    for (int i = 0;; i ^= 1) {
        const bool have = nxt.get();
        if (!err.empty()) die(err);
        if (!have) break;
        nxt = std::async(std::launch::async, fill, i ^ 1);     // parse the next batch while this one is consumed
        consume(bt[i], n);
        n += bt[i].n_reads();
    }
    return n;
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
        unsigned long long tot = 0, n = 0;
        {
            ingest::Reader rd(1);
            if (!rd.open(reads.c_str(), fasta, false)) die("Error opening compressed file: " + reads);
            ingest::Batch b;
            std::vector<uint8_t> tmp(96u << 20);
            b.bases = tmp.data(); b.cap = tmp.size();
            std::string err;
            if (rd.next_batch(b, 8u << 20, err))
                for (; n < 100 && n < b.n_reads(); n++) tot += b.off[n + 1] - b.off[n];
        }
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
    const int n_threads = (int)std::max<long>(1, threads > 0 ? threads : (long)std::thread::hardware_concurrency());
    // text window per batch; the buffers also hold a 64 MB gz block + carry.  (Pinning memory costs ~0.4 ms per MB at
    // start-up: two 192 MB buffers instead of round 1's 320 MB one + a pageable copy of the whole read set.)
    const size_t BATCH = 128u << 20, CAP = 192u << 20;
    uint8_t* pin[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++)
        if (mdbg_host_alloc_pinned(CAP, (void**)&pin[i]) != MDBG_OK) die("cudaMallocHost failed");
    const double t_ingest = now_s();
    double t_push = 0;
    unsigned long long nb_bases = 0;
    unsigned long long nb_reads = for_each_batch(reads, fasta, false, n_threads, pin, CAP, BATCH,
        [&](ingest::Batch& b, unsigned long long) {
            const double t0 = now_s();
            if (mdbg_push_reads(ctx, b.bases, b.off.data(), b.n_reads()) != MDBG_OK) die(mdbg_last_error(ctx));
            t_push += now_s() - t0;
            nb_bases += b.fill;
        });
    const double t_ingest_end = now_s();
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
        std::vector<uint32_t> counts;
        std::vector<uint64_t> coff;
        for_each_batch(read_stats, sfa, true, n_threads, pin, CAP, BATCH, [&](ingest::Batch& b, unsigned long long) {
            const uint64_t n = b.n_reads();
            coff.assign(n + 1, 0);
            uint64_t need_n = 0;
            int rc = mdbg_read_stats(ctx, b.bases, b.off.data(), n, counts.data(), coff.data(), counts.size(), &need_n);
            if (rc == MDBG_ERR_CAPACITY) {
                counts.resize(need_n + need_n / 8 + 1024);
                rc = mdbg_read_stats(ctx, b.bases, b.off.data(), n, counts.data(), coff.data(), counts.size(), &need_n);
            }
            if (rc != MDBG_OK) die(mdbg_last_error(ctx));
            std::string line;
            for (uint64_t r = 0; r < n; r++) {                                    // read_stats.rs:53-64
                line = b.ids[r] + ": ";
                for (uint64_t j = coff[r]; j < coff[r + 1]; j++) { line += std::to_string(counts[j]); line += ' '; }
                line += '\n';
                fwrite(line.data(), 1, line.size(), sf);
            }
        });
        fclose(sf);
        printf("Read stats written, exiting.\n");
        mdbg_graph_free(&g);
        mdbg_host_free_pinned(pin[0]); mdbg_host_free_pinned(pin[1]);
        mdbg_ctx_destroy(ctx);
        return 0;
    }
    const double t_graph = now_s();
    if (mdbg_write_gfa(&g, (prefix + ".gfa").c_str()) != MDBG_OK) die("Couldn't create " + prefix + ".gfa");
    const double t_gfa = now_s();
    // A large node set is written by several writers at once, one {prefix}.{t}.sequences per writer like the reference's
    // one file per worker thread (main.rs:614-630; to_basespace globs them): every batch of the second pass is walked by
    // all writers in parallel, each cutting the lines it owns.  Small outputs keep the single {prefix}.0.sequences.
    const int n_seq_writers = (g.n_seqlines >= 20000 && n_threads > 1) ? (int)std::min<long>(n_threads, 8) : 1;
    if (!no_basespace && n_seq_writers > 1) {
        std::vector<mdbg_seq_writer*> sw(n_seq_writers, nullptr);
        std::vector<std::string> sp(n_seq_writers);
        for (int t = 0; t < n_seq_writers; t++) {
            sp[t] = prefix + "." + std::to_string(t) + ".sequences";
            if (mdbg_seq_writer_open_part(&g, sp[t].c_str(), 1, (uint32_t)t, (uint32_t)n_seq_writers, &sw[t]) != MDBG_OK)
                die("Couldn't create file: " + sp[t]);
        }
        std::vector<std::string> werr(n_seq_writers);
        for_each_batch(reads, fasta, false, n_threads, pin, CAP, BATCH, [&](ingest::Batch& b, unsigned long long first) {
            std::vector<std::thread> th;
            for (int t = 0; t < n_seq_writers; t++)
                th.emplace_back([&, t] {
                    for (uint64_t r; werr[t].empty() && (r = mdbg_seq_writer_next_read(sw[t])) != UINT64_MAX && r < first + b.n_reads();) {
                        if (r < first) { werr[t] = "internal: .sequences lines out of read order"; break; }
                        const uint64_t i = r - first;
                        if (mdbg_seq_writer_read(sw[t], r, b.bases + b.off[i], b.off[i + 1] - b.off[i]) != MDBG_OK)
                            werr[t] = "Couldn't write file: " + sp[t];
                    }
                });
            for (auto& x : th) x.join();
            for (const std::string& e : werr) if (!e.empty()) die(e);
        });
        for (int t = 0; t < n_seq_writers; t++)
            if (mdbg_seq_writer_close(sw[t]) != MDBG_OK) die("Couldn't write file: " + sp[t]);
    } else if (!no_basespace) {   // second pass over the input: cut the .sequences lines (main.rs:696-707) without holding the reads
        mdbg_seq_writer* sw = nullptr;
        const std::string sp = prefix + ".0.sequences";
        if (mdbg_seq_writer_open(&g, sp.c_str(), 1, &sw) != MDBG_OK) die("Couldn't create file: " + sp);
        if (mdbg_seq_writer_next_read(sw) != UINT64_MAX)
            for_each_batch(reads, fasta, false, n_threads, pin, CAP, BATCH, [&](ingest::Batch& b, unsigned long long first) {
                for (uint64_t r; (r = mdbg_seq_writer_next_read(sw)) != UINT64_MAX && r < first + b.n_reads();) {
                    if (r < first) die("internal: .sequences lines out of read order");
                    const uint64_t i = r - first;
                    if (mdbg_seq_writer_read(sw, r, b.bases + b.off[i], b.off[i + 1] - b.off[i]) != MDBG_OK)
                        die("Couldn't write file: " + sp);
                }
            });
        if (mdbg_seq_writer_close(sw) != MDBG_OK) die("Couldn't write file: " + sp);
    }
    const double t_seq = now_s();
    if (getenv("MDBG_CLI_TIMING"))
        fprintf(stderr, "[timing] ingest+K-A %.3f s (%.3f inside mdbg_push_reads, %.2f Gbases/s over %llu bases, %d parser threads), "
                "finish %.3f s, .gfa %.3f s, .sequences %.3f s\n", t_ingest_end - t_ingest, t_push,
                (double)nb_bases / (t_ingest_end - t_ingest) / 1e9, nb_bases, n_threads, t_graph - t_ingest_end, t_gfa - t_graph,
                t_seq - t_gfa);
    printf("Number of mdBG edges: %llu\n", (unsigned long long)g.n_edges);
    if (ps > 0.0f) printf("Pre-simp = %s: %llu edges removed.\n", rust_f32(ps).c_str(), (unsigned long long)g.presimp_removed);
    mdbg_graph_free(&g);
    mdbg_host_free_pinned(pin[0]); mdbg_host_free_pinned(pin[1]);
    mdbg_ctx_destroy(ctx);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    printf("Total execution time: %.9gs\n", sec);
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    printf("Maximum RSS: %.7gGB\n", (double)ru.ru_maxrss * 1024.0 / 1e9);
    return 0;
}
