// ingest_check.cpp -- TEST INFRASTRUCTURE: drives rust-mdbg_b200/cli/ingest.hpp (the front end's parallel
// FASTA/FASTQ reader) without a GPU: prints "id<TAB>length<TAB>fnv1a(sequence)" per read, in delivery order.
//   ingest_check FILE fasta|fastq THREADS TARGET_BYTES [--count]      (--count: totals only, for timing the reader)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../rust-mdbg_b200/cli/ingest.hpp"

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const bool fasta = std::string(argv[2]) == "fasta";
    ingest::Reader rd(atoi(argv[3]));
    if (!rd.open(argv[1], fasta, true)) { fprintf(stderr, "open failed\n"); return 1; }
    std::vector<uint8_t> buf(512u << 20);
    ingest::Batch b;
    b.bases = buf.data(); b.cap = buf.size();
    std::string err;
    const bool count_only = argc > 5 && std::string(argv[5]) == "--count";
    unsigned long long n_reads = 0, n_bases = 0;
    while (rd.next_batch(b, (size_t)atoll(argv[4]), err)) {
        n_reads += b.n_reads(); n_bases += b.fill;
        if (count_only) continue;
        for (uint64_t r = 0; r < b.n_reads(); r++) {
            uint64_t h = 1469598103934665603ull;
            for (uint64_t i = b.off[r]; i < b.off[r + 1]; i++) { h ^= b.bases[i]; h *= 1099511628211ull; }
            printf("%s\t%llu\t%016llx\n", b.ids[r].c_str(), (unsigned long long)(b.off[r + 1] - b.off[r]), (unsigned long long)h);
        }
    }
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (count_only) printf("%llu reads %llu bases\n", n_reads, n_bases);
    return 0;
}
