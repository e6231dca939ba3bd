// gunzip_check.cpp -- TEST INFRASTRUCTURE: drives rust-mdbg_b200/cli/gz_inflate.hpp (the front end's gzip reader)
// on a file: writes the decompressed bytes to stdout, or prints "length fnv1a" with --hash; `--time` prints MB/s.
//   gunzip_check FILE [--hash|--time|--raw] [READ_CHUNK_BYTES] [THREADS SPAN_BYTES]
//   THREADS > 0: the parallel single-stream decoder (GzParallel).  Exit code 1 + message on stderr on a bad stream.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../rust-mdbg_b200/cli/gz_inflate.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const bool hash = argc > 2 && !strcmp(argv[2], "--hash"), timeit = argc > 2 && !strcmp(argv[2], "--time");
    const size_t chunk = argc > 3 ? (size_t)atoll(argv[3]) : (1u << 20);   // (argv[2] = --raw: plain output with this chunk)
    int fd = open(argv[1], O_RDONLY);
    if (fd < 0) return 2;
    struct stat sb;
    fstat(fd, &sb);
    const uint8_t* in = sb.st_size ? (const uint8_t*)mmap(nullptr, sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : (const uint8_t*)"";
    const int threads = argc > 5 ? atoi(argv[4]) : 0;
    const size_t span = argc > 5 ? (size_t)atoll(argv[5]) : 0;
    ingest::GzInflate z;
    ingest::GzParallel zp;
    if (threads > 0) zp.reset(in, (size_t)sb.st_size, threads, span);
    else z.reset(in, (size_t)sb.st_size);
    std::vector<uint8_t> buf(chunk);
    uint64_t h = 1469598103934665603ull, total = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const size_t n = threads > 0 ? zp.read(buf.data(), buf.size()) : z.read(buf.data(), buf.size());
        if (n == 0) break;
        total += n;
        if (hash) for (size_t i = 0; i < n; i++) { h ^= buf[i]; h *= 1099511628211ull; }
        else if (!timeit) fwrite(buf.data(), 1, n, stdout);
    }
    if (threads > 0) {
        if (!zp.error().empty()) { fprintf(stderr, "%s\n", zp.error().c_str()); return 1; }
    } else {
        if (!z.error().empty()) { fprintf(stderr, "%s\n", z.error().c_str()); return 1; }
        if (!z.eof()) { fprintf(stderr, "stopped before the end of the file\n"); return 1; }
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (hash) printf("%llu %016llx\n", (unsigned long long)total, (unsigned long long)h);
    if (timeit) printf("%llu bytes in %.3f s = %.1f MB/s\n", (unsigned long long)total, dt, total / dt / 1e6);
    return 0;
}
