// synth.cu -- synthetic HiFi-shape reads for the benchmark workloads.
// Counter-based (splitmix64 of (seed, index)), so the device generator, the host generator and
// every rank produce identical bytes without shipping gigabytes: genome base i, read r's
// length / start / strand and the substitution at (r, j) are pure functions of the seed.
// Workload tooling, not part of the timed path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#include "ctx.h"

namespace {

__host__ __device__ __forceinline__ uint64_t sm64(uint64_t x) {
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

struct Seeds { uint64_t g, r, s, t, e; };
__host__ __device__ inline Seeds seeds_of(uint64_t master) {
    Seeds k;
    k.g = sm64(master + 1); k.r = sm64(master + 2); k.s = sm64(master + 3);
    k.t = sm64(master + 4); k.e = sm64(master + 5);
    return k;
}

struct ReadGeom { uint64_t start; uint32_t strand; };
__host__ __device__ inline ReadGeom geom_of(const Seeds& k, uint64_t glen, uint64_t r, uint64_t len) {
    ReadGeom g;
    uint64_t span = glen >= len ? glen - len + 1 : 1;
    g.start = sm64(k.s ^ r) % span;
    g.strand = (uint32_t)(sm64(k.t ^ r) & 1);
    return g;
}

// base j (0-based, read orientation) of read r
__host__ __device__ inline uint8_t base_of(const Seeds& k, uint64_t glen, uint64_t err_thresh, uint64_t r,
                                           uint64_t len, const ReadGeom& g, uint64_t j) {
    uint64_t gi = g.strand ? g.start + (len - 1 - j) : g.start + j;
    if (gi >= glen) gi %= glen;
    uint32_t b = (uint32_t)(sm64(k.g ^ gi) & 3);   // 0 A, 1 C, 2 G, 3 T
    if (g.strand) b = 3 - b;                       // complement
    uint64_t e = sm64(k.e + r * 0x9e3779b97f4a7c15ULL + j);
    if (e < err_thresh) b = (b + 1 + (uint32_t)((e >> 7) % 3)) & 3;   // substitution
    return (uint8_t)("ACGT"[b]);
}

uint64_t read_len(const mdbg_synth* s, const Seeds& k, uint64_t r) {
    // Box-Muller on the host only (lengths travel to the device inside read_off)
    double u1 = ((double)(sm64(k.r ^ (2 * r)) >> 11) + 1.0) / 9007199254740993.0;
    double u2 = (double)(sm64(k.r ^ (2 * r + 1)) >> 11) / 9007199254740992.0;
    double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    double v = std::floor(s->mean_len + s->sd_len * z + 0.5);
    uint64_t lo = s->min_len, hi = std::min<uint64_t>(s->max_len, s->genome_len);
    if (v < (double)lo) return lo;
    if (v > (double)hi) return hi;
    return (uint64_t)v;
}

uint64_t err_threshold(double rate) {
    double x = rate * 18446744073709551616.0;
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)x;
}

__global__ void synth_fill_kernel(Seeds k, uint64_t glen, uint64_t err_thresh, uint64_t first_read,
                                  uint64_t n_reads, const uint64_t* __restrict__ read_off, uint64_t total,
                                  uint8_t* __restrict__ out) {
    uint64_t chunk = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t p = chunk * 16;
    if (p >= total) return;
    uint64_t lo = 0, hi = n_reads;   // last r with read_off[r] <= p
    while (hi - lo > 1) { uint64_t m = (lo + hi) >> 1; if (read_off[m] <= p) lo = m; else hi = m; }
    uint64_t r = lo;
    uint64_t rs = read_off[r], re = read_off[r + 1];
    ReadGeom g = geom_of(k, glen, first_read + r, re - rs);
    uint8_t buf[16];
    int n = 0;
    for (; n < 16 && p + n < total; n++) {
        uint64_t q = p + n;
        while (q >= re) { r++; rs = read_off[r]; re = read_off[r + 1]; g = geom_of(k, glen, first_read + r, re - rs); }
        buf[n] = base_of(k, glen, err_thresh, first_read + r, re - rs, g, q - rs);
    }
    if (n == 16) *reinterpret_cast<uint4*>(out + p) = *reinterpret_cast<uint4*>(buf);
    else for (int i = 0; i < n; i++) out[p + i] = buf[i];
}

}  // namespace

extern "C" {

uint64_t mdbg_synth_num_reads(const mdbg_synth* s, double coverage) {
    return (uint64_t)std::ceil(coverage * (double)s->genome_len / s->mean_len);
}

uint64_t mdbg_synth_plan(const mdbg_synth* s, uint64_t first_read, uint64_t n_reads, uint64_t* read_off,
                         uint64_t* start, uint8_t* strand) {
    Seeds k = seeds_of(s->seed);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < n_reads; i++) {
        uint64_t len = read_len(s, k, first_read + i);
        if (read_off) read_off[i] = acc;
        if (start || strand) {
            ReadGeom g = geom_of(k, s->genome_len, first_read + i, len);
            if (start) start[i] = g.start;
            if (strand) strand[i] = (uint8_t)g.strand;
        }
        acc += len;
    }
    if (read_off) read_off[n_reads] = acc;
    return acc;
}

void mdbg_synth_fill_host(const mdbg_synth* s, uint64_t first_read, uint64_t n_reads, const uint64_t* read_off,
                          uint8_t* bases, int threads) {
    Seeds k = seeds_of(s->seed);
    uint64_t et = err_threshold(s->error_rate);
    if (threads < 1) threads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t]() {
            for (uint64_t i = t; i < n_reads; i += threads) {
                uint64_t len = read_off[i + 1] - read_off[i];
                ReadGeom g = geom_of(k, s->genome_len, first_read + i, len);
                uint8_t* o = bases + read_off[i];
                for (uint64_t j = 0; j < len; j++) o[j] = base_of(k, s->genome_len, et, first_read + i, len, g, j);
            }
        });
    for (auto& x : th) x.join();
}

int mdbg_synth_fill_device(mdbg_ctx* c, const mdbg_synth* s, uint64_t first_read, uint64_t n_reads,
                           const uint64_t* read_off, uint8_t* d_bases, uint64_t* d_read_off) {
    if (!c || !s || !read_off) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    uint64_t total = read_off[n_reads];
    MDBG_CK(c, cudaMemcpyAsync(d_read_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->st));
    if (total) {
        uint64_t chunks = (total + 15) / 16;
        synth_fill_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, c->st>>>(
            seeds_of(s->seed), s->genome_len, err_threshold(s->error_rate), first_read, n_reads, d_read_off, total, d_bases);
        MDBG_CK(c, cudaGetLastError());
    }
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}

}  // extern "C"
