// api.cu -- context lifecycle, Entry 1 (Read::extract) and the push half of Entry 3 of the
// C ABI declared in include/mdbg.h.  Host side only orchestrates: every byte of the hot
// path is processed by the kernels in ka_*.cu .. ke_*.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include <chrono>
#include <functional>
#include <thread>

#include "ctx.h"
#include "pack_host.h"

using namespace mdbg;

static std::string g_create_err;

// ---- Pool -----------------------------------------------------------------------------------
cudaError_t Pool::alloc(size_t bytes, void** out, size_t* cap_out) {
    bytes = (bytes + 511) & ~(size_t)511;
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= bytes * 2 + (1u << 20)) {
        *out = it->second;
        *cap_out = it->first;
        free_.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {  // release the cache and retry once
        cudaGetLastError();
        trim();
        e = cudaMalloc(out, bytes);
    }
    *cap_out = bytes;
    return e;
}
void Pool::release(void* p, size_t cap) { free_.emplace(cap, p); }
void Pool::trim() {
    for (auto& kv : free_) cudaFree(kv.second);
    free_.clear();
}

// ---- small kernels ---------------------------------------------------------------------------
__global__ void widen_pos_kernel(const uint32_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t n) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
// read_off of a device-resident batch: non-decreasing, ends at n_bases, no record of 4 Gbases or more
// (positions inside a read are u32 on the device); flag bit 0 = order, bit 1 = a record too long
__global__ void check_read_off_kernel(const uint64_t* __restrict__ off, uint64_t n_reads, uint64_t n_bases,
                                      unsigned long long* flags) {
    uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    unsigned long long f = 0;
    if (r == n_reads) { if (off[r] != n_bases) f |= 1; }
    else {
        if (off[r + 1] < off[r]) f |= 1;
        else if (off[r + 1] - off[r] >= 0xFFFFFFF0ull) f |= 2;
    }
    if (r == 0 && off[0] != 0) f |= 1;
    if (f) atomicOr(flags, f);
}
__global__ void flush_kernel(uint4* p, uint64_t n, uint32_t v) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_uint4(v, v, v, v);
}

extern "C" {

const char* mdbg_version(void) { return "mdbg-b200 0.1 (sm_100a)"; }

int mdbg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

uint64_t mdbg_hash_bound(double density) { return hash_bound(density); }

const char* mdbg_last_error(const mdbg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mdbg_ctx_create(const mdbg_params* p, mdbg_ctx** out) {
    if (!p || !out) { g_create_err = "null argument"; return MDBG_ERR_BAD_ARG; }
    *out = nullptr;
    if (p->k < 2 || p->l < 2 || p->l > 64 || p->min_abundance < 1 || p->min_abundance > 65535 ||
        !(p->density > 0.0)) {
        g_create_err = "bad parameters: need k >= 2, 2 <= l <= 64, 1 <= min_abundance <= 65535, density > 0";
        return MDBG_ERR_BAD_ARG;
    }
    int n = mdbg_device_count();
    if (n <= 0) {
        g_create_err = "no CUDA device: libmdbg_b200 has no CPU fallback";
        return MDBG_ERR_NO_DEVICE;
    }
    if (p->device < 0 || p->device >= n) { g_create_err = "bad device ordinal"; return MDBG_ERR_BAD_ARG; }
    mdbg_ctx* c = new mdbg_ctx();
    c->p = *p;
    c->device = p->device;
    c->bound = hash_bound(p->density);
    c->fc = make_filter(p->l, c->bound);
    cudaError_t e = cudaSetDevice(c->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, c->device);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_sc, sizeof(Scalars));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_sc, sizeof(Scalars));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_mail, MAIL_WORDS * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_mail, MAIL_WORDS * sizeof(uint64_t));
    for (int i = 0; i < 24 && e == cudaSuccess; i++) e = cudaEventCreate(&c->ev[i]);
    for (int i = 0; i < 16 && e == cudaSuccess; i++) e = cudaEventCreate(&c->evk[i]);
    if (e != cudaSuccess) {
        g_create_err = std::string("CUDA init failed: ") + cudaGetErrorString(e);
        delete c;
        return MDBG_ERR_CUDA;
    }
    c->num_sms = prop.multiProcessorCount;
    int per_sm = ka_max_blocks_per_sm(p->hpc);
    if (per_sm < 1) per_sm = 1;
    c->ka_grid = c->num_sms * per_sm;
    // K-A variant: the bit-sliced kernel is taken only where it is instantiated (l, density); the
    // classic kernel stays the exact path for everything else and for the tiles it hands over.
    uint32_t variant = p->ka_variant;
    if (variant == 0) {
        const char* e = getenv("MDBG_KA_VARIANT");
        variant = (e && !strcmp(e, "classic")) ? 1 : 2;
    }
    if (variant > 2) { g_create_err = "bad ka_variant"; mdbg_ctx_destroy(c); return MDBG_ERR_BAD_ARG; }
    c->ka_bs = (variant == 2) && ka_bs_supported(p->l, c->bound);
    if (c->ka_bs) {
        int bs_per_sm = ka_bs_max_blocks_per_sm(p->l, p->hpc);
        if (bs_per_sm < 1) { g_create_err = "bit-sliced K-A kernel does not fit this device"; mdbg_ctx_destroy(c); return MDBG_ERR_CUDA; }
        c->ka_bs_grid = c->num_sms * bs_per_sm;
        std::vector<unsigned char> tab(KA_BS_TABLE_BYTES);
        ka_bs_tables(tab.data());
        e = cudaMalloc(&c->ka_bs_t4, KA_BS_TABLE_BYTES);
        if (e == cudaSuccess) e = cudaMemcpy(c->ka_bs_t4, tab.data(), KA_BS_TABLE_BYTES, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            g_create_err = std::string("CUDA init failed: ") + cudaGetErrorString(e);
            mdbg_ctx_destroy(c);
            return MDBG_ERR_CUDA;
        }
    }
    {   // host buffers travel 4:1 packed (pack_host.cc) unless MDBG_UPLOAD=ascii
        const char* e = getenv("MDBG_UPLOAD");      // ascii | packed | hybrid (default)
        c->upload_packed = !(e && !strcmp(e, "ascii"));
        c->upload_hybrid = !(e && !strcmp(e, "packed"));
    }
    memset(&c->tm, 0, sizeof(c->tm));
    *out = c;
    return MDBG_OK;
}

void mdbg_graph_device_free(mdbg_ctx* ctx);  // graph.cu
void mdbg_comm_release(mdbg_ctx* ctx);       // comm.cu

void mdbg_ctx_destroy(mdbg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    mdbg_graph_device_free(c);
    if (c->m_hash) cudaFree(c->m_hash);
    if (c->m_pos) cudaFree(c->m_pos);
    if (c->m_off) cudaFree(c->m_off);
    if (c->l2_flush) cudaFree(c->l2_flush);
    if (c->ka_bs_t4) cudaFree(c->ka_bs_t4);
    if (c->h_planes) cudaFreeHost(c->h_planes);
    delete c->pack_pool;
    if (c->d_sc) cudaFree(c->d_sc);
    if (c->h_sc) cudaFreeHost(c->h_sc);
    if (c->d_mail) cudaFree(c->d_mail);
    if (c->h_mail) cudaFreeHost(c->h_mail);
    mdbg_comm_release(c);
    for (int i = 0; i < 24; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 16; i++) if (c->evk[i]) cudaEventDestroy(c->evk[i]);
    c->pool.trim();
    for (auto ev : c->copy_ev) cudaEventDestroy(ev);
    for (auto& pb : c->pinned_cache) cudaFreeHost(pb.second);
    if (c->st_copy) cudaStreamDestroy(c->st_copy);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

int mdbg_ctx_set_k(mdbg_ctx* c, uint32_t k, uint32_t min_abundance, float presimp) {
    if (!c || k < 2 || min_abundance < 1 || min_abundance > 65535) {
        if (c) c->err = "bad k / min_abundance";
        return MDBG_ERR_BAD_ARG;
    }
    c->p.k = k;
    c->p.min_abundance = min_abundance;
    c->p.presimp = presimp;
    return MDBG_OK;
}

void* mdbg_stream(mdbg_ctx* c) { return c ? (void*)c->st : nullptr; }

int mdbg_reset(mdbg_ctx* c) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    c->M = c->R = c->n_bases = 0;
    c->tm.ka_launches = 0;
    c->tm.ka_ms_sum = 0;
    mdbg_graph_device_free(c);
    return MDBG_OK;
}

}  // extern "C"

// grow a persistent arena array, preserving `keep` bytes
static int grow(mdbg_ctx* c, void** p, size_t* cap, size_t need, size_t keep) {
    if (*cap >= need) return MDBG_OK;
    size_t ncap = std::max(need, *cap + *cap / 2);
    ncap = (ncap + 4095) & ~(size_t)4095;
    void* q = nullptr;
    MDBG_CK(c, cudaMalloc(&q, ncap));
    if (*p) {
        if (keep) MDBG_CK(c, cudaMemcpyAsync(q, *p, keep, cudaMemcpyDeviceToDevice, c->st));
        MDBG_CK(c, cudaStreamSynchronize(c->st));
        MDBG_CK(c, cudaFree(*p));
    }
    *p = q;
    *cap = ncap;
    return MDBG_OK;
}

static int ensure_arena(mdbg_ctx* c, uint64_t m_items, uint64_t r_items) {
    int rc;
    if ((rc = grow(c, (void**)&c->m_hash, &c->m_hash_cap, m_items * 8, c->M * 8))) return rc;
    if ((rc = grow(c, (void**)&c->m_pos, &c->m_pos_cap, m_items * 4, c->M * 4))) return rc;
    if ((rc = grow(c, (void**)&c->m_off, &c->m_off_cap, (r_items + 1) * 8, (c->R + 1) * 8))) return rc;
    c->m_cap_items = std::min(c->m_hash_cap / 8, c->m_pos_cap / 4);
    return MDBG_OK;
}

constexpr uint64_t UPLOAD_SLOTS = 8;   // chunk-sized pinned staging slots of the packed upload

// One launch of K-A covers tiles [prev tile_end, tile_end) and may first wait for an upload event.
struct KaChunk { uint64_t tile_end; cudaEvent_t wait; };

// Run K-A on a device-resident batch (or on one that is still being uploaded chunk by chunk),
// appending to the arena.
static int run_ka(mdbg_ctx* c, const uint8_t* d_bases, const uint64_t* d_read_off, uint64_t R, uint64_t B,
                  const std::vector<KaChunk>* plan = nullptr,
                  const std::function<int(size_t)>* prepare = nullptr) {   // prepare(i): make chunk i's bytes arrive
    if (((uintptr_t)d_bases & 15) != 0) { c->err = "bases must be 16-byte aligned"; return MDBG_ERR_BAD_ARG; }
    // expected minimizers: ~2.2*density of the HPC positions; start with a generous estimate and
    // re-run the batch once with the exact size if it did not fit.
    double rho = std::min(1.0, 2.6 * c->p.density + 1e-4);
    uint64_t est = (uint64_t)((double)B * rho) + 4096;
    int rc;
    bool fresh_arena = false;
    if (c->M == 0 && c->R == 0 && c->m_off == nullptr) {
        if ((rc = ensure_arena(c, est, R))) return rc;
        fresh_arena = true;   // m_off[0] = 0, written by ka_prepare's kernel
    } else if ((rc = ensure_arena(c, c->M + est, c->R + R))) return rc;

    uint64_t n_tiles = std::max<uint64_t>(1, (B + KA_TILE - 1) / KA_TILE);
    Tmp<uint64_t> tile_cnt, tile_soff, tile_excl, tile_lb, stage_hash;
    Tmp<uint32_t> stage_pos, chunk_cnt, dirty_list;
    Tmp<uint8_t> scan_tmp;
    MDBG_CK(c, tile_cnt.get(c->pool, n_tiles));
    MDBG_CK(c, tile_soff.get(c->pool, n_tiles));
    MDBG_CK(c, tile_excl.get(c->pool, n_tiles));
    MDBG_CK(c, tile_lb.get(c->pool, n_tiles + 1));
    const bool bs = c->ka_bs && n_tiles < 0xFFFFFFFFull;
    if (bs) MDBG_CK(c, dirty_list.get(c->pool, n_tiles));
    size_t scan_bytes = 0;
    MDBG_CK(c, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, tile_cnt.p, tile_excl.p, n_tiles, c->st));
    MDBG_CK(c, scan_tmp.get(c->pool, scan_bytes + 256));

    for (int attempt = 0; attempt < 2; attempt++) {
        const uint64_t stage_cap = c->m_cap_items - c->M;
        MDBG_CK(c, stage_hash.get(c->pool, stage_cap));
        MDBG_CK(c, stage_pos.get(c->pool, stage_cap));
        // All device-side state of the batch is reset by ka_prepare's KERNEL: a cudaMemsetAsync or
        // an H2D copy here may be served by the copy engine and then queues behind the bulk upload
        // of mdbg_push_reads, which serialises K-A after the copy instead of under it.
        const size_t n_launch = (plan && attempt == 0) ? plan->size() : 1;
        // counters: one tile/group counter per K-A launch, then (bit-sliced) the length of the dirty
        // list and the tile counter of the classic kernel's pass over that list
        MDBG_CK(c, chunk_cnt.get(c->pool, n_launch + 2));
        KAInit I{&c->d_sc->total_out, &c->d_sc->err_pos, &c->d_sc->stage_counter, &c->d_sc->dense_tiles,
                 chunk_cnt.p, (uint32_t)n_launch + 2, fresh_arena ? c->m_off : nullptr};
        KAArgs A{};
        A.bases = d_bases; A.read_off = d_read_off; A.n_reads = R; A.n_bases = B;
        A.l = c->p.l; A.bound = c->bound; A.fc = c->fc; A.force_dense = 0;
        A.out_hash = c->m_hash; A.out_pos = c->m_pos; A.out_read_off = c->m_off;
        A.out_base = c->M; A.out_cap = c->m_cap_items; A.read_base = c->R;
        A.total_out = &c->d_sc->total_out; A.err_pos = &c->d_sc->err_pos;
        A.dense_tiles = &c->d_sc->dense_tiles; A.tile_counter = nullptr;
        A.stage_hash = stage_hash; A.stage_pos = stage_pos; A.stage_cap = stage_cap;
        A.stage_counter = &c->d_sc->stage_counter; A.tile_cnt = tile_cnt; A.tile_soff = tile_soff;
        A.tile_lb = tile_lb; A.n_tiles = n_tiles;
        A.dirty_out = &c->d_sc->v[11];
        if (bs) {
            A.dirty_list = dirty_list; A.dirty_n = chunk_cnt.p + n_launch;
            A.bs_t4 = reinterpret_cast<const bs::T4Entry*>(c->ka_bs_t4);
            // tiles per claim: enough groups for ~8 claims per resident warp, at most 8 tiles
            uint64_t warps = (uint64_t)c->ka_bs_grid * (KA_THREADS / 32);
            uint64_t g = n_tiles / (warps * 8 + 1);
            A.bs_group = (uint32_t)std::min<uint64_t>(8, std::max<uint64_t>(1, g));
            if (const char* e = getenv("MDBG_BS_GROUP")) { long v = atol(e); if (v >= 1 && v <= 64) A.bs_group = (uint32_t)v; }
        }
        Tmp<uint32_t> dbg;
        const char* dbg_path = bs ? getenv("MDBG_BS_DEBUG_DUMP") : nullptr;
        if (dbg_path) {
            MDBG_CK(c, dbg.get(c->pool, n_tiles * 8));
            MDBG_CK(c, cudaMemsetAsync(dbg.p, 0xFF, n_tiles * 8 * sizeof(uint32_t), c->st));
            A.dbg = dbg.p;
        }
        MDBG_CK(c, cudaEventRecord(c->ev[0], c->st));
        MDBG_CK(c, ka_prepare(A, I, c->st, &c->tm.launches_push));
        if (plan && attempt == 0) {
            uint64_t tb = 0;
            size_t li = 0;
            for (const KaChunk& ch : *plan) {
                if (prepare) { int prc = (*prepare)(li); if (prc) return prc; }
                if (ch.wait) MDBG_CK(c, cudaStreamWaitEvent(c->st, ch.wait, 0));
                A.tile_begin = tb; A.tile_end = ch.tile_end; A.tile_counter = chunk_cnt.p + li++;
                if (bs) MDBG_CK(c, ka_bs_launch(A, c->p.hpc, c->ka_bs_grid, c->st, &c->tm.launches_push));
                else MDBG_CK(c, ka_launch(A, c->p.hpc, c->ka_grid, c->st, &c->tm.launches_push));
                tb = ch.tile_end;
            }
        } else {
            A.tile_begin = 0; A.tile_end = n_tiles; A.tile_counter = chunk_cnt.p;
            if (bs) MDBG_CK(c, ka_bs_launch(A, c->p.hpc, c->ka_bs_grid, c->st, &c->tm.launches_push));
            else MDBG_CK(c, ka_launch(A, c->p.hpc, c->ka_grid, c->st, &c->tm.launches_push));
        }
        if (bs) {   // the tiles the bit-sliced kernel handed over (N, long homopolymers, ...): exact kernel
            A.tile_list = dirty_list; A.tile_list_n = chunk_cnt.p + n_launch; A.tile_counter = chunk_cnt.p + n_launch + 1;
            int lg = (int)std::min<uint64_t>((uint64_t)c->ka_grid, (n_tiles + (KA_THREADS / 32) - 1) / (KA_THREADS / 32));
            MDBG_CK(c, ka_launch_list(A, c->p.hpc, lg, c->st, &c->tm.launches_push));
            A.tile_list = nullptr; A.tile_list_n = nullptr;
        }
        MDBG_CK(c, cudaEventRecord(c->ev[16], c->st));
        MDBG_CK(c, cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, tile_cnt.p, tile_excl.p, n_tiles, c->st));
        c->tm.launches_push += 2;
        MDBG_CK(c, cudaEventRecord(c->evk[10], c->st));
        MDBG_CK(c, ka_finalize(A, tile_excl, c->st, &c->tm.launches_push));
        MDBG_CK(c, cudaEventRecord(c->ev[1], c->st));
        MDBG_CK(c, cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->st));
        MDBG_CK(c, cudaStreamSynchronize(c->st));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        c->tm.ms_ka = ms;
        cudaEventElapsedTime(&c->tm.ms_ka_kernel, c->ev[0], c->ev[16]);
        cudaEventElapsedTime(&c->tm.ms_kernels[5], c->evk[10], c->ev[1]);
        c->tm.ka_ms_sum += ms;
        c->tm.ka_launches += 1;
        c->tm.ka_dense_tiles = c->h_sc->dense_tiles;
        if (dbg_path) {   // test hook: per-tile state of the bit-sliced kernel + its raw tile counts
            std::vector<uint32_t> hd(n_tiles * 8);
            std::vector<uint64_t> hc(n_tiles);
            MDBG_CK(c, cudaMemcpy(hd.data(), dbg.p, hd.size() * 4, cudaMemcpyDeviceToHost));
            MDBG_CK(c, cudaMemcpy(hc.data(), tile_cnt.p, hc.size() * 8, cudaMemcpyDeviceToHost));
            if (FILE* f = fopen(dbg_path, "wb")) {
                fwrite(&n_tiles, 8, 1, f);
                fwrite(hd.data(), 4, hd.size(), f);
                fwrite(hc.data(), 8, hc.size(), f);
                fclose(f);
            }
        }
        c->tm.ka_variant_used = bs ? 2 : 1;
        c->tm.ka_dirty_tiles = (uint32_t)c->h_sc->v[11];
        if (c->h_sc->err_pos != ~0ull) {
            char buf[160];
            snprintf(buf, sizeof buf, "Non-ACGTN nucleotide encountered! (batch byte offset %llu)", c->h_sc->err_pos);
            c->err = buf;
            return MDBG_ERR_ALPHABET;
        }
        uint64_t total = c->h_sc->total_out;
        if (total <= c->m_cap_items) {
            c->M = total;
            c->R += R;
            c->n_bases += B;
            return MDBG_OK;
        }
        if ((rc = ensure_arena(c, total + 1024, c->R + R))) return rc;  // exact size, run again
    }
    c->err = "minimizer arena overflow after resize";
    return MDBG_ERR_CAPACITY;
}

extern "C" {

int mdbg_push_reads_device(mdbg_ctx* c, const uint8_t* d_bases, const uint64_t* d_read_off,
                           uint64_t n_reads, uint64_t n_bases) {
    if (!c || !d_read_off || (n_bases && !d_bases)) { if (c) c->err = "null argument"; return MDBG_ERR_BAD_ARG; }
    MDBG_CK(c, cudaSetDevice(c->device));
    c->tm.launches_push = 0;
    c->tm.ms_h2d = 0;
    MDBG_CK(c, cudaEventRecord(c->ev[2], c->st));
    {   // the checks mdbg_push_reads makes on the host, on the device (the flag comes back with K-A's own scalars)
        MDBG_CK(c, cudaMemsetAsync(&c->d_sc->v[10], 0, 8, c->st));
        check_read_off_kernel<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, c->st>>>(d_read_off, n_reads, n_bases, &c->d_sc->v[10]);
        MDBG_CK(c, cudaGetLastError());
        MDBG_CK(c, cudaMemcpyAsync(&c->h_sc->v[10], &c->d_sc->v[10], 8, cudaMemcpyDeviceToHost, c->st));
        MDBG_CK(c, cudaStreamSynchronize(c->st));
        if (c->h_sc->v[10] & 1) { c->err = "read_off is not a non-decreasing offset array ending at n_bases"; return MDBG_ERR_BAD_ARG; }
        if (c->h_sc->v[10] & 2) { c->err = "a single record of 4 Gbases or more is not supported"; return MDBG_ERR_RANGE; }
    }
    int rc = run_ka(c, d_bases, d_read_off, n_reads, n_bases);
    if (rc) return rc;
    MDBG_CK(c, cudaEventRecord(c->ev[3], c->st));
    MDBG_CK(c, cudaEventSynchronize(c->ev[3]));
    cudaEventElapsedTime(&c->tm.ms_total_push, c->ev[2], c->ev[3]);
    return MDBG_OK;
}

// Host buffers -> device -> K-A.  Exactly one of `bases` (ASCII, 1 B/base) and `host_planes` (the caller's own 2-bit
// planes, mdbg_push_reads_packed) describes the reads.
static int push_host(mdbg_ctx* c, const uint8_t* bases, const uint32_t* host_planes, const uint64_t* read_off,
                     uint64_t n_reads) {
    MDBG_CK(c, cudaSetDevice(c->device));
    uint64_t B = read_off[n_reads];
    for (uint64_t r = 0; r < n_reads; r++) {
        if (read_off[r + 1] < read_off[r]) { c->err = "read_off is not non-decreasing"; return MDBG_ERR_BAD_ARG; }
        if (read_off[r + 1] - read_off[r] >= 0xFFFFFFF0ull) {   // positions inside a read are u32 on the device
            c->err = "a single record of 4 Gbases or more is not supported";
            return MDBG_ERR_RANGE;
        }
    }
    c->tm.launches_push = 0;
    Tmp<uint8_t> d_bases;
    Tmp<uint64_t> d_off;
    Tmp<uint32_t> d_planes;
    const bool packed = (c->upload_packed || host_planes) && B > 0;
    MDBG_CK(c, d_bases.get(c->pool, B + 64));
    MDBG_CK(c, d_off.get(c->pool, n_reads + 1));
    MDBG_CK(c, cudaEventRecord(c->ev[2], c->st));
    // Upload in 32-128 MB chunks cut at read starts on a second stream; K-A runs on the tiles whose
    // bytes have arrived, so the kernel hides behind the PCIe copy (pinned host memory).
    // (a chunk costs ~10 driver calls, a worker-pool round and one K-A launch: 32 MB for batches up to 2 GB, then
    // B / 64 up to 128 MB -- measured at config 3: 97.5 / 101.6 / 103.3 / 91.8 Gbases/s end to end at 32 / 64 / 128 / 256 MB)
    uint64_t CH = std::min<uint64_t>(128ull << 20, std::max<uint64_t>(32ull << 20, (B / 64) >> 20 << 20));
    if (const char* e = getenv("MDBG_UPLOAD_CHUNK_MB")) { long v = atol(e); if (v >= 1 && v <= 4096) CH = (uint64_t)v << 20; }
    std::vector<KaChunk> plan;
    const uint64_t n_tiles = std::max<uint64_t>(1, (B + KA_TILE - 1) / KA_TILE);
    MDBG_CK(c, cudaStreamWaitEvent(c->st_copy, c->ev[2], 0));   // the staging buffer is free again
    MDBG_CK(c, cudaEventRecord(c->ev[17], c->st_copy));
    // Every copy of the batch goes through the copy stream, the compute stream carries kernels only:
    // channels that share the copy engine are time-sliced, so one small copy on the compute stream
    // can sit behind the whole bulk upload and hold K-A back until the upload is over.
    MDBG_CK(c, cudaMemcpyAsync(d_off, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, c->st_copy));
    MDBG_CK(c, cudaEventRecord(c->ev[18], c->st_copy));
    MDBG_CK(c, cudaStreamWaitEvent(c->st, c->ev[18], 0));
    uint64_t b0 = 0;
    size_t nev = 0;
    // ---- 4:1 upload: pack on the host (worker threads), copy the bit planes, expand on the device ----
    const uint64_t n_words = (B + 31) / 32;
    std::vector<uint8_t> bad_tiles;
    // words per chunk of the packed / hybrid upload: a quarter of the ASCII chunk (8 MB of bases by default)
    uint64_t CHW = std::max<uint64_t>(2 * PACK_TILE_WORDS, (CH / 4 / 32) / PACK_TILE_WORDS * PACK_TILE_WORDS);
    // the caller's own planes need no packing pipeline: chunks of CH bytes of planes (128 Mbases by default), so that
    // a chunk's K-A launch fills the GPU and the host enqueues ~50 instead of ~900 chunks per 7 Gbases
    if (host_planes) CHW = std::max<uint64_t>(2 * PACK_TILE_WORDS, (CH / 8) / PACK_TILE_WORDS * PACK_TILE_WORDS);
    std::function<int(size_t)> prepare;
    if (packed) {
        if (!host_planes && !c->pack_pool) {
            unsigned hw = std::max<unsigned>(1, std::thread::hardware_concurrency());
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) { long v = atol(e); if (v >= 1 && v <= 64) hw = std::max<unsigned>(4, hw / (unsigned)v); }
            int nt = (int)std::min<unsigned>(32, hw);   // one process per GPU: share the host cores between the ranks
            if (const char* e = getenv("MDBG_PACK_THREADS")) { long v = atol(e); if (v >= 1 && v <= 256) nt = (int)v; }
            c->pack_pool = new PackPool(nt);
        }
        // pinned staging: a ring of UPLOAD_SLOTS chunk-sized slots (a slot is reused once the copy that read
        // it has completed), not a second copy of the batch
        const uint64_t n_chunks = (n_words + CHW - 1) / CHW;
        const size_t need = host_planes ? 0 : (size_t)std::min<uint64_t>(n_chunks, UPLOAD_SLOTS) * CHW * 8;
        if (c->h_planes_cap < need) {
            if (c->h_planes) cudaFreeHost(c->h_planes);
            c->h_planes = nullptr; c->h_planes_cap = 0;
            const size_t cap = need + need / 4 + 4096;
            MDBG_CK(c, cudaHostAlloc(&c->h_planes, cap, cudaHostAllocDefault));
            c->h_planes_cap = cap;
        }
        MDBG_CK(c, d_planes.get(c->pool, n_words * 2));
        bad_tiles.assign(n_tiles, 0);
        for (uint64_t ci = 0; ci < n_chunks; ci++) {
            if (nev == c->copy_ev.size()) {
                cudaEvent_t ev;
                MDBG_CK(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                c->copy_ev.push_back(ev);
            }
            const uint64_t wb = std::min(n_words, (ci + 1) * CHW);
            // K-A of chunk ci stops one tile short of the chunk: the look-ahead of its last tile is in the next chunk
            // (the classic kernel may walk a long homopolymer arbitrarily far ahead: it runs once, after the last chunk)
            const uint64_t te = wb == n_words ? n_tiles : (c->ka_bs ? wb / PACK_TILE_WORDS - 1 : 0);
            plan.push_back(KaChunk{te, c->copy_ev[nev]});
            nev++;
        }
        // Two resources move the batch: the host cores (packing) and the copy engine.  A chunk is packed
        // unless the copy engine has run dry, in which case it is sent as it is (ASCII) -- the engine then
        // has work for the time the cores need to pack the next chunks.  MDBG_UPLOAD=packed packs everything.
        // The decision is made on a host-side model of the engine's backlog: bytes queued so far at the link
        // rate (MDBG_PCIE_GBPS, default 50) against the measured packing time of a chunk.
        const bool hybrid = c->upload_hybrid && !host_planes;
        double link_rate = 50e9;
        if (const char* e = getenv("MDBG_PCIE_GBPS")) { double v = atof(e); if (v >= 1 && v <= 1000) link_rate = v * 1e9; }
        double ascii_rate = link_rate;   // pageable source: the driver stages ASCII chunks through its own pinned buffer
        if (bases) {
            cudaPointerAttributes pa{};
            if (cudaPointerGetAttributes(&pa, bases) != cudaSuccess) (void)cudaGetLastError();
            else if (pa.type == cudaMemoryTypeUnregistered) ascii_rate = std::min(link_rate, 12e9);
        }
        // The engine's backlog is OBSERVED, not modelled: bytes enqueued minus bytes of the chunks whose copy event
        // has completed, divided by the rate the completed chunks have actually moved at (several ranks share the
        // host's memory system and PCIe switches: the link gives half its nominal rate with 8 uploading ranks).
        // (the state of the upload lives on the heap: `prepare` is called by run_ka after this block has ended)
        struct Up {
            std::vector<uint8_t> mode;                     // per chunk: 0 undecided, 1 ASCII, 2 packed
            std::vector<double> t_pack;                    // when the chunk's packing began
            std::vector<uint64_t> enq;                     // bytes the chunk puts on the link
            size_t n_enqueued = 0, done_ptr = 0;           // chunks whose copy event has been recorded in THIS push / seen complete
            uint64_t cum_enq = 0, cum_done = 0;
            double pack_s_per_byte = 1.0 / 40e9, t_first = -1.0;
        };
        auto up = std::make_shared<Up>();
        up->mode.assign(n_chunks, 0); up->t_pack.assign(n_chunks, 0.0); up->enq.assign(n_chunks, 0);
        auto now_s = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        // A chunk is decided (ASCII or packed) one chunk AHEAD: a packed chunk starts packing on the workers at once
        // (PackPool::begin), the calling thread enqueues the copies and kernels of the previous chunk meanwhile and
        // joins the packing afterwards (PackPool::finish) -- the ~10 driver calls per chunk no longer stop 16 cores.
        auto chunk_words = [n_words, CHW](size_t ci, uint64_t& wa, uint64_t& wb) { wa = (uint64_t)ci * CHW; wb = std::min(n_words, wa + CHW); };
        auto slot_of = [c, CHW](size_t ci, uint64_t wa) {   // the chunk's slot of the staging ring, addressed as if the ring were the whole batch
            return reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(c->h_planes) + 8 * ((ci % UPLOAD_SLOTS) * CHW) - 8 * wa);
        };
        auto decide = [&, up, hybrid, link_rate, ascii_rate, now_s, chunk_words, slot_of](size_t ci) -> int {
            uint64_t wa, wb;
            chunk_words(ci, wa, wb);
            const uint64_t chunk_bytes = std::min<uint64_t>(B, wb * 32) - wa * 32;
            const double t_dec = now_s();
            while (up->done_ptr < up->n_enqueued && cudaEventQuery(c->copy_ev[up->done_ptr]) == cudaSuccess) up->cum_done += up->enq[up->done_ptr++];
            (void)cudaGetLastError();   // cudaErrorNotReady is not an error
            double rate = std::min(link_rate, ascii_rate);
            if (up->t_first >= 0 && up->cum_done >= (32u << 20) && t_dec > up->t_first)
                rate = std::min(rate, std::max(2e9, (double)up->cum_done / (t_dec - up->t_first)));
            const double backlog = (double)(up->cum_enq - up->cum_done) / rate;
            if (up->t_first < 0) up->t_first = t_dec;
            if (hybrid && backlog < up->pack_s_per_byte * (double)chunk_bytes) { up->mode[ci] = 1; return MDBG_OK; }   // the engine would run dry while we pack
            up->mode[ci] = 2;
            if (ci >= UPLOAD_SLOTS) MDBG_CK(c, cudaEventSynchronize(c->copy_ev[ci - UPLOAD_SLOTS]));   // slot free again
            up->t_pack[ci] = now_s();
            pack_parallel_begin(*c->pack_pool, bases, B, wa, wb, slot_of(ci, wa), bad_tiles.data());
            return MDBG_OK;
        };
        prepare = [&, up, decide, now_s, chunk_words, slot_of](size_t ci) -> int {
            uint64_t wa, wb;
            chunk_words(ci, wa, wb);
            if (host_planes) {         // the caller packed already: the planes of the chunk go as they are
                c->tm.upload_h2d_bytes += (wb - wa) * 8;
                MDBG_CK(c, cudaMemcpyAsync(d_planes.p + 2 * wa, host_planes + 2 * wa, (wb - wa) * 8, cudaMemcpyHostToDevice, c->st_copy));
                MDBG_CK(c, cudaEventRecord(c->copy_ev[ci], c->st_copy));
                if (wb == n_words) MDBG_CK(c, cudaEventRecord(c->ev[4], c->st_copy));
                MDBG_CK(c, cudaStreamWaitEvent(c->st, c->copy_ev[ci], 0));
                MDBG_CK(c, expand_planes(d_planes.p, d_bases.p, wa, wb, c->num_sms, c->st, &c->tm.launches_push));
                return MDBG_OK;
            }
            const uint64_t chunk_bytes = std::min<uint64_t>(B, wb * 32) - wa * 32;
            if (up->mode[ci] == 0) { int rc = decide(ci); if (rc) return rc; }
            if (up->mode[ci] == 2) {       // join the packing of this chunk and wait for the workers
                c->pack_pool->finish();
                const double per_byte = (now_s() - up->t_pack[ci]) / (double)std::max<uint64_t>(1, chunk_bytes);
                up->pack_s_per_byte = 0.5 * up->pack_s_per_byte + 0.5 * per_byte;
            }
            up->enq[ci] = up->mode[ci] == 1 ? chunk_bytes : (wb - wa) * 8;
            up->cum_enq += up->enq[ci];
            if (ci + 1 < up->mode.size()) { int rc = decide(ci + 1); if (rc) return rc; }   // the next chunk packs while this one is enqueued
            if (up->mode[ci] == 1) {
                const uint64_t off = wa * 32, end = std::min<uint64_t>(B, wb * 32);
                MDBG_CK(c, cudaMemcpyAsync(d_bases.p + off, bases + off, end - off, cudaMemcpyHostToDevice, c->st_copy));
                MDBG_CK(c, cudaEventRecord(c->copy_ev[ci], c->st_copy));
                if (wb == n_words) MDBG_CK(c, cudaEventRecord(c->ev[4], c->st_copy));
                up->n_enqueued = ci + 1;
                c->tm.upload_ascii_tiles += (uint32_t)((wb - wa + PACK_TILE_WORDS - 1) / PACK_TILE_WORDS);
                c->tm.upload_h2d_bytes += end - off;
                return MDBG_OK;        // run_ka makes the compute stream wait for copy_ev[ci]
            }
            const uint32_t* hp = slot_of(ci, wa);
            c->tm.upload_h2d_bytes += (wb - wa) * 8;
            MDBG_CK(c, cudaMemcpyAsync(d_planes.p + 2 * wa, hp + 2 * wa, (wb - wa) * 8, cudaMemcpyHostToDevice, c->st_copy));
            MDBG_CK(c, cudaEventRecord(c->copy_ev[ci], c->st_copy));
            if (wb == n_words) MDBG_CK(c, cudaEventRecord(c->ev[4], c->st_copy));
            up->n_enqueued = ci + 1;
            MDBG_CK(c, cudaStreamWaitEvent(c->st, c->copy_ev[ci], 0));
            MDBG_CK(c, expand_planes(d_planes.p, d_bases.p, wa, wb, c->num_sms, c->st, &c->tm.launches_push));
            // tiles with a byte outside ACGT travel as ASCII, over what the expansion wrote there
            const uint64_t ta = wa / PACK_TILE_WORDS, tb = (wb + PACK_TILE_WORDS - 1) / PACK_TILE_WORDS;
            for (uint64_t t = ta; t < tb;) {
                if (!bad_tiles[t]) { t++; continue; }
                uint64_t t2 = t;
                while (t2 < tb && bad_tiles[t2]) t2++;
                const uint64_t off = t * (uint64_t)KA_TILE, end = std::min<uint64_t>(B, t2 * (uint64_t)KA_TILE);
                MDBG_CK(c, cudaMemcpyAsync(d_bases.p + off, bases + off, end - off, cudaMemcpyHostToDevice, c->st));
                c->tm.upload_ascii_tiles += (uint32_t)(t2 - t);
                c->tm.upload_h2d_bytes += end - off;
                t = t2;
            }
            return MDBG_OK;
        };
        b0 = B;                        // nothing left for the ASCII chunk loop below
    }
    c->tm.upload_packed = packed ? 1 : 0;
    c->tm.upload_ascii_tiles = 0;
    c->tm.upload_h2d_bytes = packed ? 0 : B;
    while (b0 < B) {
        uint64_t target = b0 + CH;
        if (B - b0 <= CH + CH / 4 && B - b0 > CH / 2) target = B - CH / 4;   // short last chunk: short tail after the copy
        uint64_t b1 = B;
        if (target < B) {   // first read start >= target
            const uint64_t* it = std::lower_bound(read_off, read_off + n_reads + 1, target);
            b1 = *it;
        }
        MDBG_CK(c, cudaMemcpyAsync(d_bases.p + b0, bases + b0, b1 - b0, cudaMemcpyHostToDevice, c->st_copy));
        if (nev == c->copy_ev.size()) {
            cudaEvent_t ev;
            MDBG_CK(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->copy_ev.push_back(ev);
        }
        MDBG_CK(c, cudaEventRecord(c->copy_ev[nev], c->st_copy));
        plan.push_back(KaChunk{b1 >= B ? n_tiles : b1 / KA_TILE, c->copy_ev[nev]});
        nev++;
        b0 = b1;
    }
    if (plan.empty()) plan.push_back(KaChunk{n_tiles, nullptr});
    if (!packed) MDBG_CK(c, cudaEventRecord(c->ev[4], c->st_copy));
    int rc = run_ka(c, d_bases, d_off, n_reads, B, &plan, packed ? &prepare : nullptr);
    if (c->pack_pool) c->pack_pool->finish();   // (an error between a chunk's begin and its finish: nothing may outlive this call)
    if (rc == MDBG_ERR_ALPHABET && bases) {  // turn the batch offset into (read, offset) like SURVEY 5 asks
        uint64_t pos = c->h_sc->err_pos;
        uint64_t r = std::upper_bound(read_off, read_off + n_reads + 1, pos) - read_off - 1;
        char buf[200];
        snprintf(buf, sizeof buf, "Non-ACGTN nucleotide encountered! (read %llu, offset %llu, byte 0x%02x)",
                 (unsigned long long)(c->R + r), (unsigned long long)(pos - read_off[r]), bases[pos]);
        c->err = buf;
    }
    if (rc) { cudaStreamSynchronize(c->st_copy); return rc; }
    MDBG_CK(c, cudaEventRecord(c->ev[3], c->st));
    MDBG_CK(c, cudaEventSynchronize(c->ev[3]));
    MDBG_CK(c, cudaStreamSynchronize(c->st_copy));
    cudaEventElapsedTime(&c->tm.ms_h2d, c->ev[17], c->ev[4]);
    cudaEventElapsedTime(&c->tm.ms_total_push, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&c->tm.ms_ka_start, c->ev[2], c->ev[0]);
    return MDBG_OK;
}

int mdbg_push_reads(mdbg_ctx* c, const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads) {
    if (!c || !read_off || (n_reads && !bases && read_off[n_reads] != 0)) {
        if (c) c->err = "null argument";
        return MDBG_ERR_BAD_ARG;
    }
    return push_host(c, bases, nullptr, read_off, n_reads);
}

int mdbg_push_reads_packed(mdbg_ctx* c, const uint32_t* planes, const uint64_t* read_off, uint64_t n_reads) {
    if (!c || !read_off || (n_reads && !planes && read_off[n_reads] != 0)) {
        if (c) c->err = "null argument";
        return MDBG_ERR_BAD_ARG;
    }
    return push_host(c, nullptr, planes, read_off, n_reads);
}

int mdbg_get_minimizers(mdbg_ctx* c, uint64_t* hash, uint64_t* pos, uint64_t* read_off, uint64_t cap,
                        uint64_t* n_out) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    if (n_out) *n_out = c->M;
    if (read_off) {
        if (c->m_off) MDBG_CK(c, cudaMemcpyAsync(read_off, c->m_off, (c->R + 1) * 8, cudaMemcpyDeviceToHost, c->st));
        else read_off[0] = 0;
    }
    if (c->M > cap && (hash || pos)) {
        MDBG_CK(c, cudaStreamSynchronize(c->st));
        c->err = "output capacity too small";
        return MDBG_ERR_CAPACITY;
    }
    if (hash && c->M) MDBG_CK(c, cudaMemcpyAsync(hash, c->m_hash, c->M * 8, cudaMemcpyDeviceToHost, c->st));
    if (pos && c->M) {
        Tmp<uint64_t> wide;
        MDBG_CK(c, wide.get(c->pool, c->M));
        widen_pos_kernel<<<(unsigned)((c->M + 255) / 256), 256, 0, c->st>>>(c->m_pos, wide, c->M);
        MDBG_CK(c, cudaGetLastError());
        MDBG_CK(c, cudaMemcpyAsync(pos, wide, c->M * 8, cudaMemcpyDeviceToHost, c->st));
        MDBG_CK(c, cudaStreamSynchronize(c->st));
    }
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}

int mdbg_extract_minimizers(mdbg_ctx* c, const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads,
                            uint64_t* out_hash, uint64_t* out_pos, uint64_t* out_read_off, uint64_t cap,
                            uint64_t* n_out) {
    if (!c) return MDBG_ERR_BAD_ARG;
    // Stateless with respect to the pushed reads: run on a scratch arena.
    uint64_t sM = c->M, sR = c->R, sB = c->n_bases;
    uint64_t* sh = c->m_hash; uint32_t* sp = c->m_pos; uint64_t* so = c->m_off;
    size_t shc = c->m_hash_cap, spc = c->m_pos_cap, soc = c->m_off_cap;
    uint64_t smc = c->m_cap_items;
    c->M = c->R = c->n_bases = 0;
    c->m_hash = nullptr; c->m_pos = nullptr; c->m_off = nullptr;
    c->m_hash_cap = c->m_pos_cap = c->m_off_cap = 0; c->m_cap_items = 0;
    int rc = mdbg_push_reads(c, bases, read_off, n_reads);
    if (rc == MDBG_OK) rc = mdbg_get_minimizers(c, out_hash, out_pos, out_read_off, cap, n_out);
    else if (n_out) *n_out = 0;
    cudaSetDevice(c->device);
    if (c->m_hash) cudaFree(c->m_hash);
    if (c->m_pos) cudaFree(c->m_pos);
    if (c->m_off) cudaFree(c->m_off);
    c->M = sM; c->R = sR; c->n_bases = sB;
    c->m_hash = sh; c->m_pos = sp; c->m_off = so;
    c->m_hash_cap = shc; c->m_pos_cap = spc; c->m_off_cap = soc; c->m_cap_items = smc;
    return rc;
}

int mdbg_read_extract(mdbg_ctx* c, const uint8_t* seq, uint64_t len, uint64_t* out_hash, uint64_t* out_pos,
                      uint64_t cap, uint64_t* n_out) {
    uint64_t off[2] = {0, len};
    uint64_t ro[2];
    return mdbg_extract_minimizers(c, seq, off, 1, out_hash, out_pos, ro, cap, n_out);
}

int mdbg_get_timings(mdbg_ctx* c, mdbg_timings* out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    *out = c->tm;
    return MDBG_OK;
}

int mdbg_timer_start(mdbg_ctx* c) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    MDBG_CK(c, cudaEventRecord(c->ev[15], c->st));
    return MDBG_OK;
}
int mdbg_timer_stop(mdbg_ctx* c, float* ms) {
    if (!c || !ms) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    cudaEvent_t e;
    MDBG_CK(c, cudaEventCreate(&e));
    MDBG_CK(c, cudaEventRecord(e, c->st));
    MDBG_CK(c, cudaEventSynchronize(e));
    MDBG_CK(c, cudaEventElapsedTime(ms, c->ev[15], e));
    cudaEventDestroy(e);
    return MDBG_OK;
}

// ---- memory helpers ---------------------------------------------------------------------------
int mdbg_device_malloc(mdbg_ctx* c, uint64_t bytes, void** out) {
    if (!c || !out) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaMalloc(out, bytes ? bytes : 16));
    return MDBG_OK;
}
int mdbg_device_free(mdbg_ctx* c, void* p) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaFree(p));
    return MDBG_OK;
}
int mdbg_host_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return MDBG_ERR_BAD_ARG;
    if (cudaMallocHost(out, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return MDBG_ERR_CUDA; }
    return MDBG_OK;
}
int mdbg_host_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? MDBG_OK : MDBG_ERR_CUDA; }
int mdbg_memcpy_h2d(mdbg_ctx* c, void* dst, const void* src, uint64_t bytes) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}
int mdbg_memcpy_d2h(mdbg_ctx* c, void* dst, const void* src, uint64_t bytes) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->st));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}
int mdbg_sync(mdbg_ctx* c) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    MDBG_CK(c, cudaStreamSynchronize(c->st));
    return MDBG_OK;
}
int mdbg_flush_l2(mdbg_ctx* c) {
    if (!c) return MDBG_ERR_BAD_ARG;
    MDBG_CK(c, cudaSetDevice(c->device));
    if (!c->l2_flush) {
        c->l2_flush_bytes = 256u << 20;  // 2x the 126 MB L2
        MDBG_CK(c, cudaMalloc(&c->l2_flush, c->l2_flush_bytes));
    }
    static uint32_t v = 0;
    flush_kernel<<<c->num_sms * 8, 256, 0, c->st>>>((uint4*)c->l2_flush, c->l2_flush_bytes / 16, ++v);
    MDBG_CK(c, cudaGetLastError());
    return MDBG_OK;
}

}  // extern "C"
