/*
 * mdbg.h -- C ABI of libmdbg_b200.so: the B200-native (sm_100a) reads -> minimizer-space
 * de Bruijn graph hot path of ekimb/rust-mdbg, as a drop-in for that path only.
 *
 * The reference has no FFI boundary: the path sits behind in-process Rust module calls
 * (SURVEY.md 8b).  Each entry point below names the reference interface it replaces
 * (file:line into ekimb/rust-mdbg @ 077083d); INTEGRATION.md shows the `extern "C"` block
 * a Rust host would add.  Plain pointers and sizes only; no exceptions cross this boundary:
 * every function returns MDBG_OK (0) or a negative mdbg_status, and mdbg_last_error()
 * gives the message.  All compute runs in hand-written CUDA kernels; there is NO CPU
 * fallback -- without a CUDA device mdbg_ctx_create() fails with MDBG_ERR_NO_DEVICE.
 *
 * Threading: one mdbg_ctx per GPU, used from one host thread at a time (the reference calls
 * Read::extract / add_kminmer concurrently from `threads` workers, main.rs:834; here the
 * parallelism is inside the kernels).  The stateless helpers (mdbg_hash_bound,
 * mdbg_kminmer_*) are re-entrant.
 */
#ifndef MDBG_H
#define MDBG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MDBG_OK = 0,
    MDBG_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product never falls back to CPU */
    MDBG_ERR_CUDA = -2,        /* a CUDA runtime call failed                                   */
    MDBG_ERR_BAD_ARG = -3,
    MDBG_ERR_ALPHABET = -4,    /* non-ACGTN byte (the nthash crate panics there, read.rs:196)  */
    MDBG_ERR_CAPACITY = -5,    /* caller's output capacity too small; *n_out = needed          */
    MDBG_ERR_RANGE = -6,       /* an internal 32-bit index space would overflow                */
    MDBG_ERR_NCCL = -7,
    MDBG_ERR_IO = -8,
    MDBG_ERR_UNSUPPORTED = -9  /* reference mode outside the hot path (--syncmers, --uhs, ...) */
} mdbg_status;

/* &Params, src/main.rs:92-114 (the members the hot path reads).  density stays f64 and
 * presimp f32, type-exact with the reference (main.rs:98, main.rs:449).                   */
typedef struct {
    uint32_t k;              /* main.rs:95   */
    uint32_t l;              /* main.rs:94   */
    double   density;        /* main.rs:98   */
    uint32_t min_abundance;  /* DbgAbundance u16, main.rs:101; >= 1                          */
    float    presimp;        /* main.rs:449; 0 disables                                      */
    int32_t  hpc;            /* 1 = homopolymer-compress (default); 0 = --skiphpc (main.rs:507) */
    int32_t  device;         /* CUDA device ordinal for this context                         */
    int32_t  keep_bases;     /* reserved (the .sequences writer slices the caller's host copy) */
    uint32_t debug_fp_bits;  /* test hook: truncate tuple fingerprints to this many bits on the
                                first attempt to force the exact-collision path; 0 = off     */
    uint32_t bf;             /* 1 = --bf numbering (main.rs:639-655) with an ideal filter: a tuple
                                enters the table at its SECOND sighting (index order = second
                                sightings, "nodes before filter" = tuples seen >= 2 times);
                                ignored when min_abundance == 1, as in the reference            */
    uint32_t ka_variant;     /* K-A kernel: 0 = library default (= 2; env MDBG_KA_VARIANT=classic
                                selects 1), 1 = classic (ka_minimizers_kernel), 2 = bit-sliced
                                (ka_bitslice_kernel) where (l, density) allow it, classic elsewhere;
                                results are identical, only the speed differs                    */
    uint32_t reserved[5];
} mdbg_params;

typedef struct mdbg_ctx mdbg_ctx;

const char* mdbg_version(void);
int  mdbg_device_count(void);                 /* 0 when no usable CUDA device */
int  mdbg_ctx_create(const mdbg_params* p, mdbg_ctx** out);
void mdbg_ctx_destroy(mdbg_ctx* ctx);
const char* mdbg_last_error(const mdbg_ctx* ctx);   /* ctx may be NULL: last create error */
/* Change k / min_abundance / presimp between mdbg_finish calls (multi-k sweep over the
 * resident minimizer arrays, utils/multik:60-78).  l / density / hpc are fixed per context. */
int  mdbg_ctx_set_k(mdbg_ctx* ctx, uint32_t k, uint32_t min_abundance, float presimp);

/* (density as f64 * u64::MAX as f64) as u64 -- src/read.rs:183 */
uint64_t mdbg_hash_bound(double density);

/* ---- Entry 1: Read::extract (src/read.rs:85-90 -> extract_density, read.rs:176-211) -------
 * Batch form: R reads concatenated in `bases` (ASCII), read_off[R+1].  For read r the
 * minimizers are out_hash/out_pos[out_read_off[r] .. out_read_off[r+1]) == Read.transformed /
 * Read.minimizers_pos (raw coordinates).  HOST buffers; the call uploads, runs the kernel
 * and downloads.  *n_out = total minimizers (also when MDBG_ERR_CAPACITY is returned).     */
int mdbg_extract_minimizers(mdbg_ctx* ctx, const uint8_t* bases, const uint64_t* read_off,
                            uint64_t n_reads, uint64_t* out_hash, uint64_t* out_pos,
                            uint64_t* out_read_off, uint64_t cap, uint64_t* n_out);
/* Single-read convenience wrapper (what process_read_aux calls, main.rs:740). */
int mdbg_read_extract(mdbg_ctx* ctx, const uint8_t* seq, uint64_t len, uint64_t* out_hash,
                      uint64_t* out_pos, uint64_t cap, uint64_t* n_out);

/* ---- Entry 2: KmerVec value ops (src/kmer_vec.rs:16-47,73-77) ---------------------------
 * Host-side value helpers on one tuple (these are not the data-parallel path; the batch
 * canonicalisation of every window runs in the K-B kernel inside mdbg_finish /
 * mdbg_window).                                                                            */
void mdbg_kminmer_normalize(const uint64_t* in, uint32_t k, uint64_t* out, int* reversed);
void mdbg_kminmer_reverse(const uint64_t* in, uint32_t k, uint64_t* out);
void mdbg_kminmer_prefix(const uint64_t* in, uint32_t k, uint64_t* out /* k-1 */);
void mdbg_kminmer_suffix(const uint64_t* in, uint32_t k, uint64_t* out /* k-1 */);
int  mdbg_kminmer_cmp(const uint64_t* a, const uint64_t* b, uint32_t k);   /* Ord, -1/0/1 */
/* Batch windowing of already-extracted minimizers (main.rs:756-781): for every read with
 * m > k, windows i = 0..m-k.  Outputs, one per k-min-mer in (read, i) order (HOST buffers,
 * any may be NULL): canonical tuple (k u64 each), reversed flag, shift pair (untruncated),
 * read_offsets triple (main.rs:778), and kmer_read_off[R+1].                               */
int mdbg_window(mdbg_ctx* ctx, const uint64_t* hash, const uint64_t* pos,
                const uint64_t* min_read_off, uint64_t n_reads, uint64_t* out_tuple,
                uint8_t* out_reversed, uint64_t* out_shift /* 2 each */,
                uint64_t* out_offsets /* 3 each */, uint64_t* out_kmer_read_off,
                uint64_t cap, uint64_t* n_out);

/* ---- Entry 3: the node table + graph (add_kminmer main.rs:632-709, filter main.rs:922-933,
 *      GFA emission main.rs:1006-1121) ------------------------------------------------------
 * push: upload a batch of reads and run minimizer extraction; minimizers accumulate on the
 * device in global read order (reads are numbered in push order).  finish: windowing +
 * canonicalisation, hash-table count, radix sort + segmented reduce, abundance filter,
 * edge join + presimp; results under SERIAL-ORDER semantics.               */
int mdbg_push_reads(mdbg_ctx* ctx, const uint8_t* bases, const uint64_t* read_off,
                    uint64_t n_reads);
/* Same for a host that keeps its reads 2-bit packed (or packs while it parses: mdbg_pack_bases_host below):
 * planes[2w], planes[2w+1] = bit planes of the 32 bases [32w, 32w+32) of the concatenated reads, read_off as
 * above (in bases).  Only A, C, G, T can be said this way (a batch holding anything else goes through
 * mdbg_push_reads); bases past read_off[n_reads] in the last word are ignored.  A quarter of the bytes cross
 * PCIe and no host core packs inside the call; results are those of mdbg_push_reads on the same bases.     */
int mdbg_push_reads_packed(mdbg_ctx* ctx, const uint32_t* planes, const uint64_t* read_off,
                           uint64_t n_reads);
/* Same with inputs already resident in this context's device memory (16-byte aligned). */
int mdbg_push_reads_device(mdbg_ctx* ctx, const uint8_t* d_bases, const uint64_t* d_read_off,
                           uint64_t n_reads, uint64_t n_bases);
/* Forget all pushed reads (start a new job on the same context, keeping its workspace). */
int mdbg_reset(mdbg_ctx* ctx);

typedef struct {
    /* counters for the reference's stdout lines (main.rs:842,926-932,1118-1120) */
    uint64_t n_reads, n_bases, n_minimizers, n_kminmers;
    uint64_t n_distinct;       /* nodes before abundance filter */
    uint64_t n_nodes;          /* after                          */
    uint64_t n_edges, presimp_removed, n_seqlines;
    uint32_t k, l;
    /* nodes, ascending node index.  DbgEntry{index,abundance,seqlen,shift} main.rs:60       */
    uint32_t* node_index;
    uint16_t* abundance;
    uint32_t* seqlen;
    uint16_t* shift;           /* 2 per node */
    uint64_t* tuple;           /* k per node, canonical orientation */
    /* edges sorted by (n1, n2, o1, o2, overlap); o: 0 '+', 1 '-'   (L lines main.rs:1095)   */
    uint32_t* e_n1; uint8_t* e_o1; uint32_t* e_n2; uint8_t* e_o2; uint32_t* e_overlap;
    /* .sequences data lines in emission (serial) order (main.rs:696-707): node index, read,
     * raw slice [start,end), reversed, untruncated shift pair                               */
    uint32_t* q_index; uint64_t* q_read; uint64_t* q_start; uint64_t* q_end;
    uint8_t*  q_reversed; uint64_t* q_shift;   /* 2 per line */
    void* _owner;
} mdbg_graph;

/* Runs the table + graph stages and copies the result to host memory owned by *out
 * (release with mdbg_graph_free).  want_seqlines = 0 skips the q_* arrays (--no-basespace).
 * With N GPUs (mdbg_comm_init) every rank must call it; rank 0 receives the whole graph, the other
 * ranks the job-wide counters with NULL array members.                                       */
int  mdbg_finish(mdbg_ctx* ctx, int want_seqlines, mdbg_graph* out);
/* Same but leaves the graph in device memory and only returns the counters (pointer
 * members of *out are NULL): the device-resident timing path of bench.py.                  */
int  mdbg_finish_device(mdbg_ctx* ctx, mdbg_graph* out);
void mdbg_graph_free(mdbg_graph* g);

/* --read-stats (src/main.rs:939-975, src/read_stats.rs): for the reads of a second set, the abundance of
 * every k-min-mer among the nodes kept by the last mdbg_finish (0 = not a node).  HOST buffers;
 * out_counts[out_read_off[r] .. out_read_off[r+1]) belong to read r, in window order.  The pushed
 * reads and the graph stay as they are.  *n_out = number of counts (also on MDBG_ERR_CAPACITY).    */
int mdbg_read_stats(mdbg_ctx* ctx, const uint8_t* bases, const uint64_t* read_off, uint64_t n_reads,
                    uint32_t* out_counts, uint64_t* out_read_off /* n_reads+1 */, uint64_t cap,
                    uint64_t* n_out);

/* Minimizers currently resident (after push): copies to HOST buffers (any may be NULL). */
int mdbg_get_minimizers(mdbg_ctx* ctx, uint64_t* hash, uint64_t* pos, uint64_t* read_off,
                        uint64_t cap, uint64_t* n_out);

/* ---- measurement ------------------------------------------------------------------------ */
typedef struct {
    float ms_h2d, ms_ka, ms_kb, ms_kc, ms_kd, ms_ke, ms_d2h, ms_total_push, ms_total_finish;
    uint64_t launches_push, launches_finish;   /* kernels launched by the last push / finish */
    uint64_t ka_launches; float ka_ms_sum;     /* accumulated over pushes since reset        */
    uint32_t table_attempts;                   /* fingerprint seeds tried by the last finish */
    uint32_t ka_dense_tiles;                   /* tiles that took the exact (dense) path     */
    float ms_ka_kernel;                        /* ka_minimizers_kernel alone (last push)     */
    float ms_ka_start;                         /* push start -> first K-A work on the stream  */
    uint32_t ka_variant_used;                  /* 1 classic, 2 bit-sliced (last push)          */
    uint32_t ka_dirty_tiles;                   /* bit-sliced: tiles handed to the classic kernel */
    uint32_t upload_packed;                    /* last mdbg_push_reads: 1 = bases crossed PCIe as 2-bit planes */
    uint32_t upload_ascii_tiles;               /* ... 4 KiB tiles sent as ASCII instead (bytes outside ACGT, or
                                                  chunks sent unpacked because the copy engine was idle)       */
    uint64_t upload_h2d_bytes;                 /* bytes of bases that crossed PCIe in the last mdbg_push_reads */
    float ms_exchange;                         /* N > 1: the record all-to-all of the last finish (inside ms_kb..ms_kc) */
    uint64_t exchange_bytes;                   /* N > 1: bytes this GPU sent over NVLink in the last finish           */
    uint32_t exchange_p2p;                     /* N > 1: 1 = the records went straight into the owners' inboxes over NVLink
                                                  (kx_scatter_kernel over CUDA IPC mappings), 0 = ncclSend/ncclRecv      */
    float ms_kernels[8];                       /* single kernels of the last push / finish (CUDA events around each):
                                                  0 kb_records, 1 kc_insert, 2 kc_verify, 3 radix sort by slot,
                                                  4 ke_join, 5 ka_finalize, 6 kd_nodes + kd_expand, 7 reserved          */
} mdbg_timings;
int mdbg_get_timings(mdbg_ctx* ctx, mdbg_timings* out);
void* mdbg_stream(mdbg_ctx* ctx);              /* the cudaStream_t all kernels run on        */
/* CUDA-event stopwatch on that stream (the launching stream): start records an event, stop
 * records a second one, synchronises and returns the elapsed device time.                 */
int mdbg_timer_start(mdbg_ctx* ctx);
int mdbg_timer_stop(mdbg_ctx* ctx, float* ms);

/* ---- synthetic HiFi-shape reads (bench workload) ----------------------------
 * Counter-based (splitmix64) so any rank/device/CPU reproduces the same bytes.  plan: fills
 * read_off[n_reads+1] (host) for reads [first_read, first_read+n_reads) of the job; returns
 * total bases.  fill_device / fill_host write the ASCII bases.                              */
typedef struct {
    uint64_t genome_len;     /* Glen */
    double   mean_len, sd_len; uint64_t min_len, max_len;
    double   error_rate;     /* substitutions */
    uint64_t seed;
} mdbg_synth;
uint64_t mdbg_synth_num_reads(const mdbg_synth* s, double coverage);
uint64_t mdbg_synth_plan(const mdbg_synth* s, uint64_t first_read, uint64_t n_reads,
                         uint64_t* read_off, uint64_t* start, uint8_t* strand);
int mdbg_synth_fill_device(mdbg_ctx* ctx, const mdbg_synth* s, uint64_t first_read,
                           uint64_t n_reads, const uint64_t* read_off /*host*/,
                           uint8_t* d_bases, uint64_t* d_read_off);
void mdbg_synth_fill_host(const mdbg_synth* s, uint64_t first_read, uint64_t n_reads,
                          const uint64_t* read_off, uint8_t* bases, int threads);

/* device / pinned memory helpers for hosts without a CUDA binding (ctypes, Rust) */
int mdbg_device_malloc(mdbg_ctx* ctx, uint64_t bytes, void** out);
int mdbg_device_free(mdbg_ctx* ctx, void* p);
int mdbg_host_alloc_pinned(uint64_t bytes, void** out);
int mdbg_host_free_pinned(void* p);
int mdbg_memcpy_h2d(mdbg_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int mdbg_memcpy_d2h(mdbg_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int mdbg_sync(mdbg_ctx* ctx);
/* write >= L2-size bytes so the next timed iteration starts from a cold L2 */
int mdbg_flush_l2(mdbg_ctx* ctx);

/* ---- multi-GPU (one process per GPU, at most 16; reads sharded by record in contiguous ranges; every
 *      GPU windows its own reads, the k-min-mer records are bucketed by fingerprint prefix and cross
 *      NVLink in one NCCL all-to-all so that every tuple is counted on ONE owner; the 8-byte hash arenas
 *      are all-gathered so an owner can verify tuples; replaces the DashMap of src/main.rs:595,632-709;
 *      SURVEY.md 8e) ---------------------------------------------------------------------------- */
#define MDBG_NCCL_ID_BYTES 128
int mdbg_nccl_unique_id(uint8_t id[MDBG_NCCL_ID_BYTES]);             /* rank 0, then broadcast */
int mdbg_comm_init(mdbg_ctx* ctx, const uint8_t id[MDBG_NCCL_ID_BYTES], int rank, int world);
/* global index of the first read this rank pushes (default: ranks are numbered in rank order) */
int mdbg_comm_set_read_base(mdbg_ctx* ctx, uint64_t first_read);
/* host-only planning helpers (pure functions, testable without a GPU) */
void     mdbg_shard_reads(uint64_t n_reads_total, int world, int rank, uint64_t* lo, uint64_t* hi);
uint32_t mdbg_owner_of_fingerprint(uint64_t fp, int world);
uint64_t mdbg_tuple_fingerprint(const uint64_t* canonical, uint32_t k, uint64_t seed);

/* ---- file formats of the path (Appendix D of SURVEY.md) ---------------------------------- */
int mdbg_write_gfa(const mdbg_graph* g, const char* path);                   /* main.rs:1006-1117 */
/* One LZ4-frame .sequences file (stored blocks), main.rs:614-630,700-706; bases/read_off are
 * the HOST copies of the reads in global read order.                                       */
int mdbg_write_sequences(const mdbg_graph* g, const uint8_t* bases, const uint64_t* read_off,
                         const char* path, int lz4_frame);
/* Streaming form of the same writer (main.rs:696-707 writes a line while the worker still holds the read):
 * the q-entries are in serial (read, window) order, so a host that walks its reads once more in input order
 * hands over each read the writer asks for and never holds the whole read set.  *g must outlive the writer. */
typedef struct mdbg_seq_writer mdbg_seq_writer;
int      mdbg_seq_writer_open(const mdbg_graph* g, const char* path, int lz4_frame, mdbg_seq_writer** out);
/* One of n_parts writers sharing the lines of *g (the reference writes one {P}.{tid}.sequences per worker thread,
 * main.rs:614-630, and its readers glob them): part p owns the runs of 32 lines with run % n_parts == p; every part
 * writes the header.  Each part is driven like a whole writer (from its own thread if wished). */
int      mdbg_seq_writer_open_part(const mdbg_graph* g, const char* path, int lz4_frame, uint32_t part, uint32_t n_parts,
                                   mdbg_seq_writer** out);
uint64_t mdbg_seq_writer_next_read(const mdbg_seq_writer* w);      /* UINT64_MAX: every line is written */
int      mdbg_seq_writer_read(mdbg_seq_writer* w, uint64_t read_index, const uint8_t* read_bases, uint64_t read_len);
int      mdbg_seq_writer_close(mdbg_seq_writer* w);                /* MDBG_ERR_IO if a needed read never came */

/* ---- host ingest: 2-bit packing for the 4:1 upload (reference side: the reads the
 *      parser hands to Read::extract, main.rs:163-178,830-839) ----------------------------------
 * mdbg_push_reads packs host buffers like this itself (MDBG_UPLOAD=ascii turns it off, =packed packs
 * every chunk, default: chunks go unpacked whenever the copy engine would otherwise idle); the entry
 * point is public for hosts that pack while they parse.  planes[2w], planes[2w+1] = bit 1 and bit 2
 * of the 32 bases [32w, 32w+32) (A 00, C 01, T 10, G 11; bases past n_bases read as A); bad_tiles[t]
 * (may be NULL) is set to 1 when the 4096-base tile t holds a byte outside ACGT -- such tiles must
 * travel as ASCII.  threads <= 1: on the calling thread.                                        */
int mdbg_pack_bases_host(const uint8_t* bases, uint64_t n_bases, uint32_t* planes, uint8_t* bad_tiles,
                         int threads);

#ifdef __cplusplus
}
#endif
#endif
