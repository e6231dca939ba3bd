/*
 * mdbg_oracle.h -- CPU ORACLE for the reads -> mdBG hot path of ekimb/rust-mdbg.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product
 * (rust-mdbg_b200/, libmdbg_b200.so) never links, imports or calls anything here.
 *
 * It is a single-threaded, obviously-correct C++17 restatement of the reference
 * algorithm with SERIAL-ORDER semantics (one worker, reads in input order, k-min-mers
 * in increasing i), following (citations into /root/reference):
 *   src/read.rs:157-174   Read::encode_rle          -> orc_encode_rle
 *   src/read.rs:176-211   Read::extract_density     -> orc_extract
 *   crate nthash 0.5.x    NtHashIterator/ntf64/ntr64/ntc64 (NOT vendored in the
 *                         reference: Cargo.toml:26 `nthash = "*"`, no lockfile;
 *                         restated from the crate's published algorithm; pinned by
 *                         the crate's documented known-answer vectors and by the
 *                         reference-emitted .sequences lines quoted in
 *                         src/to_basespace.rs:203 and
 *                         experiments/661k_genomes/scan_genomes_minmers.py:38,
 *                         see tests/test_oracle_golden.py)
 *   src/kmer_vec.rs:16-42,73-77  KmerVec prefix/suffix/reverse/normalize/Ord
 *   src/main.rs:60        DbgEntry{index:u32, abundance:u16, seqlen:u32, shift:(u16,u16)}
 *   src/main.rs:632-709   add_kminmer (no --bf branch)
 *   src/main.rs:756-781   windowing loop (strict m > k), shift pair, read_offsets
 *   src/main.rs:922-933   abundance filter
 *   src/main.rs:1006-1121 S lines, km_index, 4-orientation edge test, presimp, overlap
 *   src/utils.rs:3-24     revcomp
 *
 * A second entry point (orc_build_mt) runs the same algorithm in the reference's
 * THREAD STRUCTURE (N workers over reads against a sharded concurrent map, then the
 * single-threaded filter + edge pass) and is used only as the timed CPU baseline;
 * like the multi-threaded reference its node indices depend on scheduling.
 */
#ifndef MDBG_ORACLE_H
#define MDBG_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t k;             /* minimizers per k-min-mer               (main.rs:95)  */
    uint32_t l;             /* l-mer length                           (main.rs:94)  */
    double   density;       /* f64, main.rs:98                                      */
    uint32_t min_abundance; /* DbgAbundance = u16, main.rs:53,101; must be >= 1     */
    float    presimp;       /* f32, main.rs:449                                     */
    int32_t  hpc;           /* 1 = homopolymer-compress (default), 0 = --skiphpc    */
    int32_t  bf;            /* 1 = --bf with an IDEAL Bloom filter (no false positives): main.rs:639-655 */
} orc_params;

/* ---- ntHash (crate nthash) --------------------------------------------------- */
/* return 0 on success, -1 if a non-ACGTN byte is met (the crate panics there).    */
int orc_ntf64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out);
int orc_ntr64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out);
int orc_ntc64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out);
/* NtHashIterator::new(seq,k) collected; out must hold len-k+1 values.
 * returns number of values, -1 bad byte, -2 k > len (the crate's Err).            */
int64_t orc_nthash_iter(const uint8_t* seq, uint64_t len, uint32_t k, uint64_t* out);

/* (density as f64 * u64::MAX as f64) as u64, read.rs:183 (Rust saturating cast).  */
uint64_t orc_hash_bound(double density);

/* Read::encode_rle, read.rs:157-174.  hpc_out/pos_out must hold len entries.
 * returns HPC length.                                                              */
uint64_t orc_encode_rle(const uint8_t* seq, uint64_t len, uint8_t* hpc_out, uint64_t* pos_out);

/* Read::extract_density, read.rs:176-211 (default path: no lmer_counts, no EC).
 * Writes up to cap (hash,pos) pairs; returns the number of minimizers (may exceed
 * cap: nothing is written past cap), -1 on a non-ACGTN byte (*bad_off = offset of the
 * first offending byte in the HASHED string's raw coordinates).                     */
int64_t orc_extract(const uint8_t* seq, uint64_t len, const orc_params* p,
                    uint64_t* out_hash, uint64_t* out_pos, uint64_t cap, uint64_t* bad_off);

/* KmerVec::normalize, kmer_vec.rs:34-39.  out may alias in only if identical.      */
void orc_normalize(const uint64_t* in, uint32_t k, uint64_t* out, int* reversed);

/* utils::revcomp, utils.rs:3-24 */
void orc_revcomp(const uint8_t* in, uint64_t len, uint8_t* out);

/* ---- whole path --------------------------------------------------------------- */
typedef struct orc_graph orc_graph;

typedef struct {
    uint64_t n_reads, n_bases, n_hpc_bases, n_minimizers, n_kminmers;
    uint64_t n_distinct;      /* "Number of nodes before abundance filter"  main.rs:926 */
    uint64_t n_nodes;         /* after filter (main.rs:928) / all if minabund==1        */
    uint64_t n_edges;         /* "Number of mdBG edges"                     main.rs:1118 */
    uint64_t presimp_removed; /* main.rs:1120                                           */
    uint64_t n_seqlines;      /* .sequences data lines (main.rs:696-707)                */
    int64_t  error;           /* 0, or -1 non-ACGTN byte                                */
    uint64_t error_read, error_offset;
} orc_stats;

/* Serial-order build.  bases = concatenated ASCII reads, read_off[R+1].            */
orc_graph* orc_build(const uint8_t* bases, const uint64_t* read_off, uint64_t R,
                     const orc_params* p);
/* Reference thread structure; timing baseline only. emit_edges=0 skips main.rs:1006+ */
orc_graph* orc_build_mt(const uint8_t* bases, const uint64_t* read_off, uint64_t R,
                        const orc_params* p, int threads);
void orc_graph_free(orc_graph* g);
void orc_graph_stats(const orc_graph* g, orc_stats* out);

/* Nodes, sorted by index ascending (n_nodes entries).  tuple: n_nodes*k u64.        */
void orc_graph_nodes(const orc_graph* g, uint32_t* index, uint16_t* abundance, uint32_t* seqlen,
                     uint16_t* shift /* 2 per node */, uint64_t* tuple);
/* Edges, sorted by (n1,n2,o1,o2,overlap); o = 0 for '+', 1 for '-'.                 */
void orc_graph_edges(const orc_graph* g, uint32_t* n1, uint8_t* o1, uint32_t* n2, uint8_t* o2,
                     uint32_t* overlap);
/* .sequences data lines in emission order: node index, read, [start,end) raw slice,
 * reversed flag, untruncated shift pair (main.rs:696-707).                          */
void orc_graph_seqlines(const orc_graph* g, uint32_t* index, uint64_t* read, uint64_t* start,
                        uint64_t* end, uint8_t* reversed, uint64_t* shift /* 2 per line */);

/* Per-read minimizers kept by the serial build (for P0 parity): offsets R+1.       */
uint64_t orc_graph_minimizers(const orc_graph* g, uint64_t* hash, uint64_t* pos, uint64_t* read_off);

/* Text writers in the CANONICAL PARITY FORM: header line, then sorted S
 * lines, then sorted L lines / sorted .sequences data lines without '#' header.     */
/* --read-stats (main.rs:939-975): abundance of every k-min-mer of a second read set among the kept nodes */
int64_t orc_read_stats(const orc_graph* g, const uint8_t* bases, const uint64_t* read_off, uint64_t R,
                       uint32_t* out_counts, uint64_t* out_off, uint64_t cap);
int orc_write_gfa(const orc_graph* g, const char* path);
int orc_write_sequences(const orc_graph* g, const uint8_t* bases, const uint64_t* read_off,
                        const char* path);

#ifdef __cplusplus
}
#endif
#endif
