// mdbg_oracle.cpp -- CPU ORACLE (test infrastructure; see mdbg_oracle.h header).
// Single-threaded restatement of the reference algorithm, written for obviousness,
// not speed.  Every function cites the reference lines it follows.
#include "mdbg_oracle.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

typedef std::vector<uint64_t> Tuple;  // KmerVec{data: Vec<u64>}  kmer_vec.rs:6-9

// ---------------------------------------------------------------- ntHash ---------
// crate nthash: seed table H, RC = H of the complement, 'N' -> 0, anything else
// panics ("Non-ACGTN nucleotide encountered!").  Plain 64-bit rotates.
const uint64_t HA = 0x3c8bfbb395c60474ULL, HC = 0x3193c18562a02b4cULL,
               HG = 0x20323ed082572324ULL, HT = 0x295549f54be24456ULL;

inline bool h_fwd(uint8_t c, uint64_t* v) {
    switch (c) {
        case 'A': *v = HA; return true;
        case 'C': *v = HC; return true;
        case 'G': *v = HG; return true;
        case 'T': *v = HT; return true;
        case 'N': *v = 0;  return true;
        default:  return false;
    }
}
inline bool h_rc(uint8_t c, uint64_t* v) {
    switch (c) {
        case 'A': *v = HT; return true;
        case 'C': *v = HG; return true;
        case 'G': *v = HC; return true;
        case 'T': *v = HA; return true;
        case 'N': *v = 0;  return true;
        default:  return false;
    }
}
inline uint64_t rol(uint64_t x, uint32_t r) { r &= 63; return r ? (x << r) | (x >> (64 - r)) : x; }
inline uint64_t ror(uint64_t x, uint32_t r) { r &= 63; return r ? (x >> r) | (x << (64 - r)) : x; }

// ntf64: XOR_j rol(H[s[i+j]], k-1-j)
bool ntf64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out) {
    uint64_t h = 0, v;
    for (uint32_t j = 0; j < k; j++) {
        if (!h_fwd(s[i + j], &v)) return false;
        h ^= rol(v, k - 1 - j);
    }
    *out = h;
    return true;
}
// ntr64: XOR_j rol(RC[s[i+j]], j)
bool ntr64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out) {
    uint64_t h = 0, v;
    for (uint32_t j = 0; j < k; j++) {
        if (!h_rc(s[i + j], &v)) return false;
        h ^= rol(v, j);
    }
    *out = h;
    return true;
}

// NtHashIterator: init with ntf64/ntr64 at 0, then the rolling update
//   fh = rol(fh,1) ^ rol(H[s[i]],k) ^ H[s[i+k]]
//   rh = ror(rh,1) ^ ror(RC[s[i]],1) ^ rol(RC[s[i+k]],k-1)
// yielding min(fh,rh) for each of the len-k+1 windows.
// returns count, -1 on bad byte (*bad = its offset), -2 if k > len.
int64_t nthash_iter(const uint8_t* seq, uint64_t len, uint32_t k, std::vector<uint64_t>& out,
                    uint64_t* bad) {
    out.clear();
    if (k > len) return -2;
    uint64_t fh = 0, rh = 0, v = 0, w = 0;
    for (uint32_t j = 0; j < k; j++)
        if (!h_fwd(seq[j], &v)) { if (bad) *bad = j; return -1; }
    ntf64(seq, 0, k, &fh);
    ntr64(seq, 0, k, &rh);
    uint64_t n = len - k + 1;
    out.reserve(n);
    out.push_back(std::min(fh, rh));
    for (uint64_t idx = 1; idx < n; idx++) {
        uint64_t i = idx - 1;
        uint8_t seqi = seq[i], seqk = seq[i + k];
        if (!h_fwd(seqk, &v)) { if (bad) *bad = i + k; return -1; }
        h_fwd(seqi, &w);
        fh = rol(fh, 1) ^ rol(w, k) ^ v;
        h_rc(seqi, &w);
        h_rc(seqk, &v);
        rh = ror(rh, 1) ^ ror(w, 1) ^ rol(v, k - 1);
        out.push_back(std::min(fh, rh));
    }
    return (int64_t)n;
}

// read.rs:183: ((density as f64) * (u64::max_value() as f64)) as u64.
// `u64::MAX as f64` rounds to 2^64; Rust float->int `as` saturates, NaN -> 0.
uint64_t hash_bound(double density) {
    double x = density * 18446744073709551616.0;
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}

// read.rs:157-174 Read::encode_rle.  pos[j] = raw index of the FIRST base of run j.
// A repeated char is only collapsed if it is in "ACTGactgNn" (read.rs:163).
void encode_rle(const uint8_t* seq, uint64_t len, std::vector<uint8_t>& hpc,
                std::vector<uint64_t>& pos) {
    hpc.clear();
    pos.clear();
    if (len == 0) {
        // the Rust code pushes prev_char '#' for an empty string; the caller then
        // sees hpc.len()=1 < l and returns an empty read (l >= 2 in practice).  We keep
        // the same observable result (no minimizers) without the sentinel.
        return;
    }
    static const char* collapsible = "ACTGactgNn";
    int prev_char = -1;  // '#'
    uint64_t prev_i = 0;
    for (uint64_t i = 0; i < len; i++) {
        int c = seq[i];
        if (c == prev_char && strchr(collapsible, c) != nullptr && c != 0) continue;
        if (prev_char != -1) {
            hpc.push_back((uint8_t)prev_char);
            pos.push_back(prev_i);
            prev_i = i;
        }
        prev_char = c;
    }
    hpc.push_back((uint8_t)prev_char);
    pos.push_back(prev_i);
}

// read.rs:176-211 Read::extract_density, default path.
// returns 0 ok, -1 bad byte (bad_off in raw coordinates).
int extract_density(const uint8_t* seq, uint64_t len, const orc_params& p,
                    std::vector<uint64_t>& hashes, std::vector<uint64_t>& positions,
                    uint64_t* n_hpc, uint64_t* bad_off) {
    hashes.clear();
    positions.clear();
    const uint64_t bound = hash_bound(p.density);
    std::vector<uint8_t> hpc;
    std::vector<uint64_t> pos_vec;
    const uint8_t* s = seq;
    uint64_t n = len;
    if (p.hpc) {  // !params.reads_already_hpc, read.rs:186-188
        encode_rle(seq, len, hpc, pos_vec);
        s = hpc.data();
        n = hpc.size();
    }
    if (n_hpc) *n_hpc = n;
    if (n < p.l) return 0;  // read.rs:193-195
    std::vector<uint64_t> hs;
    uint64_t bad = 0;
    int64_t r = nthash_iter(s, n, p.l, hs, &bad);
    if (r == -1) {
        if (bad_off) *bad_off = p.hpc ? pos_vec[bad] : bad;
        return -1;
    }
    for (uint64_t i = 0; i < hs.size(); i++) {
        if (hs[i] <= bound) {  // inclusive, read.rs:196
            positions.push_back(p.hpc ? pos_vec[i] : i);  // read.rs:206-207
            hashes.push_back(hs[i]);                       // read.rs:208
        }
    }
    return 0;
}

// kmer_vec.rs:34-39: rev = reverse(self); if self < rev {(self,false)} else {(rev,true)}
// (lexicographic Vec<u64> order; a palindromic tuple is reported reversed).
bool normalize(const Tuple& in, Tuple& out) {
    Tuple rev(in.rbegin(), in.rend());
    if (in < rev) { out = in; return false; }
    out = rev;
    return true;
}
Tuple prefix_of(const Tuple& t) { return Tuple(t.begin(), t.end() - 1); }  // kmer_vec.rs:22-26
Tuple suffix_of(const Tuple& t) { return Tuple(t.begin() + 1, t.end()); }  // kmer_vec.rs:16-20
Tuple reverse_of(const Tuple& t) { return Tuple(t.rbegin(), t.rend()); }   // kmer_vec.rs:28-32

struct TupleHash {
    size_t operator()(const Tuple& t) const {
        uint64_t h = 0x9e3779b97f4a7c15ULL;
        for (uint64_t v : t) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); h *= 0xff51afd7ed558ccdULL; }
        return (size_t)(h ^ (h >> 29));
    }
};

struct Entry {  // main.rs:60 DbgEntry
    uint32_t index;
    uint16_t abundance;
    uint32_t seqlen;
    uint16_t shift0, shift1;
};
struct SeqLine {  // what main.rs:696-707 needs to print one .sequences line
    uint32_t index;
    uint64_t read, start, end;
    uint8_t reversed;
    uint64_t shift0, shift1;
    Tuple node;
};
struct Node {
    Tuple t;
    Entry e;
};
struct Edge {
    uint32_t n1; uint8_t o1; uint32_t n2; uint8_t o2; uint32_t ov;
    bool operator<(const Edge& b) const {
        if (n1 != b.n1) return n1 < b.n1;
        if (n2 != b.n2) return n2 < b.n2;
        if (o1 != b.o1) return o1 < b.o1;
        if (o2 != b.o2) return o2 < b.o2;
        return ov < b.ov;
    }
};

}  // namespace

struct orc_graph {
    orc_params p;
    orc_stats st;
    std::vector<Node> nodes;  // sorted by index
    std::vector<Edge> edges;  // sorted
    std::vector<SeqLine> seqlines;
    std::vector<uint64_t> mhash, mpos, moff;
};

namespace {

// One k-min-mer sighting: main.rs:632-709 without --bf / --reference / EC.
// `table` plays dbg_nodes, `next_index` plays NODE_INDEX.
template <class Map>
inline void add_kminmer(Map& table, uint32_t& next_index, const orc_params& p, const Tuple& node,
                        bool reversed, uint64_t shift0, uint64_t shift1, uint64_t read,
                        uint64_t off0, uint64_t off1, uint64_t off2, std::vector<SeqLine>* lines) {
    const uint16_t minab = (uint16_t)p.min_abundance;
    uint16_t previous_abundance;
    uint32_t cur_node_index;
    auto it = table.find(node);
    const bool use_bf = p.bf && minab > 1;   // main.rs:639
    if (use_bf) {
        // main.rs:641-655 with an ideal filter: the first sighting only enters the filter (the
        // entry below, index -1 = "in the filter, not in dbg_nodes"); the second one inserts the
        // node with abundance previous_abundance+1 = 2 and consumes an index (main.rs:687-691).
        if (it == table.end()) {
            Entry e;
            e.index = 0xFFFFFFFFu; e.abundance = 0; e.seqlen = 0; e.shift0 = e.shift1 = 0;
            table.emplace(node, e);
            return;
        }
        Entry& e = it->second;
        if (e.index == 0xFFFFFFFFu) {
            previous_abundance = 1;
            e.index = next_index++;
            e.abundance = (uint16_t)(previous_abundance + 1);
            e.seqlen = (uint32_t)off2;
            e.shift0 = (uint16_t)shift0; e.shift1 = (uint16_t)shift1;
            cur_node_index = e.index;
        } else {
            cur_node_index = e.index;
            previous_abundance = e.abundance;
            if (previous_abundance == (uint16_t)(minab - 1)) {
                e.seqlen = (uint32_t)off2;
                e.shift0 = (uint16_t)shift0; e.shift1 = (uint16_t)shift1;
            }
            e.abundance = (uint16_t)(e.abundance + 1);
        }
    } else {
    if (it == table.end()) {  // main.rs:662-670: new key consumes an index, abundance 0
        Entry e;
        e.index = next_index++;
        e.abundance = 0;
        e.seqlen = (uint32_t)off2;
        e.shift0 = (uint16_t)shift0;
        e.shift1 = (uint16_t)shift1;
        it = table.emplace(node, e).first;
    }
    {  // main.rs:675-686 (contains_key is always true on this path)
        Entry& e = it->second;
        cur_node_index = e.index;
        previous_abundance = e.abundance;
        if (previous_abundance == (uint16_t)(minab - 1)) {
            e.seqlen = (uint32_t)off2;
            e.shift0 = (uint16_t)shift0;
            e.shift1 = (uint16_t)shift1;
        }
        e.abundance = (uint16_t)(e.abundance + 1);  // u16 += 1, wraps in --release
    }
    }
    if (previous_abundance >= 1 || minab == 1) {          // main.rs:693
        if (previous_abundance == (uint16_t)(minab - 1)) {  // main.rs:696
            if (lines) {
                SeqLine s;
                s.index = cur_node_index;
                s.read = read; s.start = off0; s.end = off1;
                s.reversed = reversed ? 1 : 0;
                s.shift0 = shift0; s.shift1 = shift1;
                s.node = node;
                lines->push_back(s);
            }
        }
    }
}

// main.rs:756-781: windows of one read.
template <class F>
inline void for_each_kminmer(const std::vector<uint64_t>& t, const std::vector<uint64_t>& pos,
                             const orc_params& p, F&& f) {
    const uint64_t k = p.k, l = p.l;
    if (!(t.size() > k)) return;  // strict, main.rs:756
    Tuple node, norm;
    for (uint64_t i = 0; i < t.size() - k + 1; i++) {
        node.assign(t.begin() + i, t.begin() + i + k);
        bool reversed = normalize(node, norm);
        uint64_t second = reversed ? pos[i + k - 1] - pos[i + k - 2] : pos[i + 1] - pos[i];
        uint64_t second_to_last = reversed ? pos[i + 1] - pos[i] : pos[i + k - 1] - pos[i + k - 2];
        uint64_t off0 = pos[i], off1 = pos[i + k - 1] + l, off2 = pos[i + k - 1] + 1 - pos[i] + 1;
        f(norm, reversed, second, second_to_last, off0, off1, off2);
    }
}

// main.rs:1006-1121 on the filtered node set.
void emit_graph(orc_graph* g, std::vector<Node>& nodes) {
    const orc_params& p = g->p;
    std::sort(nodes.begin(), nodes.end(), [](const Node& a, const Node& b) { return a.e.index < b.e.index; });
    // km_index: (k-1)-mer (normalised) -> list of nodes, each node pushed under its
    // prefix key then its suffix key (twice under the same key if they coincide).
    std::unordered_map<Tuple, std::vector<uint32_t>, TupleHash> km_index;
    for (uint32_t n = 0; n < nodes.size(); n++) {
        Tuple first, second;
        normalize(prefix_of(nodes[n].t), first);
        normalize(suffix_of(nodes[n].t), second);
        km_index[first].push_back(n);
        km_index[second].push_back(n);
    }
    struct Pending { uint32_t n1i, n2i; uint8_t o1, o2; uint32_t ov; };
    std::vector<Pending> vec_edges;
    std::set<std::pair<uint32_t, uint32_t>> removed;
    uint64_t presimp_removed = 0;
    const float presimp = p.presimp;
    for (uint32_t a = 0; a < nodes.size(); a++) {
        const Tuple& n1 = nodes[a].t;
        const Entry& e1 = nodes[a].e;
        Tuple rev_n1 = reverse_of(n1);
        Tuple key1, key2;
        normalize(suffix_of(n1), key1);
        normalize(prefix_of(n1), key2);
        const Tuple* keys[2] = {&key1, &key2};
        for (int kk = 0; kk < 2; kk++) {
            auto it = km_index.find(*keys[kk]);
            if (it == km_index.end()) continue;
            struct Pot { uint32_t b; uint8_t o1, o2; };
            std::vector<Pot> potential;
            for (uint32_t b : it->second) {
                const Tuple& n2 = nodes[b].t;
                Tuple rev_n2 = reverse_of(n2);
                if (suffix_of(n1) == prefix_of(n2)) potential.push_back({b, 0, 0});
                if (suffix_of(n1) == prefix_of(rev_n2)) potential.push_back({b, 0, 1});
                if (suffix_of(rev_n1) == prefix_of(n2)) potential.push_back({b, 1, 0});
                if (suffix_of(rev_n1) == prefix_of(rev_n2)) potential.push_back({b, 1, 1});
            }
            if (potential.empty()) continue;
            uint16_t abundance_max = 0;
            for (auto& q : potential) abundance_max = std::max(abundance_max, nodes[q.b].e.abundance);
            uint16_t abundance_ref = std::min(abundance_max, e1.abundance);
            for (auto& q : potential) {
                const Entry& e2 = nodes[q.b].e;
                if (presimp > 0.0f && potential.size() >= 2 &&
                    (float)e2.abundance < presimp * (float)abundance_ref) {  // main.rs:1086
                    presimp_removed++;
                    removed.insert({e1.index, e2.index});
                    continue;
                }
                uint16_t shift = q.o1 == 0 ? e1.shift0 : e1.shift1;
                uint32_t overlap = std::min((uint32_t)(e1.seqlen - (uint32_t)shift), (uint32_t)(e2.seqlen - 1));
                vec_edges.push_back({e1.index, e2.index, q.o1, q.o2, overlap});
            }
        }
    }
    g->edges.clear();
    for (auto& e : vec_edges) {
        if (presimp > 0.0f) {  // main.rs:1107-1116
            if (removed.count({e.n1i, e.n2i}) || removed.count({e.n2i, e.n1i})) continue;
        }
        g->edges.push_back({e.n1i, e.o1, e.n2i, e.o2, e.ov});
    }
    std::sort(g->edges.begin(), g->edges.end());
    g->st.n_edges = g->edges.size();
    g->st.presimp_removed = presimp > 0.0f ? presimp_removed : 0;
    g->nodes.swap(nodes);
    g->st.n_nodes = g->nodes.size();
}

}  // namespace

extern "C" {

int orc_ntf64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out) { return ntf64(s, i, k, out) ? 0 : -1; }
int orc_ntr64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out) { return ntr64(s, i, k, out) ? 0 : -1; }
int orc_ntc64(const uint8_t* s, uint64_t i, uint32_t k, uint64_t* out) {
    uint64_t f, r;
    if (!ntf64(s, i, k, &f) || !ntr64(s, i, k, &r)) return -1;
    *out = std::min(f, r);
    return 0;
}
int64_t orc_nthash_iter(const uint8_t* seq, uint64_t len, uint32_t k, uint64_t* out) {
    std::vector<uint64_t> v;
    int64_t r = nthash_iter(seq, len, k, v, nullptr);
    if (r < 0) return r;
    memcpy(out, v.data(), v.size() * 8);
    return r;
}
uint64_t orc_hash_bound(double density) { return hash_bound(density); }

uint64_t orc_encode_rle(const uint8_t* seq, uint64_t len, uint8_t* hpc_out, uint64_t* pos_out) {
    std::vector<uint8_t> h;
    std::vector<uint64_t> p;
    encode_rle(seq, len, h, p);
    if (hpc_out) memcpy(hpc_out, h.data(), h.size());
    if (pos_out) memcpy(pos_out, p.data(), p.size() * 8);
    return h.size();
}

int64_t orc_extract(const uint8_t* seq, uint64_t len, const orc_params* p, uint64_t* out_hash,
                    uint64_t* out_pos, uint64_t cap, uint64_t* bad_off) {
    std::vector<uint64_t> h, ps;
    uint64_t bad = 0;
    if (extract_density(seq, len, *p, h, ps, nullptr, &bad) != 0) {
        if (bad_off) *bad_off = bad;
        return -1;
    }
    uint64_t n = std::min<uint64_t>(h.size(), cap);
    if (out_hash) memcpy(out_hash, h.data(), n * 8);
    if (out_pos) memcpy(out_pos, ps.data(), n * 8);
    return (int64_t)h.size();
}

void orc_normalize(const uint64_t* in, uint32_t k, uint64_t* out, int* reversed) {
    Tuple a(in, in + k), o;
    bool r = normalize(a, o);
    memcpy(out, o.data(), k * 8);
    if (reversed) *reversed = r ? 1 : 0;
}

void orc_revcomp(const uint8_t* in, uint64_t len, uint8_t* out) {  // utils.rs:3-24
    for (uint64_t i = 0; i < len; i++) {
        uint8_t c = in[len - 1 - i], o;
        switch (c) {
            case 'a': o = 't'; break; case 'c': o = 'g'; break; case 't': o = 'a'; break;
            case 'g': o = 'c'; break; case 'u': o = 'a'; break; case 'A': o = 'T'; break;
            case 'C': o = 'G'; break; case 'T': o = 'A'; break; case 'G': o = 'C'; break;
            case 'U': o = 'A'; break; default: o = 'N';
        }
        out[i] = o;
    }
}

orc_graph* orc_build(const uint8_t* bases, const uint64_t* read_off, uint64_t R, const orc_params* pp) {
    orc_graph* g = new orc_graph();
    g->p = *pp;
    memset(&g->st, 0, sizeof(g->st));
    const orc_params& p = g->p;
    std::unordered_map<Tuple, Entry, TupleHash> table;  // dbg_nodes, main.rs:595
    uint32_t next_index = 0;                            // NODE_INDEX, main.rs:598
    std::vector<uint64_t> t, pos;
    g->moff.push_back(0);
    g->st.n_reads = R;
    for (uint64_t r = 0; r < R; r++) {
        const uint8_t* s = bases + read_off[r];
        uint64_t len = read_off[r + 1] - read_off[r];
        uint64_t nh = 0, bad = 0;
        g->st.n_bases += len;
        if (extract_density(s, len, p, t, pos, &nh, &bad) != 0) {
            g->st.error = -1; g->st.error_read = r; g->st.error_offset = bad;
            return g;
        }
        g->st.n_hpc_bases += nh;
        g->mhash.insert(g->mhash.end(), t.begin(), t.end());
        g->mpos.insert(g->mpos.end(), pos.begin(), pos.end());
        g->moff.push_back(g->mhash.size());
        for_each_kminmer(t, pos, p, [&](const Tuple& node, bool rev, uint64_t s0, uint64_t s1,
                                        uint64_t o0, uint64_t o1, uint64_t o2) {
            g->st.n_kminmers++;
            add_kminmer(table, next_index, p, node, rev, s0, s1, r, o0, o1, o2, &g->seqlines);
        });
    }
    g->st.n_minimizers = g->mhash.size();
    g->st.n_seqlines = g->seqlines.size();
    std::vector<Node> nodes;
    nodes.reserve(table.size());
    uint64_t in_table = 0;
    for (auto& kv : table) {
        if (kv.second.index == 0xFFFFFFFFu) continue;   // --bf: seen once, only in the filter
        in_table++;
        // main.rs:922-929: retain(abundance >= minabund) only when minabund > 1
        if (p.min_abundance > 1 && kv.second.abundance < (uint16_t)p.min_abundance) continue;
        nodes.push_back({kv.first, kv.second});
    }
    g->st.n_distinct = in_table;
    emit_graph(g, nodes);
    return g;
}

// Same algorithm in the reference's thread structure (main.rs:834: `threads` workers pull
// reads, each does extract -> windows -> add_kminmer against one shared concurrent map;
// then the single-threaded filter + edge pass).  Used only as the timed CPU baseline.
orc_graph* orc_build_mt(const uint8_t* bases, const uint64_t* read_off, uint64_t R,
                        const orc_params* pp, int threads) {
    orc_graph* g = new orc_graph();
    g->p = *pp;
    memset(&g->st, 0, sizeof(g->st));
    const orc_params& p = g->p;
    if (threads < 1) threads = 1;
    const int NSHARD = 1024;  // DashMap is a sharded RwLock<HashMap>; same idea
    struct Shard { std::mutex mu; std::unordered_map<Tuple, Entry, TupleHash> m; };
    std::vector<Shard> shards(NSHARD);
    std::atomic<uint32_t> node_index(0);
    std::atomic<uint64_t> next_read(0), n_min(0), n_kmm(0), n_hpc(0), n_bases(0);
    std::atomic<int> err(0);
    TupleHash hasher;
    auto worker = [&]() {
        std::vector<uint64_t> t, pos;
        uint64_t lm = 0, lk = 0, lh = 0, lb = 0;
        for (;;) {
            uint64_t r0 = next_read.fetch_add(16);
            if (r0 >= R) break;
            uint64_t r1 = std::min(R, r0 + 16);
            for (uint64_t r = r0; r < r1; r++) {
                const uint8_t* s = bases + read_off[r];
                uint64_t len = read_off[r + 1] - read_off[r];
                uint64_t nh = 0, bad = 0;
                lb += len;
                if (extract_density(s, len, p, t, pos, &nh, &bad) != 0) { err = 1; continue; }
                lh += nh;
                lm += t.size();
                for_each_kminmer(t, pos, p, [&](const Tuple& node, bool rev, uint64_t s0, uint64_t s1,
                                                uint64_t o0, uint64_t o1, uint64_t o2) {
                    lk++;
                    Shard& sh = shards[hasher(node) % NSHARD];
                    std::lock_guard<std::mutex> lock(sh.mu);
                    uint32_t idx_dummy = 0;
                    auto it = sh.m.find(node);
                    if (it == sh.m.end()) idx_dummy = node_index.fetch_add(1);
                    uint32_t local_next = idx_dummy;
                    add_kminmer(sh.m, local_next, p, node, rev, s0, s1, r, o0, o1, o2, nullptr);
                });
            }
        }
        n_min += lm; n_kmm += lk; n_hpc += lh; n_bases += lb;
    };
    std::vector<std::thread> th;
    for (int i = 0; i < threads; i++) th.emplace_back(worker);
    for (auto& x : th) x.join();
    g->st.n_reads = R;
    g->st.n_bases = n_bases; g->st.n_hpc_bases = n_hpc;
    g->st.n_minimizers = n_min; g->st.n_kminmers = n_kmm;
    if (err) g->st.error = -1;
    std::vector<Node> nodes;
    uint64_t distinct = 0;
    for (auto& sh : shards) {
        distinct += sh.m.size();
        for (auto& kv : sh.m) {
            if (p.min_abundance > 1 && kv.second.abundance < (uint16_t)p.min_abundance) continue;
            nodes.push_back({kv.first, kv.second});
        }
        sh.m.clear();
    }
    g->st.n_distinct = distinct;
    emit_graph(g, nodes);
    return g;
}

void orc_graph_free(orc_graph* g) { delete g; }
void orc_graph_stats(const orc_graph* g, orc_stats* out) { *out = g->st; }

void orc_graph_nodes(const orc_graph* g, uint32_t* index, uint16_t* abundance, uint32_t* seqlen,
                     uint16_t* shift, uint64_t* tuple) {
    const uint32_t k = g->p.k;
    for (size_t i = 0; i < g->nodes.size(); i++) {
        const Node& n = g->nodes[i];
        if (index) index[i] = n.e.index;
        if (abundance) abundance[i] = n.e.abundance;
        if (seqlen) seqlen[i] = n.e.seqlen;
        if (shift) { shift[2 * i] = n.e.shift0; shift[2 * i + 1] = n.e.shift1; }
        if (tuple) memcpy(tuple + i * k, n.t.data(), k * 8);
    }
}
void orc_graph_edges(const orc_graph* g, uint32_t* n1, uint8_t* o1, uint32_t* n2, uint8_t* o2,
                     uint32_t* overlap) {
    for (size_t i = 0; i < g->edges.size(); i++) {
        const Edge& e = g->edges[i];
        n1[i] = e.n1; o1[i] = e.o1; n2[i] = e.n2; o2[i] = e.o2; overlap[i] = e.ov;
    }
}
void orc_graph_seqlines(const orc_graph* g, uint32_t* index, uint64_t* read, uint64_t* start,
                        uint64_t* end, uint8_t* reversed, uint64_t* shift) {
    for (size_t i = 0; i < g->seqlines.size(); i++) {
        const SeqLine& s = g->seqlines[i];
        index[i] = s.index; read[i] = s.read; start[i] = s.start; end[i] = s.end;
        reversed[i] = s.reversed; shift[2 * i] = s.shift0; shift[2 * i + 1] = s.shift1;
    }
}
uint64_t orc_graph_minimizers(const orc_graph* g, uint64_t* hash, uint64_t* pos, uint64_t* read_off) {
    if (hash) memcpy(hash, g->mhash.data(), g->mhash.size() * 8);
    if (pos) memcpy(pos, g->mpos.data(), g->mpos.size() * 8);
    if (read_off) memcpy(read_off, g->moff.data(), g->moff.size() * 8);
    return g->mhash.size();
}

// Canonical parity form of {prefix}.gfa: header (main.rs:1011), sorted S lines
// (main.rs:1021), sorted L lines (main.rs:1095,1113).
// --read-stats, main.rs:939-975: for every k-min-mer of every read of a second read set, the
// abundance of its canonical tuple among the nodes that survived the abundance filter
// (`dbg_nodes` at that point, main.rs:922-933), 0 when absent (main.rs:961-966).
// out_off[r] = index of the first count of read r (R+1 entries); returns the number of counts, or
// -1 on an illegal base (the crate panics there); writes at most cap counts.
int64_t orc_read_stats(const orc_graph* g, const uint8_t* bases, const uint64_t* read_off, uint64_t R,
                       uint32_t* out_counts, uint64_t* out_off, uint64_t cap) {
    std::unordered_map<Tuple, uint16_t, TupleHash> ab;
    ab.reserve(g->nodes.size() * 2);
    for (const Node& n : g->nodes) ab[n.t] = n.e.abundance;
    std::vector<uint64_t> t, pos;
    uint64_t n = 0;
    for (uint64_t r = 0; r < R; r++) {
        if (out_off) out_off[r] = n;
        uint64_t nh = 0, bad = 0;
        if (extract_density(bases + read_off[r], read_off[r + 1] - read_off[r], g->p, t, pos, &nh, &bad) != 0) return -1;
        for_each_kminmer(t, pos, g->p, [&](const Tuple& node, bool, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t) {
            auto it = ab.find(node);
            if (n < cap && out_counts) out_counts[n] = it == ab.end() ? 0u : (uint32_t)it->second;
            n++;
        });
    }
    if (out_off) out_off[R] = n;
    return (int64_t)n;
}

int orc_write_gfa(const orc_graph* g, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "H\tVN:Z:1.0\n");
    std::vector<std::string> lines;
    char buf[256];
    for (auto& n : g->nodes) {
        snprintf(buf, sizeof buf, "S\t%u\t*\tLN:i:%u\tKC:i:%u\n", n.e.index, n.e.seqlen, (unsigned)n.e.abundance);
        lines.push_back(buf);
    }
    std::sort(lines.begin(), lines.end());
    for (auto& s : lines) fputs(s.c_str(), f);
    lines.clear();
    for (auto& e : g->edges) {
        snprintf(buf, sizeof buf, "L\t%u\t%c\t%u\t%c\t%uM\n", e.n1, e.o1 ? '-' : '+', e.n2, e.o2 ? '-' : '+', e.ov);
        lines.push_back(buf);
    }
    std::sort(lines.begin(), lines.end());
    for (auto& s : lines) fputs(s.c_str(), f);
    fclose(f);
    return 0;
}

// Canonical parity form of the {prefix}.*.sequences data lines (main.rs:702):
// "{index}\t{:?Vec<u64>}\t{seq}\t*\t*\t({s0}, {s1})", sorted, '#' header dropped.
int orc_write_sequences(const orc_graph* g, const uint8_t* bases, const uint64_t* read_off,
                        const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    std::vector<std::string> lines;
    for (auto& s : g->seqlines) {
        std::string line = std::to_string(s.index) + "\t[";
        for (size_t j = 0; j < s.node.size(); j++) {
            if (j) line += ", ";
            line += std::to_string(s.node[j]);
        }
        line += "]\t";
        std::string seq((const char*)bases + read_off[s.read] + s.start, s.end - s.start);
        if (s.reversed) {
            std::string rc(seq.size(), 'N');
            orc_revcomp((const uint8_t*)seq.data(), seq.size(), (uint8_t*)&rc[0]);
            seq.swap(rc);
        }
        line += seq;
        line += "\t*\t*\t(" + std::to_string(s.shift0) + ", " + std::to_string(s.shift1) + ")\n";
        lines.push_back(line);
    }
    std::sort(lines.begin(), lines.end());
    for (auto& s : lines) fputs(s.c_str(), f);
    fclose(f);
    return 0;
}

}  // extern "C"
