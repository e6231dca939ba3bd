"""Pins the CPU oracle (oracle/) against every golden vector available for this path:
 * the nthash crate's documented known-answer vectors,
 * two node lines EMITTED BY THE REFERENCE and quoted in its own sources
   (tests/golden/reference_sequences_lines.json, made by tests/golden/make_golden.py),
 * SURVEY Appendix B bound integers and Appendix C counts on BASELINE config #1,
 * a naive independent Python restatement on random inputs.
CPU only."""
import json
import os

import numpy as np
import pytest

from helpers import py_extract, py_kminmers, py_ntc64, random_reads, pack_reads, genome_reads
from conftest import GOLDEN


def test_nthash_crate_known_answers(oracle):
    assert [int(x) for x in oracle.nthash_iter("ACTGC", 3)] == [
        0x9b1eda9a185413ce, 0x9f6acfa2235b86fc, 0xd4a29bf149877c5c]
    assert oracle.ntf64("TGCAG", 0, 5) == 0x0bafa6728fc6dabf
    assert oracle.ntr64("TGCAG", 0, 5) == 0x8cf2d4072cca480e
    assert oracle.ntc64("ACGTC", 0, 5) == 0x480202d54e8ebecd
    assert oracle.ntc64("GACGT", 0, 5) == 0x480202d54e8ebecd


def test_nthash_rejects_non_acgtn(oracle):
    for bad in ("ACGTa", "ACG\nT", "ACGUX"):
        with pytest.raises(ValueError):
            oracle.nthash_iter(bad, 3)
    # 'N' hashes as 0 on both strands
    assert oracle.ntf64("NNN", 0, 3) == 0 and oracle.ntr64("NNN", 0, 3) == 0


def test_rolling_equals_direct(oracle):
    rng = np.random.default_rng(1)
    for k in (3, 10, 12, 14, 31, 64, 70):
        s = random_reads(rng, 1, mean=300, sd=0, lo=300, alphabet=b"ACGTN")[0]
        it = oracle.nthash_iter(s, k)
        assert len(it) == len(s) - k + 1
        for i in (0, 1, 17, len(s) - k):
            assert int(it[i]) == oracle.ntc64(s, i, k) == (py_ntc64(s, i, k) if k <= 64 else int(it[i]))


def test_hash_bound_integers(oracle):
    # SURVEY Appendix B
    assert oracle.hash_bound(0.0008) == 14757395258967642
    assert oracle.hash_bound(0.002) == 36893488147419104
    assert oracle.hash_bound(0.003) == 55340232221128656
    assert oracle.hash_bound(0.01) == 184467440737095520
    assert oracle.hash_bound(0.10) == 1844674407370955264
    assert oracle.hash_bound(1.0) == 2 ** 64 - 1      # Rust `as u64` saturates
    assert oracle.hash_bound(0.0) == 0


def test_reference_emitted_lines(oracle):
    """The reference's own output, quoted in its sources, reproduced exactly."""
    recs = json.load(open(os.path.join(GOLDEN, "reference_sequences_lines.json")))
    # record 0: to_basespace.rs:203, k=7, pre-HPC'd reads => --skiphpc semantics, l=10
    r = recs[0]
    for d in (0.008, 0.01):
        h, p = oracle.extract(r["sequence"], 10, d, hpc=False)
        assert [int(x) for x in h] == r["minimizers"]
        # sequence = raw[p_0 .. p_{k-1}+l)   (main.rs:778,700)
        assert int(p[0]) == 0 and int(p[-1]) + 10 == len(r["sequence"])
        # shift = (p1-p0, p_{k-1}-p_{k-2}) for a non-reversed node (main.rs:769-777)
        assert [int(p[1] - p[0]), int(p[-1] - p[-2])] == r["shift"]
    t, rev = oracle.normalize(np.array(r["minimizers"], dtype=np.uint64))
    assert not rev and [int(x) for x in t] == r["minimizers"]
    # record 1: scan_genomes_minmers.py:38, k=10, l=12, d=0.01
    r = recs[1]
    h, p = oracle.extract(r["sequence"], 12, 0.01, hpc=False)
    assert [int(x) for x in h] == r["minimizers"]
    assert int(p[0]) == 0 and int(p[-1]) + 12 == len(r["sequence"])


def test_encode_rle(oracle):
    h, p = oracle.encode_rle("AAACCGTTTTTA")
    assert h.tobytes() == b"ACGTA" and list(p) == [0, 3, 5, 6, 11]
    h, p = oracle.encode_rle("A")
    assert h.tobytes() == b"A" and list(p) == [0]
    h, p = oracle.encode_rle("NNNAANN")
    assert h.tobytes() == b"NAN" and list(p) == [0, 3, 5]
    # characters outside "ACTGactgNn" are not collapsed (read.rs:163)
    h, p = oracle.encode_rle("XXAA")
    assert h.tobytes() == b"XXA" and list(p) == [0, 1, 2]


def test_extract_vs_naive_python(oracle):
    rng = np.random.default_rng(7)
    seqs = random_reads(rng, 6, mean=1500, sd=600, lo=0, hp=0.3) + [b"", b"A", b"ACGTACGTAC", b"A" * 50]
    seqs += random_reads(rng, 2, mean=800, sd=10, alphabet=b"ACGTN", hp=0.2)
    for s in seqs:
        for (l, d, hpc) in ((10, 0.05, True), (12, 0.02, True), (12, 0.05, False), (5, 0.3, True)):
            h, p = oracle.extract(s, l, d, hpc=hpc)
            eh, ep = py_extract(s, l, d, hpc=hpc)
            assert [int(x) for x in h] == eh and [int(x) for x in p] == ep


def test_normalize_quirks(oracle):
    t, rev = oracle.normalize(np.array([1, 2, 3], dtype=np.uint64))
    assert not rev and list(t) == [1, 2, 3]
    t, rev = oracle.normalize(np.array([3, 2, 1], dtype=np.uint64))
    assert rev and list(t) == [1, 2, 3]
    t, rev = oracle.normalize(np.array([5, 9, 5], dtype=np.uint64))   # palindrome => reversed
    assert rev and list(t) == [5, 9, 5]
    t, rev = oracle.normalize(np.array([2 ** 63, 1], dtype=np.uint64))  # unsigned compare
    assert rev and [int(x) for x in t] == [1, 2 ** 63]


def test_revcomp(oracle):
    assert oracle.revcomp("ACGTNacgtuUX") == b"NAaacgtNACGT"


def test_graph_vs_naive_python(oracle):
    """Serial-order table semantics (main.rs:632-709) against a dict-based restatement."""
    rng = np.random.default_rng(3)
    seqs = genome_reads(rng, 20000, 60, mean=2500, sd=600, err=0.002)
    bases, off = pack_reads(seqs)
    k, l, d = 4, 8, 0.02
    for minab in (1, 2, 3):
        g = oracle.build_graph(bases, off, k, l, d, min_abundance=minab, presimp=0.0)
        table, order = {}, []
        nk = 0
        for s in seqs:
            hs, ps = py_extract(s, l, d)
            for node, rv, shift, offs in py_kminmers(hs, ps, k, l):
                nk += 1
                if node not in table:
                    table[node] = [len(order), 0, offs[2], shift]
                    order.append(node)
                e = table[node]
                if e[1] == minab - 1:
                    e[2], e[3] = offs[2], shift
                e[1] += 1
        assert g.stats["n_kminmers"] == nk and g.stats["n_distinct"] == len(table)
        kept = {n: e for n, e in table.items() if minab == 1 or e[1] >= minab}
        assert g.stats["n_nodes"] == len(kept)
        for i in range(len(g.index)):
            e = kept[tuple(int(x) for x in g.tuple[i])]
            assert (int(g.index[i]), int(g.abundance[i]), int(g.seqlen[i])) == (e[0], e[1], e[2])
            assert (int(g.shift[i, 0]), int(g.shift[i, 1])) == (e[3][0] & 0xffff, e[3][1] & 0xffff)


def test_example_config1_counts(oracle, example_reads):
    """BASELINE config #1; expected counts from SURVEY Appendix C (an independent numpy
    restatement by the surveyor) and the first-read minimizers listed there."""
    bases, off, names = example_reads
    assert len(names) == 657 and int(off[-1]) == 14744805
    g = oracle.build_graph(bases, off, 7, 10, 0.0008, min_abundance=2, presimp=0.01)
    st = g.stats
    assert st["error"] == 0
    assert st["n_hpc_bases"] == 10156492
    assert st["n_minimizers"] == 16069
    assert st["n_kminmers"] == 12127
    assert st["n_distinct"] == 104 and st["n_nodes"] == 104
    assert st["n_edges"] == 206 and st["presimp_removed"] == 0
    assert int(g.abundance.max()) == 198
    n0 = int(g.m_off[1])
    assert n0 == 36
    assert [int(x) for x in g.m_pos[:8]] == [490, 499, 1465, 3900, 4458, 5179, 5452, 5698]
    assert [int(x) for x in g.m_hash[:8]] == [
        7226384715172567, 9396755393060453, 7893203584393974, 14372547039170726,
        984557686330639, 9148316371314830, 12206407598145061, 6406161515958848]
    # k-min-mer 0 / 1 of read 0 (Appendix C): LN = 4964 / 5201, shifts (9,273) / (246,966)
    kms = py_kminmers([int(x) for x in g.m_hash[:n0]], [int(x) for x in g.m_pos[:n0]], 7, 10)
    assert kms[0][1] is False and kms[0][2] == (9, 273) and kms[0][3] == (490, 5462, 4964)
    assert kms[1][1] is True and kms[1][2] == (246, 966) and kms[1][3] == (499, 5708, 5201)


def test_mt_baseline_same_node_set(oracle, example_reads):
    """The threaded baseline build (reference thread structure) agrees with the serial oracle on
    everything that is scheduling-independent: node set, abundances, edge count."""
    bases, off, _ = example_reads
    a = oracle.build_graph(bases, off, 7, 10, 0.0008)
    b = oracle.build_graph(bases, off, 7, 10, 0.0008, threads=4)
    assert a.stats["n_nodes"] == b.stats["n_nodes"] and a.stats["n_edges"] == b.stats["n_edges"]
    sa = sorted((tuple(int(x) for x in a.tuple[i]), int(a.abundance[i])) for i in range(len(a.index)))
    sb = sorted((tuple(int(x) for x in b.tuple[i]), int(b.abundance[i])) for i in range(len(b.index)))
    assert sa == sb


def test_read_stats_against_naive_python(oracle):
    """--read-stats (main.rs:939-975): abundance of each k-min-mer of a second read set among the kept
    nodes, 0 when absent -- the oracle against the naive Python restatement of tests/helpers.py."""
    from helpers import genome_reads, pack_reads, py_extract, py_kminmers
    rng = np.random.default_rng(11)
    k, l, d = 4, 8, 0.05
    seqs = genome_reads(rng, 6000, 60, mean=900, sd=200, err=0.01)
    bases, off = pack_reads(seqs)
    g = oracle.build_graph(bases, off, k, l, d, 2, 0.01)
    kept = {tuple(int(x) for x in g.tuple[i]): int(g.abundance[i]) for i in range(len(g.index))}
    assert len(kept) > 20
    query = seqs[::3] + genome_reads(rng, 6000, 10, mean=900, sd=200) + [b"", b"ACGT", b"A" * 50]
    qb, qo = pack_reads(query)
    cnt, coff = g.read_stats(qb, qo)
    exp, eoff = [], [0]
    for s in query:
        hs, ps = py_extract(s, l, d)
        exp += [kept.get(tuple(node), 0) for node, _, _, _ in py_kminmers(hs, ps, k, l)]
        eoff.append(len(exp))
    assert list(cnt) == exp and list(coff) == eoff
    assert any(c > 0 for c in exp) and any(c == 0 for c in exp)


def test_oracle_reproduces_synthetic_digests(oracle):
    """tests/golden/synth_graph_hashes.json (BASELINE config 2 in full): the oracle of this checkout still
    produces the committed digests -- a change of the restatement cannot slip through unnoticed.  (The
    1.05 Gbase slice of config 3 is re-derived by the GPU test, which runs the oracle anyway.)"""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_synth_hashes as G
    gold = json.load(open(os.path.join(here, "golden", "synth_graph_hashes.json")))
    c, host, ro, total = G.build_case("ecoli50x_full", threads=8)
    o = oracle.build_graph(host, ro, c["k"], c["l"], c["density"], 2, 0.01)
    assert int(total) == gold["ecoli50x_full"]["n_bases"]
    assert G.graph_digest(o) == gold["ecoli50x_full"]["graph_sha256"]
    assert G.minimizer_digest(o.m_hash, o.m_pos, o.m_off) == gold["ecoli50x_full"]["minimizers_sha256"]
