"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI,
against the CPU oracle on the same inputs.  Bit-exact: integer hashes, positions, indices."""
import os

import numpy as np
import pytest

from helpers import genome_reads, pack_reads, random_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mdbg():
    import rust_mdbg_b200
    if rust_mdbg_b200.ffi.lib().mdbg_device_count() < 1:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return rust_mdbg_b200


def oracle_minimizers(oracle, seqs, l, d, hpc=True):
    hs, ps, off = [], [], [0]
    for s in seqs:
        h, p = oracle.extract(s, l, d, hpc=hpc)
        hs.append(h); ps.append(p); off.append(off[-1] + len(h))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint64)
    return cat(hs), cat(ps), np.array(off, np.uint64)


def check_extract(mdbg, oracle, seqs, l, d, hpc=True):
    bases, off = pack_reads(seqs)
    with mdbg.Context(mdbg.Params(k=5, l=l, density=d, hpc=hpc)) as ctx:
        h, p, mo = ctx.extract_minimizers(bases, off)
        dense = ctx.timings()["ka_dense_tiles"]
    eh, ep, eo = oracle_minimizers(oracle, seqs, l, d, hpc)
    assert np.array_equal(mo, eo), "per-read minimizer offsets differ"
    assert np.array_equal(h, eh), "hashes differ"
    assert np.array_equal(p, ep), "positions differ"
    return dense


EDGE_READS = [b"", b"A", b"ACGTACGTAC", b"A" * 5000, b"AC" * 4000, b"ACG" * 3000, b"ACGTACGTACG",
              b"", b"", b"T" * 70000, b"ACGT" * 5000 + b"A" * 300 + b"CGTA" * 100]


@pytest.mark.parametrize("l,d", [(10, 0.0008), (12, 0.003), (12, 0.002), (14, 0.01), (5, 0.02), (15, 0.005)])
def test_extract_random_reads(mdbg, oracle, l, d):
    rng = np.random.default_rng(100 + l)
    seqs = random_reads(rng, 40, mean=9000, sd=4000, lo=0, hi=40000, hp=0.25) + EDGE_READS
    seqs += random_reads(rng, 300, mean=40, sd=30, lo=0, hi=200)          # many tiny reads in one tile
    dense = check_extract(mdbg, oracle, seqs, l, d)
    if d <= 0.005 and l <= 14:   # l = 15 is outside the filter's range: exact path by design
        assert dense <= 16   # only the short-period repeat reads may overflow the candidate queue


def test_extract_skiphpc(mdbg, oracle):
    rng = np.random.default_rng(5)
    seqs = random_reads(rng, 30, mean=7000, sd=3000, hp=0.3) + EDGE_READS
    check_extract(mdbg, oracle, seqs, 12, 0.003, hpc=False)


def test_extract_with_N(mdbg, oracle):
    rng = np.random.default_rng(6)
    seqs = random_reads(rng, 20, mean=8000, sd=2000, hp=0.2)
    out = []
    for i, s in enumerate(seqs):
        a = bytearray(s)
        for _ in range(i % 5):
            j = int(rng.integers(0, len(a)))
            n = int(rng.integers(1, 40))
            a[j:j + n] = b"N" * len(a[j:j + n])
        out.append(bytes(a))
    check_extract(mdbg, oracle, out + [b"N" * 3000, b"ACGT" * 10 + b"N" + b"ACGT" * 10], 12, 0.01)


def test_extract_dense_modes(mdbg, oracle):
    """Densities / l outside the filter's range take the exact per-position path."""
    rng = np.random.default_rng(8)
    seqs = random_reads(rng, 10, mean=3000, sd=1000, hp=0.2) + EDGE_READS[:6]
    check_extract(mdbg, oracle, seqs, 12, 0.10)     # CLI default density (main.rs:439)
    check_extract(mdbg, oracle, seqs, 12, 1.0)
    check_extract(mdbg, oracle, seqs, 21, 0.005)    # l > 15
    check_extract(mdbg, oracle, seqs, 31, 0.02)


def test_extract_rejects_bad_alphabet(mdbg):
    with mdbg.Context(mdbg.Params(k=5, l=10, density=0.01)) as ctx:
        good = b"ACGTTGCATGCATGACTGACTAGCTAGCATCGATCAGCTACGACTAGC" * 10
        for bad in (good[:100] + b"a" + good[100:], good + b"\n" + good, b"ACGTRYACGTACGTACGATCGATCGATGCATGC"):
            with pytest.raises(mdbg.MdbgError) as e:
                ctx.read_extract(bad)
            assert e.value.code == -4
        # a read too short to be hashed is never inspected by the reference (read.rs:193)
        h, p = ctx.read_extract(b"ACGXACG")
        assert len(h) == 0


def test_extract_example_config1(mdbg, oracle, example_reads):
    bases, off, _ = example_reads
    with mdbg.Context(mdbg.Params(k=7, l=10, density=0.0008)) as ctx:
        h, p, mo = ctx.extract_minimizers(bases, off)
    g = oracle.build_graph(bases, off, 7, 10, 0.0008)
    assert len(h) == 16069
    assert np.array_equal(h, g.m_hash) and np.array_equal(p, g.m_pos) and np.array_equal(mo, g.m_off)


def compare_graph(g, o, check_seqlines=True):
    for key in ("n_reads", "n_minimizers", "n_kminmers", "n_distinct", "n_nodes", "n_edges", "presimp_removed"):
        assert g.stats[key] == o.stats[key], key
    assert np.array_equal(g.index, o.index)
    assert np.array_equal(g.abundance, o.abundance)
    assert np.array_equal(g.seqlen, o.seqlen)
    assert np.array_equal(g.shift, o.shift)
    assert np.array_equal(g.tuple, o.tuple)
    for a in ("e_n1", "e_n2", "e_o1", "e_o2", "e_ov"):
        assert np.array_equal(getattr(g, a), getattr(o, a)), a
    if check_seqlines:
        assert g.stats["n_seqlines"] == o.stats["n_seqlines"]
        for a in ("q_index", "q_read", "q_start", "q_end", "q_rev", "q_shift"):
            assert np.array_equal(getattr(g, a), getattr(o, a)), a


def test_graph_example_config1(mdbg, oracle, example_reads, tmp_path):
    """BASELINE config #1 end to end, including the written .gfa / .sequences (sorted-line equality)."""
    bases, off, _ = example_reads
    with mdbg.Context(mdbg.Params(k=7, l=10, density=0.0008, min_abundance=2, presimp=0.01)) as ctx:
        ctx.push_reads(bases, off)
        cg = ctx.finish_raw()
        g = mdbg.Graph(cg)
        L = mdbg.ffi.lib()
        import ctypes
        gfa, seq = str(tmp_path / "x.gfa"), str(tmp_path / "x.sequences")
        assert L.mdbg_write_gfa(ctypes.byref(cg), gfa.encode()) == 0
        assert L.mdbg_write_sequences(ctypes.byref(cg), bases.ctypes.data, off.ctypes.data, seq.encode(), 0) == 0
        ctx.graph_free(cg)
    o = oracle.build_graph(bases, off, 7, 10, 0.0008, 2, 0.01)
    assert g.stats["n_nodes"] == 104 and g.stats["n_edges"] == 206
    compare_graph(g, o)
    ogfa, oseq = str(tmp_path / "o.gfa"), str(tmp_path / "o.sequences")
    o.write_gfa(ogfa); o.write_sequences(oseq)
    canon = lambda path, skip: sorted(x for x in open(path) if not x.startswith(skip))
    assert open(gfa).readline() == "H\tVN:Z:1.0\n"
    assert canon(gfa, "H") == canon(ogfa, "H")
    assert canon(seq, "#") == canon(oseq, "#")


@pytest.mark.parametrize("k,l,d,minab,presimp,err", [
    (4, 8, 0.02, 2, 0.0, 0.0), (5, 10, 0.01, 2, 0.01, 0.002), (7, 10, 0.01, 1, 0.5, 0.003),
    (10, 12, 0.02, 3, 0.3, 0.001), (21, 12, 0.05, 2, 0.01, 0.001), (3, 6, 0.05, 2, 0.4, 0.01)])
def test_graph_random_genomes(mdbg, oracle, k, l, d, minab, presimp, err):
    rng = np.random.default_rng(k * 100 + l)
    seqs = genome_reads(rng, 60000, 300, mean=6000, sd=2500, err=err) + [b"", b"ACGT"]
    bases, off = pack_reads(seqs)
    with mdbg.Context(mdbg.Params(k=k, l=l, density=d, min_abundance=minab, presimp=presimp)) as ctx:
        ctx.push_reads(bases, off)
        g = ctx.finish()
    o = oracle.build_graph(bases, off, k, l, d, minab, presimp)
    assert o.stats["n_nodes"] > 0
    compare_graph(g, o)


def test_graph_two_pushes_and_multik(mdbg, oracle):
    """Batches accumulate in read order; k can be swept over the resident minimizers."""
    rng = np.random.default_rng(77)
    seqs = genome_reads(rng, 40000, 200, mean=5000, sd=2000, err=0.002)
    bases, off = pack_reads(seqs)
    b1, o1 = pack_reads(seqs[:90]); b2, o2 = pack_reads(seqs[90:])
    with mdbg.Context(mdbg.Params(k=5, l=10, density=0.01)) as ctx:
        ctx.push_reads(b1, o1); ctx.push_reads(b2, o2)
        for k in (5, 8, 12):
            ctx.set_k(k)
            compare_graph(ctx.finish(), oracle.build_graph(bases, off, k, 10, 0.01))
        ctx.reset()
        ctx.set_k(6)
        ctx.push_reads(b2, o2)
        compare_graph(ctx.finish(), oracle.build_graph(b2, o2, 6, 10, 0.01))


def test_graph_fingerprint_collision_path(mdbg, oracle):
    """Force fingerprint collisions (8-bit fingerprints on the first attempt): the table must
    detect them on the tuples and retry with a new seed, and the result stays exact."""
    rng = np.random.default_rng(9)
    seqs = genome_reads(rng, 30000, 150, mean=5000, sd=1500, err=0.002)
    bases, off = pack_reads(seqs)
    with mdbg.Context(mdbg.Params(k=6, l=10, density=0.01, debug_fp_bits=8)) as ctx:
        ctx.push_reads(bases, off)
        g = ctx.finish()
        assert ctx.timings()["table_attempts"] == 2
    compare_graph(g, oracle.build_graph(bases, off, 6, 10, 0.01))


def test_window_entry2(mdbg, oracle):
    from helpers import py_kminmers
    rng = np.random.default_rng(11)
    seqs = random_reads(rng, 12, mean=6000, sd=3000, hp=0.2) + [b"ACGT" * 3]
    k, l, d = 6, 10, 0.01
    with mdbg.Context(mdbg.Params(k=k, l=l, density=d)) as ctx:
        bases, off = pack_reads(seqs)
        h, p, mo = ctx.extract_minimizers(bases, off)
        tup, rev, sh, of, ko = ctx.window(h, p, mo)
        r0 = mdbg.Read.extract("r0", seqs[0], ctx)
        kms = r0.read_to_kmers(ctx)
    exp = []
    for r in range(len(seqs)):
        exp += py_kminmers([int(x) for x in h[int(mo[r]):int(mo[r + 1])]], [int(x) for x in p[int(mo[r]):int(mo[r + 1])]], k, l)
    assert len(exp) == len(rev)
    for i, (node, rv, shift, offs) in enumerate(exp):
        assert tuple(int(x) for x in tup[i]) == node and bool(rev[i]) == rv
        assert (int(sh[i, 0]), int(sh[i, 1])) == shift and tuple(int(x) for x in of[i]) == offs
    e0 = py_kminmers([int(x) for x in r0.transformed], [int(x) for x in r0.minimizers_pos], k, l)
    assert [(tuple(int(x) for x in a.data), b, c, d_) for a, b, c, d_ in kms] == e0


def test_synth_device_equals_host(mdbg):
    s = mdbg.Synth(genome_len=200000, mean_len=3000, sd_len=900, min_len=200, max_len=12000, error_rate=0.01)
    n = s.num_reads(3.0)
    ro, total = s.plan(5, n)
    host = s.fill_host(5, n, ro)
    with mdbg.Context(mdbg.Params(k=5, l=10, density=0.01)) as ctx:
        d_b = ctx.device_malloc(total + 16); d_o = ctx.device_malloc((n + 1) * 8)
        s.fill_device(ctx, 5, n, ro, d_b, d_o)
        dev = np.zeros(total, np.uint8)
        ctx.d2h(dev, d_b)
        ctx.device_free(d_b); ctx.device_free(d_o)
    assert np.array_equal(host, dev)
    assert set(np.unique(host)) <= set(b"ACGT")


def test_full_size_properties(mdbg, oracle):
    """BASELINE config 2 shape (synthetic E. coli-like, k=21 l=12 d=0.003) at a size the oracle
    cannot check directly in the GPU-suite's budget: size-independent properties + a sampled
    oracle check + determinism (two runs byte-identical)."""
    s = mdbg.Synth(genome_len=1000000)
    n = s.num_reads(20.0)
    ro, total = s.plan(0, n)
    host = s.fill_host(0, n, ro)
    P = mdbg.Params(k=21, l=12, density=0.003)
    with mdbg.Context(P) as ctx:
        ctx.push_reads(host, ro)
        h, p, mo = ctx.get_minimizers()
        g1 = ctx.finish()
        g2 = ctx.finish()
    # sampled per-read oracle check
    for r in range(0, n, max(1, n // 40)):
        eh, ep = oracle.extract(host[int(ro[r]):int(ro[r + 1])], 12, 0.003)
        assert np.array_equal(h[int(mo[r]):int(mo[r + 1])], eh) and np.array_equal(p[int(mo[r]):int(mo[r + 1])], ep)
    bound = mdbg.minimizers.hash_bound(0.003)
    assert h.max() <= bound
    lens = np.diff(ro.astype(np.int64)); m = np.diff(mo.astype(np.int64))
    assert np.all(np.diff(mo.astype(np.int64)) >= 0)
    for r in range(n):   # positions strictly increase inside a read and stay inside it
        pr = p[int(mo[r]):int(mo[r + 1])].astype(np.int64)
        assert np.all(np.diff(pr) > 0) and (len(pr) == 0 or pr[-1] + 12 <= lens[r])
    assert g1.stats["n_kminmers"] == int(np.where(m > 21, m - 21 + 1, 0).sum())
    # node table invariants
    assert np.all(np.diff(g1.index.astype(np.int64)) > 0) and g1.index.max() < g1.stats["n_distinct"]
    assert np.all(g1.abundance >= 2)
    fwd_le_rev = [tuple(t) <= tuple(t[::-1]) for t in g1.tuple[:2000].tolist()]
    assert all(fwd_le_rev)
    # edges reference existing nodes; determinism
    idx = set(g1.index.tolist())
    assert set(g1.e_n1.tolist()) <= idx and set(g1.e_n2.tolist()) <= idx
    for a in ("index", "abundance", "seqlen", "shift", "tuple", "e_n1", "e_n2", "e_o1", "e_o2", "e_ov"):
        assert np.array_equal(getattr(g1, a), getattr(g2, a))


def test_multi_gpu_equals_oracle(mdbg):
    """N-GPU graph (reads sharded by record, NCCL all-to-all by fingerprint range) == oracle.
    Needs >= 2 GPUs on the box; tests/multi_gpu_check.py is the program run under torchrun."""
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _lz4_stored_frame_decode(raw):
    """Minimal LZ4 frame reader (only what a conforming decoder needs for stored blocks)."""
    assert raw[:4] == b"\x04\x22\x4d\x18"
    flg, bd = raw[4], raw[5]
    assert flg >> 6 == 1 and not (flg & 0x08) and not (flg & 0x04) and not (flg & 0x01)   # v1, no size/checksum/dict
    import struct

    def xxh32(data, seed=0):
        P1, P2, P3, P4, P5, M = 2654435761, 2246822519, 3266489917, 668265263, 374761393, 0xFFFFFFFF
        rot = lambda x, r: ((x << r) | (x >> (32 - r))) & M
        h = (seed + P5 + len(data)) & M
        for b in data:
            h = (rot((h + b * P5) & M, 11) * P1) & M
        h ^= h >> 15; h = (h * P2) & M; h ^= h >> 13; h = (h * P3) & M; h ^= h >> 16
        return h
    assert raw[6] == (xxh32(raw[4:6]) >> 8) & 0xFF, "bad frame header checksum"
    out, p = bytearray(), 7
    while True:
        (sz,) = struct.unpack_from("<I", raw, p); p += 4
        if sz == 0:
            break
        assert sz & 0x80000000, "block is not stored"
        sz &= 0x7FFFFFFF
        out += raw[p:p + sz]; p += sz
    assert p == len(raw)
    return bytes(out)


def test_cli_example_config1(mdbg, oracle, example_reads, tmp_path):
    """The flag-compatible front end on BASELINE config #1 (README.md:40 of the reference):
    same stdout counters, sorted .gfa lines and LZ4-framed .sequences lines as the oracle."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rust-mdbg_b200", "rust-mdbg")
    prefix = str(tmp_path / "example")
    r = subprocess.run([exe, os.path.join(root, "tests", "golden", "config1_reads.fa.gz"), "-k", "7", "--density", "0.0008",
                        "-l", "10", "--minabund", "2", "--prefix", prefix], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for line in ("Format: FASTA", "Parsing input sequences...", "Number of reads: 657",
                 "Number of nodes before abundance filter: 104", "Number of nodes after abundance filter: 104",
                 "Number of mdBG edges: 206", "Pre-simp = 0.01: 0 edges removed."):
        assert line in r.stdout, (line, r.stdout)
    assert "Converted reads to k-min-mers." in r.stderr
    bases, off, _ = example_reads
    o = oracle.build_graph(bases, off, 7, 10, 0.0008, 2, 0.01)
    ogfa, oseq = str(tmp_path / "o.gfa"), str(tmp_path / "o.sequences")
    o.write_gfa(ogfa); o.write_sequences(oseq)
    canon = lambda text, skip: sorted(x for x in text.splitlines(True) if not x.startswith(skip))
    assert canon(open(prefix + ".gfa").read(), "H") == canon(open(ogfa).read(), "H")
    seq = _lz4_stored_frame_decode(open(prefix + ".0.sequences", "rb").read()).decode()
    assert seq.startswith("# k = 7\n# l = 10\n# Structure of remaining of the file:\n# [node name]\t[list of minimizers]")
    assert canon(seq, "#") == canon(open(oseq).read(), "#")
    # refused modes fail loudly
    r = subprocess.run([exe, "x.fa", "--syncmers"], capture_output=True, text=True)
    assert r.returncode != 0 and "outside the reads->mdBG hot path" in r.stderr


@pytest.mark.parametrize("k,l,d,minab", [(5, 10, 0.01, 2), (8, 10, 0.02, 3), (6, 10, 0.01, 1)])
def test_graph_bf_numbering(mdbg, oracle, k, l, d, minab):
    """--bf (main.rs:639-655) with an ideal filter: tuples enter the table at their second
    sighting; with minabund == 1 the flag is ignored, as in the reference."""
    rng = np.random.default_rng(4242 + k)
    seqs = genome_reads(rng, 50000, 250, mean=5000, sd=2000, err=0.004)
    bases, off = pack_reads(seqs)
    with mdbg.Context(mdbg.Params(k=k, l=l, density=d, min_abundance=minab, presimp=0.01, bf=True)) as ctx:
        ctx.push_reads(bases, off)
        g = ctx.finish()
    o = oracle.build_graph(bases, off, k, l, d, minab, 0.01, bf=True)
    plain = oracle.build_graph(bases, off, k, l, d, minab, 0.01)
    assert o.stats["n_nodes"] == plain.stats["n_nodes"] > 0
    if minab > 1:
        assert o.stats["n_distinct"] < plain.stats["n_distinct"]
    compare_graph(g, o)


def test_degenerate_inputs(mdbg, oracle):
    """No reads, only empty reads, reads too short to window: every stage must cope with zero items."""
    with mdbg.Context(mdbg.Params(k=5, l=10, density=0.01)) as ctx:
        for seqs in ([], [b""], [b"", b"", b""], [b"ACGT"], [b"ACGTTGCAAC" * 30]):
            ctx.reset()
            bases, off = pack_reads(seqs)
            ctx.push_reads(bases, off)
            h, p, mo = ctx.get_minimizers()
            eh, ep, eo = oracle_minimizers(oracle, seqs, 10, 0.01)
            assert np.array_equal(h, eh) and np.array_equal(p, ep) and np.array_equal(mo, eo)
            g = ctx.finish()
            o = oracle.build_graph(bases, off, 5, 10, 0.01)
            compare_graph(g, o)
        # an empty batch between two real ones keeps the read numbering
        rng = np.random.default_rng(3)
        seqs = genome_reads(rng, 20000, 60, mean=4000, sd=1000, err=0.002)
        ctx.reset()
        b1, o1 = pack_reads(seqs[:30]); b0, o0 = pack_reads([]); b2, o2 = pack_reads(seqs[30:])
        ctx.push_reads(b1, o1); ctx.push_reads(b0, o0); ctx.push_reads(b2, o2)
        bases, off = pack_reads(seqs)
        compare_graph(ctx.finish(), oracle.build_graph(bases, off, 5, 10, 0.01))


@pytest.mark.parametrize("mode,gbps", [("ascii", None), ("packed", None), ("hybrid", "1000"), ("hybrid", "3"), ("hybrid", None)])
def test_upload_modes_equal_oracle(mdbg, oracle, mode, gbps, monkeypatch):
    """mdbg_push_reads moves host buffers as ASCII, as 2-bit planes packed on the host and expanded on the
    device, or as a mix decided chunk by chunk: all three must give the oracle's minimizers and graph.
    256 KiB chunks so that a 3 MB batch crosses many chunk boundaries (N, homopolymers and tiny reads on them)."""
    monkeypatch.setenv("MDBG_UPLOAD", mode)
    monkeypatch.setenv("MDBG_UPLOAD_CHUNK_MB", "1")
    if gbps:
        monkeypatch.setenv("MDBG_PCIE_GBPS", gbps)
    else:
        monkeypatch.delenv("MDBG_PCIE_GBPS", raising=False)
    rng = np.random.default_rng(77)
    seqs = genome_reads(rng, 200000, 260, mean=11000, sd=3000, err=0.002)
    seqs[3] = seqs[3][:500] + b"N" * 37 + seqs[3][537:]
    seqs[40] = b"A" * 300000 + seqs[40]                     # a homopolymer across a chunk boundary
    seqs[41] = b"N" * 5000
    seqs[100:100] = [b"", b"ACGT", b"T" * 9000]
    bases, off = pack_reads(seqs)
    assert len(bases) > 10 * 262144
    k, l, d = 8, 12, 0.003
    with mdbg.Context(mdbg.Params(k=k, l=l, density=d, min_abundance=2, presimp=0.01)) as ctx:
        ctx.push_reads(bases, off)
        tm = ctx.timings()
        h, p, mo = ctx.get_minimizers()
        g = ctx.finish()
    assert tm["upload_packed"] == (0 if mode == "ascii" else 1)
    if mode == "packed":
        assert 0 < tm["upload_ascii_tiles"] < 20            # only the tiles holding N
        assert tm["upload_h2d_bytes"] < len(bases) // 3
    o = oracle.build_graph(bases, off, k, l, d, 2, 0.01)
    assert np.array_equal(h, o.m_hash) and np.array_equal(p, o.m_pos) and np.array_equal(mo, o.m_off)
    compare_graph(g, o, check_seqlines=False)


@pytest.mark.parametrize("chunk_mb", ["1", None])
def test_push_reads_packed_equals_oracle(mdbg, oracle, chunk_mb, monkeypatch):
    """mdbg_push_reads_packed: the caller's own 2-bit planes (mdbg_pack_bases_host) instead of ASCII bases -- same
    minimizers and graph as the oracle on the bases they encode; a quarter of the bytes cross PCIe.  Two pushes
    (the second one starts in the middle of a plane word of the first batch's length), homopolymers across chunk
    boundaries, tiny and empty reads, a batch whose length is no multiple of 32."""
    if chunk_mb:
        monkeypatch.setenv("MDBG_UPLOAD_CHUNK_MB", chunk_mb)
    else:
        monkeypatch.delenv("MDBG_UPLOAD_CHUNK_MB", raising=False)
    rng = np.random.default_rng(78)
    seqs = genome_reads(rng, 150000, 200, mean=9000, sd=3000, err=0.002)
    seqs[40] = b"A" * 300000 + seqs[40]
    seqs[100:100] = [b"", b"ACGT", b"T" * 9000, b"G"]
    half = len(seqs) // 2
    b0, o0 = pack_reads(seqs[:half])
    b1, o1 = pack_reads(seqs[half:] + [b"ACGTTGCAAC" * 3 + b"A"])
    assert len(b0) % 32 != 0 and len(b1) % 32 != 0
    k, l, d = 8, 12, 0.003
    with mdbg.Context(mdbg.Params(k=k, l=l, density=d, min_abundance=2, presimp=0.01)) as ctx:
        for b, o in ((b0, o0), (b1, o1)):
            planes, bad = mdbg.pack_bases(b, threads=4)
            assert not bad.any()
            ctx.push_reads_packed(planes, o)
            tm = ctx.timings()
            assert tm["upload_packed"] == 1 and tm["upload_ascii_tiles"] == 0
            assert tm["upload_h2d_bytes"] == 8 * ((len(b) + 31) // 32)
        h, p, mo = ctx.get_minimizers()
        g = ctx.finish()
    bases, off = pack_reads(seqs + [b"ACGTTGCAAC" * 3 + b"A"])
    o = oracle.build_graph(bases, off, k, l, d, 2, 0.01)
    assert np.array_equal(h, o.m_hash) and np.array_equal(p, o.m_pos) and np.array_equal(mo, o.m_off)
    compare_graph(g, o, check_seqlines=False)


def test_push_reads_packed_degenerate(mdbg):
    with mdbg.Context(mdbg.Params(k=5, l=12, density=0.003)) as ctx:
        ctx.push_reads_packed(np.zeros(0, np.uint32), np.zeros(1, np.uint64))            # no reads
        ctx.push_reads_packed(np.zeros(0, np.uint32), np.zeros(4, np.uint64))            # three empty reads
        planes, _ = mdbg.pack_bases(np.frombuffer(b"ACGTT", np.uint8))
        ctx.push_reads_packed(planes, np.array([0, 5], np.uint64))                        # shorter than l
        h, p, mo = ctx.get_minimizers()
        assert len(h) == 0 and list(mo) == [0, 0, 0, 0, 0]
        with pytest.raises(mdbg.MdbgError):
            ctx.push_reads_packed(planes, np.array([0, 5, 3], np.uint64))                 # offsets must not decrease


def test_push_argument_limits(mdbg):
    """The documented limits of a push are refused before any byte is touched: offsets that decrease, a record of
    4 Gbases or more (positions inside a read are u32 on the device) -- on the host entries and, checked by a kernel,
    on the device-resident entry."""
    small = np.frombuffer(b"ACGT" * 64, np.uint8).copy()
    with mdbg.Context(mdbg.Params(k=5, l=12, density=0.003)) as ctx:
        planes = mdbg.pack_bases(small)[0]
        for ro in ([0, 200, 100], [0, 1 << 32], [0, 100, (1 << 32) + 200]):
            o = np.array(ro, np.uint64)
            for push in (lambda: ctx.push_reads(small, o),
                         lambda: ctx.push_reads_packed_ptr(mdbg.ffi.ptr(planes), mdbg.ffi.ptr(o), len(ro) - 1)):
                with pytest.raises(mdbg.MdbgError) as ei:
                    push()
                assert ei.value.code in (-3, -6)              # MDBG_ERR_BAD_ARG / MDBG_ERR_RANGE
        d_b = ctx.device_malloc(4096); d_o = ctx.device_malloc(64)
        ctx.h2d(d_b, small)
        for ro, n_bases in (([0, 200, 100], 100), ([0, 128, 200], 256), ([0, 1 << 32], 1 << 32)):
            ctx.h2d(d_o, np.array(ro, np.uint64))
            with pytest.raises(mdbg.MdbgError) as ei:
                ctx.push_reads_device(d_b, d_o, len(ro) - 1, n_bases)
            assert ei.value.code in (-3, -6)
        # the context is still usable
        ctx.push_reads(small, np.array([0, 256], np.uint64))
        h, p, mo = ctx.get_minimizers()
        assert list(mo)[0] == 0 and len(mo) == 2
        ctx.device_free(d_b); ctx.device_free(d_o)

