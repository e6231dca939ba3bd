"""CPU-side checks of the product (no GPU needed): the C-ABI library loads and exports every
symbol include/mdbg.h declares, host-side value helpers mirror the reference's KmerVec, the
filter constants of the K-A kernel are a true superset test, and there is no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def mdbg():
    import __graft_entry__ as ge
    ge.build()
    import rust_mdbg_b200
    return rust_mdbg_b200


def test_header_symbols_exported(mdbg):
    hdr = open(os.path.join(ROOT, "include", "mdbg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mdbg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    L = ctypes.CDLL(mdbg.ffi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(mdbg.ffi.SYMBOLS), declared ^ set(mdbg.ffi.SYMBOLS)


def test_no_cpu_fallback(mdbg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mdbg.MdbgError) as e:
        mdbg.Context(mdbg.Params(k=7, l=10, density=0.01))
    assert e.value.code == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rust-mdbg_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower(), os.path.join(dp, f)


def test_hash_bound_matches_reference_cast(mdbg, oracle):
    for d in (0.0008, 0.002, 0.003, 0.01, 0.1, 0.5, 1.0, 2.0, 1e-12):
        assert mdbg.minimizers.hash_bound(d) == oracle.hash_bound(d)


def test_kmervec_surface(mdbg, oracle):
    K = mdbg.KmerVec
    rng = np.random.default_rng(0)
    for _ in range(200):
        k = int(rng.integers(2, 40))
        t = rng.integers(0, 2 ** 63, k, dtype=np.uint64) * 2 + rng.integers(0, 2, k, dtype=np.uint64)
        if rng.random() < 0.2:
            t[k // 2:] = t[:k - k // 2][::-1]       # palindromes
        a = K.make_from(t)
        n, rev = a.normalize()
        on, orev = oracle.normalize(t)
        assert rev == orev and np.array_equal(n.data, on)
        assert np.array_equal(a.reverse().data, t[::-1])
        assert np.array_equal(a.prefix().data, t[:-1]) and np.array_equal(a.suffix().data, t[1:])
        b = K.make_from(t[::-1])
        assert (a < b) == (tuple(t.tolist()) < tuple(t[::-1].tolist()))
        assert (a == b) == (tuple(t.tolist()) == tuple(t[::-1].tolist()))
    assert K.make_from([1, 22, 333]).print_as_string() == "[1, 22, 333]"


def test_shard_and_owner_helpers(mdbg):
    L = mdbg.ffi.lib()
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            got, prev = 0, 0
            for r in range(world):
                lo, hi = ctypes.c_uint64(), ctypes.c_uint64()
                L.mdbg_shard_reads(n, world, r, ctypes.byref(lo), ctypes.byref(hi))
                assert lo.value == prev and hi.value >= lo.value
                prev = hi.value
                got += hi.value - lo.value
            assert got == n and prev == n
    rng = np.random.default_rng(1)
    fps = rng.integers(0, 2 ** 63, 4000, dtype=np.uint64) * 2
    for world in (1, 2, 3, 8):
        own = np.array([L.mdbg_owner_of_fingerprint(int(f), world) for f in fps])
        assert own.min() >= 0 and own.max() < world
        assert np.all(np.diff(own[np.argsort(fps)]) >= 0)     # prefix (range) partition
        if world > 1:
            assert np.bincount(own, minlength=world).min() > 4000 / world / 2


def test_filter_is_superset(oracle):
    """Python model of the 32-bit rolling filter used by the K-A kernel (mdbg_common.cuh):
    every true minimizer must pass, and the rolled state must equal the closed form."""
    H = {0: 0x3c8bfbb395c60474, 1: 0x3193c18562a02b4c, 2: 0x295549f54be24456, 3: 0x20323ed082572324}  # A C T G
    RC = {0: H[2], 1: H[3], 2: H[0], 3: H[1]}
    M32 = 0xFFFFFFFF
    rng = np.random.default_rng(2)
    for l, d in ((10, 0.0008), (12, 0.003), (12, 0.002), (15, 0.01), (5, 0.02)):
        bound = oracle.hash_bound(d)
        bh = bound >> 32
        fth, gm, gth = bh | ((1 << (l - 1)) - 1), M32 >> (l - 1), bh >> (l - 1)
        gz = gm & ~((1 << gth.bit_length()) - 1)      # exact bits of G that must be zero (make_filter)
        s = rng.integers(0, 4, 30000)
        txt = bytes(b"ACTG"[c] for c in s)
        hashes = oracle.nthash_iter(txt, l)
        F = G = 0
        for j in range(l):                      # phantom A's
            F ^= ((H[0] >> 32) << (l - 1 - j)) & M32
            G ^= (RC[0] >> 32) >> (l - 1 - j)
        hist = [0] * l
        npass = 0
        for i, c in enumerate(s):
            out = hist[i % l]; hist[i % l] = int(c)
            F = ((F << 1) ^ (((H[out] >> 32) << l) & M32) ^ (H[int(c)] >> 32)) & M32
            G = (G >> 1) ^ ((RC[out] >> 32) >> l) ^ (RC[int(c)] >> 32)
            if i >= l - 1:
                hv = int(hashes[i - l + 1])
                passed = F <= fth or (G & gz) == 0
                assert not ((G & gm) <= gth) or (G & gz) == 0
                npass += passed
                if hv <= bound:
                    assert passed, (l, d, i)
        true = int((hashes <= np.uint64(bound)).sum())
        assert true <= npass <= true * 2.2 + 20, (l, d, true, npass)


def test_alphabet_predicate_model():
    """Word-parallel A/C/G/T check of the K-A kernel (bad_accumulate in ka_minimizers.cu), modelled
    bit for bit: byte j of the result is non-zero iff byte j of the word is not one of ACGT."""
    M = 0xFFFFFFFF

    def badword(w):
        s4, s3, s2 = (w << 4) & M, (w << 3) & M, (w << 2) & M
        k, y, x = w ^ 0x40404040, (~(s4 ^ w)) & M, (w ^ (s2 & ~s3)) & M
        return (k & 0xE8E8E8E8) | ((y | x) & 0x10101010)

    for lane in range(4):
        for c in range(256):
            for fill in b"ACGT":
                w = 0
                for j in range(4):
                    w |= (c if j == lane else fill) << (8 * j)
                bw = badword(w)
                for j in range(4):
                    assert (((bw >> (8 * j)) & 0xFF) != 0) == (j == lane and chr(c) not in "ACGT")


def _lz4_stored_frame_decode(raw):
    import struct

    def xxh32(data, seed=0):
        P1, P2, P3, P5, M = 2654435761, 2246822519, 3266489917, 374761393, 0xFFFFFFFF
        rot = lambda x, r: ((x << r) | (x >> (32 - r))) & M
        h = (seed + P5 + len(data)) & M
        for b in data:
            h = (rot((h + b * P5) & M, 11) * P1) & M
        h ^= h >> 15; h = (h * P2) & M; h ^= h >> 13; h = (h * P3) & M; h ^= h >> 16
        return h
    assert raw[:4] == b"\x04\x22\x4d\x18" and raw[4] == 0x60 and raw[5] == 0x70
    assert raw[6] == (xxh32(raw[4:6]) >> 8) & 0xFF
    out, p = bytearray(), 7
    while True:
        (sz,) = struct.unpack_from("<I", raw, p); p += 4
        if sz == 0:
            break
        assert sz & 0x80000000
        out += raw[p:p + (sz & 0x7FFFFFFF)]; p += sz & 0x7FFFFFFF
    assert p == len(raw)
    return bytes(out)


def _liblz4_frame_decode(raw):
    """LZ4 frame -> bytes with the SYSTEM liblz4 (LZ4F_decompress), not this repository's reader."""
    lz4 = ctypes.CDLL("liblz4.so.1")
    lz4.LZ4F_createDecompressionContext.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint]
    lz4.LZ4F_createDecompressionContext.restype = ctypes.c_size_t
    lz4.LZ4F_decompress.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
    lz4.LZ4F_decompress.restype = ctypes.c_size_t
    lz4.LZ4F_isError.argtypes = [ctypes.c_size_t]
    lz4.LZ4F_freeDecompressionContext.argtypes = [ctypes.c_void_p]
    ctx = ctypes.c_void_p()
    assert not lz4.LZ4F_isError(lz4.LZ4F_createDecompressionContext(ctypes.byref(ctx), 100))
    src = ctypes.create_string_buffer(raw, len(raw))
    dst = ctypes.create_string_buffer(1 << 20)
    out, pos = bytearray(), 0
    while pos < len(raw):
        dn, sn = ctypes.c_size_t(len(dst)), ctypes.c_size_t(len(raw) - pos)
        rc = lz4.LZ4F_decompress(ctx, dst, ctypes.byref(dn), ctypes.byref(src, pos), ctypes.byref(sn), None)
        assert not lz4.LZ4F_isError(rc), "liblz4 rejects the frame"
        out += dst.raw[:dn.value]
        pos += sn.value
        if rc == 0 and sn.value == 0 and dn.value == 0:
            break
    assert rc == 0, "frame not complete"
    lz4.LZ4F_freeDecompressionContext(ctx)
    return bytes(out)


def _parse_sequences_like_reference(data):
    """Line grammar as utils/parse_sequences_file.py:21-34 of the reference reads it."""
    k = l = 0
    node_minims, kmer_to_seq, minim_shift = {}, {}, {}
    for line in data.decode().splitlines():
        if line.startswith("#"):
            if line.startswith("# k = "):
                k = int(line.split()[-1])
            if line.startswith("# l = "):
                l = int(line.split()[-1])
            continue
        spl = line.split()
        seq_id = spl[0]
        minims = tuple(map(lambda x: int(x.strip("[").strip("]").replace(",", "")), spl[1:-5]))
        assert spl[-4] == "*" and spl[-3] == "*"
        seq = spl[-5]
        shifts = tuple(map(lambda x: int(x.strip("(").strip(")").replace(",", "")), spl[-2:]))
        node_minims[seq_id] = minims
        kmer_to_seq[minims] = seq
        minim_shift[seq_id] = shifts
    return k, l, node_minims, kmer_to_seq, minim_shift


def test_file_writers_on_host(mdbg, oracle, example_reads, tmp_path):
    """mdbg_write_gfa / mdbg_write_sequences are host code: feed them a graph (the oracle's, through
    the C struct) and compare with the oracle's own canonical text -- .gfa S/L grammar
    (main.rs:1021,1095), .sequences header + lines (main.rs:625-628,702), LZ4 frame validity."""
    bases, off, _ = example_reads
    o = oracle.build_graph(bases, off, 7, 10, 0.0008, 2, 0.01)
    F = mdbg.ffi
    cg = F.CGraph()
    keep = []

    def put(name, arr):
        a = np.ascontiguousarray(arr)
        keep.append(a)
        setattr(cg, name, a.ctypes.data)
    cg.n_nodes, cg.n_edges, cg.n_seqlines = len(o.index), len(o.e_n1), len(o.q_index)
    cg.k, cg.l = 7, 10
    for name, arr in (("node_index", o.index), ("abundance", o.abundance), ("seqlen", o.seqlen), ("shift", o.shift),
                      ("tuple", o.tuple), ("e_n1", o.e_n1), ("e_o1", o.e_o1), ("e_n2", o.e_n2), ("e_o2", o.e_o2),
                      ("e_overlap", o.e_ov), ("q_index", o.q_index), ("q_read", o.q_read), ("q_start", o.q_start),
                      ("q_end", o.q_end), ("q_reversed", o.q_rev), ("q_shift", o.q_shift)):
        put(name, arr)
    L = F.lib()
    gfa, seq, seqz = str(tmp_path / "p.gfa"), str(tmp_path / "p.sequences"), str(tmp_path / "p.lz4.sequences")
    assert L.mdbg_write_gfa(ctypes.byref(cg), gfa.encode()) == 0
    assert L.mdbg_write_sequences(ctypes.byref(cg), bases.ctypes.data, off.ctypes.data, seq.encode(), 0) == 0
    assert L.mdbg_write_sequences(ctypes.byref(cg), bases.ctypes.data, off.ctypes.data, seqz.encode(), 1) == 0
    ogfa, oseq = str(tmp_path / "o.gfa"), str(tmp_path / "o.sequences")
    o.write_gfa(ogfa); o.write_sequences(oseq)
    text = open(gfa).read().splitlines(True)
    assert text[0] == "H\tVN:Z:1.0\n"
    n_s = sum(1 for x in text if x.startswith("S"))
    assert all(x.startswith("S") for x in text[1:1 + n_s]) and all(x.startswith("L") for x in text[1 + n_s:])
    assert sorted(text[1:]) == sorted(open(ogfa).read().splitlines(True)[1:])
    plain = open(seq).read()
    assert plain.startswith("# k = 7\n# l = 10\n# Structure of remaining of the file:\n"
                            "# [node name]\t[list of minimizers]\t[sequence of node]\t[abundance]\t[origin]\t[shift]\n")
    body = sorted(x for x in plain.splitlines(True) if not x.startswith("#"))
    assert body == sorted(open(oseq).read().splitlines(True))
    assert _lz4_stored_frame_decode(open(seqz, "rb").read()).decode() == plain
    # an independent consumer: the system's liblz4 (what lzzzz / python-lz4 wrap: to_basespace.rs:233,
    # utils/parse_sequences_file.py:12) decodes the frame, then the reference's parser logic reads the lines
    assert _liblz4_frame_decode(open(seqz, "rb").read()).decode() == plain
    k, l, node_minims, kmer_to_seq, minim_shift = _parse_sequences_like_reference(_liblz4_frame_decode(open(seqz, "rb").read()))
    assert (k, l) == (7, 10) and len(node_minims) == len(o.q_index)
    for i in range(len(o.index)):
        name = str(int(o.index[i]))
        assert node_minims[name] == tuple(int(x) for x in o.tuple[i])
    for q in range(len(o.q_index)):
        name = str(int(o.q_index[q]))
        raw = bytes(bases[int(off[int(o.q_read[q])]) + int(o.q_start[q]):int(off[int(o.q_read[q])]) + int(o.q_end[q])])
        if o.q_rev[q]:
            raw = raw[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))
        assert kmer_to_seq[node_minims[name]] == raw.decode()
        assert minim_shift[name] == (int(o.q_shift[q][0]), int(o.q_shift[q][1]))
    # the streaming writer (what the front end uses in its second pass over the input) gives the same file
    W = F.vp()
    seqs = str(tmp_path / "p.stream.sequences")
    assert L.mdbg_seq_writer_open(ctypes.byref(cg), seqs.encode(), 1, ctypes.byref(W)) == 0
    n_calls = 0
    while True:
        r = L.mdbg_seq_writer_next_read(W)
        if r == 0xFFFFFFFFFFFFFFFF:
            break
        assert L.mdbg_seq_writer_read(W, r, bases.ctypes.data + int(off[r]), int(off[r + 1] - off[r])) == 0
        n_calls += 1
    assert L.mdbg_seq_writer_close(W) == 0 and n_calls > 0
    assert open(seqs, "rb").read() == open(seqz, "rb").read()
    # several writers sharing the lines (one file per part, like the reference's one file per worker thread): every
    # file carries the header, together they hold exactly the lines of the single file
    all_lines = sorted(x for x in plain.splitlines(True) if not x.startswith("#"))
    for n_parts in (1, 2, 3, 7):
        got = []
        for part in range(n_parts):
            W = F.vp()
            pth = str(tmp_path / ("p.%d_of_%d.sequences" % (part, n_parts)))
            assert L.mdbg_seq_writer_open_part(ctypes.byref(cg), pth.encode(), 0, part, n_parts, ctypes.byref(W)) == 0
            last = -1
            while True:
                r = L.mdbg_seq_writer_next_read(W)
                if r == 0xFFFFFFFFFFFFFFFF:
                    break
                assert r > last or last == -1
                last = r
                assert L.mdbg_seq_writer_read(W, r, bases.ctypes.data + int(off[r]), int(off[r + 1] - off[r])) == 0
            assert L.mdbg_seq_writer_close(W) == 0
            txt = open(pth).read()
            assert txt.startswith("# k = 7\n# l = 10\n")
            got += [x for x in txt.splitlines(True) if not x.startswith("#")]
        assert sorted(got) == all_lines and len(got) == len(all_lines)
    W = F.vp()
    assert L.mdbg_seq_writer_open_part(ctypes.byref(cg), str(tmp_path / "bad").encode(), 0, 3, 3, ctypes.byref(W)) != 0


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints one JSON line
    with the contract's keys; it needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["impl"] == "reference" and j["unit"] == "Gbases/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
