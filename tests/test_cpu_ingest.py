"""CPU tests of the front end's parallel FASTA/FASTQ reader (rust-mdbg_b200/cli/ingest.hpp, driven by
tests/model/ingest_check.cpp): records, ids and sequences must equal a plain Python parse for plain and gzip
files, multi-line FASTA, CRLF, FASTQ whose quality lines start with '@', tiny batches (many windows cut at
record starts) and many parser threads."""
import gzip
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "model", "ingest_check")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", EXE, os.path.join(HERE, "model", "ingest_check.cpp"), "-lz"])
    return EXE


def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def run(exe, path, fmt, threads, target):
    out = subprocess.run([exe, path, fmt, str(threads), str(target)], capture_output=True, text=True, check=True).stdout
    return [tuple(x.split("\t")) for x in out.splitlines()]


def expected(records):
    return [(i, str(len(s)), "%016x" % fnv(s)) for i, s in records]


def make_records(rng, n, mean=3000):
    recs = []
    for i in range(n):
        ln = int(max(0, rng.normal(mean, mean / 2))) if i % 17 else 0
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), ln, p=[.24, .24, .24, .24, .04]))
        recs.append(("read%d/%d" % (i, ln), s))
    return recs


@pytest.mark.parametrize("gz", [False, True])
@pytest.mark.parametrize("width,crlf", [(0, False), (60, False), (70, True)])
def test_fasta(exe, tmp_path, gz, width, crlf):
    rng = np.random.default_rng(5 + width)
    recs = make_records(rng, 700)
    nl = b"\r\n" if crlf else b"\n"
    text = bytearray()
    for i, s in recs:
        text += b">" + i.encode() + b" some description" + nl
        if width:
            for j in range(0, len(s), width):
                text += s[j:j + width] + nl
        else:
            text += s + nl
    p = str(tmp_path / ("x.fa.gz" if gz else "x.fa"))
    with (gzip.open(p, "wb", compresslevel=1) if gz else open(p, "wb")) as f:
        f.write(bytes(text))
    exp = expected(recs)
    for threads, target in ((1, 1 << 30), (7, 1 << 30), (4, 50000), (16, 3 << 20)):
        assert run(exe, p, "fasta", threads, target) == exp


@pytest.mark.parametrize("gz", [False, True])
def test_fastq_with_at_quality_lines(exe, tmp_path, gz):
    rng = np.random.default_rng(9)
    recs = make_records(rng, 600, mean=2500)
    text = bytearray()
    for n, (i, s) in enumerate(recs):
        q = bytes(rng.integers(33, 74, len(s), dtype=np.uint8))
        if n % 3 == 0 and len(q):
            q = b"@" + q[1:]                       # a quality line that looks like a header
        if n % 5 == 0 and len(q) > 1:
            q = q[:1] + b"+" + q[2:]
        text += b"@" + i.encode() + b" d\n" + s + b"\n+\n" + q + b"\n"
    p = str(tmp_path / ("x.fq.gz" if gz else "x.fq"))
    with (gzip.open(p, "wb", compresslevel=1) if gz else open(p, "wb")) as f:
        f.write(bytes(text))
    exp = expected(recs)
    for threads, target in ((1, 1 << 30), (5, 1 << 30), (8, 40000), (16, 2 << 20)):
        assert run(exe, p, "fastq", threads, target) == exp


def test_degenerate_files(exe, tmp_path):
    p = str(tmp_path / "empty.fa")
    open(p, "wb").close()
    assert run(exe, p, "fasta", 4, 1 << 20) == []
    p = str(tmp_path / "one.fa")
    with open(p, "wb") as f:
        f.write(b">r1\nACGT")                          # no trailing newline
    assert run(exe, p, "fasta", 4, 1 << 20) == expected([("r1", b"ACGT")])
    p = str(tmp_path / "example.fa.gz")               # the reference's shipped example (multi-member-free gzip)
    src = os.path.join(HERE, "golden", "config1_reads.fa.gz")
    from helpers import load_fasta
    bases, off, names = load_fasta(src)
    got = run(exe, src, "fasta", 8, 1 << 22)
    assert len(got) == 657
    assert [int(x[1]) for x in got] == [int(off[i + 1] - off[i]) for i in range(657)]
    assert got[0][2] == "%016x" % fnv(bytes(bases[int(off[0]):int(off[1])]))


def test_gzip_reader_paths_agree_and_damage_is_an_error(exe, tmp_path):
    """The front end's own gzip decoder (cli/gz_inflate.hpp) and zlib's gzread (MDBG_GZ_ZLIB=1) deliver the same
    records for single- and multi-member files; a damaged or truncated .gz is an error, not a shorter read set."""
    rng = np.random.default_rng(21)
    recs = make_records(rng, 900, mean=4000)
    text = b"".join(b">" + i.encode() + b"\n" + s + b"\n" for i, s in recs)
    exp = expected(recs)
    one = str(tmp_path / "one.fa.gz")
    open(one, "wb").write(gzip.compress(text, 6))
    cut = [0]
    while cut[-1] < len(text):                                   # members cut at arbitrary bytes (bgzip does that too)
        cut.append(min(len(text), cut[-1] + int(rng.integers(1, 90000))))
    many = str(tmp_path / "many.fa.gz")
    open(many, "wb").write(b"".join(gzip.compress(text[a:b], int(rng.integers(1, 10))) for a, b in zip(cut, cut[1:])))
    for path in (one, many):
        for env in ({}, {"MDBG_GZ_ZLIB": "1"}):
            out = subprocess.run([exe, path, "fasta", "5", str(1 << 20)], capture_output=True, text=True, check=True,
                                 env={**os.environ, **env}).stdout
            assert [tuple(x.split("\t")) for x in out.splitlines()] == exp
    good = open(one, "rb").read()
    flipped = bytearray(good); flipped[len(good) // 2] ^= 0x10
    for name, blob in (("truncated", good[:len(good) * 2 // 3]), ("flipped", bytes(flipped)), ("no_trailer", good[:-8])):
        bad = str(tmp_path / (name + ".fa.gz"))
        open(bad, "wb").write(blob)
        r = subprocess.run([exe, bad, "fasta", "3", str(1 << 20)], capture_output=True, text=True)
        assert r.returncode == 1 and "gzip" in r.stderr, (name, r.returncode, r.stderr[-200:])


def bgzf_member(chunk, level=6):
    import struct
    import zlib
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    d = c.compress(chunk) + c.flush()
    bsize = 12 + 6 + len(d) + 8
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize - 1) + d +
            struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))


def test_block_gzip_is_inflated_in_parallel(exe, tmp_path):
    """bgzip / BGZF files (independent members that carry their compressed size): a block's worth of members is
    inflated by all threads at once; same records as the serial decoder (MDBG_GZ_SERIAL=1) and as zlib; a file that
    stops being BGZF half way (plain gzip members appended) goes on through the serial decoder; a damaged member is
    an error."""
    rng = np.random.default_rng(22)
    recs = make_records(rng, 1500, mean=5000)
    text = b"".join(b">" + i.encode() + b"\n" + s + b"\n" for i, s in recs)
    exp = expected(recs)
    members = [bgzf_member(text[a:a + 65280], int(rng.integers(1, 10))) for a in range(0, len(text), 65280)]
    pure = str(tmp_path / "pure.fa.gz")
    open(pure, "wb").write(b"".join(members) + bgzf_member(b""))            # with the BGZF end-of-file marker
    half = len(members) // 2
    cut = sum(len(m) for m in members[:half])
    tail_text = text[half * 65280:]
    mixed = str(tmp_path / "mixed.fa.gz")
    open(mixed, "wb").write(b"".join(members[:half]) + gzip.compress(tail_text[:100000], 6) + gzip.compress(tail_text[100000:], 1))
    for path in (pure, mixed):
        for env in ({}, {"MDBG_GZ_SERIAL": "1"}, {"MDBG_GZ_ZLIB": "1"}):
            for threads, target in ((1, 1 << 30), (6, 1 << 20), (16, 200000)):
                out = subprocess.run([exe, path, "fasta", str(threads), str(target)], capture_output=True, text=True, check=True,
                                     env={**os.environ, **env}).stdout
                assert [tuple(x.split("\t")) for x in out.splitlines()] == exp, (path, env, threads)
    blob = bytearray(open(pure, "rb").read())
    blob[cut + 40] ^= 0x04                                                 # inside the deflate data of a member in the middle
    bad = str(tmp_path / "bad.fa.gz")
    open(bad, "wb").write(bytes(blob))
    r = subprocess.run([exe, bad, "fasta", "4", str(1 << 20)], capture_output=True, text=True)
    assert r.returncode == 1 and "gzip" in r.stderr, (r.returncode, r.stderr[-200:])


def test_plain_gzip_is_inflated_in_parallel(exe, tmp_path):
    """One gzip stream above the size threshold (lowered here) goes through GzParallel: same records as the serial
    decoder and as zlib for FASTA and FASTQ, small spans, many threads; damage is an error."""
    rng = np.random.default_rng(23)
    recs = make_records(rng, 1200, mean=6000)
    fa = b"".join(b">" + i.encode() + b"\n" + s + b"\n" for i, s in recs)
    fq = b"".join(b"@" + i.encode() + b"\n" + s + b"\n+\n" + bytes(rng.integers(33, 74, len(s), dtype=np.uint8)) + b"\n" for i, s in recs)
    exp = expected(recs)
    par = {"MDBG_GZ_PAR_MIN": "0", "MDBG_GZ_SPAN": "65536"}
    for name, text, fmt in (("x.fa.gz", fa, "fasta"), ("x.fq.gz", fq, "fastq")):
        p = str(tmp_path / name)
        open(p, "wb").write(gzip.compress(text, 6))
        for env in (par, {**par, "MDBG_GZ_SPAN": "1000000"}, {"MDBG_GZ_SERIAL": "1"}, {"MDBG_GZ_ZLIB": "1"}):
            for threads, target in ((3, 1 << 30), (8, 1 << 20), (16, 300000)):
                out = subprocess.run([exe, p, fmt, str(threads), str(target)], capture_output=True, text=True, check=True,
                                     env={**os.environ, **env}).stdout
                assert [tuple(x.split("\t")) for x in out.splitlines()] == exp, (name, env, threads)
    good = open(str(tmp_path / "x.fa.gz"), "rb").read()
    flipped = bytearray(good); flipped[len(good) // 3] ^= 0x20
    for name, blob in (("truncated", good[:len(good) // 2]), ("flipped", bytes(flipped))):
        bad = str(tmp_path / (name + ".fa.gz"))
        open(bad, "wb").write(blob)
        r = subprocess.run([exe, bad, "fasta", "6", str(1 << 20)], capture_output=True, text=True, env={**os.environ, **par})
        assert r.returncode == 1 and "gzip" in r.stderr, (name, r.returncode, r.stderr[-200:])

