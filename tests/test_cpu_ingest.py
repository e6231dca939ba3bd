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

