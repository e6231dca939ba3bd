"""CPU tests of the front end's gzip reader (rust-mdbg_b200/cli/gz_inflate.hpp, driven by tests/model/gunzip_check.cpp):
its output must equal zlib's (Python's gzip / zlib modules) byte for byte on stored, fixed-Huffman and dynamic
blocks, every compression level, multi-member files (bgzip-like), header options, long matches at distance 1 and
32768, inputs around every internal buffer size -- and every damaged stream must be refused with an error, never
decoded to something else."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "model", "gunzip_check")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", EXE, os.path.join(HERE, "model", "gunzip_check.cpp"), "-lz"])
    return EXE


def gunzip(exe, path, chunk=None, par=None):
    args = [exe, path]
    if chunk or par:
        args += ["--raw", str(chunk or (1 << 20))]
    if par:
        args += [str(par[0]), str(par[1])]          # the parallel single-stream decoder: threads, span bytes
    r = subprocess.run(args, capture_output=True, timeout=300)
    return r.returncode, r.stdout, r.stderr.decode()


def dna(rng, n, repeat=0.0):
    b = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    if repeat and n > 2000:                       # copies of earlier stretches: long matches
        for _ in range(int(repeat * n / 500)):
            L = int(rng.integers(20, 600)); s = int(rng.integers(0, n - L)); d = int(rng.integers(0, n - L))
            b[d:d + L] = b[s:s + L]
    return b.tobytes()


def fasta(rng, n_reads, mean):
    out = []
    for i in range(n_reads):
        L = max(1, int(rng.normal(mean, mean / 4)))
        s = dna(rng, L, 0.2)
        out.append(b">read_%d some description\n" % i)
        out.extend(s[j:j + 80] + b"\n" for j in range(0, L, 80))
    return b"".join(out)


def raw_deflate(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


def member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flg=0, extra=b"", name=b"", comment=b"", hcrc=False):
    hdr = bytearray(b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\x03")
    if flg & 4:
        hdr += struct.pack("<H", len(extra)) + extra
    if flg & 8:
        hdr += name + b"\0"
    if flg & 16:
        hdr += comment + b"\0"
    if flg & 2:
        hdr += struct.pack("<H", zlib.crc32(bytes(hdr)) & 0xFFFF)
    return bytes(hdr) + raw_deflate(data, level, strategy) + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


def check(exe, tmp_path, blob, expect, name="x.gz", chunk=None):
    p = str(tmp_path / name)
    open(p, "wb").write(blob)
    rc, out, err = gunzip(exe, p, chunk)
    assert rc == 0, err
    assert out == expect, (len(out), len(expect))


@pytest.mark.parametrize("level", [0, 1, 2, 4, 6, 9])
def test_levels_fasta(exe, tmp_path, level):
    rng = np.random.default_rng(level)
    text = fasta(rng, 120, 15000)
    check(exe, tmp_path, gzip.compress(text, compresslevel=level), text)


@pytest.mark.parametrize("strategy", [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
def test_strategies(exe, tmp_path, strategy):
    rng = np.random.default_rng(11)
    text = fasta(rng, 40, 9000) + bytes(rng.integers(0, 256, 70000, dtype=np.uint8)) + b"A" * 100000 + fasta(rng, 5, 3000)
    check(exe, tmp_path, member(text, 6, strategy), text)


@pytest.mark.parametrize("n", [0, 1, 2, 5, 257, 258, 259, 32767, 32768, 32769, (1 << 20) - 1, 1 << 20, (1 << 20) + 1,
                               (1 << 20) + 258, 3 * (1 << 20) + 12345])
def test_sizes_around_the_buffers(exe, tmp_path, n):
    rng = np.random.default_rng(n % 1000)
    text = dna(rng, n, 0.3)
    check(exe, tmp_path, gzip.compress(text, 6), text)
    check(exe, tmp_path, gzip.compress(text, 0), text, "stored.gz")


def test_long_matches_and_far_distances(exe, tmp_path):
    rng = np.random.default_rng(5)
    unit = dna(rng, 32768)
    text = b"A" * 300000 + b"AC" * 100000 + b"ACG" * 70000 + unit + unit + unit[:100] + dna(rng, 5000) + unit[-300:] * 50
    for level in (1, 6, 9):
        check(exe, tmp_path, gzip.compress(text, level), text)


def test_binary_data_uses_long_codes(exe, tmp_path):
    rng = np.random.default_rng(6)
    # a skewed byte distribution gives code lengths up to 15 (sub-tables)
    p = np.array([2.0 ** -(i // 8) for i in range(256)]); p /= p.sum()
    text = bytes(rng.choice(256, 600000, p=p).astype(np.uint8))
    for level in (1, 6, 9):
        check(exe, tmp_path, gzip.compress(text, level), text)


def test_multi_member_and_header_fields(exe, tmp_path):
    rng = np.random.default_rng(7)
    parts = [fasta(rng, 3, 2000), b"", dna(rng, 70000, 0.5), b"x", fasta(rng, 20, 30000)]
    blob = (member(parts[0], 6, flg=8, name=b"reads.fa") + member(parts[1], 6) + member(parts[2], 9, flg=4 | 16, extra=b"BC\x02\x00\x10\x00", comment=b"hello") +
            member(parts[3], 1, flg=2) + member(parts[4], 6, flg=2 | 4 | 8 | 16, extra=b"", name=b"n", comment=b"c"))
    text = b"".join(parts)
    check(exe, tmp_path, blob, text)
    check(exe, tmp_path, blob + b"\0" * 37, text, "padded.gz")          # zero padding after the last member
    # bgzip-like: many small members
    blob = b"".join(member(text[i:i + 4000], 6) for i in range(0, len(text), 4000))
    check(exe, tmp_path, blob, text, "bg.gz")


def test_read_chunk_sizes(exe, tmp_path):
    rng = np.random.default_rng(8)
    text = fasta(rng, 60, 20000)
    blob = gzip.compress(text, 6)
    for chunk in (1, 7, 4096, 65536, 1 << 20, (1 << 20) + 1, 5 << 20):
        if chunk == 1 and len(text) > 300000:
            check(exe, tmp_path, gzip.compress(text[:200000], 6), text[:200000], chunk=chunk)
        else:
            check(exe, tmp_path, blob, text, chunk=chunk)


def test_damaged_streams_are_refused(exe, tmp_path):
    rng = np.random.default_rng(9)
    text = fasta(rng, 30, 12000)
    good = gzip.compress(text, 6)
    bad = {}
    bad["truncated"] = good[:len(good) // 2]
    bad["no_trailer"] = good[:-8]
    bad["short_trailer"] = good[:-3]
    bad["crc"] = good[:-8] + bytes([good[-8] ^ 1]) + good[-7:]
    bad["isize"] = good[:-1] + bytes([good[-1] ^ 0x40])
    bad["magic"] = b"\x1f\x8c" + good[2:]
    bad["method"] = good[:2] + b"\x07" + good[3:]
    bad["reserved_flag"] = good[:3] + b"\x20" + good[4:]
    bad["garbage_after"] = good + b"garbage"
    bad["block_type_3"] = good[:10] + bytes([good[10] | 0x06]) + good[11:]
    for k, pos in enumerate(rng.integers(12, len(good) - 10, 40)):     # bit flips inside the deflate data
        b = bytearray(good); b[int(pos)] ^= 1 << int(rng.integers(0, 8)); bad["flip%d" % k] = bytes(b)
    n_wrong = 0
    for name, blob in bad.items():
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(blob)
        rc, out, err = gunzip(exe, p)
        try:
            ref = gzip.decompress(blob)
            if name == "reserved_flag":            # Python's header parser ignores them; zlib (and this reader) refuse
                ref = None
        except Exception:
            ref = None
        if ref is None:
            assert rc == 1 and err.startswith("gzip:") or "stopped" in err, (name, rc, err)
        else:                                      # a flip zlib itself does not notice cannot be noticed here either
            assert rc == 0 and out == ref, name
            n_wrong += 1
    assert n_wrong <= 2
    # a distance that reaches before the start of the member's data: hand-made fixed-Huffman block
    # BFINAL=1 BTYPE=01, literal 'A' (0x41 -> code 0x71, 8 bits), length 3 (code 257: 0000001), distance 2 (code 1: 00001), EOB
    bits = "1" + "10"                             # final, fixed (BTYPE bits LSB first: 01 -> "10")
    bits += format(0x30 + 0x41, "08b")            # literal A, MSB first
    bits += "0000001" + "00001" + "0000000"       # length 3, distance code 1 (= 2), end of block
    by = bytearray()
    for i in range(0, len(bits), 8):
        chunk = bits[i:i + 8].ljust(8, "0")
        by.append(int(chunk[::-1], 2))
    blob = b"\x1f\x8b\x08\0\0\0\0\0\0\x03" + bytes(by) + struct.pack("<II", 0, 4)
    p = str(tmp_path / "far.gz")
    open(p, "wb").write(blob)
    rc, out, err = gunzip(exe, p)
    assert rc == 1 and "distance" in err, (rc, err)


def test_damaged_streams_under_sanitizers(tmp_path):
    """The decoder built with AddressSanitizer + UBSan on truncated, bit-flipped and random streams: every one is either
    decoded or refused with a message -- no out-of-bounds access, no undefined shift, no crash."""
    san = str(tmp_path / "gunzip_san")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        "-o", san, os.path.join(HERE, "model", "gunzip_check.cpp"), "-lz"], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no sanitizer runtime in this image: " + r.stderr[-200:])
    rng = np.random.default_rng(123)
    good = gzip.compress(fasta(rng, 8, 6000), 6)
    small = gzip.compress(b"ACGT" * 10, 9)
    cases = [good[:int(c)] for c in rng.integers(0, len(good), 60)]
    for pos in rng.integers(0, len(good), 120):
        b = bytearray(good); b[int(pos)] ^= 1 << int(rng.integers(0, 8)); cases.append(bytes(b))
    cases += [b"\x1f\x8b\x08\x00\0\0\0\0\0\x03" + bytes(rng.integers(0, 256, int(rng.integers(1, 3000)), dtype=np.uint8)) for _ in range(40)]
    for pos in range(10, len(small)):
        b = bytearray(small); b[pos] ^= 1 << (pos % 8); cases.append(bytes(b))
    p = str(tmp_path / "f.gz")
    for n, blob in enumerate(cases):
        open(p, "wb").write(blob)
        for extra in ([], ["1048576", "4", "65536"]) if n % 3 == 0 else ([],):      # every third one through GzParallel too
            r = subprocess.run([san, p, "--hash"] + extra, capture_output=True, text=True, timeout=60)
            assert r.returncode in (0, 1), (r.returncode, r.stderr[-500:])
            assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-800:]
            if r.returncode == 1:
                assert r.stderr.startswith("gzip:") or "stopped" in r.stderr, r.stderr[-300:]


PAR = [(1, 1 << 16), (3, 1 << 16), (8, 70000), (5, 1 << 20), (16, 200000)]


@pytest.mark.parametrize("level", [1, 6, 9])
def test_parallel_single_stream(exe, tmp_path, level):
    """GzParallel: spans of ONE gzip stream decoded by several threads from block starts they find themselves, unknown
    windows as 16-bit markers resolved afterwards, spans joined only where they meet exactly: same bytes as zlib for every
    thread count and span size (tiny spans: a join every few blocks)."""
    rng = np.random.default_rng(30 + level)
    text = fasta(rng, 400, 12000)
    p = str(tmp_path / "x.gz")
    open(p, "wb").write(gzip.compress(text, level))
    for par in PAR:
        rc, out, err = gunzip(exe, p, par=par)
        assert rc == 0, (par, err)
        assert out == text, (par, len(out), len(text))


def test_parallel_odd_streams(exe, tmp_path):
    """What the block finder does not look for or the marker scheme is not made for must still come out right: stored and
    fixed-Huffman blocks, incompressible and extremely compressible stretches (the serial decoder takes over mid-member),
    several members, an empty member first, tiny files, binary data with long codes."""
    rng = np.random.default_rng(40)
    dna_text = fasta(rng, 150, 15000)
    noise = bytes(rng.integers(0, 256, 600000, dtype=np.uint8))
    p8 = np.array([2.0 ** -(i // 8) for i in range(256)]); p8 /= p8.sum()
    skew = bytes(rng.choice(256, 500000, p=p8).astype(np.uint8))
    cases = {
        "stored": gzip.compress(dna_text, 0),
        "fixed": member(dna_text, 6, zlib.Z_FIXED),
        "huffman_only": member(dna_text, 6, zlib.Z_HUFFMAN_ONLY),
        "rle": member(dna_text + b"A" * 3000000 + dna_text, 6, zlib.Z_RLE),
        "mixed": gzip.compress(dna_text + noise + dna_text[:200000] + b"ACGT" * 2000000 + dna_text[-300000:] + skew, 6),
        "very_compressible": gzip.compress(b"A" * 50000000 + dna_text, 6),
        "members": member(dna_text[:700000], 6) + member(b"", 6) + member(dna_text[700000:], 1) + member(noise, 6),
        "empty_first": member(b"", 6) + gzip.compress(dna_text, 6),
        "tiny": gzip.compress(b"ACGT\n", 6),
        "empty": gzip.compress(b"", 6),
        "padded": gzip.compress(dna_text, 6) + b"\0" * 100,
    }
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(blob)
        ref = dna_text if name == "padded" else gzip.decompress(blob)
        for par in ((4, 1 << 16), (7, 300000), (2, 4 << 20)):
            rc, out, err = gunzip(exe, p, par=par)
            assert rc == 0, (name, par, err)
            assert out == ref, (name, par, len(out), len(ref))


def test_parallel_damaged_streams_are_refused(exe, tmp_path):
    rng = np.random.default_rng(41)
    text = fasta(rng, 200, 12000)
    good = gzip.compress(text, 6)
    bad = {"truncated": good[:len(good) // 2], "no_trailer": good[:-8], "crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],
           "isize": good[:-1] + bytes([good[-1] ^ 0x40]), "garbage_after": good + b"garbage"}
    for k, pos in enumerate(rng.integers(12, len(good) - 10, 60)):
        b = bytearray(good); b[int(pos)] ^= 1 << int(rng.integers(0, 8)); bad["flip%d" % k] = bytes(b)
    for name, blob in bad.items():
        p = str(tmp_path / (name + ".gz"))
        open(p, "wb").write(blob)
        try:
            ref = gzip.decompress(blob)
        except Exception:
            ref = None
        for par in ((4, 1 << 16), (3, 500000)):
            rc, out, err = gunzip(exe, p, par=par)
            if ref is None:
                assert rc == 1 and "gzip" in err, (name, par, rc, err)
            else:
                assert rc == 0 and out == ref, (name, par)

