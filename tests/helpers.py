"""Shared helpers for the test-suite (plain Python/numpy; no product code)."""
import gzip

import numpy as np

# --- an independent, deliberately naive Python restatement used as a second opinion on small
# --- inputs (crate nthash direct formulas; read.rs:157-211; kmer_vec.rs:34-39) ---------------
H = {ord('A'): 0x3c8bfbb395c60474, ord('C'): 0x3193c18562a02b4c, ord('G'): 0x20323ed082572324,
     ord('T'): 0x295549f54be24456, ord('N'): 0}
COMP = {ord('A'): ord('T'), ord('C'): ord('G'), ord('G'): ord('C'), ord('T'): ord('A'), ord('N'): ord('N')}
M64 = (1 << 64) - 1


def rol(x, r):
    r &= 63
    return ((x << r) | (x >> (64 - r))) & M64 if r else x


def py_ntc64(s, i, k):
    f = r = 0
    for j in range(k):
        f ^= rol(H[s[i + j]], k - 1 - j)
        r ^= rol(H[COMP[s[i + j]]], j)
    return min(f, r)


def py_hpc(s):
    out, pos = [], []
    for i, c in enumerate(s):
        if i == 0 or c != s[i - 1]:
            out.append(c)
            pos.append(i)
    return bytes(out), pos


def py_extract(s, l, density, hpc=True):
    """-> (hashes, raw positions) by the direct (non-rolling) formulas."""
    s = bytes(s)
    bound = int(float(density) * 18446744073709551616.0)
    if hpc:
        h, pos = py_hpc(s)
    else:
        h, pos = s, list(range(len(s)))
    hs, ps = [], []
    for i in range(len(h) - l + 1):
        v = py_ntc64(h, i, l)
        if v <= bound:
            hs.append(v)
            ps.append(pos[i])
    return hs, ps


def py_kminmers(hs, ps, k, l):
    """main.rs:756-781 -> list of (canonical tuple, reversed, shift pair, (off0, off1, off2))."""
    out = []
    if not len(hs) > k:
        return out
    for i in range(len(hs) - k + 1):
        t = tuple(hs[i:i + k])
        rv = t[::-1]
        reversed_ = not (t < rv)
        node = rv if reversed_ else t
        a = ps[i + 1] - ps[i]
        b = ps[i + k - 1] - ps[i + k - 2]
        shift = (b, a) if reversed_ else (a, b)
        out.append((node, reversed_, shift, (ps[i], ps[i + k - 1] + l, ps[i + k - 1] - ps[i] + 2)))
    return out


def load_fasta(path):
    op = gzip.open if path.endswith(".gz") else open
    names, seqs, cur = [], [], []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if names:
                    seqs.append(b"".join(cur))
                names.append(line[1:].split()[0].decode())
                cur = []
            else:
                cur.append(line)
    if names:
        seqs.append(b"".join(cur))
    return pack_reads(seqs) + (names,)


def pack_reads(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
    return bases, off


def random_reads(rng, n, mean=3000, sd=1000, lo=0, hi=20000, alphabet=b"ACGT", hp=0.0):
    """Random reads; hp = probability of repeating the previous base (homopolymers)."""
    seqs = []
    al = np.frombuffer(alphabet, dtype=np.uint8)
    for _ in range(n):
        ln = int(min(hi, max(lo, rng.normal(mean, sd))))
        s = al[rng.integers(0, len(al), ln)]
        if hp > 0 and ln > 1:
            rep = rng.random(ln) < hp
            rep[0] = False
            idx = np.where(~rep, np.arange(ln), 0)
            np.maximum.accumulate(idx, out=idx)
            s = s[idx]
        seqs.append(s.tobytes())
    return seqs


def genome_reads(rng, glen, n, mean=3000, sd=800, err=0.0, lo=200):
    """Reads sampled from a random genome (both strands) so that k-min-mers repeat."""
    al = np.frombuffer(b"ACGT", dtype=np.uint8)
    g = rng.integers(0, 4, glen).astype(np.uint8)
    seqs = []
    for _ in range(n):
        ln = int(min(glen, max(lo, rng.normal(mean, sd))))
        st = int(rng.integers(0, glen - ln + 1))
        s = g[st:st + ln].copy()
        if err > 0:
            e = rng.random(ln) < err
            s[e] = (s[e] + rng.integers(1, 4, int(e.sum()))) & 3
        if rng.integers(0, 2):
            s = (3 - s)[::-1]
        seqs.append(al[s].tobytes())
    return seqs
