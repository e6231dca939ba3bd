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
