"""CPU tests of the 4:1 upload encoding: host packer (rust-mdbg_b200/csrc/pack_host.cc, through the
C ABI entry mdbg_pack_bases_host) against a numpy restatement, and the device expansion arithmetic
(csrc/expand_math.h, compiled into tests/model) as its exact inverse."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def L():
    import rust_mdbg_b200
    return rust_mdbg_b200.ffi.lib()


@pytest.fixture(scope="module")
def model():
    subprocess.check_call(["make", "-C", ROOT, "-s", "model"], stdout=subprocess.DEVNULL)
    m = ctypes.CDLL(os.path.join(HERE, "model", "libka_bitslice_model.so"))
    m.bs_model_expand4.restype = ctypes.c_uint32
    m.bs_model_expand4.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    return m


def np_planes(b):
    n = len(b)
    nw = (n + 31) // 32
    x = np.full(nw * 32, ord("A"), np.uint8)
    x[:n] = b
    a = ((x >> 1) & 1).reshape(nw, 32).astype(np.uint64)
    c = ((x >> 2) & 1).reshape(nw, 32).astype(np.uint64)
    w = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    out = np.zeros(nw * 2, np.uint32)
    out[0::2] = (a * w).sum(axis=1).astype(np.uint32)
    out[1::2] = (c * w).sum(axis=1).astype(np.uint32)
    return out


@pytest.mark.parametrize("n,threads", [(0, 1), (1, 1), (31, 1), (32, 3), (33, 1), (4096, 2), (4097, 4), (1 << 20, 1),
                                       ((1 << 20) + 77, 8), (3 * 4096 * 64 + 5, 5)])
def test_pack_matches_numpy(L, n, threads):
    rng = np.random.default_rng(n + threads)
    b = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    nt = (n + 4095) // 4096
    bad_pos = []
    if n > 100:
        for p in rng.integers(0, n, 3):
            b[p] = rng.choice(np.frombuffer(b"NacgtRY\n\x00\xff", np.uint8))
            bad_pos.append(int(p))
    planes = np.full(2 * ((n + 31) // 32) + 2, 0xDEADBEEF, np.uint32)
    bad = np.full(nt + 1, 7, np.uint8)
    bb = b if n else np.zeros(1, np.uint8)
    assert L.mdbg_pack_bases_host(bb.ctypes.data, n, planes.ctypes.data, bad.ctypes.data, threads) == 0
    assert np.array_equal(planes[:-2], np_planes(b))
    assert planes[-1] == 0xDEADBEEF and planes[-2] == 0xDEADBEEF and bad[nt] == 7
    exp_bad = np.zeros(nt, np.uint8)
    for p in bad_pos:
        exp_bad[p // 4096] = 1
    assert np.array_equal(bad[:nt], exp_bad)


@pytest.mark.parametrize("env", [{}, {"MDBG_PACK_NO_AVX512": "1"}, {"MDBG_PACK_NO_AVX2": "1"}])
def test_every_byte_value_is_classified(env):
    """All 256 byte values, one per tile, at every position of a 64-byte group: only A, C, G, T leave a tile clean
    (AVX-512, AVX2 and SSSE3/scalar code paths: the dispatch is cached per process, hence the subprocess)."""
    code = r"""
import sys, numpy as np
sys.path.insert(0, @ROOT@)
import rust_mdbg_b200
L = rust_mdbg_b200.ffi.lib()
for shift in (0, 1, 15, 16, 31, 32, 33, 63):
    n = 256 * 4096
    b = np.full(n, ord("A"), np.uint8)
    b[1::2] = ord("G"); b[2::5] = ord("C"); b[3::7] = ord("T")
    pos = np.arange(256) * 4096 + 1024 + shift
    b[pos] = np.arange(256, dtype=np.uint8)
    planes = np.zeros(2 * (n // 32) + 2, np.uint32); bad = np.zeros(257, np.uint8)
    assert L.mdbg_pack_bases_host(b.ctypes.data, n, planes.ctypes.data, bad.ctypes.data, 2) == 0
    exp = np.ones(256, np.uint8); exp[[ord(c) for c in "ACGT"]] = 0
    assert np.array_equal(bad[:256], exp), (shift, np.nonzero(bad[:256] != exp)[0])
    assert planes[2 * (pos[65] // 32)] >> (pos[65] % 32) & 1 == 0 and planes[2 * (pos[67] // 32)] >> (pos[67] % 32) & 1 == 1
print("ok")
""".replace("@ROOT@", repr(ROOT))
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, env={**os.environ, **env}, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-500:], r.stderr[-2000:])


def test_expand_is_the_inverse(model):
    codes = "ACTG"
    for a in range(16):
        for b in range(16):
            w = model.bs_model_expand4(a | 0xABCDEF0, b | 0x1234560)      # only the low nibbles count
            got = bytes((w >> (8 * k)) & 0xFF for k in range(4))
            exp = bytes(ord(codes[((a >> k) & 1) | (((b >> k) & 1) << 1)]) for k in range(4))
            assert got == exp
