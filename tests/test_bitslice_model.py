"""CPU tests of the bit-sliced K-A variant: the kernel body (rust-mdbg_b200/csrc/ka_bitslice_body.h)
is compiled by g++ into tests/model/libka_bitslice_model.so, where 32 host threads per warp execute
it with emulated warp primitives, and compared with the oracle bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import H, pack_reads, py_ntc64, random_reads, rol

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TILE = 4096


@pytest.fixture(scope="module")
def model():
    """The kernel body compiled for the CPU (tests/model/)."""
    subprocess.check_call(["make", "-C", ROOT, "-s", "model"], stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(os.path.join(HERE, "model", "libka_bitslice_model.so"))
    vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
    L.bs_model_run.restype = ctypes.c_int
    L.bs_model_run.argtypes = [vp, vp, u64, u64, u32, u64, ctypes.c_int, u32, ctypes.c_int, u64, u64,
                               vp, vp, vp, vp, u64, vp, vp, vp, vp, vp]
    L.bs_model_filter.restype = u32
    L.bs_model_filter.argtypes = [u32] * 5
    L.bs_model_exact.restype = u64
    L.bs_model_exact.argtypes = [u32] * 3
    L.bs_model_planes.argtypes = [vp] * 4
    L.bs_model_pext.argtypes = [u32, vp, vp]
    L.bs_model_select.restype = u32
    L.bs_model_select.argtypes = [u32, u32]
    return L


CODE = {ord('A'): 0, ord('C'): 1, ord('T'): 2, ord('G'): 3}
COMPL = {ord('A'): ord('T'), ord('C'): ord('G'), ord('G'): ord('C'), ord('T'): ord('A')}


def test_planes_and_alphabet(model):
    rng = np.random.default_rng(1)
    al = np.frombuffer(b"ACGT", np.uint8)
    for it in range(200):
        s = al[rng.integers(0, 4, 32)].copy()
        if it % 4 == 3:
            s[int(rng.integers(0, 32))] = int(rng.integers(0, 256))
        a, b, bad = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        model.bs_model_planes(s.ctypes.data, ctypes.byref(a), ctypes.byref(b), ctypes.byref(bad))
        ok = all(int(c) in CODE for c in s)
        assert (bad.value == 0) == ok
        if ok:
            assert a.value == sum((CODE[int(c)] & 1) << i for i, c in enumerate(s))
            assert b.value == sum((CODE[int(c)] >> 1) << i for i, c in enumerate(s))


def test_pext_and_select(model):
    rng = np.random.default_rng(2)
    special = [0, 0xFFFFFFFF, 1, 0x80000000, 0x55555555]
    for it in range(500):
        m = int(rng.integers(0, 1 << 32)) if it % 5 else special[it // 5 % 5]
        x, y = int(rng.integers(0, 1 << 32)), int(rng.integers(0, 1 << 32))
        cx, cy = ctypes.c_uint32(x), ctypes.c_uint32(y)
        model.bs_model_pext(m, ctypes.byref(cx), ctypes.byref(cy))
        bits = [i for i in range(32) if (m >> i) & 1]
        assert cx.value == sum(((x >> b) & 1) << j for j, b in enumerate(bits))
        assert cy.value == sum(((y >> b) & 1) << j for j, b in enumerate(bits))
        for k, b in enumerate(bits):
            assert model.bs_model_select(m, k) == b


@pytest.mark.parametrize("l", [10, 12, 14])
def test_filter_and_exact_hash(model, l):
    """filter bit k set <=> top 8 bits of fh or of rh of window k are zero; exact hash == ntc64."""
    rng = np.random.default_rng(l)
    al = np.frombuffer(b"ACGT", np.uint8)
    hits = 0
    for it in range(300):
        s = bytes(al[rng.integers(0, 4, 64)])
        a = sum((CODE[c] & 1) << i for i, c in enumerate(s))
        b = sum((CODE[c] >> 1) << i for i, c in enumerate(s))
        got = model.bs_model_filter(l, a & 0xFFFFFFFF, a >> 32, b & 0xFFFFFFFF, b >> 32)
        for k in range(25):
            f = r = 0
            for j in range(l):
                f ^= rol(H[s[k + j]], l - 1 - j)
                r ^= rol(H[COMPL[s[k + j]]], j)
            exp = (f >> 56) == 0 or (r >> 56) == 0
            assert ((got >> k) & 1) == int(exp), (it, k)
            hits += exp
            if k < 3:
                av, bv = (a >> k) & ((1 << l) - 1), (b >> k) & ((1 << l) - 1)
                assert model.bs_model_exact(l, av, bv) == py_ntc64(s, k, l)
    assert hits > 0


def run_model(model, oracle, seqs, l, d, hpc=True, group=4, n_warps=1, max_dirty=None, return_dirty=False,
              launches=None):
    """Run the emulated kernel over the batch and compare every clean tile with the oracle:
    its (hash, pos) slice in order, and the (tile, rank) it leaves for every read starting in it.
    Returns (dirty tiles, minimizers checked, tiles)."""
    bases, off = pack_reads(seqs)
    B, R = int(off[-1]), len(seqs)
    n_tiles = max(1, (B + TILE - 1) // TILE)
    bound = oracle.lib().orc_hash_bound(d)
    cap = max(4096, B // 4)
    tile_cnt = np.full(n_tiles, 0xDEAD, np.uint64)
    tile_soff = np.zeros(n_tiles, np.uint64)
    sh, sp = np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    oro = np.full(R + 1, 2**64 - 1, np.uint64)
    dl, dn, st = np.zeros(n_tiles, np.uint32), ctypes.c_uint32(0), ctypes.c_uint64(0)
    bb = bases if B else np.zeros(1, np.uint8)
    # one launch over all tiles, or several over consecutive tile ranges (the chunks of an overlapped upload)
    cuts = [0] + sorted(set(min(max(1, int(c)), n_tiles) for c in (launches or []))) + [n_tiles]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a >= b:
            continue
        rc = model.bs_model_run(bb.ctypes.data, off.ctypes.data, R, B, l, bound, int(hpc), group, n_warps, a, b,
                                tile_cnt.ctypes.data, tile_soff.ctypes.data, sh.ctypes.data, sp.ctypes.data, cap,
                                oro.ctypes.data, dl.ctypes.data, ctypes.byref(dn), ctypes.byref(st), None)
        assert rc == 0
    dirty = set(int(x) for x in dl[:dn.value])
    assert len(dirty) == dn.value
    per_tile = [[] for _ in range(n_tiles)]          # oracle minimizers by the tile of their start
    for r, s in enumerate(seqs):
        h, p = oracle.extract(s, l, d, hpc=hpc)
        for hv, pv in zip(h, p):
            ab = int(off[r]) + int(pv)
            per_tile[ab // TILE].append((ab, int(hv), int(pv)))
    n_checked = 0
    for t in range(n_tiles):
        if t in dirty:
            assert tile_cnt[t] == 0xDEAD, "a dirty tile must not emit"
            continue
        exp = per_tile[t]
        cnt, so = int(tile_cnt[t]), int(tile_soff[t])
        assert cnt == len(exp), (t, cnt, len(exp))
        assert [int(x) for x in sh[so:so + cnt]] == [e[1] for e in exp], t
        assert [int(x) for x in sp[so:so + cnt]] == [e[2] for e in exp], t
        n_checked += cnt
    assert int(st.value) == sum(len(per_tile[t]) for t in range(n_tiles) if t not in dirty)
    # read r (and the end sentinel r = R) is reported by the tile j with j*TILE <= read_off[r] < (j+1)*TILE,
    # the last tile taking everything at or past its end (ka_tile_lb_kernel)
    for r in range(R + 1):
        x = int(off[r])
        j = min(x // TILE, n_tiles - 1)
        if j in dirty:
            continue
        rank = sum(1 for e in per_tile[j] if e[0] < x)
        assert int(oro[r]) == (j << 32) | rank, (r, x, j, int(oro[r]) >> 32, int(oro[r]) & 0xFFFFFFFF, rank)
    if max_dirty is not None:
        assert len(dirty) <= max_dirty, sorted(dirty)
    return (dirty if return_dirty else len(dirty)), n_checked, n_tiles


@pytest.mark.parametrize("l,d", [(10, 0.0008), (12, 0.003), (12, 0.002), (14, 0.003)])
def test_model_random_reads(model, oracle, l, d):
    rng = np.random.default_rng(100 + l)
    seqs = random_reads(rng, 12, mean=9000, sd=4000, lo=0, hi=30000, hp=0.25)
    nd, n, nt = run_model(model, oracle, seqs, l, d, max_dirty=0)
    assert n > 50


EDGE_READS = [b"", b"A", b"ACGTACGTAC", b"A" * 5000, b"AC" * 4000, b"ACG" * 3000, b"ACGTACGTACG",
              b"", b"", b"T" * 9000, b"ACGT" * 5000 + b"A" * 300 + b"CGTA" * 100]


@pytest.mark.parametrize("group,n_warps", [(1, 1), (2, 3), (4, 2), (7, 1)])
def test_model_edge_reads(model, oracle, group, n_warps):
    """Homopolymers, short-period repeats, empty and tiny reads: clean tiles must be exact, the rest
    must be handed over (dirty), for every group size / number of concurrent warps."""
    rng = np.random.default_rng(7)
    seqs = random_reads(rng, 6, mean=7000, sd=3000, hp=0.3) + EDGE_READS
    seqs += random_reads(rng, 200, mean=40, sd=30, lo=0, hi=200)          # many tiny reads in one tile
    seqs += random_reads(rng, 3, mean=6000, sd=100)
    nd, n, nt = run_model(model, oracle, seqs, 12, 0.003, group=group, n_warps=n_warps)
    assert nd < nt // 2 and n > 100


def test_model_skiphpc(model, oracle):
    rng = np.random.default_rng(5)
    seqs = random_reads(rng, 8, mean=7000, sd=3000, hp=0.3) + EDGE_READS[:7]
    nd, n, nt = run_model(model, oracle, seqs, 12, 0.003, hpc=False)
    assert n > 100


def test_model_non_acgt_goes_dirty(model, oracle):
    """N (hashes as 0) and illegal bytes have no 2-bit code: their tile, and the tile below whose
    look-ahead they are, must be handed to the exact kernel; every other tile stays exact."""
    rng = np.random.default_rng(6)
    seqs = [bytearray(s) for s in random_reads(rng, 5, mean=20000, sd=10, hp=0.2)]
    seqs[1][5000:5010] = b"N" * 10
    seqs[3][777] = ord("N")
    seqs[4][4096 * 3 - 1] = ord("N")
    seqs = [bytes(s) for s in seqs]
    bases, _ = pack_reads(seqs)
    bad_tiles = set(int(i) // TILE for i in np.nonzero(bases == ord("N"))[0])
    dirty, n, nt = run_model(model, oracle, seqs, 12, 0.003, return_dirty=True)
    assert bad_tiles <= dirty <= bad_tiles | set(t - 1 for t in bad_tiles)
    assert n > 100


@pytest.mark.parametrize("nbytes", [0, 1, 11, 12, 4095, 4096, 4097, 8192, 4096 * 3 + 31, 4096 * 3 + 33])
def test_model_batch_sizes(model, oracle, nbytes):
    """Batches that end on / next to tile and look-ahead boundaries (one read, and cut in three)."""
    rng = np.random.default_rng(nbytes)
    al = np.frombuffer(b"ACGT", np.uint8)
    s = al[rng.integers(0, 4, nbytes)].tobytes()
    run_model(model, oracle, [s], 10, 0.003, max_dirty=0)
    run_model(model, oracle, [s[:nbytes // 3], b"", s[nbytes // 3:nbytes // 2], s[nbytes // 2:]], 10, 0.003,
              group=2, max_dirty=0)


def test_model_read_starts_on_tile_edges(model, oracle):
    rng = np.random.default_rng(9)
    al = np.frombuffer(b"ACGT", np.uint8)
    lens = [4096, 4096 - 5, 5, 4096 + 20, 4096 * 2 - 20, 10, 11, 12, 13, 4096 - 46, 6000]
    seqs = [al[rng.integers(0, 4, n)].tobytes() for n in lens]
    # make neighbouring reads share the base across the boundary (a run must still start there)
    seqs = [seqs[0]] + [seqs[i - 1][-1:] + s[1:] for i, s in enumerate(seqs) if i > 0]
    for g in (1, 3):
        run_model(model, oracle, seqs, 12, 0.003, group=g, max_dirty=0)


def adversarial_batch(rng, n_tiles_target):
    """Reads whose starts/ends crowd the tile edges and the 32-byte look-ahead, homopolymer-rich
    stretches (look-ahead with too few runs), N next to tile edges, batch end next to an edge."""
    al = np.frombuffer(b"ACGT", np.uint8)
    seqs, total = [], 0
    target = n_tiles_target * TILE + int(rng.choice([-33, -32, -31, -1, 0, 1, 12, 31, 32, 33, 2000]))
    while total < target:
        kind = rng.integers(0, 6)
        to_edge = (-total) % TILE
        if kind == 0:
            ln = to_edge + int(rng.integers(-14, 15))          # end next to the coming tile edge
        elif kind == 1:
            ln = to_edge + int(rng.integers(0, 45))            # end inside the look-ahead of the tile
        elif kind == 2:
            ln = int(rng.integers(0, 30))                      # tiny (often shorter than l)
        else:
            ln = int(rng.integers(200, 9000))
        ln = max(0, min(ln, target - total))
        hp = float(rng.choice([0.0, 0.25, 0.6, 0.8]))
        s = al[rng.integers(0, 4, ln)]
        if hp > 0 and ln > 1:
            rep = rng.random(ln) < hp
            rep[0] = False
            idx = np.where(~rep, np.arange(ln), 0)
            np.maximum.accumulate(idx, out=idx)
            s = s[idx]
        s = bytearray(s.tobytes())
        if seqs and s and rng.random() < 0.5:
            prev = next((q for q in reversed(seqs) if q), b"")
            if prev:
                s[0] = prev[-1]                                # same base across the read boundary
        if ln > 0 and rng.random() < 0.04:
            e = (-total) % TILE                                # N on the last / first byte of a tile
            for q in (e - 1, e):
                if 0 <= q < ln and rng.random() < 0.5:
                    s[q] = ord("N")
        seqs.append(bytes(s))
        total += ln
        if ln == 0 and rng.random() < 0.5:
            seqs.append(b"")
    return seqs


@pytest.mark.parametrize("seed", range(12))
def test_model_adversarial(model, oracle, seed):
    """Highest density the variant supports (bound just below 2^56) so that the rare events --
    a minimizer whose window crosses a tile edge, a read start inside a look-ahead, a look-ahead
    that runs out of runs -- happen many times per batch."""
    rng = np.random.default_rng(1000 + seed)
    l = [10, 12, 14][seed % 3]
    seqs = adversarial_batch(rng, int(rng.integers(3, 40)))
    dirty, n, nt = run_model(model, oracle, seqs, l, 0.0039, group=int(rng.integers(1, 6)),
                             n_warps=int(rng.integers(1, 4)), hpc=(seed % 4 != 3), return_dirty=True)
    assert n > 0


def passing_lmer(rng, l, bound, ok):
    """A random l-mer without repeated neighbours whose canonical ntHash is <= bound and ok(s)."""
    al = b"ACGT"
    while True:
        s = bytearray()
        while len(s) < l:
            c = al[int(rng.integers(0, 4))]
            if not s or s[-1] != c:
                s.append(c)
        s = bytes(s)
        if ok(s) and py_ntc64(s, 0, l) <= bound:
            return s


def rand_seq(rng, n):
    return np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes()


def test_model_n_before_tile_edge(model, oracle):
    """The byte before a tile is N: the first base of the tile starts a run whatever its code."""
    rng = np.random.default_rng(21)
    s = bytearray(rand_seq(rng, 64 * TILE))
    bound = oracle.lib().orc_hash_bound(0.0039)
    for t in range(2, 64, 3):
        s[t * TILE - 1] = ord("N")
        w = passing_lmer(rng, 12, bound, lambda x: True)      # the window that starts the tile is a minimizer
        s[t * TILE:t * TILE + 12] = w
        if s[t * TILE + 12] == w[-1]:
            s[t * TILE + 12] = b"ACGT"[(b"ACGT".index(w[-1]) + 1) % 4]
        s[t * TILE + 127] = w[0]          # = what lane 0 would compare with if it looked at itself
    dirty, n, nt = run_model(model, oracle, [bytes(s)], 12, 0.0039, group=4, return_dirty=True)
    bad = set(range(1, 63, 3))
    assert bad <= dirty <= bad | set(t - 1 for t in bad)


@pytest.mark.parametrize("delta", [1, 5, 17, 31])
def test_model_phantom_window_past_batch_end(model, oracle, delta):
    """The batch ends inside the look-ahead of a top tile with l-1 runs that, continued by the 'A'
    filler, would hash below the bound: no minimizer may come out of the filler."""
    l, d = 12, 0.0039
    bound = oracle.lib().orc_hash_bound(d)
    rng = np.random.default_rng(delta)
    for _ in range(6):
        w = passing_lmer(rng, l, bound, lambda s: s[-1:] == b"A")
        x = w[:-1]                                  # l-1 runs; the filler would supply the final A
        B = TILE + delta
        head = bytearray(rand_seq(rng, B - len(x)))
        if head[-1] == x[0]:
            head[-1] = b"ACGT"[(b"ACGT".index(x[0]) + 1) % 4]
        run_model(model, oracle, [bytes(head) + x], l, d, group=1, max_dirty=0)
        run_model(model, oracle, [bytes(head[:100]), bytes(head[100:]) + x], l, d, group=1, max_dirty=0)


def test_model_window_across_long_homopolymer(model, oracle):
    """A minimizer whose window spans a homopolymer longer than a tile: the tiles inside the run have
    no run start, the look-ahead of the tile before them is short although the read goes on -- the
    tile must be handed over, not silently lose the minimizer."""
    l, d = 12, 0.0039
    bound = oracle.lib().orc_hash_bound(d)
    rng = np.random.default_rng(33)
    seqs = []
    for i in range(10):
        j = int(rng.integers(3, 9))
        w = passing_lmer(rng, l, bound, lambda s: s[j:j + 1] == b"A")
        pre = bytearray(rand_seq(rng, int(rng.integers(3000, 5000))))
        if pre[-1] == w[0]:
            pre[-1] = b"ACGT"[(b"ACGT".index(w[0]) + 1) % 4]
        post = bytearray(rand_seq(rng, int(rng.integers(100, 3000))))
        if post[0] == w[-1]:
            post[0] = b"ACGT"[(b"ACGT".index(w[-1]) + 1) % 4]
        seqs.append(bytes(pre) + w[:j] + b"A" * int(rng.integers(4200, 9000)) + w[j + 1:] + bytes(post))
    bases, off = pack_reads(seqs)
    # the oracle does find those minimizers
    for r, s in enumerate(seqs):
        h, p = oracle.extract(s, l, d)
        assert len(h) > 0
    dirty, n, nt = run_model(model, oracle, seqs, l, d, group=3, return_dirty=True)
    assert len(dirty) >= 1 and n > 0


@pytest.mark.parametrize("group", [1, 4])
def test_model_read_start_inside_look_ahead(model, oracle, group):
    """An l-mer that would be a minimizer straddles a tile edge AND a read boundary lying in the
    look-ahead of the lower tile: it belongs to no read and must not be emitted (top-tile and
    carried look-ahead)."""
    l, d = 12, 0.0039
    bound = oracle.lib().orc_hash_bound(d)
    rng = np.random.default_rng(55 + group)
    seqs, total = [], 0
    for i in range(24):
        w = passing_lmer(rng, l, bound, lambda s: True)
        edge = (total // TILE + 2) * TILE                  # a tile edge comfortably ahead
        delta = int(rng.integers(1, 11))                   # the next read starts delta bytes after the edge
        j = int(rng.integers(1, l - delta)) + delta        # bases of w before the read boundary (> delta)
        start_w = edge + delta - j                         # w[0] lies before the edge: owned by the lower tile
        pre = bytearray(rand_seq(rng, start_w - total))
        if pre and pre[-1] == w[0]:
            pre[-1] = b"ACGT"[(b"ACGT".index(w[0]) + 1) % 4]
        seqs.append(bytes(pre) + w[:j])
        total += len(seqs[-1])
        assert total == edge + delta
        tail = bytearray(rand_seq(rng, int(rng.integers(50, 3000))))
        if tail[0] == w[-1]:
            tail[0] = b"ACGT"[(b"ACGT".index(w[-1]) + 1) % 4]
        seqs.append(w[j:] + bytes(tail))
        total += len(seqs[-1])
    run_model(model, oracle, seqs, l, d, group=group, max_dirty=0)


@pytest.mark.parametrize("seed", range(4))
def test_model_chunked_launches(model, oracle, seed):
    """Several launches over consecutive tile ranges sharing the staging arrays and the dirty list, as
    run_ka issues them while an upload is in flight: same result as one launch."""
    rng = np.random.default_rng(300 + seed)
    seqs = adversarial_batch(rng, 30)
    nt = (sum(len(s) for s in seqs) + TILE - 1) // TILE
    cuts = sorted(int(x) for x in rng.integers(1, nt, 4))
    run_model(model, oracle, seqs, 12, 0.0039, group=int(rng.integers(1, 5)), n_warps=2, launches=cuts)
