"""GPU check of the bit-sliced K-A variant (run as a script in its own process so that a fault in an
experimental kernel cannot poison the CUDA context of the other GPU tests):

    python tests/bitslice_gpu_check.py [--quick] [--out FILE]

Through the C ABI with ka_variant = 2: per-read minimizers against the oracle on the parity cases of
test_gpu_parity.py (random + edge reads, N, --skiphpc, BASELINE config 1), the whole graph of config 1,
then classic == bit-sliced on a large synthetic batch with the K-A kernel times of both.  Prints one
JSON line; exit code 0 only if every comparison was bit-exact and the variant really ran."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_py  # noqa: E402
import rust_mdbg_b200 as M  # noqa: E402
from helpers import load_fasta, pack_reads, random_reads  # noqa: E402

EDGE_READS = [b"", b"A", b"ACGTACGTAC", b"A" * 5000, b"AC" * 4000, b"ACG" * 3000, b"ACGTACGTACG",
              b"", b"", b"T" * 70000, b"ACGT" * 5000 + b"A" * 300 + b"CGTA" * 100]


def oracle_minimizers(seqs, l, d, hpc=True):
    hs, ps, off = [], [], [0]
    for s in seqs:
        h, p = oracle_py.extract(s, l, d, hpc=hpc)
        hs.append(h); ps.append(p); off.append(off[-1] + len(h))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint64)
    return cat(hs), cat(ps), np.array(off, np.uint64)


def check_extract(seqs, l, d, hpc=True, expect_variant=2):
    bases, off = pack_reads(seqs)
    with M.Context(M.Params(k=5, l=l, density=d, hpc=hpc, ka_variant=2)) as ctx:
        h, p, mo = ctx.extract_minimizers(bases, off)
        tm = ctx.timings()
    eh, ep, eo = oracle_minimizers(seqs, l, d, hpc)
    assert tm["ka_variant_used"] == expect_variant, tm
    assert np.array_equal(mo, eo), "per-read minimizer offsets differ"
    assert np.array_equal(h, eh), "hashes differ"
    assert np.array_equal(p, ep), "positions differ"
    return int(tm["ka_dirty_tiles"]), len(h)


def main():
    quick = "--quick" in sys.argv
    out = {"ok": False, "cases": []}
    oracle_py.lib()
    if M.ffi.lib().mdbg_device_count() < 1:
        print(json.dumps({"ok": False, "error": "no CUDA device"}))
        return 2
    t_start = time.time()
    # 1. parity cases
    for l, d in [(12, 0.003), (10, 0.0008), (14, 0.002), (12, 0.002)]:
        rng = np.random.default_rng(100 + l)
        seqs = random_reads(rng, 40, mean=9000, sd=4000, lo=0, hi=40000, hp=0.25) + EDGE_READS
        seqs += random_reads(rng, 300, mean=40, sd=30, lo=0, hi=200)
        nd, n = check_extract(seqs, l, d)
        out["cases"].append({"case": "random+edge", "l": l, "d": d, "minimizers": n, "dirty_tiles": nd})
    rng = np.random.default_rng(5)
    seqs = random_reads(rng, 30, mean=7000, sd=3000, hp=0.3) + EDGE_READS
    nd, n = check_extract(seqs, 12, 0.003, hpc=False)
    out["cases"].append({"case": "skiphpc", "minimizers": n, "dirty_tiles": nd})
    rng = np.random.default_rng(6)
    seqs = random_reads(rng, 20, mean=8000, sd=2000, hp=0.2)
    withn = []
    for i, s in enumerate(seqs):
        a = bytearray(s)
        for _ in range(i % 5):
            j = int(rng.integers(0, len(a))); n_ = int(rng.integers(1, 40))
            a[j:j + n_] = b"N" * len(a[j:j + n_])
        withn.append(bytes(a))
    nd, n = check_extract(withn + [b"N" * 3000, b"ACGT" * 10 + b"N" + b"ACGT" * 10], 12, 0.003)
    out["cases"].append({"case": "with N", "minimizers": n, "dirty_tiles": nd})
    # outside the variant's range: must fall back to the classic kernel, same results
    nd, n = check_extract(random_reads(rng, 5, mean=3000, sd=500), 12, 0.10, expect_variant=1)
    out["cases"].append({"case": "d=0.10 -> classic", "minimizers": n})
    # illegal byte: same error as the classic kernel
    with M.Context(M.Params(k=5, l=10, density=0.003, ka_variant=2)) as ctx:
        good = b"ACGTTGCATGCATGACTGACTAGCTAGCATCGATCAGCTACGACTAGC" * 10
        try:
            ctx.read_extract(good[:100] + b"a" + good[100:])
            raise AssertionError("illegal byte accepted")
        except M.MdbgError as e:
            assert e.code == -4
    # 2. BASELINE config 1: minimizers and the whole graph
    bases, off, _ = load_fasta(os.path.join(HERE, "golden", "config1_reads.fa.gz"))
    with M.Context(M.Params(k=7, l=10, density=0.0008, min_abundance=2, presimp=0.01, ka_variant=2)) as ctx:
        ctx.push_reads(bases, off)
        h, p, mo = ctx.get_minimizers()
        g = ctx.finish()
        tm = ctx.timings()
    o = oracle_py.build_graph(bases, off, 7, 10, 0.0008, 2, 0.01)
    assert tm["ka_variant_used"] == 2
    assert len(h) == 16069 and np.array_equal(h, o.m_hash) and np.array_equal(p, o.m_pos) and np.array_equal(mo, o.m_off)
    for a in ("index", "abundance", "seqlen", "shift", "tuple", "e_n1", "e_n2", "e_o1", "e_o2", "e_ov"):
        assert np.array_equal(getattr(g, a), getattr(o, a)), a
    out["cases"].append({"case": "config1", "minimizers": len(h), "nodes": int(g.stats["n_nodes"]),
                         "edges": int(g.stats["n_edges"]), "dirty_tiles": int(tm["ka_dirty_tiles"])})
    # 3. classic == bit-sliced on a large batch, with the K-A kernel time of both
    nb = (40 if quick else 250) * 1000 * 1000
    rng = np.random.default_rng(77)
    nreads = nb // 15000
    lens = np.clip(rng.normal(15000, 4000, nreads), 1000, 60000).astype(np.int64)
    off = np.zeros(nreads + 1, np.uint64); np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(off[-1]), dtype=np.uint8)]
    res = {}
    for name, var in (("classic", 1), ("bitslice", 2)):
        with M.Context(M.Params(k=21, l=12, density=0.003, ka_variant=var)) as ctx:
            best = 1e9
            for it in range(3):
                ctx.reset()
                ctx.push_reads(bases, off)
                best = min(best, ctx.timings()["ms_ka_kernel"])
            tm = ctx.timings()
            res[name] = ctx.get_minimizers() + (best, int(tm["ka_dirty_tiles"]), int(tm["ka_variant_used"]))
    for i in range(3):
        assert np.array_equal(res["classic"][i], res["bitslice"][i]), "classic != bit-sliced on the large batch"
    assert res["bitslice"][5] == 2
    B = int(off[-1])
    out["large"] = {"bases": B, "minimizers": len(res["classic"][0]),
                    "classic_ms": res["classic"][3], "bitslice_ms": res["bitslice"][3],
                    "classic_GBps": B * 1.056e-6 / res["classic"][3], "bitslice_GBps": B * 1.056e-6 / res["bitslice"][3],
                    "dirty_tiles": res["bitslice"][4]}
    out["ok"] = True
    out["seconds"] = round(time.time() - t_start, 1)
    line = json.dumps(out)
    print(line)
    if "--out" in sys.argv:
        with open(sys.argv[sys.argv.index("--out") + 1], "w") as f:
            f.write(line + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
