"""GPU check of mdbg_read_stats / `rust-mdbg --read-stats` (src/main.rs:939-975) against the oracle, in a
process of its own (the entry point is new and has not run on hardware when this file was written).

    python tests/read_stats_gpu_check.py

Exit code 0 only if the counts, the per-read offsets, the CLI's .read_stats file and the state of the
context afterwards (same minimizers, same graph) are all as the oracle says."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
import oracle_py  # noqa: E402
import rust_mdbg_b200 as M  # noqa: E402
from helpers import genome_reads, pack_reads, random_reads  # noqa: E402


def main():
    oracle_py.lib()
    if M.ffi.lib().mdbg_device_count() < 1:
        print(json.dumps({"ok": False, "error": "no CUDA device"}))
        return 2
    out = {"ok": False, "cases": []}
    for (k, l, d, minab) in [(7, 10, 0.01, 2), (10, 12, 0.003, 2), (5, 10, 0.02, 1)]:
        rng = np.random.default_rng(k)
        seqs = genome_reads(rng, 120000, 300, mean=6000, sd=1500, err=0.002)
        bases, off = pack_reads(seqs)
        query = seqs[::4] + random_reads(rng, 20, mean=5000, sd=1000) + [b"", b"ACGT", b"A" * 3000, seqs[0][:50]]
        qb, qo = pack_reads(query)
        with M.Context(M.Params(k=k, l=l, density=d, min_abundance=minab, presimp=0.01)) as ctx:
            ctx.push_reads(bases, off)
            g1 = ctx.finish()
            m1 = ctx.get_minimizers()
            cnt, coff = ctx.read_stats(qb, qo)
            cnt2, coff2 = ctx.read_stats(qb, qo)                 # twice: the arena is rolled back
            m2 = ctx.get_minimizers()
            g2 = ctx.finish()
        o = oracle_py.build_graph(bases, off, k, l, d, minab, 0.01)
        ecnt, eoff = o.read_stats(qb, qo)
        assert np.array_equal(coff, eoff), "per-read offsets differ"
        assert np.array_equal(cnt, ecnt), "counts differ"
        assert np.array_equal(cnt2, ecnt) and np.array_equal(coff2, eoff)
        for a, b in zip(m1, m2):
            assert np.array_equal(a, b), "the resident minimizers changed"
        for a in ("index", "abundance", "tuple", "e_n1", "e_n2"):
            assert np.array_equal(getattr(g1, a), getattr(g2, a)) and np.array_equal(getattr(g1, a), getattr(o, a)), a
        out["cases"].append({"k": k, "l": l, "d": d, "minab": minab, "counts": int(len(cnt)),
                             "nonzero": int((cnt > 0).sum())})
    # the front end: same flags as the reference, output {file}.read_stats, no .gfa
    k, l, d = 7, 10, 0.01
    rng = np.random.default_rng(99)
    seqs = genome_reads(rng, 60000, 120, mean=5000, sd=1000, err=0.002)
    query = seqs[::5] + random_reads(rng, 5, mean=3000, sd=500)
    with tempfile.TemporaryDirectory() as td:
        rp, qp = os.path.join(td, "reads.fa"), os.path.join(td, "query.fa")
        with open(rp, "wb") as f:
            for i, s in enumerate(seqs):
                f.write(b">r%d some description\n" % i + s + b"\n")
        with open(qp, "wb") as f:
            for i, s in enumerate(query):
                f.write(b">q%d other words\n" % i + s[:70] + b"\n" + s[70:] + b"\n")
        r = subprocess.run([os.path.join(ROOT, "rust-mdbg_b200", "rust-mdbg"), rp, "-k", str(k), "-l", str(l),
                            "--density", str(d), "--minabund", "2", "--prefix", os.path.join(td, "x"),
                            "--read-stats", qp], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "Read stats written, exiting." in r.stdout and "Stats module initialized." in r.stdout
        assert not os.path.exists(os.path.join(td, "x.gfa"))
        got = open(qp + ".read_stats").read().splitlines()
    bases, off = pack_reads(seqs)
    o = oracle_py.build_graph(bases, off, k, l, d, 2, 0.01)
    qb, qo = pack_reads(query)
    ecnt, eoff = o.read_stats(qb, qo)
    exp = ["q%d: %s" % (i, "".join("%d " % c for c in ecnt[int(eoff[i]):int(eoff[i + 1])])) for i in range(len(query))]
    assert got == [e.rstrip("\n") for e in exp], "CLI .read_stats differs"
    out["cli_lines"] = len(got)
    out["ok"] = True
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
