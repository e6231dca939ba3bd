#!/usr/bin/env python3
"""Graph digests of the serial oracle on the synthetic BASELINE workloads at (near) bench size:
config 2 in full (ecoli50x, 251 Mbases) and the first 70,000 reads of config 3 (dmel50x, 1.05 Gbases).

    python tests/golden/make_synth_hashes.py        -> tests/golden/synth_graph_hashes.json

The digests pin the oracle (tests/test_oracle_golden.py re-derives them on the CPU) and give the GPU
tests a second, oracle-independent check at sizes where printing arrays is useless."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

CASES = {
    "ecoli50x_full": dict(genome_len=5_000_000, n_reads=None, coverage=50.0, k=21, l=12, density=0.003),
    "dmel50x_first70k": dict(genome_len=140_000_000, n_reads=70000, coverage=50.0, k=35, l=12, density=0.002),
}
ARRAYS = ("index", "abundance", "seqlen", "shift", "tuple", "e_n1", "e_n2", "e_o1", "e_o2", "e_ov")


def graph_digest(g):
    h = hashlib.sha256()
    for a in ARRAYS:
        h.update(a.encode())
        h.update(np.ascontiguousarray(getattr(g, a)).tobytes())
    return h.hexdigest()


def minimizer_digest(h_, p_, off_):
    h = hashlib.sha256()
    for a in (h_, p_, off_):
        h.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return h.hexdigest()


def build_case(name, threads=8):
    import rust_mdbg_b200 as m
    c = CASES[name]
    s = m.Synth(genome_len=c["genome_len"])
    n = c["n_reads"] or s.num_reads(c["coverage"])
    ro, total = s.plan(0, n)
    host = s.fill_host(0, n, ro, threads=threads)
    return c, host, ro, total


def main():
    import oracle_py
    out = {}
    for name in CASES:
        c, host, ro, total = build_case(name)
        o = oracle_py.build_graph(host, ro, c["k"], c["l"], c["density"], 2, 0.01)
        out[name] = {"params": c, "n_bases": int(total), "stats": {k: int(v) for k, v in o.stats.items()},
                     "graph_sha256": graph_digest(o), "minimizers_sha256": minimizer_digest(o.m_hash, o.m_pos, o.m_off)}
        print(name, out[name]["stats"], out[name]["graph_sha256"][:16])
        o.close()
    with open(os.path.join(HERE, "synth_graph_hashes.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
