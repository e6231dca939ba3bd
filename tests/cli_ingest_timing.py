"""Writes BASELINE config 2's synthetic reads as a plain FASTA file and times the `rust-mdbg` front
end on it (host ingest + GPU path + file writers): the file-to-file number next to bench.py's
memory-to-memory ones.  Usage: python tests/cli_ingest_timing.py [out_dir]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import rust_mdbg_b200 as m


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
    s = m.Synth(genome_len=5_000_000)
    n = s.num_reads(50.0)
    ro, total = s.plan(0, n)
    bases = s.fill_host(0, n, ro, threads=32)
    fa = os.path.join(out, "ecoli50x.fa")
    t = time.time()
    with open(fa, "wb") as f:
        for r in range(n):
            f.write(b">r%d\n" % r)
            f.write(bases[int(ro[r]):int(ro[r + 1])].tobytes())
            f.write(b"\n")
    print("wrote %s: %d reads, %d bases in %.1f s" % (fa, n, total, time.time() - t))
    exe = os.path.join(ROOT, "rust-mdbg_b200", "rust-mdbg")
    for extra in ([], ["--no-basespace"]):
        t = time.time()
        r = subprocess.run([exe, fa, "-k", "21", "-l", "12", "--density", "0.003", "--minabund", "2", "--prefix",
                            os.path.join(out, "ecoli50x")] + extra, capture_output=True, text=True)
        dt = time.time() - t
        tail = [x for x in r.stdout.splitlines() if x.startswith(("Number of", "Total execution", "Maximum RSS"))]
        print("rust-mdbg %s: rc=%d wall %.2f s = %.2f Gbases/s | %s" % (" ".join(extra) or "(default)", r.returncode, dt,
                                                                        total / dt / 1e9, "; ".join(tail)))


if __name__ == "__main__":
    main()
