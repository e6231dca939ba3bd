"""Writes synthetic reads as FASTA files (plain and .gz) and times the `rust-mdbg` front end on them (parallel
host ingest + GPU path + file writers): the file-to-file numbers next to bench.py's memory-to-memory ones, and
next to the CPU restatement of the reference algorithm on the same reads (all host threads).
Usage: python tests/cli_ingest_timing.py [out_dir] [--big]
  default: BASELINE config 2 (251 Mbases); --big adds a 2.5 Gbase file (first 167k reads of config 3): the run
  where process start-up and CUDA context creation (~0.5 s) stop dominating."""
import gzip
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import rust_mdbg_b200 as m


def write_fasta(path, bases, ro, n):
    # one big buffer: ">r<i>\n" + sequence + "\n" per read
    parts = []
    for r in range(n):
        parts.append(b">r%d\n" % r)
        parts.append(bases[int(ro[r]):int(ro[r + 1])].tobytes())
        parts.append(b"\n")
    data = b"".join(parts)
    with open(path, "wb") as f:
        f.write(data)
    return data


def run_cli(exe, fa, prefix, k, d, extra, total):
    t = time.time()
    env = dict(os.environ, MDBG_CLI_TIMING="1")
    r = subprocess.run([exe, fa, "-k", str(k), "-l", "12", "--density", str(d), "--minabund", "2", "--prefix", prefix] + extra,
                       capture_output=True, text=True, env=env)
    dt = time.time() - t
    tail = [x for x in r.stdout.splitlines() if x.startswith(("Number of", "Total execution", "Maximum RSS"))]
    timing = [x for x in r.stderr.splitlines() if x.startswith("[timing]")]
    print("rust-mdbg %s %s: rc=%d wall %.2f s = %.2f Gbases/s file->.gfa | %s | %s" %
          (os.path.basename(fa), " ".join(extra) or "(default)", r.returncode, dt, total / dt / 1e9, "; ".join(tail), " ".join(timing)))
    sys.stdout.flush()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    out = args[0] if args else "/tmp"
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(ROOT, "rust-mdbg_b200", "rust-mdbg")
    cases = [("ecoli50x", 5_000_000, None, 21, 0.003)]
    if "--big" in sys.argv:
        cases.append(("dmel50x_first167k", 140_000_000, 167000, 35, 0.002))
    for name, glen, n_reads, k, d in cases:
        s = m.Synth(genome_len=glen)
        n = n_reads or s.num_reads(50.0)
        ro, total = s.plan(0, n)
        bases = s.fill_host(0, n, ro, threads=32)
        fa = os.path.join(out, name + ".fa")
        t = time.time()
        data = write_fasta(fa, bases, ro, n)
        print("wrote %s: %d reads, %d bases in %.1f s" % (fa, n, total, time.time() - t))
        run_cli(exe, fa, os.path.join(out, name), k, d, ["--no-basespace"], total)     # warm page cache + first CUDA init
        run_cli(exe, fa, os.path.join(out, name), k, d, ["--no-basespace"], total)
        run_cli(exe, fa, os.path.join(out, name), k, d, [], total)
        if name == "ecoli50x":
            gz = fa + ".gz"
            with gzip.open(gz, "wb", compresslevel=1) as f:
                f.write(data)
            run_cli(exe, gz, os.path.join(out, name + "_gz"), k, d, ["--no-basespace"], total)
        del data
        # the CPU restatement of the reference algorithm on the same reads, memory to memory (no parsing, no files)
        import oracle_py
        cores = len(os.sched_getaffinity(0))
        t = time.time()
        g = oracle_py.build_graph(bases, ro, k, 12, d, 2, 0.01, threads=cores)
        dt = time.time() - t
        print("reference-algorithm CPU restatement (%d threads), memory to memory: %.2f s = %.2f Gbases/s; nodes %d edges %d" %
              (cores, dt, total / dt / 1e9, g.stats["n_nodes"], g.stats["n_edges"]))
        g.close()


if __name__ == "__main__":
    main()
