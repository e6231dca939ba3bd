#!/usr/bin/env python
"""Extracts the golden vectors this repo pins its oracle with FROM THE REFERENCE TREE
(/root/reference, present only in the build container) into small committed fixtures.

  reference_sequences_lines.json
      the two `.sequences` / unitig lines the reference authors pasted into comments as
      examples of their program's OUTPUT:
        src/to_basespace.rs:203                            (k=7 node, minimizers + sequence + shift)
        experiments/661k_genomes/scan_genomes_minmers.py:38  (k=10 node, minimizers + sequence)
      They pin: the ntHash restatement (17 exact 64-bit hashes), the inclusive density
      threshold (no other l-mer of the sequence may be selected), the slice convention
      seq = raw[p_i .. p_{i+k-1}+l) and the shift pair (main.rs:769-778).
  config1_reads.fa.gz
      BASELINE config #1 input: the 657 reads of the reference's example/reads-0.00.fa.gz,
      re-compressed (test DATA, needed on the GPU box where /root/reference does not exist).

Run:  python tests/golden/make_golden.py      (needs /root/reference)
"""
import gzip
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    out = []
    src = open(os.path.join(REF, "src/to_basespace.rs")).read().split("\n")[202]
    m = re.search(r"// (\d+)\s+\[([^\]]+)\]\s+([ACGT]+)\s+\*\s+\*\s+\((\d+), (\d+)\)", src)
    idx, mins, seq, s0, s1 = m.groups()
    out.append({"source": "src/to_basespace.rs:203", "node": int(idx),
                "minimizers": [int(x) for x in mins.split(",")], "sequence": seq,
                "shift": [int(s0), int(s1)]})
    src = open(os.path.join(REF, "experiments/661k_genomes/scan_genomes_minmers.py")).read().split("\n")[37]
    m = re.search(r"# (\d+)\s+\[([^\]]+)\]\s+([ACGT]+)\s+([ACGT]+)\s+([ACGT]+)", src)
    idx, mins, a, b, c = m.groups()
    out.append({"source": "experiments/661k_genomes/scan_genomes_minmers.py:38", "node": int(idx),
                "minimizers": [int(x) for x in mins.split(",")], "sequence": a + b + c,
                "parts": [a, b, c]})
    with open(os.path.join(HERE, "reference_sequences_lines.json"), "w") as f:
        json.dump(out, f, indent=1)
    data = gzip.open(os.path.join(REF, "example/reads-0.00.fa.gz"), "rb").read()
    with gzip.GzipFile(os.path.join(HERE, "config1_reads.fa.gz"), "wb", compresslevel=9, mtime=0) as f:
        f.write(data)
    print("wrote", len(out), "golden lines + example reads")


if __name__ == "__main__":
    main()
