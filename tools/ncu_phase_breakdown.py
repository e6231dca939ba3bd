#!/usr/bin/env python3
"""Per-phase instruction / stall-sample breakdown of one ka_bitslice_kernel launch from an ncu report.

    ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -c 1 -f -o gpurun_out/bs \\
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e          (on the GPU box)
    python tools/ncu_phase_breakdown.py gpurun_out/bs.ncu-rep [n_tiles]          (anywhere ncu is installed)

Every SASS instruction is attributed to the outermost source line it was inlined into: lines of
ka_bitslice_body.h map to phases by the `// ---- Pn` markers found in the file itself, inlined code that
carries only a ka_bitslice_math.h line maps to the function it belongs to.  Prints warp instructions per
tile, share of the stall samples and the opcode mix per phase (the kernel is ALU-pipe bound: LOP3 / SHF /
ISETP / SEL counts are what to shrink)."""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BODY = os.path.join(ROOT, "rust-mdbg_b200", "csrc", "ka_bitslice_body.h")
MATH = os.path.join(ROOT, "rust-mdbg_b200", "csrc", "ka_bitslice_math.h")


def body_phases():
    marks = []
    for i, line in enumerate(open(BODY), 1):
        m = re.match(r"\s*// ---- (.*?) -*$", line)
        if m:
            marks.append((i, m.group(1).strip()))
        m = re.match(r"(?:template.*)?BS_DEV \w[\w:<> ]* (\w+)\(", line)
        if m:
            marks.append((i, "fn " + m.group(1)))
    return sorted(marks)


def math_functions():
    marks = []
    for i, line in enumerate(open(MATH), 1):
        m = re.match(r"(?:MDBG_HDC?|inline|template.*MDBG_HD) [\w:<> ]*?(\w+)\(", line)
        if m:
            marks.append((i, "math " + m.group(1)))
    return sorted(marks)


def lookup(marks, line):
    name = "?"
    for l, n in marks:
        if line >= l:
            name = n
    return name


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, cur_line, byaddr = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if not r or r[0] == "Line No" or len(r) <= 20:
            continue
        if r[0] != "":
            try:
                cur_line = int(r[0])
            except ValueError:
                cur_line = None
        elif r[2].startswith("0x"):
            byaddr.setdefault(int(r[2], 16), []).append((cur_file, cur_line, r[3].strip(), int(r[6] or 0), int(r[7] or 0)))
    bp, mf = body_phases(), math_functions()
    agg, tot_i, tot_s = {}, 0, 0
    for recs in byaddr.values():
        rec = recs[0]
        body = [x for x in recs if x[0] == "ka_bitslice_body.h" and x[1]]
        math = [x for x in recs if x[0] == "ka_bitslice_math.h" and x[1]]
        if math and lookup(mf, math[-1][1]) not in ("math fsr", "math fsl", "math popc32", "math low_mask"):
            ph = lookup(mf, math[-1][1])
        elif body:
            ph = lookup(bp, body[0][1])
        elif math:
            ph = lookup(mf, math[-1][1])
        else:
            ph = "other (" + (rec[0] or "?") + ")"
        op = (rec[2].split()[1] if rec[2].startswith("@") else rec[2].split()[0]).split(".")[0]
        d = agg.setdefault(ph, [0, 0, {}])
        d[0] += rec[4]; d[1] += rec[3]
        d[2][op] = d[2].get(op, 0) + rec[4]
        tot_i += rec[4]; tot_s += rec[3]
    n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 61236
    print("warp instructions %d (%.0f per tile of 4096 bases, %d tiles), stall samples %d, static %d" %
          (tot_i, tot_i / n_tiles, n_tiles, tot_s, len(byaddr)))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        mix = " ".join("%s:%d" % (o, c / n_tiles) for o, c in sorted(v[2].items(), key=lambda x: -x[1])[:6])
        print("%-44s %6.0f/tile %5.1f%%  samples %5.1f%%  %s" % (k[:44], v[0] / n_tiles, 100 * v[0] / tot_i,
                                                              100 * v[1] / max(1, tot_s), mix))


if __name__ == "__main__":
    main()
