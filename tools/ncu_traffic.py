#!/usr/bin/env python3
"""DRAM traffic of one K-A launch from an `ncu --set full` report -> profiles/ka_traffic_<variant>.json.

    ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -s 1 -c 1 -f -o gpurun_out/ka \\
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra     (on the GPU box)
    python tools/ncu_traffic.py gpurun_out/ka.ncu-rep bitslice dmel50x                 (anywhere ncu is installed)

The file records the hash of the kernel sources (tools/ka_hash.py): bench.py drops `roofline.traffic` as
stale as soon as the kernel changes."""
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ka_hash import ROOT, ka_source_hash  # noqa: E402


def main():
    rep, variant, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    vals = rows[2] if len(rows) > 2 else rows[1]       # row 1 holds the units
    units = rows[1]
    get = lambda name: (vals[hdr.index(name)], units[hdr.index(name)])

    def to_bytes(v, u):
        x = float(v.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
    out = {"kernel": get("Kernel Name")[0] if "Kernel Name" in hdr else variant, "ka_variant": 2 if variant == "bitslice" else 1,
           "workload": workload, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "gpu_time_us": get("gpu__time_duration.sum"), "source_hash": ka_source_hash(variant),
           "report": os.path.basename(rep)}
    name = "ka_traffic_bitslice.json" if variant == "bitslice" else "ka_traffic.json"
    with open(os.path.join(ROOT, "profiles", name), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
