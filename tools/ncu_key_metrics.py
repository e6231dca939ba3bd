#!/usr/bin/env python3
"""Key raw metrics of every kernel launch in an `ncu --set full` report -> JSON (the numbers profiles/*_SUMMARY.md quotes).

    python tools/ncu_key_metrics.py gpurun_out/x.ncu-rep profiles/rNN_x_ncu_metrics.json "what was captured"
"""
import csv
import json
import re
import subprocess
import sys

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed\.sum\.per_cycle_elapsed|"
                  r"sm__pipe_(alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(lsu|xu)\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__inst_executed\.sum|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__warps_eligible\.avg\.per_cycle_active)$")


def main():
    rep, out, what = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        m = {}
        for i, name in enumerate(hdr):
            if KEEP.match(name) and i < len(vals):
                v = vals[i]
                try:
                    v = float(v.replace(",", ""))
                except ValueError:
                    pass
                m[name] = [v, units[i]] if units[i] else v
        kernels.append(m)
    with open(out, "w") as f:
        json.dump({"report": rep, "what": what, "kernels": kernels}, f, indent=1)
    for m in kernels:
        print(m.get("Kernel Name"), m.get("gpu__time_duration.sum"))


if __name__ == "__main__":
    main()
