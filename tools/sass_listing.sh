#!/bin/bash
# SASS listings + opcode histograms of the hot kernels (from the built objects) -> profiles/
#   tools/sass_listing.sh r02
set -e
R=${1:-r02}
cd "$(dirname "$0")/.."
K=$(cuobjdump -elf rust-mdbg_b200/csrc/ka_bitslice.o | grep -o "_ZN4mdbg[A-Za-z0-9_]*ka_bitslice_kernelILi12ELb1EEEvNS_6KAArgsE" | head -1)
cuobjdump -sass -fun "$K" rust-mdbg_b200/csrc/ka_bitslice.o | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' > profiles/${R}_ka_bitslice_L12_hpc.sass
{
  echo "# opcode histograms (static SASS, sm_100a) of the hot kernels; made by tools/sass_listing.sh"
  for spec in "ka_bitslice.o:ka_bitslice_kernelILi12ELb1" "graph.o:kb_records_kernel" "graph.o:kc_insert_kernel" "graph.o:kc_verify_kernel" "graph.o:ke_join_kernelILi0" "graph.o:kd_expand_kernel" "ka_minimizers.o:ka_finalize_kernel"; do
    obj=${spec%%:*}; pat=${spec##*:}
    fn=$(cuobjdump -elf rust-mdbg_b200/csrc/$obj | grep -o "_Z[A-Za-z0-9_]*${pat}[A-Za-z0-9_]*" | sort -u | head -1)
    echo; echo "## $fn ($obj)"
    cuobjdump -sass -fun "$fn" rust-mdbg_b200/csrc/$obj | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T] )?//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END {print ""}'
    cuobjdump -sass -fun "$fn" rust-mdbg_b200/csrc/$obj | grep -cE "UTMALDG|UBLKCP|UTC.*MMA|HMMA" | sed 's/^/tensor-core \/ TMA instructions: /'
  done
} > profiles/${R}_sass_opcode_histograms.txt
grep -A3 "ka_bitslice_kernelILi12ELb1" rust-mdbg_b200/csrc/ka_bitslice.o.ptxas.log | head -6 > profiles/${R}_ka_bitslice_ptxas.txt || true
