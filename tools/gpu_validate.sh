#!/bin/bash
# One-GPU validation + profiling pass behind profiles/rNN_* (run on the B200 box from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh gpurun_out/val'
# then, here:  python tools/ncu_traffic.py gpurun_out/val/ka_dmel.ncu-rep bitslice dmel50x
#              python tools/ncu_phase_breakdown.py gpurun_out/val/ka_dmel.ncu-rep <tiles>
#              python tools/ncu_key_metrics.py gpurun_out/val/ka_dmel.ncu-rep profiles/rNN_ka_bitslice_ncu_metrics.json
set -x
O=${1:-gpurun_out/val}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/nvidia_smi.csv
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench_dmel.json 2> $O/bench_dmel.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --workload ecoli50x --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/bench_ecoli.json 2> $O/bench_ecoli.err
# every launch of two steps (shares of the step); numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_dmel.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -s 1 -c 1 -f -o $O/ka_dmel python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"kb_records|kc_insert|kc_verify|ke_join" -s 4 -c 4 -f -o $O/graph_dmel python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_graph.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python tests/bitslice_gpu_check.py --quick > $O/bs_check.json 2> $O/bs_check.err
ls -la $O
