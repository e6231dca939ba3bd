set -x
mkdir -p gpurun_out/r2e
timeout 600 python tests/bitslice_gpu_check.py --quick > gpurun_out/r2e/bs_check.json 2> gpurun_out/r2e/bs_check.err; echo rc=$?
python bench.py --workload dmel50x --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2e/bench_dmel.json 2> gpurun_out/r2e/bench_dmel.err
ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -s 1 -c 1 -f -o gpurun_out/r2e/ka python bench.py --workload ecoli50x --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2e/ncu.log 2>&1
