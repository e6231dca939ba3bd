set -x
mkdir -p gpurun_out/r2q
python bench.py --steps 20 --warmup 5 > gpurun_out/r2q/bench_dmel.json 2> gpurun_out/r2q/bench_dmel.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q/launches_dmel.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2q/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -s 1 -c 1 -f -o gpurun_out/r2q/ka_dmel python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2q/ncu_full.log 2>&1
python tests/bitslice_gpu_check.py --quick > gpurun_out/r2q/bs_check.json 2> gpurun_out/r2q/bs_check.err
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r2q/nvidia_smi.csv
