set -x
mkdir -p gpurun_out/r2f
./tools/ubench/pipes > gpurun_out/r2f/pipes.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest_gpu.log
tail -5 gpurun_out/r2f/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f/bench_dmel.json 2> gpurun_out/r2f/bench_dmel.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f/launches_dmel.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2f/ncu_dmel.log 2>&1
