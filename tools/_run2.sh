set -x
mkdir -p gpurun_out/r2b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest_gpu.log
tail -5 gpurun_out/r2b/pytest_gpu.log
python bench.py --workload dmel50x --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2b/bench_dmel.json 2> gpurun_out/r2b/bench_dmel.err
python bench.py --workload ecoli50x --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2b/bench_ecoli.json 2> gpurun_out/r2b/bench_ecoli.err
