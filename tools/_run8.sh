set -x
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest_gpu.log
tail -4 gpurun_out/r2j/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j/bench_dmel.json 2> gpurun_out/r2j/bench_dmel.err
python bench.py --workload ecoli50x --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j/bench_ecoli.json 2> gpurun_out/r2j/bench_ecoli.err
python tests/cli_ingest_timing.py /tmp/cli > gpurun_out/r2j/cli_ingest_timing.txt 2>&1
