set -x
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest_gpu.log
tail -4 gpurun_out/r2h/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2h/bench_dmel.json 2> gpurun_out/r2h/bench_dmel.err
python tests/cli_ingest_timing.py /tmp/cli --big > gpurun_out/r2h/cli_ingest_timing.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:ka_bitslice -s 1 -c 1 -f -o gpurun_out/r2h/ka_dmel python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r2h/ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h/smoke.log 2>&1
