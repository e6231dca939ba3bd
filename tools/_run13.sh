set -x
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p/pytest_gpu.log
tail -4 gpurun_out/r2p/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p/bench_dmel.json 2> gpurun_out/r2p/bench_dmel.err
MDBG_PACK_NO_AVX512=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2p/bench_dmel_avx2.json 2> gpurun_out/r2p/bench_dmel_avx2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p/bench_reference.json 2> gpurun_out/r2p/bench_reference.err
