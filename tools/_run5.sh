mkdir -p gpurun_out/r2g
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r2g/mgpu2.log 2>&1; echo rc=$? >> gpurun_out/r2g/mgpu2.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2g/mgpu2.log | tail -6
MDBG_COMM2=1 timeout 300 $T --master-port 29519 tests/multi_gpu_check.py > gpurun_out/r2g/mgpu2_comm2.log 2>&1; echo rc=$? >> gpurun_out/r2g/mgpu2_comm2.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2g/mgpu2_comm2.log | tail -6
timeout 300 $T --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2g/bench2_dmel.json 2> gpurun_out/r2g/bench2_dmel.err
MDBG_COMM2=1 timeout 300 $T --master-port 29520 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2g/bench2_dmel_comm2.json 2> gpurun_out/r2g/bench2_dmel_comm2.err
tail -2 gpurun_out/r2g/bench2_dmel_comm2.err
cut -c1-200 gpurun_out/r2g/bench2_dmel.json; cut -c1-200 gpurun_out/r2g/bench2_dmel_comm2.json
