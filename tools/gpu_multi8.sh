#!/bin/bash
# Eight-GPU runs behind profiles/r02_bench_n8_* (gpurun --gpus 8): multi_gpu_check, dmel50x shards, human52x_per8 + multi-k sweep
mkdir -p gpurun_out/r2o
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/r2o/mgpu8.log 2>&1; echo rc=$? >> gpurun_out/r2o/mgpu8.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2o/mgpu8.log | tail -5
timeout 400 $T --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2o/bench8_dmel.json 2> gpurun_out/r2o/bench8_dmel.err
tail -1 gpurun_out/r2o/bench8_dmel.err | cut -c1-200
timeout 600 $T --master-port 29543 bench.py --gpus 8 --workload human52x_per8 --steps 3 --warmup 2 --sweep-k 10,15,20,25,30,35,40 > gpurun_out/r2o/bench8_human.json 2> gpurun_out/r2o/bench8_human.err
for f in bench8_dmel bench8_human; do python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2o/$f.json"))
    print("$f", round(j["value"],1), round(j["ms_per_step"],3), j.get("parity_vs_single_gpu"), {k:round(v,3) for k,v in j["stage_ms_per_step"].items()}, "e2e", j["e2e"] and round(j["e2e"]["value"],1))
except Exception as e: print("$f ERR", e)
PY
done
