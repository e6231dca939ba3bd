#!/bin/bash
# Two-GPU check (gpurun --gpus 2): N-GPU graph == oracle, bench line with parity_vs_single_gpu
set -x
O=gpurun_out/r3j; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29540 tests/multi_gpu_check.py > $O/mgpu2.log 2>&1; echo rc=$? >> $O/mgpu2.log
grep -E "case|rc=|MISMATCH|rror" $O/mgpu2.log | tail -6
timeout 400 $T --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err; tail -2 $O/bench2.err | cut -c1-300
python - <<'PY'
import json
j=json.load(open("gpurun_out/r3j/bench2.json")); e=j.get("e2e") or {}
print("N=2 value %.1f ms %.3f parity %s e2e %s" % (j["value"], j["ms_per_step"], j.get("parity_vs_single_gpu"), e.get("value")))
print({k:round(v,3) for k,v in j["stage_ms_per_step"].items()})
PY
