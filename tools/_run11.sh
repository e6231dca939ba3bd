mkdir -p gpurun_out/r2n
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29540 tests/multi_gpu_check.py > gpurun_out/r2n/mgpu4.log 2>&1; echo rc=$? >> gpurun_out/r2n/mgpu4.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2n/mgpu4.log | tail -5
timeout 400 $T --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2n/bench8_dmel.json 2> gpurun_out/r2n/bench8_dmel.err
tail -1 gpurun_out/r2n/bench8_dmel.err | cut -c1-200
MDBG_COMM2=0 timeout 300 $T --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-parity > gpurun_out/r2n/bench8_dmel_nocomm2.json 2> gpurun_out/r2n/bench8_dmel_nocomm2.err
timeout 600 $T --master-port 29543 bench.py --gpus 8 --workload human52x_per8 --steps 3 --warmup 2 --no-e2e --no-parity > gpurun_out/r2n/bench8_human.json 2> gpurun_out/r2n/bench8_human.err
for f in bench8_dmel bench8_dmel_nocomm2 bench8_human; do python - <<PY
import json
try:
    j=json.load(open("gpurun_out/r2n/$f.json"))
    print("$f", round(j["value"],1), round(j["ms_per_step"],3), j.get("parity_vs_single_gpu"), {k:round(v,3) for k,v in j["stage_ms_per_step"].items()})
    print("   kern", [(k["kernel"][:12], round(k["avg_ms"],3)) for k in j["kernel_rooflines"]])
except Exception as e: print("$f ERR", e)
PY
done
