mkdir -p gpurun_out/r2l
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29517 tests/multi_gpu_check.py > gpurun_out/r2l/mgpu2.log 2>&1; echo rc=$? >> gpurun_out/r2l/mgpu2.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2l/mgpu2.log | tail -6
MDBG_P2P=0 timeout 300 $T --master-port 29519 tests/multi_gpu_check.py > gpurun_out/r2l/mgpu2_nccl.log 2>&1; echo rc=$? >> gpurun_out/r2l/mgpu2_nccl.log
grep -E "case|rc=|MISMATCH|rror" gpurun_out/r2l/mgpu2_nccl.log | tail -6
i=0
for envs in "X=1" "MDBG_COMM2=1" "MDBG_P2P=0"; do
i=$((i+1))
env $envs timeout 300 $T --master-port $((29530+i)) bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/r2l/b$i.json 2> gpurun_out/r2l/b$i.err
tail -1 gpurun_out/r2l/b$i.err | cut -c1-200
python - <<PY
import json
j=json.load(open("gpurun_out/r2l/b$i.json"))
print("$envs", round(j["ms_per_step"],3), j.get("parity_vs_single_gpu"), {k:round(v,3) for k,v in j["stage_ms_per_step"].items()})
PY
done
