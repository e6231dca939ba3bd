mkdir -p gpurun_out/r2k
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for envs in "X=1" "NCCL_MIN_P2P_NCHANNELS=16" "NCCL_MIN_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_P2P_NET_CHUNKSIZE=524288" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_BUFFSIZE=16777216"; do
i=$((i+1))
env $envs MDBG_COMM2=1 timeout 300 $T --master-port $((29530+i)) bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-parity > gpurun_out/r2k/b$i.json 2> gpurun_out/r2k/b$i.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2k/b$i.json"))
print("$envs", round(j["ms_per_step"],3), {k:round(v,3) for k,v in j["stage_ms_per_step"].items()})
PY
done
