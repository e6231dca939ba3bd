set -x
mkdir -p gpurun_out/r3d
B="python bench.py --no-cpu-baseline --no-e2e"
$B --steps 10 --warmup 3 > gpurun_out/r3d/bench.json 2> gpurun_out/r3d/bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r3d/bench.json")); x=(j.get("extra") or {}).get("ecoli50x") or {}
print("value %.1f ms %.3f ka_ms %.3f frac %.4f | ecoli %s ka %s" % (j["value"], j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["roofline"]["frac"], x.get("value"), x.get("ka_kernel_ms")))
for k in j["kernel_rooflines"]: print("  %-40s %.3f ms" % (k["kernel"], k["avg_ms"]))
print(j["stage_ms_per_step"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
