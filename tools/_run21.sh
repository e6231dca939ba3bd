set -x
O=gpurun_out/r3g; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or upload" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r3g/bench.json")); e=j["e2e"]
print("value %.1f e2e %.1f traffic %s" % (j["value"], e["value"], j["roofline"]["traffic"]))
print(e.get("packed_input"))
PY
