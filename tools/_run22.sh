set -x
O=gpurun_out/r3h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or limits" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; tail -3 $O/bench.err
MDBG_UPLOAD_CHUNK_MB=8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c8.json 2> $O/bench_c8.err
python - <<'PY'
import json
for f in ("bench","bench_c8"):
    j=json.load(open("gpurun_out/r3h/%s.json"%f)); e=j["e2e"]
    print(f, "value %.1f e2e %.1f" % (j["value"], e["value"]), e["stage_ms_per_step"])
    print("   ", e.get("packed_input"))
PY
