set -x
mkdir -p gpurun_out/r3c
B="python bench.py --no-cpu-baseline --no-e2e"
for v in rows rows3 rows3u2 rows rows3; do
MDBG_LIB=$PWD/build/libmdbg_b200_$v.so $B --steps 10 --warmup 3 > gpurun_out/r3c/$v.json 2> gpurun_out/r3c/$v.err
python - <<PY
import json
j=json.load(open("gpurun_out/r3c/$v.json")); x=(j.get("extra") or {}).get("ecoli50x") or {}
print("$v", "value %.1f ms %.3f ka_ms %.3f frac %.4f dirty %s | ecoli %s ka %s" % (j["value"], j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["roofline"]["frac"], j["roofline"]["ka_dirty_tiles"], x.get("value"), x.get("ka_kernel_ms")))
PY
done
MDBG_LIB=$PWD/build/libmdbg_b200_rows3.so python tests/bitslice_gpu_check.py --quick > gpurun_out/r3c/bs_check_rows3.json 2> gpurun_out/r3c/bs_check_rows3.err; echo check rc=$?
