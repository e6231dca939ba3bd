set -x
mkdir -p gpurun_out/r3e
B="python bench.py --no-cpu-baseline --no-extra"
for i in 1 2; do
MDBG_LIB=$PWD/build/libmdbg_b200_oldpack.so $B --steps 10 --warmup 3 > gpurun_out/r3e/e2e_old_$i.json 2> gpurun_out/r3e/e2e_old_$i.err
$B --steps 10 --warmup 3 > gpurun_out/r3e/e2e_new_$i.json 2> gpurun_out/r3e/e2e_new_$i.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3e/e2e*.json")):
    j=json.load(open(f)); e=j["e2e"]
    print(f.split("/")[-1], "value %.1f ms %.3f e2e %.1f h2d %.2f bytes %d" % (j["value"], j["ms_per_step"], e["value"], e["stage_ms_per_step"]["h2d"], e["h2d_bytes_per_step"]))
j=json.load(open("gpurun_out/r3e/e2e_new_2.json"))
for k in j["kernel_rooflines"]: print("  %-40s %.3f ms" % (k["kernel"], k["avg_ms"]))
print(j["stage_ms_per_step"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
