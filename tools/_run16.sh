set -x
mkdir -p gpurun_out/r3b
B="python bench.py --no-cpu-baseline --no-e2e"
python tests/bitslice_gpu_check.py --quick > gpurun_out/r3b/bs_check.json 2> gpurun_out/r3b/bs_check.err; echo check rc=$?
tail -c 600 gpurun_out/r3b/bs_check.json
$B --steps 10 --warmup 3 > gpurun_out/r3b/stream.json 2> gpurun_out/r3b/stream.err
for v in rows mb9 u2mb8; do
MDBG_LIB=$PWD/build/libmdbg_b200_$v.so $B --steps 10 --warmup 3 > gpurun_out/r3b/$v.json 2> gpurun_out/r3b/$v.err
done
MDBG_LIB=$PWD/build/libmdbg_b200_mb9.so python tests/bitslice_gpu_check.py --quick > gpurun_out/r3b/bs_check_mb9.json 2> gpurun_out/r3b/bs_check_mb9.err; echo check rc=$?
for g in 2 4 16; do
MDBG_BS_GROUP=$g $B --no-extra --steps 10 --warmup 3 > gpurun_out/r3b/stream_g$g.json 2> gpurun_out/r3b/stream_g$g.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3b/*.json")):
    if "check" in f: continue
    try:
        j=json.load(open(f)); x=(j.get("extra") or {}).get("ecoli50x") or {}
        print(f.split("/")[-1], "value %.1f ms %.3f ka_ms %.3f frac %.4f dirty %s | ecoli %s ka %s" % (j["value"], j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["roofline"]["frac"], j["roofline"]["ka_dirty_tiles"], x.get("value"), x.get("ka_kernel_ms")))
    except Exception as ex: print(f, "ERR", ex)
PY
