set -x
mkdir -p gpurun_out/r3a
B="python bench.py --no-cpu-baseline"
# e2e with / without the packer's software prefetch (same box, back to back, twice)
for i in 1 2; do
MDBG_PACK_PREFETCH=0 $B --steps 10 --warmup 3 --no-extra > gpurun_out/r3a/e2e_pf0_$i.json 2> gpurun_out/r3a/e2e_pf0_$i.err
$B --steps 10 --warmup 3 --no-extra > gpurun_out/r3a/e2e_pf2k_$i.json 2> gpurun_out/r3a/e2e_pf2k_$i.err
done
MDBG_PACK_PREFETCH=4096 $B --steps 10 --warmup 3 --no-extra > gpurun_out/r3a/e2e_pf4k_1.json 2> gpurun_out/r3a/e2e_pf4k_1.err
# K-A vs resident warps per SM: 24 (default), 20, 16, 12
for pad in 0 4096 10240 22000; do
MDBG_BS_SMEM_PAD=$pad $B --no-e2e --no-extra --steps 10 --warmup 3 > gpurun_out/r3a/occ_pad$pad.json 2> gpurun_out/r3a/occ_pad$pad.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3a/*.json")):
    try:
        j=json.load(open(f)); e=j.get("e2e") or {}
        print(f.split("/")[-1], "value %.1f ms %.3f ka_ms %.3f frac %.4f e2e %s h2d %s" % (j["value"], j["ms_per_step"], j["roofline"]["avg_launch_ms"], j["roofline"]["frac"], e.get("value"), (e.get("stage_ms_per_step") or {}).get("h2d")))
    except Exception as ex: print(f, "ERR", ex)
PY
nproc; lscpu | grep -i "model name"
