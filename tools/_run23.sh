set -x
O=gpurun_out/r3i; mkdir -p $O
for mb in 32 64 128 256; do
MDBG_UPLOAD_CHUNK_MB=$mb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_c$mb.json 2> $O/bench_c$mb.err
done
python - <<'PY'
import json
for mb in (32,64,128,256):
    j=json.load(open("gpurun_out/r3i/bench_c%d.json"%mb)); e=j["e2e"]
    print(mb, "e2e %.1f ms %.2f h2d %.2f bytes %d ascii_tiles %s" % (e["value"], e["ms_per_step"], e["stage_ms_per_step"]["h2d"], e["h2d_bytes_per_step"], e["upload"].split(";")[1][:40]))
PY
