mkdir -p gpurun_out/r2i
for v in "" _u2 _u4; do
MDBG_LIB=rust-mdbg_b200/libmdbg_b200$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i/bench_dmel$v.json 2> gpurun_out/r2i/bench_dmel$v.err
python - <<PY
import json
j=json.load(open("gpurun_out/r2i/bench_dmel$v.json"))
print("$v", j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["avg_launch_ms"], j["extra"]["ecoli50x"]["ka_kernel_ms"])
PY
done
