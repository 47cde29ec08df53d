mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t14.txt 2>&1
tail -15 gpurun_out/r02_t14.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_v6.json 2> gpurun_out/r02_bench_c3_v6.err
tail -3 gpurun_out/r02_bench_c3_v6.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_v6.json"))
print("kernel ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
PY
