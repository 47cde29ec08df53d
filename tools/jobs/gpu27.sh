mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t27.txt 2>&1
tail -4 gpurun_out/r02_t27.txt
for w in c4s c4fe c4se; do
  timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_hs.json 2> gpurun_out/r02_bench_${w}_hs.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_${w}_hs.json"))
    print("$w", "kernel ms", round(d["ms_per_step"],3), d["config"].get("kernel_variant"))
except Exception as e:
    print("$w ERR", e)
PY
done
