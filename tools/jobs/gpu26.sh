mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t26.txt 2>&1
tail -5 gpurun_out/r02_t26.txt
for w in c4s c4se c4 c4fe c8 c6 c7; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_kc.json 2> gpurun_out/r02_bench_${w}_kc.err
  tail -1 gpurun_out/r02_bench_${w}_kc.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_${w}_kc.json"))
    print("$w", "kernel ms", round(d["ms_per_step"],3), d["config"].get("kernel_variant"), "e2e", round(d["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$w ERR", e)
PY
done
