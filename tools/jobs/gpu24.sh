mkdir -p gpurun_out
for w in c4s c4se c4 c4fe c8 c3; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_at2.json 2> gpurun_out/r02_bench_${w}_at2.err
  tail -1 gpurun_out/r02_bench_${w}_at2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_bench_${w}_at2.json"))
    print("$w", "kernel ms", round(d["ms_per_step"],3), d["config"].get("kernel_variant"), {k: round(v*1e3,3) for k,v in (d["config"].get("autotune_s") or {}).items()}, "e2e", round(d["e2e"]["ms_per_step"],2))
except Exception as e:
    print("$w ERR", e)
PY
done
python -m pytest tests -m gpu -q -k "autotune or specialized or refill" 2>&1 | tail -3
