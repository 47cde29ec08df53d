mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t16.txt 2>&1
tail -12 gpurun_out/r02_t16.txt
for w in c6 c7 c8; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_spec.json 2> gpurun_out/r02_bench_${w}_spec.err
  tail -2 gpurun_out/r02_bench_${w}_spec.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${w}_spec.json"))
print("$w", "kernel ms", d["ms_per_step"], d["config"].get("kernel_variant"), d["config"].get("autotune_s"), "e2e", d["e2e"]["ms_per_step"], "cells", d["config"]["cells_per_gpu"])
PY
done
