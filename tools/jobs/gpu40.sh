mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "c3mr or multirate or hanford" 2>&1 | tail -3
timeout 600 python bench.py --workload c3mr --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c3mr_mr2.json 2> gpurun_out/r02_bench_c3mr_mr2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3mr_mr2.json"))
print("c3mr kernel ms %.3f e2e %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
