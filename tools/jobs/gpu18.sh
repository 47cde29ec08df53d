mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t19.txt 2>&1
tail -8 gpurun_out/r02_t19.txt
for w in c4fe c4s c4 c8; do
  for ord in 1 0; do
  PFRX_CELL_ORDER=$ord timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_ord${ord}.json 2> gpurun_out/r02_bench_${w}_ord${ord}.err
  tail -2 gpurun_out/r02_bench_${w}_ord${ord}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${w}_ord${ord}.json"))
print("$w order=$ord", "kernel ms", d["ms_per_step"], d["config"].get("kernel_variant"), d["config"].get("autotune_s"), "e2e", d["e2e"]["ms_per_step"])
PY
  done
done
