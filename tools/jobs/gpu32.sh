mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "os_step or os_block or resident or multi_step" 2>&1 | tail -3
for one in 0 1; do
  if [ $one -eq 1 ]; then export PFRX_OS_ONE_KERNEL_STREAM=1; else unset PFRX_OS_ONE_KERNEL_STREAM; fi
  for w in c3 c5; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_dual${one}.json 2> gpurun_out/r02_bench_${w}_dual${one}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_${w}_dual${one}.json"))
print("$w one_stream=$one kernel ms %.3f e2e ms %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
  done
done
