mkdir -p gpurun_out
for ch in 0 24 32; do
  if [ $ch -eq 0 ]; then unset PFRX_OS_CHUNKS; else export PFRX_OS_CHUNKS=$ch; fi
  timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_ch${ch}b.json 2> gpurun_out/r02_bench_c3_ch${ch}b.err
  tail -1 gpurun_out/r02_bench_c3_ch${ch}b.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_ch${ch}b.json"))
print("chunks=$ch kernel ms %.3f e2e ms %.3f device_vectors %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("device_vectors")))
PY
done
