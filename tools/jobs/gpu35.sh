mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_10k_final.json 2> gpurun_out/r02_bench_c2_10k_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c2_10k_final.json"))
print("c2 10k cells: kernel ms %.4f value %.3e e2e ms %.4f full %.4f variant %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e_full_state"]["ms_per_step"], d["config"]["kernel_variant"]))
PY
