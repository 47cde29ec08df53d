mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "update_auxvars or constraint_equilibration or zero_immobile or c4_clm or c4s_elm or refill" > gpurun_out/r02_t12.txt 2>&1
tail -12 gpurun_out/r02_t12.txt
timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_v4.json 2> gpurun_out/r02_bench_c3_v4.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_v4.json"))
print("kernel ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
PY
