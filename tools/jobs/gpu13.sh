mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t13.txt 2>&1
tail -12 gpurun_out/r02_t13.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c3_v5.json 2> gpurun_out/r02_bench_c3_v5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c3_v5.json"))
print("kernel ms", d["ms_per_step"], "value %.3e" % d["value"], "frac", d["roofline"]["frac"], d["config"]["kernel_variant"])
print("e2e", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"])
print("e2e_full", d["e2e_full_state"]["ms_per_step"], "parity", d["parity_sample"]["max_rel_err"], "cpu %.3e" % d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/r02_bench_c3_v5.err
