mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "constraint_equilibration or zero_immobile or c4" > gpurun_out/r02_t11.txt 2>&1
tail -5 gpurun_out/r02_t11.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c3_v3.json 2> gpurun_out/r02_bench_c3_v3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_c3_v3.json"))
print("kernel ms", d["ms_per_step"], "value %.3e" % d["value"], "frac", d["roofline"]["frac"], d["config"]["kernel_variant"], d["config"].get("autotune_s"))
print("e2e", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("chunks"))
print("e2e_full", d["e2e_full_state"]["ms_per_step"], "parity", d["parity_sample"], "cpu %.3e" % d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/r02_bench_c3_v3.err
for ch in 4 8 16; do
PFRX_OS_CHUNKS=$ch timeout 600 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_ch$ch.json 2> gpurun_out/r02_bench_c3_ch$ch.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_ch$ch.json"))
print("chunks $ch kernel ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
PY
done
