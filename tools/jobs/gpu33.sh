mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t33.txt 2>&1
tail -4 gpurun_out/r02_t33.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.txt 2>&1; tail -4 gpurun_out/r02_smoke_final.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c3_final2.json 2> gpurun_out/r02_bench_c3_final2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_final2.json"))
print("kernel ms %.3f value %.3e e2e ms %.3f frac %.3f dev_vec %s clocks %s" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["e2e"].get("device_vectors", {}).get("ms_per_step"), d["clocks"]))
print(d["roofline"].get("ncu"))
PY
