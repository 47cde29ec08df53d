mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err
tail -2 gpurun_out/r02_bench_c3_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/r02_bench_reference_arm_final.err
for w in c3mr c4 c4s c4se c4fe c5 c6 c7 c8; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_final.json 2> gpurun_out/r02_bench_${w}_final.err
done
timeout 600 python bench.py --workload c2 --cells 16777216 --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_final.json 2> gpurun_out/r02_bench_c2_final.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_final.json")):
    try:
        d = json.load(open(f))
        e = d.get("e2e") or {}
        print(f.split("/")[-1], "ms", round(d.get("ms_per_step", 0), 3), "value %.3e" % d.get("value", 0), d.get("config", {}).get("kernel_variant"), "frac", round((d.get("roofline") or {}).get("frac", 0), 3), "e2e ms", e.get("ms_per_step"), "link", e.get("host_link"), e.get("link_floor_ms"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
