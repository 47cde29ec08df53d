mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_c3_final.json 2> gpurun_out/r02_bench_c3_final.err
tail -2 gpurun_out/r02_bench_c3_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_final.json 2> gpurun_out/r02_bench_reference_arm_final.err
tail -c 600 gpurun_out/r02_bench_reference_arm_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_c3_final.log 2>&1
PFRX_SPEC_VARIANT=k1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c3_k1_final python bench.py --cells 303104 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c3_final.log 2>&1
PFRX_SPEC_VARIANT=w1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c4fe_w1_final python bench.py --workload c4fe --cells 151552 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c4fe_final.log 2>&1
PFRX_SPEC_VARIANT=s1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c8_s1_final python bench.py --workload c8 --cells 303104 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c8_final.log 2>&1
for w in c3mr c4 c4s c4se c4fe c5 c6 c7 c8; do
  timeout 600 python bench.py --workload $w --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_final.json 2> gpurun_out/r02_bench_${w}_final.err
done
timeout 600 python bench.py --workload c2 --cells 16777216 --no-cpu --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_final.json 2> gpurun_out/r02_bench_c2_final.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*_final.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms", round(d.get("ms_per_step", 0), 3), "value %.3e" % d.get("value", 0), d.get("config", {}).get("kernel_variant"), "frac", round((d.get("roofline") or {}).get("frac", 0), 3), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out/*.ncu-rep
