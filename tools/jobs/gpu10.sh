mkdir -p gpurun_out
PFRX_SPEC_VARIANT=w1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c4fe_w1 python bench.py --workload c4fe --cells 75776 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c4fe.log 2>&1
tail -2 gpurun_out/r02_ncu_c4fe.log
for v in w1 q1; do
PFRX_SPEC_VARIANT=$v timeout 300 python bench.py --workload c4fe --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r02_s_c4fe_$v.json 2> gpurun_out/r02_s_c4fe_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_s_c4fe_$v.json"))
print("c4fe $v", d["config"]["kernel"], "ms", round(d["ms_per_step"], 3), "cells/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 3), d["roofline"].get("newton_its_per_cell"))
PY
done
