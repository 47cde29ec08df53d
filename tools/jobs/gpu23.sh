mkdir -p gpurun_out
for mw in 6 8; do
PFRX_SPEC_MAXWARPS=$mw python - <<PY
from pflotran_elm_interface_b200 import specialize, workloads
for nm in ("c4s","c4","c4fe"):
    cfg = workloads.by_name(nm, ncell=1).cfg
    p = specialize.build(cfg, warps=1, style="refill", force=True)
    import re
    log = open(p[:-6]+".log").read()
    print(nm, re.findall(r"Used \d+ registers.*", log)[-1:], re.findall(r"\d+ bytes spill stores", log)[-1:])
PY
for w in c4s c4 c4fe; do
  PFRX_SPEC_MAXWARPS=$mw PFRX_SPEC_VARIANT=q1 timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r02_occ_${w}_q1_mw${mw}.json 2> gpurun_out/r02_occ_${w}_q1_mw${mw}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_occ_${w}_q1_mw${mw}.json"))
    print("$w q1 maxwarps=$mw", "kernel ms", round(d["ms_per_step"],3), d["config"]["kernel"])
except Exception as e:
    print("$w $mw ERR", e)
PY
done
done
