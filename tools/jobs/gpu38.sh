mkdir -p gpurun_out
for lw in 2 4; do
PFRX_SPEC_LOCKSTEP_WARPS=$lw python - <<PY
from pflotran_elm_interface_b200 import specialize, workloads
for nm in ("c4s","c4","c4fe"):
    cfg = workloads.by_name(nm, ncell=1).cfg
    p = specialize.build(cfg, warps=1, style="refill", force=True)
PY
for w in c4s c4 c4fe; do
  PFRX_SPEC_LOCKSTEP_WARPS=$lw PFRX_SPEC_VARIANT=q1 timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r02_lw_${w}_${lw}.json 2> gpurun_out/r02_lw_${w}_${lw}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_lw_${w}_${lw}.json"))
    print("$w q1 lockstep_warps=$lw", "kernel ms", round(d["ms_per_step"],3), d["config"]["kernel"])
except Exception as e:
    print("$w $lw ERR", e)
PY
done
done
