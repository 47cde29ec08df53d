mkdir -p gpurun_out
for mw in 4 6 8; do
for w in c4s c4 c4fe; do
  for v in q1; do
  PFRX_SPEC_MAXWARPS=$mw PFRX_SPEC_VARIANT=$v timeout 600 python bench.py --workload $w --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/r02_occ_${w}_${v}_mw${mw}.json 2> gpurun_out/r02_occ_${w}_${v}_mw${mw}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_occ_${w}_${v}_mw${mw}.json"))
    print("$w $v maxwarps=$mw", "kernel ms", round(d["ms_per_step"],3), d["config"]["kernel"])
except Exception as e:
    print("$w $v $mw ERR", e)
PY
  done
done
done
