mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specialized or hanford or skeleton or refill or less_travelled or os_step" > gpurun_out/r02_t7.txt 2>&1
tail -3 gpurun_out/r02_t7.txt
one() {  # label, env..., -- bench args
  label=$1; shift
  env "$@" PFRX_SPEC_VARIANT=${VAR:-k1} timeout 300 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 $BARGS > gpurun_out/r02_s_$label.json 2> gpurun_out/r02_s_$label.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02_s_$label.json"))
    print("$label", d["config"]["kernel_variant"], d["config"]["kernel"], "ms", round(d["ms_per_step"], 3), "cells/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$label failed", e)
PY
}
rebuild() {  # workload, env...
  wl=$1; shift
  env "$@" python -c "
import sys; sys.path.insert(0, '.')
from pflotran_elm_interface_b200 import workloads, specialize
specialize.build(workloads.by_name('$wl', ncell=1).cfg, warps=1, style='lockstep', force=True)"
}
BARGS="" one c3_k1_sec A=1
BARGS="--workload c5" one c5_k1_sec A=1
BARGS="--workload c2 --cells 16777216" one c2_k1_sec A=1
VAR=s1 BARGS="--workload c2 --cells 16777216" one c2_s1_sec A=1
rebuild c3 PFRX_SPEC2_KTAB=0
BARGS="" one c3_k1_sec_noktab PFRX_SPEC2_KTAB=0
for hot in 36 40 44; do
  rebuild c3 PFRX_SPEC2_HOT=$hot
  BARGS="" one c3_k1_sec_hot$hot PFRX_SPEC2_HOT=$hot
done
rebuild c3 A=1
