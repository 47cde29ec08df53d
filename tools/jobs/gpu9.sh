mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "constraint_equilibration or specialized or hanford or skeleton or os_step" > gpurun_out/r02_t9.txt 2>&1
tail -25 gpurun_out/r02_t9.txt
for v in "" "--workload c5"; do
PFRX_SPEC_VARIANT=k1 timeout 300 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 $v > gpurun_out/r02_s_v9.json 2> gpurun_out/r02_s_v9.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_s_v9.json"))
print("$v", d["config"]["kernel_variant"], "ms", round(d["ms_per_step"], 3), "cells/s %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 3))
PY
done
