mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_8gpu_final.json 2> gpurun_out/r02_bench_c3_8gpu_final.err
tail -2 gpurun_out/r02_bench_c3_8gpu_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_8gpu_final.json"))
print("N=8 kernel ms", d["ms_per_step"], "value %.3e" % d["value"], "e2e ms", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"])
print("c5_baseline", d.get("c5_baseline"))
PY
