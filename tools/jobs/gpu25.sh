mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_2gpu_final.json 2> gpurun_out/r02_bench_c3_2gpu_final.err
tail -3 gpurun_out/r02_bench_c3_2gpu_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm_2gpu_final.json 2> gpurun_out/r02_bench_reference_arm_2gpu_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_c3_2gpu_final.json"))
print("N=2 kernel ms", d["ms_per_step"], "value %.3e" % d["value"], "e2e ms", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"])
print("c5_baseline", d.get("c5_baseline"))
r = json.load(open("gpurun_out/r02_bench_reference_arm_2gpu_final.json"))
print("ref arm", r.get("value"), r.get("cpu_baseline", {}).get("cores"))
PY
