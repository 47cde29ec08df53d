mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
tail -c 2500 gpurun_out/r02_bench_2gpu.json; echo
tail -5 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_2gpu.json"))
print("N=2 kernel ms", d["ms_per_step"], "value %.3e" % d["value"], "e2e ms", d["e2e"]["ms_per_step"], "%.3e" % d["e2e"]["value"])
print("c5_baseline", d["c5_baseline"])
print("cpu", d["cpu_baseline"])
PY
