mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_8gpu_final.json 2> gpurun_out/r02_bench_c3_8gpu_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_4gpu_final.json 2> gpurun_out/r02_bench_c3_4gpu_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_2gpu_final.json 2> gpurun_out/r02_bench_c3_2gpu_final.err
python - <<PY
import json
for n in (2, 4, 8):
    d = json.load(open(f"gpurun_out/r02_bench_c3_{n}gpu_final.json"))
    e = d["e2e"]
    print("N=%d kernel ms %.3f value %.3e e2e ms %.2f link %s floor %s c5 %s" % (n, d["ms_per_step"], d["value"], e["ms_per_step"], (e.get("host_link") or {}).get("GBps_each_direction_per_rank"), e.get("link_floor_ms"), (d.get("c5_baseline") or {}).get("ms_per_step")))
PY
