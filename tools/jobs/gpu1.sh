mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "specialized or refill or skeleton or autotune or os_step or hanford" > gpurun_out/r02_t1.txt 2>&1
tail -15 gpurun_out/r02_t1.txt
for v in s1 k1 q1 w1; do PFRX_SPEC_VARIANT=$v timeout 300 python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r02_b_c3_$v.json 2> gpurun_out/r02_b_c3_$v.err; tail -c 600 gpurun_out/r02_b_c3_$v.json; echo; done
PFRX_SPEC_VARIANT=k1 timeout 300 python bench.py --workload c5 --no-e2e --no-cpu --steps 3 --warmup 3 > gpurun_out/r02_b_c5_k1.json 2> gpurun_out/r02_b_c5_k1.err; tail -c 400 gpurun_out/r02_b_c5_k1.json
