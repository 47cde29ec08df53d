mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c3mr_s1 python bench.py --workload c3mr --cells 151552 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c3mr.log 2>&1
tail -c 400 gpurun_out/r02_ncu_c3mr.log
ls -la gpurun_out/r02_c3mr_s1.ncu-rep
