mkdir -p gpurun_out
PFRX_SPEC_VARIANT=q1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c4fe_q1_final python bench.py --workload c4fe --cells 151552 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c4fe_q1_final.log 2>&1
PFRX_SPEC_VARIANT=q1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c4s_q1_final python bench.py --workload c4s --cells 606208 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c4s_q1_final.log 2>&1
grep -o '"sum_newton_iterations": [0-9]*' gpurun_out/r02_ncu_c4fe_q1_final.log | head -1
grep -o '"sum_newton_iterations": [0-9]*' gpurun_out/r02_ncu_c4s_q1_final.log | head -1
ls -la gpurun_out/*_q1_final.ncu-rep
