mkdir -p gpurun_out
PFRX_SPEC_VARIANT=k1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c3_k1_v3 python bench.py --cells 303104 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c3_v3.log 2>&1
tail -2 gpurun_out/r02_ncu_c3_v3.log
cp pflotran_elm_interface_b200/csrc/_spec/spec_d295660e39211017_k1f2.cubin gpurun_out/r02_c3_k1_v3.cubin
cp pflotran_elm_interface_b200/csrc/_spec/spec_d295660e39211017_k1f2.cu gpurun_out/r02_c3_k1_v3.cu
