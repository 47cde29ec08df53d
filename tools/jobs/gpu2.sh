# full GPU parity suite, default bench (the driver's command), launch list and one full ncu capture of the C3 kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.txt 2>&1
tail -12 gpurun_out/r02_t2.txt
timeout 900 python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
tail -c 3000 gpurun_out/r02_bench_c3.json; echo
tail -5 gpurun_out/r02_bench_c3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
tail -c 800 gpurun_out/r02_bench_ref.json; echo
PFRX_SPEC_VARIANT=k1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfrx_spec_kernel -s 1 -c 1 -o gpurun_out/r02_c3_k1 python bench.py --cells 303104 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02_ncu_c3.log 2>&1
tail -3 gpurun_out/r02_ncu_c3.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_c3.log 2>&1
tail -2 gpurun_out/r02_launches_c3.log
