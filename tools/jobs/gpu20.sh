mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_t20.txt 2>&1
tail -6 gpurun_out/r02_t20.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c3_final2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_c3_final2.log 2>&1
tail -c 300 gpurun_out/r02_launches_c3_final2.log
wc -l gpurun_out/r02_launches_c3_final2.csv
