mkdir -p gpurun_out
PFRX_TPC=1 python -m pytest tests -m gpu -q -k "not lane_variants and not c2_calcite_column" 2>&1 | grep -E "^E  +Assertion|^E  +pflo|^FAILED|passed|failed" > gpurun_out/t6.log; tail -12 gpurun_out/t6.log
run() { # workload extra
  PFRX_TPC=1 PFRX_THREADS=$2 python bench.py --workload $1 --steps 2 --warmup 1 --no-e2e --no-cpu $3 > gpurun_out/s.json 2>gpurun_out/s.err || tail -3 gpurun_out/s.err
  python -c "
import json
d=json.load(open('gpurun_out/s.json')); r=d['roofline']
print('$1 TPC T=$2', '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['kernel'], 'fp64 %.4f hbm %.4f'%(r['frac_fp64'], r['frac_hbm']))
"
}
run c3 64 "--cells 1048576"
run c3 32 "--cells 1048576"
run c4 128
run c4 64
run c2 128 "--cells 4194304"
run c2 64 "--cells 4194304"
