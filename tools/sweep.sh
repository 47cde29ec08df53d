mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  +Assertion|^E  +pflo|^FAILED|passed|failed" > gpurun_out/t4.log; tail -4 gpurun_out/t4.log
run() { # workload lanes threads extra
  PFRX_LANES=$2 PFRX_THREADS=$3 python bench.py --workload $1 --steps 2 --warmup 1 --no-e2e --no-cpu $4 > gpurun_out/s.json 2>gpurun_out/s.err || tail -3 gpurun_out/s.err
  python -c "
import json
d=json.load(open('gpurun_out/s.json')); r=d['roofline']
print('$1 L=$2 T=$3', '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['kernel'], 'fp64 %.4f hbm %.4f'%(r['frac_fp64'], r['frac_hbm']))
"
}
for L in 16 8; do for T in 128 64; do run c3 $L $T "--cells 1048576"; done; done
for L in 16 8; do run c4 $L 128; run c4 $L 64; done
run c2 1 128 "--cells 4194304"
run c2 1 64 "--cells 4194304"
run c2 4 128 "--cells 4194304"
