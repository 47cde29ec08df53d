# thread-per-cell kernel: block size against resident threads per SM (shared memory bound)
mkdir -p gpurun_out
run() { # workload threads
  PFRX_THREADS=$2 python bench.py --workload $1 --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/s.json 2>gpurun_out/s.err || tail -3 gpurun_out/s.err
  python -c "
import json
d=json.load(open('gpurun_out/s.json'))
print('$1 T=$2', '%.3e'%d['value'], 'ms %.2f'%d['ms_per_step'], d['config']['kernel'])
"
}
for w in c8 c7 c6 c3mr; do for t in 128 96 64 32; do run $w $t; done; done
