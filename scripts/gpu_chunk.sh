#!/bin/bash
# VEGAS-mode k_points against the work-list chunk size (points per launch): bench.py --workload W with GRB_CHUNK
mkdir -p gpurun_out
for w in c1 c4 c5; do for ch in 8388608 2097152 1048576 262144; do
GRB_CHUNK=$ch python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ch_${w}_$ch.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_ch_${w}_$ch.json').read().strip().splitlines()[-1]); print('$w chunk $ch: %.4e pts/s ms/step %.3f launch %.3f launches %d e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_avg'], d['gpu_launches'], d['e2e']['value']))"
done; done
