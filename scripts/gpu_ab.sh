#!/bin/bash
# A/B of library builds on one box: GRB_LIB=<variant> python bench.py --workload W.  usage: bash scripts/gpu_ab.sh "<workloads>" "<variants>"
mkdir -p gpurun_out
for w in ${1:-c1}; do for v in ${2:-work}; do
lib=graniitti_b200/libgrb_$v.so; [ $v = work ] && lib=graniitti_b200/libgrb.so
GRB_LIB=$PWD/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_ab_${w}_$v.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_ab_${w}_$v.json').read().strip().splitlines()[-1]); print('$w $v: %.4e pts/s ms/step %.3f launch %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_avg'], d['e2e']['value']))"
done; done
