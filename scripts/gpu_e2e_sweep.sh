#!/bin/bash
# e2e (grb_eval_rng with pinned host buffers) vs points per pipeline slice (GRB_HOST_SLICE), bench.py --workload W
out=gpurun_out/r02_e2e_slice_sweep.txt
echo "# bench.py --workload W --steps 5 (B200): e2e = grb_eval_rng with host buffers, 2^22 points per call; GRB_HOST_SLICE = points per pipeline slice (the first two slices are a quarter and a half of it)" > $out
for w in ${WL:-c2 c1 c5}; do for s in ${SL:-262144 524288 1048576 2097152 4194304}; do
  GRB_HOST_SLICE=$s python bench.py --workload $w --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w slice $s:', '%.4e pts/s  e2e %.4e (with caller tape %.4e)' % (d['value'], d['e2e']['value'], d['e2e']['with_caller_tape']['value']))" | tee -a $out
done; done
