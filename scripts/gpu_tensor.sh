#!/bin/bash
# tensor-pomeron 2xS node check: parity tests that touch the tensor amplitudes + the c2t / c2 bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "TENSOR or tensor or c2t" > gpurun_out/pytest_gpu_tensor.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_tensor.log
for w in c2t c2; do
python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_t_$w.json 2> gpurun_out/bench_t_$w.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_t_$w.json').read().strip().splitlines()[-1]); print('$w: %.4e pts/s ms/step %.3f launch %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_avg'], d['e2e']['value']))"
done
