#!/bin/bash
# resonance amplitudes: parity tests that touch them + the c1 / c3 bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "res or RES or c1_ or c3_ or rescon or batched or two_engines" > gpurun_out/pytest_gpu_res.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_res.log
for w in c1 c3; do
python bench.py --workload $w --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_r_$w.json 2> gpurun_out/bench_r_$w.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r_$w.json').read().strip().splitlines()[-1]); print('$w: %.4e pts/s ms/step %.3f launch %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['launch_ms_avg'], d['e2e']['value']))"
done
python scripts/bench_variants.py "RES+CON" 2>/dev/null | tail -1
