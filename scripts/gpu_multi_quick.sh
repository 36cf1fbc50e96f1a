#!/bin/bash
# short multi-GPU sanity pass (N GPUs of one box): sharded-VEGAS tests + two bench lines.  usage: gpurun --gpus N -- 'bash scripts/gpu_multi_quick.sh N'
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for cfg in "weak_native --comm native" "strong_torch --scaling strong" "c1_weak --workload c1"; do
  set -- $cfg; name=$1; shift
  $TR bench.py --gpus $N --steps 6 --warmup 3 --no-cpu "$@" > gpurun_out/bench_q${N}_$name.json 2> gpurun_out/bench_q${N}_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_q${N}_$name.json').read().strip().splitlines()[-1]); t=d['turnaround_ms_per_step']
    print('$name N=%d %s %s: %.4e pts/s ms/step %.3f (kernels %.3f exchange %.3f) e2e %.3e' % (d['n_gpus'], d['scaling'], d['config']['workload_id'], d['value'], d['ms_per_step'], t['device_kernels'], t['exchange_fetch_sync'], d['e2e']['value']))
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/bench_q${N}_$name.err').read()[-1200:])"
done
