#!/bin/bash
# Multi-GPU pass (one box, N GPUs): sharded-VEGAS tests, weak / strong scaling, torch / native communicator.
# usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N [tag]'
N=${1:-2}; TAG=${2:-m$N}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N | tr '\n' ';'; echo
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { # name, args...
  name=$1; shift
  $TR bench.py --gpus $N --steps 8 --warmup 3 --no-cpu "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_$name.json").read().strip().splitlines()[-1])
    t = d["turnaround_ms_per_step"]
    print("%-22s N=%d %s %s: %.4e pts/s  ms/step %.3f (kernels %.3f, exchange+fetch %.3f, host refine %.3f)  e2e %.3e  wall/step %.3f" % ("$name", d["n_gpus"], d["scaling"], d["config"]["workload_id"], d["value"], d["ms_per_step"], t["device_kernels"], t["exchange_fetch_sync"], t["host_grid_refinement"], d["e2e"]["value"], d["wall_ms_per_step"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
}
run weak_torch
run weak_native --comm native
run strong_torch --scaling strong
run strong_native --scaling strong --comm native
run c5_weak --workload c5
run c5_strong_native --workload c5 --scaling strong --comm native
python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_${TAG}_n1.json').read().strip().splitlines()[-1]); print('N=1 weak c2: %.4e pts/s ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
