#!/bin/bash
# k_points CTA-size sweep (GRB_KP_THREADS) over the unscreened workloads + the headline
mkdir -p gpurun_out
for wl in c4 c3 c1 c5 c1t c5t c4q c2; do for th in 0 128 64 32; do
  GRB_KP_THREADS=$th python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_kp_${wl}_$th.json 2> gpurun_out/bench_kp.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_kp_${wl}_$th.json").read().strip().splitlines()[-1])
    print("$wl threads $th: %.4e pts/s  ms/step %.3f  kernels %.3f" % (d["value"], d["ms_per_step"], d["turnaround_ms_per_step"]["device_kernels"]))
except Exception as e:
    print("FAILED $wl $th", e); print(open("gpurun_out/bench_kp.err").read()[-600:])
PY
done; done
