#!/bin/bash
# k_points / k_screen overlap sweep (bench c2, short): GRB_OVERLAP x GRB_CHUNK
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "vegas or golden or seeded" 2>&1 | tail -3
for ov in 0 3 2; do for ch in 8388608 4194304 2097152; do
  GRB_OVERLAP=$ov GRB_CHUNK=$ch python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_ov${ov}_$ch.json 2> gpurun_out/bench_ov.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_ov${ov}_$ch.json").read().strip().splitlines()[-1])
    print("overlap $ov chunk $ch: %.4e pts/s  ms/step %.3f  screen launch_ms %.3f share %.3f sigma %.5f" % (d["value"], d["ms_per_step"], d["roofline"]["launch_ms_avg"], d["roofline"]["kernel_share_of_step"], d["run"]["sigma_last_iteration_ub"]))
except Exception as e:
    print("FAILED $ov $ch", e); print(open("gpurun_out/bench_ov.err").read()[-800:])
PY
done; done
