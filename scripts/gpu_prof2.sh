#!/bin/bash
# One profiling pass on the GPU box: quick parity, a short bench of one workload, the ncu launch list and one full capture
# (source page with per-instruction execution counts) of one kernel.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_prof2.sh <tag> <workload> <kernel-regex> [quick|noparity]'
TAG=${1:-p}; WL=${2:-c2}; KRX=${3:-k_screen}; MODE=${4:-quick}
mkdir -p gpurun_out
if [ "$MODE" = quick ]; then
  python -m pytest tests/test_gpu_parity.py tests/test_fastmath.py -x -q -k "golden or seeded or fast_node or fastmath or folded or exact" 2>&1 | tail -4
fi
python bench.py --workload $WL --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_$WL.json 2> gpurun_out/bench_${TAG}_$WL.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_$WL.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$WL", "%.4e pts/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], "kernel", r["kernel"], "launch_ms %.3f" % r["launch_ms_avg"], "frac %.3f" % r["frac"])
except Exception as e:
    print("$WL FAILED", e); print(open("gpurun_out/bench_${TAG}_$WL.err").read()[-1500:])
PY
ncu --set full --clock-control none --import-source on -k regex:$KRX -s 2 -c 1 -f -o /tmp/prof_$TAG \
    python bench.py --workload $WL --steps 1 --warmup 3 --ncalls 4194304 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>/dev/null
python scripts/ncu_opcode_mix.py gpurun_out/prof_${TAG}_source.csv 2>/dev/null | tail -8
