#!/bin/bash
# Round-2 GPU pass A: all GPU tests (incl. the full-size parity sets), the headline bench (both arms), every workload.
TAG=${1:-r02a}
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_fullsize.txt
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -25 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_c2.json 2> gpurun_out/bench_$TAG.err; tail -c 2500 gpurun_out/bench_${TAG}_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}_c2.json 2>> gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_ref_${TAG}_c2.json
for w in c1 c1t c2t c3 c4 c4q c5 c5t; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$w.json 2>> gpurun_out/bench_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_$w.json").read().strip().splitlines()[-1])
    print("$w", "%.3e pts/s" % d["value"], "e2e %.3e" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
except Exception as e:
    print("$w FAILED", e)
PY
done
tail -5 gpurun_out/bench_$TAG.err
