#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list and one full capture of the dominant kernel.
# Usage (from the repo root):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_screen -s 2 -c 1 -f -o /tmp/prof_screen_$TAG \
    python bench.py --steps 1 --warmup 3 --ncalls 4194304 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
# the report itself can exceed the 64 MiB pull limit: extract the pages on the box, bring the report only if small
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_screen_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page source --csv > gpurun_out/prof_screen_${TAG}_source.csv 2>/dev/null
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page details > gpurun_out/prof_screen_${TAG}_details.txt 2>/dev/null
[ $(stat -c %s /tmp/prof_screen_$TAG.ncu-rep) -lt 30000000 ] && cp /tmp/prof_screen_$TAG.ncu-rep gpurun_out/
du -sh gpurun_out
