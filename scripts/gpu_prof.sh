#!/bin/bash
# One ncu --set full capture of the dominant kernel (k_screen) + launch list.  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_prof.sh <tag>'
TAG=${1:-p}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-k_screen} -s ${KSKIP:-2} -c 1 -f -o /tmp/prof_screen_$TAG \
    python bench.py --steps 1 --warmup 3 --ncalls 4194304 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_screen_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page source --csv > gpurun_out/prof_screen_${TAG}_source.csv 2>/dev/null
ncu -i /tmp/prof_screen_$TAG.ncu-rep --page details > gpurun_out/prof_screen_${TAG}_details.txt 2>/dev/null
python scripts/ncu_summary.py gpurun_out/prof_screen_${TAG}_raw.csv | head -40
