#!/bin/bash
# Quick GPU-box pass: parity tests + bench (own arm only).  Usage: gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag> [pytest -k expr]'
TAG=${1:-q}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  python -m pytest tests -m gpu -q -k "$KEXPR" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
else
  python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
fi
tail -25 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1200 gpurun_out/bench_$TAG.json
python scripts/parity_stats.py > gpurun_out/parity_stats_$TAG.txt 2>&1; tail -32 gpurun_out/parity_stats_$TAG.txt
