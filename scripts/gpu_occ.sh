#!/bin/bash
# occupancy probe of the fast Regge node: bench with 3 / 4 / 5 CTAs per SM
for m in 3 4 5; do
  GRB_SCREEN_MINB=$m python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('MINB=$m', d['ms_per_step'], d['roofline']['launch_ms_avg'])"
done
