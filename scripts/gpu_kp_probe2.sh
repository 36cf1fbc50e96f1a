#!/bin/bash
# k_points probes (B200): GRB_KP_SYNC bits -- 1 CTA-wide tile rounds, 2 __syncwarp() per tile, 4 histogram by global RED
out=gpurun_out/r02_kp_probe2.txt
: > $out
one() {  # label, workload
  python bench.py --workload $2 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', '%.4e pts/s ms/step %.3f e2e %.3e sigma %.6g' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['run'].get('sigma', 0)))" | tee -a $out
}
for w in ${WL:-c4 c1 c1t c5 c3}; do for k in ${KS:-0 1 2 4 6 5 3 7}; do GRB_KP_SYNC=$k one "sync=$k" $w; done; done
