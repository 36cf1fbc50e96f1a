#!/bin/bash
# k_points probes (B200): histogram cost, CTA-synchronised tile rounds, one 256-thread CTA per SM.
# Variant libraries are built by hand (graniitti_b200.build.build_lib(out=..., extra=...)); absent ones are skipped.
out=gpurun_out/r02_kp_probe.txt
: > $out
one() {  # label, workload
  python bench.py --workload $2 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', '%.4e pts/s ms/step %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))" | tee -a $out
}
for w in c4 c1 c3 c5 c1t c4q c2; do one "product" $w; done
for w in c4 c1 c3 c4q c1t; do GRB_KP_SYNC=1 one "product_sync" $w; done
if [ -f graniitti_b200/libgrb_nohist.so ]; then for w in c4 c1 c1t; do GRB_LIB=$PWD/graniitti_b200/libgrb_nohist.so one "nohist" $w; done; fi
if [ -f graniitti_b200/libgrb_f256.so ]; then
  for w in c4 c1 c3 c4q; do GRB_LIB=$PWD/graniitti_b200/libgrb_f256.so one "f256" $w; GRB_KP_SYNC=1 GRB_LIB=$PWD/graniitti_b200/libgrb_f256.so one "f256_sync" $w; done
fi
