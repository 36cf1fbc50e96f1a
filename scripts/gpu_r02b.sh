#!/bin/bash
# round-2 side measurements: variants table, e2e with the step's own batch size, memcheck of small runs of every kernel family
mkdir -p gpurun_out
( timeout 420 python scripts/bench_variants.py > gpurun_out/r02_variants.txt ) 2> gpurun_out/variants.err; cat gpurun_out/r02_variants.txt; tail -3 gpurun_out/variants.err
( time python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_e2e_default.json 2>/dev/null ) 2>&1 | grep real
python -c "
import json; d=json.loads(open('gpurun_out/bench_e2e_default.json').read().strip().splitlines()[-1]); print('default e2e batch %d: value %.4e e2e %.4e tape %.4e' % (d['e2e']['points_per_step_per_gpu'], d['value'], d['e2e']['value'], d['e2e']['with_caller_tape']['value']))"
cat > /tmp/mc.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import graniitti_b200 as g
from helpers import CARDS, GOLDEN
msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
for card, proc, pom, n in (("c2_alice.json", None, 1, 20000), ("c2_alice.json", "PP[CONTENSOR]<F> -> pi+ pi-", 1, 8000), ("c1_test13.json", None, 0, 20000),
                           ("c4_cms_mumu_lowmass.json", None, 0, 20000), ("c5_cms_4pi.json", None, 0, 20000),
                           ("c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi-", 1, 4000), ("c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi- pi+ pi-", 1, 1500),
                           ("c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi- pi+ pi-", 0, 8000)):
    E = g.Engine(card=os.path.join(CARDS, card), process=proc, pomloop=pom)
    if E.desc.screening: E.set_eikonal(msa[0], msa[1], msa[2])
    dim, legs = E.desc.lipsdim, E.desc.n_legs
    xmat = np.tile(((np.arange(128) + 1) / 128.0)[:, None], (1, dim))
    if legs >= 4: xmat[:, 4:3 + legs] *= 0.3
    a = E.vegas_iterate(xmat, 0, n, 7, 1)
    b = E.vegas_iterate(xmat, 0, n, 7, 1, mode=g._abi.GRB_MODE_GENERATE, maxf=max(a.maxf, 1e-300))
    print(card, proc, "evaluations", a.evaluations, "fid", a.n_fid_ok, "fsum", a.fsum)
    E.close()
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/mc.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/r02_memcheck.txt
