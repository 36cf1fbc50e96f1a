#!/bin/bash
# Round-2 final pass on one B200: GPU tests, every bench workload (both arms), ncu launch list, full captures of the
# dominant kernels (k_screen fast node, generic k_screen, k_points Regge / RES+CON / HELAS), SASS mixes, parity statistics.
TAG=r02
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_fullsize.txt
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -q --durations=10 ) > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -16 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in c2 c1 c1t c2t c3 c4 c4q c5 c5t; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2>> gpurun_out/bench_$TAG.err
  python bench.py --workload $w --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_$w.json 2>> gpurun_out/bench_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_$w.json").read().strip().splitlines()[-1])
    r = json.loads(open("gpurun_out/${TAG}_bench_reference_$w.json").read().strip().splitlines()[-1])
    rf = d["roofline"]
    print("%-4s %.3e pts/s  e2e %.3e  ref-arm %.3e (%d cores)  e2e/ref %.0f  kernel %s frac %.3f launch %.3f ms  acc %.3f" % ("$w", d["value"], d["e2e"]["value"], r["value"], r["cpu_baseline"]["cores"], d["e2e"]["value"] / r["value"], rf["kernel"], rf["frac"], rf["launch_ms_avg"], d["run"]["accepted_fraction"]))
except Exception as e:
    print("$w FAILED", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
prof() { # key workload kernel-regex [env]
  key=$1; wl=$2; krx=$3; shift 3
  env "$@" ncu --set full --clock-control none --import-source on -k regex:$krx -s 2 -c 1 -f -o /tmp/prof_$key \
      python bench.py --workload $wl --steps 1 --warmup 3 --ncalls 4194304 --no-cpu > gpurun_out/ncu_full_$key.log 2>&1
  ncu -i /tmp/prof_$key.ncu-rep --page raw --csv > gpurun_out/prof_${key}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$key.ncu-rep --page source --csv > gpurun_out/prof_${key}_source.csv 2>/dev/null
  ncu -i /tmp/prof_$key.ncu-rep --page details > gpurun_out/${TAG}_${key}_details.txt 2>/dev/null
  python scripts/ncu_opcode_mix.py gpurun_out/prof_${key}_source.csv > gpurun_out/${TAG}_${key}_opcode_mix.txt 2>/dev/null
  python scripts/ncu_sass_mix.py $wl:${krx} gpurun_out/prof_${key}_source.csv gpurun_out/prof_${key}_raw.csv gpurun_out/ncu_full_$key.log
}
prof k_screen_c2 c2 k_screen GRB_DUMMY=1
prof k_screen_c2t c2t k_screen GRB_DUMMY=1
prof k_points_c2 c2 k_points GRB_DUMMY=1
prof k_points_c1 c1 k_points GRB_DUMMY=1
prof k_points_c4 c4 k_points GRB_DUMMY=1
prof k_points_c5 c5 k_points GRB_DUMMY=1
cp profiles/r02_sass_mix.json gpurun_out/r02_sass_mix.json
GRB_SCREEN_NOFAST=1 ncu --set full --clock-control none --import-source on -k regex:k_screen -s 2 -c 1 -f -o /tmp/prof_generic \
    python bench.py --workload c2 --steps 1 --warmup 3 --ncalls 4194304 --no-cpu > gpurun_out/ncu_full_generic.log 2>&1
ncu -i /tmp/prof_generic.ncu-rep --page details > gpurun_out/${TAG}_k_screen_generic_details.txt 2>/dev/null
ncu -i /tmp/prof_generic.ncu-rep --page source --csv > gpurun_out/prof_generic_source.csv 2>/dev/null
python scripts/ncu_opcode_mix.py gpurun_out/prof_generic_source.csv > gpurun_out/${TAG}_k_screen_generic_opcode_mix.txt 2>/dev/null
python scripts/parity_stats.py > gpurun_out/${TAG}_parity_stats.txt 2>&1; tail -3 gpurun_out/${TAG}_parity_stats.txt
( timeout 300 python scripts/bench_variants.py > gpurun_out/${TAG}_variants.txt ) 2> gpurun_out/variants.err; tail -4 gpurun_out/${TAG}_variants.txt
# compute-sanitizer memcheck over small runs of every kernel family (integration and generation mode)
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
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/mc.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_memcheck.txt

# the second bench pass uses the freshly measured SASS mixes (executed flops, fp64 slot fraction)
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2>> gpurun_out/bench_$TAG.err; tail -c 1800 gpurun_out/${TAG}_bench_n1.json
du -sh gpurun_out
