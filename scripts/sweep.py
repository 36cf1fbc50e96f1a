"""Scratch perf/parity sweep over build variants and launch parameters (one subprocess per config)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ONE = r'''
import os, sys, numpy as np, time
sys.path.insert(0, "%s"); sys.path.insert(0, "%s/tests")
import graniitti_b200 as g
from helpers import *
msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
E = g.Engine(card=os.path.join(CARDS, "c2_alice.json")); E.set_eikonal(msa[0], msa[1], msa[2])
u, tape = make_points(40000, 6, 8, 7, adapted=True)
o = oracle_eval(E.desc, Eikonal(msa[0], msa[1], msa[2]), u, tape)
w, fl, _ = E.eval(u.T.copy(), tape.T.copy())
nz = o["w"] != 0
rel = np.abs(w[nz] - o["w"][nz]) / np.abs(o["w"][nz])
ok = np.array_equal(fl, o["flags"]) and np.array_equal(w != 0, nz)
xmat = np.tile(((np.arange(128) + 1) / 128.0)[:, None], (1, 6))
N = 1 << 24
best = 1e9
for r in range(4):
    E.vegas_iterate(xmat, 0, N, 1, r)
    tot, scr, n = E.last_timing_kernels()
    best = min(best, tot)
print("RESULT flags_ok=%%s med=%%.2e p99=%%.2e max=%%.2e frac>1e-10=%%.1e | ms=%%.2f screen_ms=%%.2f pts/s=%%.4g" %% (ok, np.median(rel), np.quantile(rel, .99), rel.max(), np.mean(rel > 1e-10), best, scr, N / best * 1e3))
''' % (ROOT, ROOT)
configs = []
L = os.path.join(ROOT, "graniitti_b200", "libgrb.so")
for minb in (4, 3, 5, 6):
    configs.append({"GRB_LIB": L, "GRB_SCREEN_MINB": str(minb)})
for c in configs:
    env = dict(os.environ); env.update(c)
    r = subprocess.run([sys.executable, "-c", ONE], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print({k: os.path.basename(v) for k, v in c.items()}, line[0] if line else r.stderr[-500:], flush=True)
