"""Reference facts for the @R[name]{...} on-the-flight resonance overrides (src/MGraniitti.cc:482-566): the resonance
objects of the UNMODIFIED reference after ReadInput (mass, width, decay coupling g_decay computed from the overridden
values, spin density matrix), printed by oracle/_ref/gra/bin/ref_eval info.  Run here: python tests/golden/make_res_override_ref.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import GOLDEN, ref_info  # noqa: E402

PROCESS = "PP[RES+CON]<F> -> pi+ pi- @R[f0_980]{M:1.02,W:0.09} @R[f2_1270]{JZ0:0.2,JZ2:0.8,W:0.2} @R[rho_770]{g0:1.5}"
out = {"card": "./input/test.json", "process": PROCESS, "res": {}}
for ln in ref_info("./input/test.json", PROCESS, 0):
    f = ln.split()
    if f[0] != "res":
        continue
    kv = {k: f[f.index(k) + 1] for k in ("pdg", "mass", "width", "spinX2", "BR", "gdecay")}
    g = f.index("g")
    r = f.index("rho")
    n = int(f[r + 1])
    vals = [float(x) for x in f[r + 2:r + 2 + 2 * n * n]]
    out["res"][f[1]] = dict(pdg=int(kv["pdg"]), mass=float(kv["mass"]), width=float(kv["width"]), spinX2=int(kv["spinX2"]),
                            BR=float(kv["BR"]), gdecay=float(kv["gdecay"]), g=[float(f[g + 1]), float(f[g + 2])],
                            rho_re=vals[0::2], rho_im=vals[1::2])
json.dump(out, open(os.path.join(GOLDEN, "res_override_ref.json"), "w"), indent=1)
print("wrote", len(out["res"]), "resonances")
