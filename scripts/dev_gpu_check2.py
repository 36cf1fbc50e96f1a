import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from helpers import *
import graniitti_b200 as g
msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
for name in sys.argv[1:]:
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    card = {"./tests/processes/ALICECUTS.json": "c2_alice.json", "./input/test.json": "c1_test13.json", "./tests/processes/CMSCUTS.json": "c5_cms_4pi.json"}[str(z["card"])]
    E = g.Engine(card=os.path.join(CARDS, card), process=str(z["process"]), pomloop=int(z["pomloop"]))
    if E.desc.screening: E.set_eikonal(msa[0], msa[1], msa[2])
    w, fl, dump = E.eval(z["u"].T.copy(), z["tape"].T.copy(), want_dump=True)
    kin = (z["flags"] & 1) == 1
    print(name, "kin eq", np.array_equal(fl & 1, z["flags"] & 1), "flags eq", np.array_equal(fl[kin], z["flags"][kin]))
    nz = z["W"] != 0
    print(" zero pattern", np.array_equal(w != 0, nz), nz.sum(), (w != 0).sum())
    m = nz & (w != 0)
    rel = np.abs(w[m] - z["W"][m]) / np.abs(z["W"][m])
    print(" rel med %.3g p99 %.3g max %.3g" % (np.median(rel), np.quantile(rel, .99), rel.max()))
    k = np.flatnonzero(m)[:4]
    print(" W dev", w[k], "ref", z["W"][k], "ratio", w[k] / z["W"][k])
    h = z["hamp"][k, 0]
    print(" hamp ref", h, "dev", dump[10][k] + 1j * dump[11][k])
    E.close()
