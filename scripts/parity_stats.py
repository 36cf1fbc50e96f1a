"""Prints device-vs-reference-golden statistics per golden file (run on the GPU box): python scripts/parity_stats.py [substr]"""
import glob, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graniitti_b200 as g
from helpers import CARDS, GOLDEN, card_variant
from test_oracle_golden import CARD_OF
msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
for f in sorted(glob.glob(os.path.join(GOLDEN, "c*.npz"))):
    if len(sys.argv) > 1 and sys.argv[1] not in f:
        continue
    z = np.load(f)
    try:
        nstars = int(z["nstars"]) if "nstars" in z.files else 0          # NSTARS edited in the run card of the golden
        card = card_variant(os.path.join(CARDS, CARD_OF[str(z["card"])]), nstars)
        if "qed_alpha" in z.files and int(z["qed_alpha"]):                # a tune with PARAM_STRUCTURE.QED_alpha != ZERO
            d = g.parse_card(card, str(z["process"]), int(z["pomloop"]))
            d.qed_alpha = int(z["qed_alpha"])
            E = g.Engine(desc=d)
        else:
            E = g.Engine(card=card, process=str(z["process"]), pomloop=int(z["pomloop"]))
    except g.GrbError as e:
        print(os.path.basename(f), "ENGINE:", e); continue
    if E.desc.screening:
        E.set_eikonal(msa[0], msa[1], msa[2])
    w, fl, dump = E.eval(z["u"].T.copy(), z["tape"].T.copy(), want_dump=True)
    nz = z["W"] != 0
    kin = (z["flags"] & 1) == 1
    rel = np.abs(w[nz] - z["W"][nz]) / np.abs(z["W"][nz])
    k = z["kappa"][nz]
    print(f"{os.path.basename(f):44s} n={nz.sum():5d} flags_ok={np.array_equal(fl[kin], z['flags'][kin])} zero_ok={np.array_equal(w != 0, nz)} "
          f"median={np.median(rel):.2e} p99={np.percentile(rel, 99):.2e} max={rel.max():.2e} frac>1e-10={np.mean(rel > 1e-10):.4f} "
          f"frac>max(1e-10,8k)={np.mean(rel > np.maximum(1e-10, 8 * k)):.4f} frac(k>1e-11)={np.mean(k > 1e-11):.4f} max rel/k={np.max(rel / np.maximum(k, 1e-13)):.2f}")
    E.close()
