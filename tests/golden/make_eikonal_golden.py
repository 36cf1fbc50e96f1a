"""Golden eikonal tables from the UNMODIFIED reference compiled in this container (oracle/_ref): the proton
opacity table MBT (as the reference caches it under eikonal/MBT_*, 15 significant digits) and the screening
amplitude MSA, at sqrt(s) = 13 TeV and 7 TeV, plus the reference's printed sigma_tot/el/inel.

Run here (where /root/reference exists):   python tests/golden/make_eikonal_golden.py
msa_pp_13000.npy (full MSA, written by make_golden.py) is the table the parity tests screen with; this script
adds eikonal_ref.npz with every 10th row of the other tables (rows are independent integrals)."""
import glob
import os
import re
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import GOLDEN, REF_GR, REF_TREE, have_ref, ref_msa  # noqa: E402

CASES = {"13000": "./tests/processes/ALICECUTS.json", "7000": "./tests/LHC_TEVATRON_RHIC/ALICE19_2pi_E0.json"}


def main():
    if not have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run oracle/build_ref.sh")
    out = {}
    for e, card in CASES.items():
        hdr, eik = ref_msa(card, "PP[CON]<F> -> pi+ pi-")
        f = glob.glob(os.path.join(REF_TREE, "eikonal", f"MBT_2212_2212_{e}_*"))
        g = glob.glob(os.path.join(REF_TREE, "eikonal", f"MSA_2212_2212_{e}_*"))
        assert len(f) == 1 and len(g) == 1, (f, g)
        mbt = np.loadtxt(f[0], delimiter=",")
        out[f"mbt_{e}"] = mbt[::10].T.copy()
        out[f"msa_{e}"] = np.stack([eik.x, eik.re, eik.im])[:, ::10].copy()
        out[f"mbt_name_{e}"] = os.path.basename(f[0])
        out[f"msa_name_{e}"] = os.path.basename(g[0])
        r = subprocess.run([REF_GR, "-i", card, "-p", "PP[CON]<F> -> pi+ pi-", "-l", "true", "-n", "0", "-h", "0"],
                           cwd=REF_TREE, capture_output=True, text=True, timeout=600)
        xs = [float(re.search(rf"{k} xs:\s+([0-9.]+) mb", r.stdout).group(1)) for k in ("Total", "Elastic", "Inelastic")]
        out[f"xs_mb_{e}"] = np.array(xs)
        print(e, out[f"mbt_name_{e}"], out[f"msa_name_{e}"], xs)
    np.savez_compressed(os.path.join(GOLDEN, "eikonal_ref.npz"), **out)


if __name__ == "__main__":
    main()
