"""Throughput of one device VEGAS iteration for every BASELINE configuration and the main variants (secondary numbers
next to bench.py's headline): points/s on a flat grid and on the grid of a short adaptive run, with the reference CLI's
"Integrand sampling frequency" (tests/golden/sigma_ref.json, 8 host cores of the build container) beside it where stored.
    python scripts/bench_variants.py > profiles/rNN_variants.txt        (on a GPU box)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graniitti_b200 as g  # noqa: E402
from helpers import CARDS, GOLDEN  # noqa: E402

CASES = [  # label, card, process, pomloop, ncalls, sigma_ref key
    ("C2 headline  PP[CON]<F> pi+pi- screened", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 1, 1 << 23, "c2_alice_pipi_con_screened"),
    ("C2           PP[CON]<F> pi+pi- born", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 0, 1 << 23, "c2_alice_pipi_con_born"),
    ("C2 variant   PP[CONTENSOR]<F> pi+pi- screened", "c2_alice.json", "PP[CONTENSOR]<F> -> pi+ pi-", 1, 1 << 22, "c2_alice_pipi_tensor_screened"),
    ("C1           PP[RES+CON]<F> pi+pi- born", "c1_test13.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 1 << 23, "c1_test_pipi_rescon_born"),
    ("C1 variant   PP[CONTENSOR]<F> pi+pi- born", "c1_test13.json", "PP[CONTENSOR]<F> -> pi+ pi-", 0, 1 << 23, "c1_test_pipi_tensor_born"),
    ("C3           PP[RES+CON]<F> pi+pi- ATLAS born", "c3_atlas.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 1 << 23, "c3_atlas_pipi_rescon_born"),
    ("C4           yy[CON]<F> mu+mu- born", "c4_cms_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", 0, 1 << 23, "c4_cms_mumu_yycon_born"),
    ("C4 variant   yy[QED]<F> mu+mu- born", "c4_cms_mumu_lowmass.json", "yy[QED]<F> -> mu+ mu-", 0, 1 << 23, None),
    ("C5           PP[CON]<F> rho rho -> 4pi born", "c5_cms_4pi.json", "PP[CON]<F> -> rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}", 0, 1 << 23, "c5_cms_rhorho_4pi_con_born"),
    ("C5 variant   PP[CONTENSOR]<F> rho rho -> 4pi born", "c5_cms_4pi.json", "PP[CONTENSOR]<F> -> rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}", 0, 1 << 22, None),
    ("<C> class    PP[CON]<C> pi+pi- screened", "c2_alice.json", "PP[CON]<C> -> pi+ pi-", 1, 1 << 22, "cc_alice_pipi_con_screened"),
    ("<C> class    PP[CON]<C> pi+pi- born", "c2_alice.json", "PP[CON]<C> -> pi+ pi-", 0, 1 << 23, "cc_alice_pipi_con_born"),
    ("2xF          PP[CONTENSOR]<F> p pbar screened", "c2_alice.json", "PP[CONTENSOR]<F> -> p+ p-", 1, 1 << 21, None),
]


def main():
    msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
    ref = json.load(open(os.path.join(GOLDEN, "sigma_ref.json")))
    print(f"{'configuration':52s} {'points/iter':>11s} {'ms/iter':>9s} {'points/s':>10s} {'accepted':>8s} {'reference CLI points/s (8 cores)':>32s} {'ratio':>8s}")
    for label, card, proc, pom, ncalls, key in CASES:
        E = g.Engine(card=os.path.join(CARDS, card), process=proc, pomloop=pom)
        if E.desc.screening:
            E.set_eikonal(msa[0], msa[1], msa[2])
        r = E.integrate(ncall=1 << 20, iters=3, precision=0.05, seed=1)          # an adapted grid, as in production
        dim = E.desc.lipsdim
        xmat = np.frombuffer(r.xmat, dtype=np.float64)[:128 * dim].reshape(128, dim).copy()
        best = 1e30
        for it in range(4):
            a = E.vegas_iterate(xmat, 0, ncalls, 99, 10 + it)
            ms, _ = E.last_timing()
            best = min(best, ms)
        rate = ncalls / (best * 1e-3)
        rr = ref[key]["ref_points_per_s"] if key and key in ref else None
        print(f"{label:52s} {ncalls:11d} {best:9.2f} {rate:10.3e} {a.n_fid_ok / a.evaluations:8.3f} "
              f"{(('%.3e' % rr) if rr else '-'):>32s} {(('%.0f' % (rate / rr)) if rr else '-'):>8s}")
        E.close()


if __name__ == "__main__":
    main()
