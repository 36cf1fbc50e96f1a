"""Throughput of one device VEGAS iteration for the amplitudes and phase-space classes that bench.py's workloads do not cover
(secondary numbers next to the bench lines): screened generic loops (2xV, 2xF, resonances, gamma gamma), the <C> class, the
four- and six-leg ladders, forward excitation.  Grid: a short adaptive run where the integration converges, else (ladders:
propagator poles inside the phase space, the reference's own integration fails there) a uniform grid with the intermediate
kt compressed to < 0.75 GeV as in the tests.
    python scripts/bench_variants.py > profiles/rNN_variants.txt        (on a GPU box)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graniitti_b200 as g  # noqa: E402
from helpers import CARDS, GOLDEN, card_variant  # noqa: E402

RHO4PI = "rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}"
CASES = [  # label, card, process, pomloop, ncalls, nstars
    ("<F> Regge 2pi born", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 0, 1 << 23, 0),
    ("<F> Regge 2pi screened (headline kernel)", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 1, 1 << 23, 0),
    ("<F> Regge KK (CON-) screened", "c2_alice.json", "PP[CON-]<F> -> K+ K-", 1, 1 << 22, 0),
    ("<F> tensor 2xS screened", "c2_alice.json", "PP[CONTENSOR]<F> -> pi+ pi-", 1, 1 << 22, 0),
    ("<F> tensor 2xV rho rho (stable) screened", "c5_cms_4pi.json", "PP[CONTENSOR]<F> -> rho(770)0 rho(770)0", 1, 1 << 20, 0),
    ("<F> tensor 2xF p pbar screened", "c2_alice.json", "PP[CONTENSOR]<F> -> p+ p-", 1, 1 << 21, 0),
    ("<F> RES+CON 2pi screened", "c1_test13.json", "PP[RES+CON]<F> -> pi+ pi-", 1, 1 << 20, 0),
    ("<F> yy[CON] mu mu screened", "c4_cms_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", 1, 1 << 20, 0),
    ("<F> Regge 2pi N* (NSTARS 2) born", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 0, 1 << 22, 2),
    ("<F> Regge 2pi N* (NSTARS 2) screened", "c2_alice.json", "PP[CON]<F> -> pi+ pi-", 1, 1 << 21, 2),
    ("<C> Regge 2pi born", "c2_alice.json", "PP[CON]<C> -> pi+ pi-", 0, 1 << 23, 0),
    ("<C> Regge 2pi screened", "c2_alice.json", "PP[CON]<C> -> pi+ pi-", 1, 1 << 22, 0),
    ("<C> tensor 2xS screened", "c2_alice.json", "PP[CONTENSOR]<C> -> pi+ pi-", 1, 1 << 21, 0),
    ("<C> rho rho -> 4 pi born", "c5_cms_4pi.json", "PP[CON]<C> -> " + RHO4PI, 0, 1 << 22, 0),
    ("<C> ladder ME6 4 pi born", "c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi-", 0, 1 << 22, 0),
    ("<C> ladder ME6 4 pi screened", "c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi-", 1, 1 << 19, 0),
    ("<C> ladder ME8 6 pi born", "c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi- pi+ pi-", 0, 1 << 21, 0),
    ("<C> ladder ME8 6 pi screened", "c5_cms_4pi.json", "PP[CON]<C> -> pi+ pi- pi+ pi- pi+ pi-", 1, 1 << 16, 0),
]


def main():
    msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
    only = sys.argv[1:]
    print(f"{'configuration':44s} {'dim':>3s} {'points/iter':>11s} {'ms/iter':>9s} {'points/s':>10s} {'accepted':>8s} {'launches':>8s}  grid")
    for label, card, proc, pom, ncalls, nstars in CASES:
        if only and not any(o in label for o in only):
            continue
        path = os.path.join(CARDS, card)
        if nstars:
            path = card_variant(path, nstars)
        E = g.Engine(card=path, process=proc, pomloop=pom)
        if E.desc.screening:
            E.set_eikonal(msa[0], msa[1], msa[2])
        dim, legs = E.desc.lipsdim, E.desc.n_legs
        if legs >= 4:   # ladders: uniform grid, intermediate kt below 0.75 GeV
            xmat = np.tile(((np.arange(128) + 1) / 128.0)[:, None], (1, dim))
            xmat[:, 4:3 + legs] *= 0.3
            grid = "uniform, kt < 0.75 GeV"
        else:
            r = E.integrate(ncall=min(ncalls, 1 << 20), iters=3, precision=0.05, seed=1)          # an adapted grid, as in production
            xmat = np.frombuffer(r.xmat, dtype=np.float64)[:128 * dim].reshape(128, dim).copy()
            grid = "adapted (3 iterations)"
        best, nl = 1e30, 0
        for it in range(3):
            a = E.vegas_iterate(xmat, 0, ncalls, 99, 10 + it)
            ms, nl = E.last_timing()
            best = min(best, ms)
        rate = ncalls / (best * 1e-3)
        print(f"{label:44s} {dim:3d} {ncalls:11d} {best:9.2f} {rate:10.3e} {a.n_fid_ok / a.evaluations:8.3f} {nl:8d}  {grid}", flush=True)
        E.close()


if __name__ == "__main__":
    main()
