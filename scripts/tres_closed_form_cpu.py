#!/usr/bin/env python
"""CPU-only, full-size check of the closed-form tensor-pomeron resonance algebra the kernels run (graniitti_b200/csrc/grb_tres.cuh,
compiled by the HOST compiler through tests/host_tres.cc) against the oracle's explicit-tensor restatement -- which reproduces
the unmodified reference bit for bit (tests/test_oracle_golden.py) -- on 2^17 seeded phase-space points per set.
Per point: |A|^2 of the closed form against |A|^2 of the oracle; points beyond 1e-10 get the 1-ulp input conditioning kappa of
the oracle's weight (tests/fullsize.py: oracle_kappa) and are judged against max(1e-10, 8 kappa), like every GPU parity test.
Output: profiles/r02_tres_closed_form_cpu.txt.      python scripts/tres_closed_form_cpu.py [log2 n]"""
import ctypes as C
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import graniitti_b200 as g  # noqa: E402
import helpers  # noqa: E402
import fullsize  # noqa: E402
import test_tensor_res as T  # noqa: E402

SETS = [("f2(1270), g0 only", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f2_1270:1}"),
        ("f2(1270), seven couplings", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f2_1270:1} @R[f2_1270]{g0:0.3,g1:1.0,g2:0.7,g3:-0.6,g4:0.5,g5:0.4,g6:-0.2}"),
        ("f0(980) + f0(1500) + f2(1270)", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f0_980:1,f0_1500:1,f2_1270:1}"),
        ("K+ K-: f0(980) + f2(1525) + f0(1710)", "PP[RESTENSOR]<F> -> K+ K- @RES{f0_980:1,f2_1525:1,f0_1710:1}")]


def _chunk(args):
    proc, u, tape = args
    d = g.parse_card(os.path.join(helpers.CARDS, "c5_cms_4pi.json"), proc, 0)
    o, cap = T.captured(d, u, tape)
    return o["w"], cap


def main():
    n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
    helpers.load_oracle()   # build once, before the workers start
    so = os.path.join(ROOT, "tests", "_host_tres.so")
    H = C.CDLL(so)
    lines = ["# closed form (grb_tres.cuh, host-compiled, double) vs the oracle's explicit tensors (= the reference, bit for bit); %d seeded points per set," % n,
             "# CMS cuts, u0, u1 in [0.85, 1], system mass below 3.1 GeV; |dA^2| / A^2; kappa = 1-ulp input conditioning of the oracle's weight for the points beyond 1e-10",
             "%-40s %8s %8s | %9s %9s %9s | %8s %8s %10s" % ("set", "points", "W != 0", "median", "p99", "max", "> 1e-10", "fail", "max kappa")]
    nproc = os.cpu_count()
    for label, proc in SETS:
        d = g.parse_card(os.path.join(helpers.CARDS, "c5_cms_4pi.json"), proc, 0)
        u, tape = helpers.make_points(n, d.lipsdim, 8, 20261018, adapted=True)
        u[:, 5] *= 0.12
        with mp.Pool(nproc) as pool:
            parts = pool.map(_chunk, [(proc, u[i::nproc], tape[i::nproc]) for i in range(nproc)])
        w = np.zeros(n)
        cap = np.zeros((n, 34))
        for i, (wi, ci) in enumerate(parts):
            w[i::nproc], cap[i::nproc] = wi, ci
        nz = w != 0
        A = T.closed_form(H, d, cap[nz])
        ref = cap[nz, 32] + 1j * cap[nz, 33]
        rel = np.abs(np.abs(A) ** 2 - np.abs(ref) ** 2) / np.abs(ref) ** 2
        above = rel > 1e-10
        kap = np.zeros(rel.shape)
        if above.any():
            idx = np.flatnonzero(nz)[above]
            kap[above] = fullsize.oracle_kappa(d, None, u[idx], tape[idx])
        fail = rel > np.maximum(1e-10, 8 * kap)
        lines.append("%-40s %8d %8d | %9.2e %9.2e %9.2e | %8d %8d %10.2e" % (label, n, nz.sum(), np.median(rel), np.quantile(rel, 0.99), rel.max(),
                                                                         above.sum(), fail.sum(), kap.max()))
        print(lines[-1], flush=True)
    open(os.path.join(ROOT, "profiles", "r02_tres_closed_form_cpu.txt"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
