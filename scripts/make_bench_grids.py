#!/usr/bin/env python
"""Converged VEGAS grids of the bench workloads (run on a GPU box): the grid after grb_integrate's burn-in + integration
phases, so that the three legs of bench.py (device iterations, end-to-end grb_eval with host buffers, CPU reference arm)
all see points of the SAME distribution -- the one the metric is quoted on.
    gpurun -- 'python scripts/make_bench_grids.py gpurun_out/bench_grids.npz'      then copy to tests/golden/"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import graniitti_b200 as g  # noqa: E402

out = {}
msa = np.load(bench.MSA)
for wl, (card, _, process, pom, *_rest) in bench.WORKLOADS.items():
    E = g.Engine(card=os.path.join(bench.CARDS, card), process=process, pomloop=pom)
    if pom:
        E.set_eikonal(msa[0], msa[1], msa[2])
    r = E.integrate(ncall=1 << 22, iters=5, precision=0.002, seed=12345678)
    dim = E.desc.lipsdim
    out[wl] = np.frombuffer(r.xmat, dtype=np.float64)[:128 * dim].reshape(128, dim).copy()
    print(f"{wl}: sigma = {r.sigma:.6e} +- {r.sigma_err:.2e} b, chi2 {r.chi2:.2f}, {r.iterations} iterations, "
          f"accepted {r.n_fid_ok / r.evaluations:.3f}", flush=True)
    E.close()
np.savez_compressed(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "bench_grids.npz"), **out)
