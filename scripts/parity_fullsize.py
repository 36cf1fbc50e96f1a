#!/usr/bin/env python
"""Prints the full-size parity table (tests/fullsize.py) for all BASELINE configurations; run on a GPU box:
    gpurun -- 'python scripts/parity_fullsize.py > gpurun_out/r02_parity_fullsize.txt'"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import fullsize  # noqa: E402
import graniitti_b200 as grb  # noqa: E402
import helpers  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
msa = np.load(os.path.join(helpers.GOLDEN, "msa_pp_13000.npy"))
print(f"# full-size parity, n = {n} per set, host cores = {os.cpu_count()}, reference compiled in: {helpers.have_ref()}")
print(fullsize.HEADER, flush=True)
for case in fullsize.CASES:
    if len(sys.argv) > 2 and case[0] not in sys.argv[2:]:
        continue
    fullsize.run_case(grb, msa, case, n=n, log=lambda s: print(s, flush=True))
