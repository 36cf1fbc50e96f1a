"""grb_vegas_optimize_grid (host) against the reference's own VEGASData::OptimizeGrid / Rebin (include/Graniitti/MVEGAS.h:
87-142) on the same grids and <f^2> histograms: bit-equal refined grids.  Goldens: tests/golden/optimize_grid_ref.npz
(make_optimize_grid_ref.py, from the unmodified header through oracle/_ref/gra/bin/ref_vegas)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from graniitti_b200 import _abi

Z = np.load(os.path.join(GOLDEN, "optimize_grid_ref.npz"))
NAMES = sorted(k[:-4] for k in Z.files if k.endswith("_ref"))


@pytest.mark.parametrize("name", NAMES)
def test_optimize_grid_bit_equal_to_reference(grb, name):
    x, f2, lam, ref = Z[name + "_xmat"].copy(), Z[name + "_f2mat"], float(Z[name + "_lambda"]), Z[name + "_ref"]
    bins, dim = x.shape
    f2p = np.zeros((bins, _abi.GRB_MAX_DIM))
    f2p[:, :dim] = f2
    grb.optimize_grid(x, f2p, lam)
    assert np.array_equal(x, ref), np.abs(x - ref).max()
    assert np.all(np.diff(x, axis=0) >= 0) and np.all(x[-1] == 1.0)
