"""Golden grids for grb_vegas_optimize_grid: the UNMODIFIED reference's VEGASData::OptimizeGrid (header-only MVEGAS.h,
driven by oracle/_ref/gra/bin/ref_vegas) on four (xmat, f2mat) inputs: a peaked histogram on a uniform grid, a realistic
accumulated histogram on an adapted grid, a near-empty histogram (few populated bins, exact zeros elsewhere: the EPS and
NaN / Inf guards) and a second refinement applied on top of the first.   Run here: python tests/golden/make_optimize_grid_ref.py"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import GOLDEN, REF_TREE  # noqa: E402

BIN = os.path.join(REF_TREE, "bin", "ref_vegas")


def ref_optimize(xmat, f2mat, lam):
    bins, dim = xmat.shape
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "i"), os.path.join(td, "o")
        with open(fin, "wb") as f:
            f.write(struct.pack("<qqd", bins, dim, lam))
            f.write(np.ascontiguousarray(xmat).tobytes())
            f.write(np.ascontiguousarray(f2mat).tobytes())
        subprocess.check_call([BIN, fin, fout])
        return np.fromfile(fout).reshape(bins, dim)


def main():
    rng = np.random.default_rng(11)
    out = {}
    bins, dim = 128, 6
    uni = np.tile(((np.arange(bins) + 1) / bins)[:, None], (1, dim))
    c = (np.arange(bins) + 0.5) / bins
    peaked = np.stack([np.exp(-0.5 * ((c - m) / s) ** 2) * a + 1e-12 for m, s, a in
                       ((0.9, 0.03, 1e-9), (0.95, 0.02, 3e-10), (0.5, 0.3, 1e-11), (0.2, 0.1, 2e-8), (0.5, 0.05, 1e-9), (0.1, 0.2, 1e-13))], axis=1)
    cases = [("peaked_uniform", uni, peaked, 1.5)]
    adapted = np.vstack([np.sort(rng.random((bins - 1, dim)), axis=0), np.ones((1, dim))])
    cases.append(("random_adapted", adapted, rng.random((bins, dim)) ** 4 * 1e-10, 1.5))
    sparse = np.zeros((bins, dim))
    sparse[[5, 6, 90], 0] = [1e-20, 3e-21, 2e-19]
    sparse[64, 1] = 1e-15
    sparse[:, 2] = 1e-40                        # below EPS everywhere
    sparse[10:14, 4] = rng.random(4) * 1e-12    # (columns 3 and 5 stay exactly zero)
    cases.append(("near_empty", uni, sparse, 1.5))
    cases.append(("bins32_lambda1", np.tile(((np.arange(32) + 1) / 32)[:, None], (1, 14)), rng.random((32, 14)) ** 6, 1.0))
    for name, x, f2, lam in cases:
        out[name + "_xmat"], out[name + "_f2mat"], out[name + "_lambda"] = x, f2, lam
        out[name + "_ref"] = ref_optimize(x, f2, lam)
    x2 = ref_optimize(out["peaked_uniform_ref"], peaked * 1.7, 1.5)
    out["second_pass_xmat"], out["second_pass_f2mat"], out["second_pass_lambda"], out["second_pass_ref"] = out["peaked_uniform_ref"], peaked * 1.7, 1.5, x2
    np.savez_compressed(os.path.join(GOLDEN, "optimize_grid_ref.npz"), **out)
    print("wrote", sorted(k for k in out if k.endswith("_ref")))


if __name__ == "__main__":
    main()
