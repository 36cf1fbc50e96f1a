"""Device parity on the goldens under tests/golden/late/: reference vectors that were generated after this round's GPU budget
was spent, so the CUDA path has been checked on them only through the host-compiled device header (tests/test_tensor_res.py) and
the oracle (tests/test_oracle_golden.py), never on a GPU by the builder.  The file sorts last on purpose: with `pytest -x` a
failure here cannot hide the result of any other GPU test.  Same protocol as tests/test_gpu_parity.py."""
import glob
import os

import pytest

from helpers import GOLDEN
from test_gpu_parity import run_golden, run_sigma

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN, "late", "c*.npz")))


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_device_vs_reference_golden_late(grb, msa, path):
    run_golden(grb, msa, path)


SIGMA_LATE = [  # name in tests/golden/sigma_ref.json, card, ncall, iters, precision
    ("c5_cms_pipi_restensor_f2_born", "c5_cms_4pi.json", 1 << 21, 5, 0.002),
    ("c5_cms_pipi_rescontensor_born", "c5_cms_4pi.json", 1 << 21, 5, 0.002),
]


@pytest.mark.parametrize("name,card,ncall,iters,prec", SIGMA_LATE, ids=[c[0] for c in SIGMA_LATE])
def test_cross_section_matches_reference_late(grb, msa, name, card, ncall, iters, prec):
    """sigma of the tensor-pomeron resonance processes: full adaptive integration on the device against the unmodified
    reference CLI (tests/golden/sigma_ref.json), within the combined Monte Carlo error."""
    run_sigma(grb, msa, name, card, ncall, iters, prec)
