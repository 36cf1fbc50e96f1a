"""Full-size parity: for every BASELINE configuration 2^20 points drawn through the converged
VEGAS grid of that configuration and 2^20 flat points, CUDA path (through the C ABI) against the CPU oracle on all host
cores and -- where oracle/_ref travelled to this box -- against the unmodified reference on a subsample.

Bar: integer decisions bit-exact; weights |dW|/W <= max(1e-10, 8 kappa(u)) with kappa the measured 1-ulp conditioning of
that very point (reference's own where available), and the fraction of points that need the kappa term bounded.
The table of all sets is appended to gpurun_out/r02_parity_fullsize.txt (copied to profiles/ after a GPU run).

GRB_FULLSIZE_N overrides the set size (default 2^20)."""
import os

import pytest

import fullsize

pytestmark = pytest.mark.gpu

N = int(os.environ.get("GRB_FULLSIZE_N", 1 << 20))
TABLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_parity_fullsize.txt")


def _log(line):
    os.makedirs(os.path.dirname(TABLE), exist_ok=True)
    new = not os.path.exists(TABLE)
    with open(TABLE, "a") as f:
        if new:
            f.write(fullsize.HEADER + "\n")
        f.write(line + "\n")


@pytest.mark.parametrize("case", fullsize.CASES, ids=fullsize.CASE_IDS)
def test_fullsize_parity(grb, msa, case):
    for s, (allow, med, _floor) in zip(fullsize.run_case(grb, msa, case, n=N, log=_log), case[7]):
        assert s["n"] == N
        assert s["kin_mismatch"] == 0 and s["cut_mismatch"] == 0 and s["zero_mismatch"] == 0, s
        assert s["n_fail"] == 0, s                       # every point inside max(1e-10, 8 kappa) (floor: fullsize.CASES)
        if s["n_valid"]:
            assert s["n_above"] <= allow * s["n_valid"] + 2, s   # how many points may need the conditioning term
            assert s["median"] < med, s
        if s.get("ref_n"):
            assert s["ref_flag_mismatch"] == 0 and s["ref_n_fail"] == 0, s
        if "iter_fsum_rel" in s:
            assert s["iter_fsum_rel"] < 1e-11 and s["iter_maxf_equal"], s
