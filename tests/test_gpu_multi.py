"""Multi-GPU tests (need >= 2 visible GPUs; skipped otherwise): the torch-free NCCL route of the C ABI
(grb_comm_unique_id / grb_comm_init / grb_allreduce inside grb_integrate)."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_native_nccl_two_ranks_match_one_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "nccl_native_test.py")], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "NCCL native path OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def _run_multi(n, workload, comm):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "scripts", "multi_sigma.py"), "--workload", workload, "--comm", comm]
    if n == 1:
        cmd = [sys.executable, os.path.join(ROOT, "scripts", "multi_sigma.py"), "--workload", workload, "--comm", comm]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    f = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1].split()
    return float(f[f.index("sigma") + 1]), float(f[f.index("err") + 1])


@pytest.mark.parametrize("comm", ["native", "torch"])
def test_config5_sharded_vegas_under_torchrun(comm):
    """BASELINE configs[4] (rho rho -> 4 pi, 8-GPU sharded VEGAS): grb_integrate on every visible GPU under torchrun, the
    per-iteration exchange through the library's own NCCL communicator / torch.distributed (one collective per iteration);
    all ranks agree bit for bit (checked inside the script), sigma equals the one-rank run within the MC error."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s1, e1 = _run_multi(1, "c5", comm)
    sn, en = _run_multi(n, "c5", comm)
    assert abs(sn - s1) < 4 * (e1 ** 2 + en ** 2) ** 0.5, (s1, e1, sn, en)
