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
