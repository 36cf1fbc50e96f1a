"""Sharded VEGAS under torchrun (one rank per GPU): BASELINE configs[4] (PP[CON]<F> -> rho rho -> 4 pi) or any bench
workload integrated by grb_integrate with the per-iteration exchange over torch.distributed (--comm torch) or the library's
own NCCL communicator (--comm native).  Every rank must end with the same sigma and grid; rank 0 prints one RESULT line.
    torchrun --nproc-per-node N scripts/multi_sigma.py --workload c5 --comm native"""
import argparse
import ctypes
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import graniitti_b200 as g  # noqa: E402
from graniitti_b200 import _abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c5")
ap.add_argument("--comm", default="torch", choices=["torch", "native"])
ap.add_argument("--ncall", type=int, default=1 << 22)
ap.add_argument("--precision", type=float, default=0.002)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
card, _, process, pom = bench.WORKLOADS[a.workload][:4]
E = g.Engine(card=os.path.join(bench.CARDS, card), process=process, pomloop=pom, device=local)
if pom:
    msa = np.load(bench.MSA)
    E.set_eikonal(msa[0], msa[1], msa[2])
reduce = None
if world > 1 and a.comm == "native":
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(g.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    E.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
elif world > 1:
    ext = torch.cuda.ExternalStream(E.stream_ptr(), device=torch.device("cuda", local))
    nacc = _abi.ACCUM_LEN

    def reduce(ptr):
        with torch.cuda.stream(ext):
            g.allreduce_accum(torch.as_tensor(g.DevArray(ptr, nacc + 2 * world), device=torch.device("cuda", local)), rank, world)
        return 0
r = E.integrate(ncall=a.ncall, iters=5, precision=a.precision, seed=20261018, rank=rank, nranks=world, reduce=reduce)
dim = E.desc.lipsdim
grid = bytes(r.xmat)[:128 * dim * 8]
h = hashlib.sha1(grid + ctypes.c_double(r.sigma)).hexdigest()
if world > 1:
    hs = [None] * world
    dist.all_gather_object(hs, h)
    assert len(set(hs)) == 1, "ranks disagree on sigma / grid: " + str(hs)
if rank == 0:
    print(f"RESULT workload {a.workload} world {world} comm {a.comm} sigma {r.sigma:.10e} err {r.sigma_err:.4e} chi2 {r.chi2:.3f} "
          f"iterations {r.iterations} evaluations {r.evaluations:.0f} device_ms {r.device_ms:.2f} seconds {r.seconds:.3f}", flush=True)
E.close()
if world > 1:
    dist.destroy_process_group()
