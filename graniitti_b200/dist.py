"""Per-iteration exchange of the VEGAS accumulators over torch.distributed (NCCL on GPUs, gloo in the CPU
tests): the multi-GPU equivalent of the reference's mutex-guarded accumulation (src/MGraniitti.cc:1389-1419).

Layout of grb_accum (include/grb.h): 10 sum-reduced scalars, then (maxW, maxf) max-reduced, then the
f / f^2 histograms (sum-reduced)."""
from . import _abi


class DevArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def allreduce_accum(t):
    """In-place all-reduce of a float64 tensor holding one grb_accum (any device): one SUM over the block, one
    MAX over the two maxima."""
    import torch.distributed as dist
    lo, hi = _abi.ACCUM_NSUM, _abi.ACCUM_NSUM + 2
    mx = t[lo:hi].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[lo:hi] = mx
    return t
