"""Per-iteration exchange of the VEGAS accumulators over torch.distributed (NCCL on GPUs, gloo in the CPU
tests): the multi-GPU equivalent of the reference's mutex-guarded accumulation (src/MGraniitti.cc:1389-1419).

Layout of grb_accum (include/grb.h): 10 sum-reduced scalars, then (maxW, maxf) max-reduced, then the
f / f^2 histograms (sum-reduced)."""
from . import _abi


class DevArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def allreduce_accum(t, rank=None, world=None):
    """In-place all-reduce of a float64 tensor holding one grb_accum (any device).

    If the tensor carries the 2 x world tail slots of the engine's accumulator buffer (grb_vegas_iterate_async returns a
    pointer to sizeof(grb_accum) / 8 + 128 doubles) ONE sum all-reduce does everything: every rank parks its two maxima in
    its own slots, zeroes the others, and folds the gathered maxima back afterwards -- the same scheme as the library's
    grb_allreduce.  A bare accumulator block takes one SUM over the block and one MAX over the two maxima."""
    import torch
    import torch.distributed as dist
    lo, hi = _abi.ACCUM_NSUM, _abi.ACCUM_NSUM + 2
    nacc = _abi.ACCUM_LEN
    if rank is not None and world is not None and t.numel() >= nacc + 2 * world:
        tail = t[nacc:nacc + 2 * world]
        tail.zero_()
        tail[2 * rank:2 * rank + 2] = t[lo:hi]
        dist.all_reduce(t[:nacc + 2 * world], op=dist.ReduceOp.SUM)
        t[lo:hi] = torch.amax(tail.view(world, 2), dim=0)
        return t
    mx = t[lo:hi].clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    t[lo:hi] = mx
    return t
