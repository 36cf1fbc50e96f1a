"""world_size-2 test (gloo, CPU) of the host-side multi-GPU logic: call-range sharding, the accumulator
all-reduce (sum block + max pair) and the redundant per-rank grid refinement.  The per-rank accumulators are
produced here by the CPU oracle on the device's own Philox point stream (tests/philox_ref.py) -- the product
itself never computes on the CPU."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import CARDS, ROOT

BINS, DIM, NCALLS, SEED = 32, 6, 1501, 4242   # odd call count: exercises the remainder-to-rank-0 rule


def host_accumulate(desc, xmat, first, n, it):
    """What k_points accumulates for calls [first, first+n) of iteration `it` (unscreened)."""
    from helpers import oracle_eval
    from philox_ref import draw, philox_points
    from graniitti_b200 import _abi
    u, vw, tape = philox_points(xmat, first, n, SEED, it, ntape=8)
    o = oracle_eval(desc, None, u, tape)
    f = o["w"] * vw
    acc = np.zeros(ctypes.sizeof(_abi.Accum) // 8)
    acc[0], acc[1], acc[2] = f.sum(), (f * f).sum(), n
    acc[3] = ((o["flags"] & 1) != 0).sum()
    acc[10], acc[11] = o["w"].max(initial=0.0), f.max(initial=0.0)
    fm = acc[12:12 + _abi.GRB_MAX_BINS * _abi.GRB_MAX_DIM].reshape(_abi.GRB_MAX_BINS, _abi.GRB_MAX_DIM)
    f2m = acc[12 + _abi.GRB_MAX_BINS * _abi.GRB_MAX_DIM:].reshape(_abi.GRB_MAX_BINS, _abi.GRB_MAX_DIM)
    for k in range(n):
        if f[k] == 0.0:
            continue
        for j in range(DIM):
            zn = draw(SEED, first + k, it, j) * BINS + 1.0
            b = max(1, min(int(zn), BINS)) - 1
            fm[b, j] += f[k]
            f2m[b, j] += f[k] * f[k]
    return acc


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import graniitti_b200 as g
    from graniitti_b200 import _abi
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = g.parse_card(os.path.join(CARDS, "c2_alice.json"), None, 0)
    xmat = np.tile(((np.arange(BINS) + 1) / BINS)[:, None], (1, DIM))
    grids, sums = [], []
    for it in range(2):
        first, n = g.local_calls(NCALLS, rank, world)
        acc = host_accumulate(desc, xmat, first, n, it)
        if it == 0:   # the bare block: one SUM + one MAX collective
            t = torch.from_numpy(acc)
            g.allreduce_accum(t)
        else:         # the engine's buffer layout with the per-rank maxima slots: ONE collective
            t = torch.from_numpy(np.concatenate([acc, np.full(2 * world, 7.0)]))
            g.allreduce_accum(t, rank, world)
        a = t.numpy()[:len(acc)]
        nb = _abi.GRB_MAX_BINS * _abi.GRB_MAX_DIM
        g.optimize_grid(xmat, a[12 + nb:].reshape(_abi.GRB_MAX_BINS, _abi.GRB_MAX_DIM)[:BINS].copy(), 1.5)
        grids.append(xmat.copy()); sums.append(a[:12].copy())
    q.put((rank, grids, sums))
    dist.destroy_process_group()


def run_world(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=300) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_local_calls_partition(grb):
    for calls, R in ((1501, 2), (10, 3), (7, 8), (1 << 24, 8), (5, 1)):
        spans = [grb.local_calls(calls, r, R) for r in range(R)]
        assert sum(n for _, n in spans) == calls
        pos = 0
        for first, n in spans:          # contiguous, ordered, remainder on rank 0 (MVEGAS.h:59-69)
            assert first == pos
            pos += n
        assert spans[0][1] == calls // R + calls % R


def test_two_ranks_match_one_rank(grb):
    pytest.importorskip("torch")
    two = run_world(2, 29533)
    one = run_world(1, 29534)
    (_, g0, s0), (_, g1, s1) = two
    (_, gref, sref) = one[0]
    for it in range(2):
        assert np.array_equal(g0[it], g1[it]), "ranks ended with different grids"
        assert np.array_equal(s0[it], s1[it])
        # sharded == unsharded up to fp64 summation order; maxima and counters exactly
        assert np.allclose(s0[it][:2], sref[it][:2], rtol=1e-12, atol=0)
        assert s0[it][2] == NCALLS and s0[it][3] == sref[it][3]
        if it == 0:   # identical grid: maxima exactly; later grids differ by the summation-order noise
            assert s0[it][10] == sref[it][10] and s0[it][11] == sref[it][11]
        else:
            assert np.allclose(s0[it][10:12], sref[it][10:12], rtol=1e-12, atol=0)
        assert np.allclose(g0[it], gref[it], rtol=1e-9, atol=1e-12)
    assert not np.array_equal(g0[0], g0[1])
