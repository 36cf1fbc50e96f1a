"""2-rank check of the torch-free NCCL path of the C ABI (grb_comm_unique_id / grb_comm_init / grb_allreduce via
grb_integrate): both ranks must end with the same sigma and grid, equal to a 1-rank run within MC error."""
import os, sys, ctypes as C, multiprocessing as mp
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def worker(rank, world, idfile, q):
    import graniitti_b200 as g
    from helpers import CARDS, GOLDEN
    L = g.lib()
    msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
    E = g.Engine(card=os.path.join(CARDS, "c2_alice.json"), device=rank)
    E.set_eikonal(msa[0], msa[1], msa[2])
    if world > 1:
        buf = C.create_string_buffer(128)
        if rank == 0:
            assert L.grb_comm_unique_id(buf) == 0
            open(idfile + ".tmp", "wb").write(buf.raw); os.rename(idfile + ".tmp", idfile)
        else:
            import time
            while not os.path.exists(idfile): time.sleep(0.05)
            buf = C.create_string_buffer(open(idfile, "rb").read(), 128)
        rc = L.grb_comm_init(E._h, rank, world, buf)
        assert rc == 0, L.grb_last_error(E._h)
    r = E.integrate(ncall=1 << 21, iters=4, precision=0.002, rank=rank, nranks=world, seed=7)
    q.put((rank, r.sigma, r.sigma_err, r.chi2, r.iterations, bytes(r.xmat)[:128 * 6 * 8], r.device_ms))
    E.close()

if __name__ == "__main__":
    mp.set_start_method("spawn")
    out = {}
    for world in (1, 2):
        idfile = f"/tmp/grb_nccl_id_{os.getpid()}_{world}"
        q = mp.Queue()
        ps = [mp.Process(target=worker, args=(r, world, idfile, q)) for r in range(world)]
        [p.start() for p in ps]
        res = sorted(q.get(timeout=600) for _ in ps)
        [p.join() for p in ps]
        out[world] = res
        for r in res: print("world", world, "rank", r[0], "sigma %.6e +- %.2e chi2 %.2f iters %d device_ms %.1f" % (r[1], r[2], r[3], r[4], r[6]))
    a, b = out[2]
    assert a[1] == b[1] and a[5] == b[5], "ranks disagree"
    s1 = out[1][0]
    assert abs(a[1] - s1[1]) < 4 * np.hypot(a[2], s1[2])
    print("NCCL native path OK")
