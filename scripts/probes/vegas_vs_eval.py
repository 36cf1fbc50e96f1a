"""Probe: k_points in VEGAS mode (Philox + grid map + histogram accumulation) vs parity mode (device-resident coordinates,
no accumulation) on the same point distribution.  (Round 2: the gap this probe showed -- 2.2e8 vs 8.8e8 points/s for
yy -> l+ l- -- was the warp staying split after the histogram's compare-and-swap loops; profiles/r02_kp_probe2.txt.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench, graniitti_b200 as g
from philox_ref import philox_points_np
N = 1 << 23
for wl in sys.argv[1:] or ["c4", "c1", "c3", "c5", "c1t"]:
    card, _, process, pom, ntape = bench.WORKLOADS[wl][:5]
    E = g.Engine(card=os.path.join(bench.CARDS, card), process=process, pomloop=pom)
    if E.desc.screening:
        msa = np.load(bench.MSA)
        E.set_eikonal(msa[0], msa[1], msa[2])
    xmat, _ = bench.workload_grid(wl, 128, E.desc.lipsdim)
    for _ in range(2):
        a = E.vegas_iterate(xmat, 0, N, 5, 1)
    ms_v, _ = E.last_timing()
    u, vw, tape = philox_points_np(xmat, 0, 1 << 21, 5, 1, ntape)
    du = torch.tensor(np.ascontiguousarray(u.T), device="cuda")
    dw = torch.empty(du.shape[1], dtype=torch.float64, device="cuda"); df = torch.empty(du.shape[1], dtype=torch.uint8, device="cuda")
    for _ in range(2):
        E.eval_rng_raw(du.shape[1], du, du.shape[0], 5, 1, 0, dw, df)
    ms_e, _ = E.last_timing()
    print(f"{wl}: VEGAS mode {N / ms_v * 1e3:.3e} pts/s ({ms_v:.2f} ms per 2^23), parity mode {du.shape[1] / ms_e * 1e3:.3e} pts/s; accepted {a.n_fid_ok / a.evaluations:.3f} / {float((dw != 0).float().mean()):.3f}")
    E.close()
