#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: fp64 screened event weights per second.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): continuum
pp -> p pi+ pi- p at sqrt(s) = 13 TeV, `PP[CON]<F> -> pi+ pi-` on ALICE-style fiducial cuts with the eikonal
screening loop on (ND = 1: 16 x 19 = 304 loop nodes per accepted point), MODELPARAM TUNE0.

A "step" is one VEGAS iteration over one batch of synthetic phase-space points per GPU: Philox-generated
points through the VEGAS grid -> kinematics -> cuts -> Born amplitude -> screening loop -> accumulators,
then the per-iteration exchange (all-reduce of the accumulators over NCCL for N > 1) and the host grid
refinement.  `value` = points of all ranks / device time (max over ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ncalls P] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

`--impl reference` times the reference's own CPU implementation of the same path (the unmodified
GRANIITTI sources compiled into oracle/_ref, driven per point through MProcess::EventWeight on all host
cores; falls back to the C port oracle/grb_oracle.c when oracle/_ref is absent).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CARD = os.path.join(ROOT, "tests", "cards", "c2_alice.json")
REF_CARD = "./tests/processes/ALICECUTS.json"
PROCESS = "PP[CON]<F> -> pi+ pi-"
MSA = os.path.join(ROOT, "tests", "golden", "msa_pp_13000.npy")
WORKLOAD = "configs[1]: PP[CON]<F> -> pi+ pi-, 13 TeV, ALICE cuts (|eta|<0.9, pt>0.15), POMLOOP on, ND=1 (304 nodes)"
METRIC = "fp64 screened event weights/sec"
UNIT = "points/s"

# SURVEY.md section 8(d) algorithmic flop model (transcendental = 20 flop, FMA = 2)
F_KIN, F_BORN, F_NODE = 600.0, 650.0, 700.0
# DRAM traffic of k_screen per work-list item, from the committed `ncu --set full` capture
# (profiles/r01_k_screen_summary.txt: dram read 281.3 MB + write 17.1 MB for the 1 879 000 items of one launch
# = 4 194 304 points x accepted fraction 0.448; algorithmic: 22 doubles + index + bins = 188 B per item)
SCREEN_DRAM_BYTES_PER_ITEM = (281.289984e6 + 17.106176e6) / (4194304 * 0.4480448544)


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


# ---- clocks sampling ------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = upper half of the samples
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU reference arm -----------------------------------------------------------------------------------
def cpu_reference_sample(points_per_proc, nproc, seed):
    """One bounded sample of the workload on the host cores through the reference's own EventWeight.
    Returns (points, seconds = max over processes of the in-loop time, kind)."""
    import helpers
    u_t = [helpers.make_points(points_per_proc, 6, 8, seed + 17 * k, adapted=True) for k in range(nproc)]
    if helpers.have_ref():
        td = tempfile.mkdtemp(prefix="grb_ref_")
        procs = []
        for k, (u, tape) in enumerate(u_t):
            fin, fout = os.path.join(td, f"in{k}.bin"), os.path.join(td, f"out{k}.bin")
            with open(fin, "wb") as f:
                f.write(np.array([points_per_proc, 6, 8], dtype=np.int64).tobytes())
                f.write(np.hstack([u, tape]).tobytes())
            procs.append(subprocess.Popen([helpers.REF_BIN, "eval", REF_CARD, PROCESS, "1", fin, fout, "0"],
                                          cwd=helpers.REF_TREE, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                          text=True))
        secs = []
        for p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                raise RuntimeError("ref_eval failed")
            for ln in out.splitlines():
                if ln.startswith("TIME "):
                    secs.append(float(ln.split()[1]))
        return points_per_proc * nproc, max(secs), "reference"
    # port: the C restatement, one process per core
    import multiprocessing as mp
    with mp.Pool(nproc) as pool:
        secs = pool.map(_port_worker, [(u, tape) for (u, tape) in u_t])
    return points_per_proc * nproc, max(secs), "port"


def _port_worker(args):
    import helpers
    import graniitti_b200 as g
    u, tape = args
    msa = np.load(MSA)
    d = g.parse_card(CARD)
    eik = helpers.Eikonal(msa[0], msa[1], msa[2])
    helpers.oracle_eval(d, eik, u[:64], tape[:64])
    t0 = time.perf_counter()
    helpers.oracle_eval(d, eik, u, tape)
    return time.perf_counter() - t0


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    nproc = os.cpu_count() or 1
    ppp = args.ref_points
    for w in range(args.warmup):
        cpu_reference_sample(max(ppp // 8, 64), nproc, 100 + w)
    tot_pts, tot_s, kind = 0, 0.0, "reference"
    for k in range(args.steps):
        n, s, kind = cpu_reference_sample(ppp, nproc, 1000 + k)
        tot_pts += n
        tot_s += s
    val = tot_pts / tot_s
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_step": ppp * nproc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nproc, "kind": kind,
                             "sample": f"{args.steps} x {ppp * nproc} synthetic phase-space points through "
                                       f"MProcess::EventWeight, {nproc} processes"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import graniitti_b200 as g
    from graniitti_b200 import _abi

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    msa = np.load(MSA)
    E = g.Engine(card=CARD, device=local)
    E.set_eikonal(msa[0], msa[1], msa[2])
    ext = torch.cuda.ExternalStream(E.stream_ptr(), device=torch.device("cuda", local))
    bins, dim, ncalls = 128, 6, args.ncalls
    xmat = np.tile(((np.arange(bins) + 1) / bins)[:, None], (1, dim))
    import ctypes
    nacc = ctypes.sizeof(_abi.Accum) // 8
    seed = 12345678
    state = {"xmat": xmat, "acc": None, "launches": 0, "ms_kernels": 0.0, "ms_screen": 0.0, "n_screen": 0,
             "fid": 0.0, "evals": 0.0}

    def step(it, timed):
        ptr = E.vegas_iterate_async(state["xmat"], rank * ncalls, ncalls, seed, it)   # weak scaling: ncalls per GPU
        if world > 1:
            with torch.cuda.stream(ext):
                g.allreduce_accum(torch.as_tensor(g.DevArray(ptr, nacc), device=torch.device("cuda", local)))
        acc = E.vegas_fetch()                      # D2H of the accumulators (33 KB) + stream sync
        f2 = np.frombuffer(acc.f2mat, dtype=np.float64).reshape(_abi.GRB_MAX_BINS, _abi.GRB_MAX_DIM).copy()
        g.optimize_grid(state["xmat"], f2, 1.5)    # identical on every rank
        state["acc"] = acc
        if timed:
            tot, scr, nscr = E.last_timing_kernels()
            _, nl = E.last_timing()
            state["launches"] += nl
            state["ms_kernels"] += tot
            state["ms_screen"] += scr
            state["n_screen"] += nscr
            state["fid"] += acc.n_fid_ok
            state["evals"] += acc.evaluations

    for w in range(args.warmup):
        step(w, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    t0 = time.perf_counter()
    for k in range(args.steps):
        step(args.warmup + k, True)
    ev1.record(ext)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the same metric through the C ABI with HOST buffers (seam A x n, grb_eval) --------------------
    n_e2e = args.e2e_points
    rng = np.random.default_rng(1234 + rank)
    hu = torch.empty((dim, n_e2e), dtype=torch.float64).pin_memory()
    ht = torch.empty((4, n_e2e), dtype=torch.float64).pin_memory()
    hw = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    hf = torch.empty(n_e2e, dtype=torch.uint8).pin_memory()
    hu.numpy()[:] = rng.random((dim, n_e2e))
    hu.numpy()[0:2] = 0.85 + 0.15 * hu.numpy()[0:2]
    ht.numpy()[:] = rng.random((4, n_e2e))
    for _ in range(2):
        E.eval_raw(n_e2e, hu.numpy(), dim, ht.numpy(), 4, hw.numpy(), hf.numpy())
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        E.eval_raw(n_e2e, hu.numpy(), dim, ht.numpy(), 4, hw.numpy(), hf.numpy())   # H2D + kernels + D2H, synchronous
    e2e_s = time.perf_counter() - t0
    e2e_launches = E.last_timing()[1] * e2e_steps

    red = torch.tensor([dev_ms, wall_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_s = [float(x) for x in red.cpu()]

    if rank == 0:
        total_points = float(ncalls) * world * args.steps
        value = total_points / (dev_ms * 1e-3)
        # roofline of the dominant kernel k_screen (fp64 CUDA-core bound; tensor cores unused by design)
        peak = g.measure_dfma_peak(local)
        n_items = state["fid"] / world      # accepted points per rank (counters are all-reduced)
        n_nodes = (E.desc.loop_nkt + 1) * (E.desc.loop_nphi + 1)
        flops_screen = n_items * n_nodes * F_NODE
        ach = flops_screen / (state["ms_screen"] * 1e-3) / 1e12 if state["ms_screen"] > 0 else 0.0
        p_acc = state["fid"] / state["evals"] if state["evals"] else 0.0
        f_point = F_KIN + p_acc * (F_BORN + n_nodes * F_NODE)
        acc = state["acc"]
        sigma = acc.fsum / (ncalls * world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "points_per_step_per_gpu": ncalls, "vegas_bins": bins, "dim": dim,
                           "l2": "inputs are generated on the device per step (Philox counters), nothing is re-read "
                                 "across steps; the per-chunk work list (151 MB) exceeds L2",
                           "timing": "CUDA events on the engine stream around K steps, max over ranks",
                           "accepted_fraction": p_acc, "loop_nodes": n_nodes,
                           "sigma_last_iteration_ub": sigma * 1e6},
                "wall_ms_per_step": wall_ms / args.steps,
                "clocks": clocks,
                "gpu_launches": int(state["launches"]),
                "e2e": {"value": n_e2e * world * e2e_steps / e2e_s, "unit": UNIT,
                        "h2d_bytes_per_step": int((dim + 4) * 8 * n_e2e), "d2h_bytes_per_step": int(9 * n_e2e),
                        "points_per_step_per_gpu": n_e2e, "steps": e2e_steps, "gpu_launches": int(e2e_launches),
                        "api": "grb_eval (seam A x n) with pinned host buffers, H2D+D2H inside the timed region"},
                "roofline": {"bound": "fp64", "kernel": "k_screen", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                             "frac": ach / peak if peak else None,
                             "traffic": SCREEN_DRAM_BYTES_PER_ITEM * n_items / max(state["n_screen"], 1),
                             "traffic_unit": "bytes per launch (ncu dram read+write per work-list item x items per launch; "
                                             "profiles/r01_k_screen_summary.txt)",
                             "peak_source": "measured live: dependent-free DFMA chains on all SMs (grb_measure_dfma_peak); "
                                            "MEASURED_PEAKS.json has no fp64 figure",
                             "flops_per_node": F_NODE, "nodes_per_accepted_point": n_nodes,
                             "flops_per_point_model": f_point,
                             "whole_path_tflops": f_point * value / 1e12,
                             "kernel_share_of_step": state["ms_screen"] / state["ms_kernels"] if state["ms_kernels"] else None,
                             "launch_ms_avg": state["ms_screen"] / max(state["n_screen"], 1)},
                }
        if world == 1 and not args.no_cpu:
            nproc = os.cpu_count() or 1
            n, s, kind = cpu_reference_sample(args.cpu_points, nproc, 77)
            line["cpu_baseline"] = {"value": n / s, "unit": UNIT, "cores": nproc, "kind": kind,
                                    "sample": f"{n} synthetic phase-space points of the same workload through "
                                              f"MProcess::EventWeight, {nproc} processes, in-loop time"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    E.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ncalls", type=int, default=1 << 24, help="VEGAS points per step per GPU")
    ap.add_argument("--e2e-points", type=int, default=1 << 22)
    ap.add_argument("--cpu-points", type=int, default=6000, help="cpu_baseline sample: points per host process")
    ap.add_argument("--ref-points", type=int, default=3000, help="--impl reference: points per host process per step")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
