#!/usr/bin/env python
"""bench.py -- benchmark of the hot path: fp64 event weights per second.

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on): continuum
pp -> p pi+ pi- p at sqrt(s) = 13 TeV, `PP[CON]<F> -> pi+ pi-` on ALICE-style fiducial cuts with the eikonal
screening loop on (ND = 1: 19 kt x 15 folded phi nodes = 285 executed loop nodes per accepted point), MODELPARAM TUNE0.
`--workload c1|c1t|c2|c2t|c3|c4|c4q|c5|c5t` selects the other BASELINE configurations (`c1r`, `c2r`: the tensor-pomeron resonance processes) (same JSON line, per-workload roofline).

A "step" is one VEGAS iteration over one batch of synthetic phase-space points: Philox-generated points through the VEGAS
grid -> kinematics -> cuts -> Born amplitude -> (screening loop) -> accumulators, then the per-iteration exchange
(all-reduce of the accumulators over NCCL for N > 1) and the grid refinement.  `value` = points of all ranks / device
time (max over ranks).  `--scaling weak` (default): --ncalls points per GPU; `--scaling strong`: --ncalls points in total,
split over the ranks like the reference splits one iteration over its threads (MVEGAS.h:59-69).  `--comm torch` runs the
exchange through torch.distributed (NCCL), `--comm native` through the library's own communicator (grb_comm_init /
grb_allreduce, what a C++ host uses).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--ncalls P] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU)

`--impl reference` times the reference's own CPU implementation of the same path (the unmodified GRANIITTI sources
compiled into oracle/_ref, driven per point through MProcess::EventWeight on all host cores; falls back to the C port
oracle/grb_oracle.c when oracle/_ref is absent).
"""
import argparse
import ctypes
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

CARDS = os.path.join(ROOT, "tests", "cards")
MSA = os.path.join(ROOT, "tests", "golden", "msa_pp_13000.npy")
UNIT = "points/s"
RHO4PI = "rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}"

# name -> card of this repo, the reference's card, process, POMLOOP, tape length of the e2e / CPU legs, CPU sample sizes
# (points per host process: cpu_baseline, --impl reference per step), description
WORKLOADS = {
    "c1": ("c1_test13.json", "./input/test.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 4, 150000, 100000,
           "configs[0]: input/test.json as shipped, PP[RES+CON]<F> -> pi+ pi-, 13 TeV, screening off"),
    "c1t": ("c1_test13.json", "./input/test.json", "PP[CONTENSOR]<F> -> pi+ pi-", 0, 4, 300000, 200000,
            "configs[0] as described: PP[CONTENSOR]<F> -> pi+ pi- (tensor pomeron continuum), 13 TeV, screening off"),
    "c2": ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 1, 4, 40000, 20000,
           "configs[1]: PP[CON]<F> -> pi+ pi-, 13 TeV, ALICE cuts (|eta|<0.9, pt>0.15), POMLOOP on, ND=1"),
    "c2t": ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CONTENSOR]<F> -> pi+ pi-", 1, 4, 4000, 2000,
            "configs[1] variant: PP[CONTENSOR]<F> -> pi+ pi-, 13 TeV, ALICE cuts, POMLOOP on, ND=1"),
    "c3": ("c3_atlas.json", "./tests/run_atlas_multi/ATLAS.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 4, 200000, 150000,
           "configs[2]: ATLAS 13 TeV PP[RES+CON]<F> -> pi+ pi- (8 resonances + continuum), screening off"),
    "c4": ("c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", 0, 4,
           300000, 200000, "configs[3]: yy[CON]<F> -> mu+ mu- (kt-EPA flux x MG5 HELAS), 13 TeV, CMS cuts, screening off"),
    "c4q": ("c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[QED]<F> -> mu+ mu-", 0, 4,
            40000, 20000, "configs[3] variant: yy[QED]<F> -> mu+ mu- (tensor-pomeron code, QED branch), screening off"),
    "c5": ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CON]<F> -> " + RHO4PI, 0, 16, 300000, 200000,
           "configs[4]: PP[CON]<F> -> rho rho -> 4 pi (decay chains), 13 TeV, CMS cuts, screening off"),
    "c5t": ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CONTENSOR]<F> -> " + RHO4PI, 0, 16, 60000, 40000,
            "configs[4] variant: PP[CONTENSOR]<F> -> rho rho -> 4 pi (tensor pomeron 2xV), screening off"),
    # tensor-pomeron resonances (MTensorPomeron::ME3; no converged grid committed yet: uniform grid, see workload_grid)
    "c1r": ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f2_1270:1}", 0, 4, 4000, 2000,
            "tensor-pomeron resonance: PP[RESTENSOR]<F> -> pi+ pi- via f2(1270), 13 TeV, CMS cuts, screening off"),
    "c2r": ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[RES+CONTENSOR]<F> -> pi+ pi- @RES{f0_980:1,f2_1270:1}", 1, 4, 16, 8,
            "tensor-pomeron resonances + continuum: PP[RES+CONTENSOR]<F> -> pi+ pi- (f0(980), f2(1270)), ALICE cuts, POMLOOP on, ND=1"),
}

# Algorithmic flop model, SURVEY.md section 8(d) convention (add/sub/mul/div/sqrt/compare = 1, FMA = 2, transcendental
# = 20): per point F_kin (always), F_born (accepted points), F_node (per executed loop node).  profiles/r02_opcount.json
# holds the op counts measured on the oracle restatement (scripts/oracle_opcount.py); the survey's estimates are the
# fallback and are labelled as such.
SURVEY_FLOPS = {"c1": (600, 24000, 0), "c1t": (600, 900, 0), "c2": (600, 650, 700), "c2t": (600, 900, 900), "c3": (600, 24000, 0),
                "c4": (600, 8000, 0), "c4q": (600, 8000, 0), "c5": (1800, 650, 0), "c5t": (1800, 50000, 0),
                "c1r": (1300, 800, 0), "c2r": (1300, 1700, 1700)}   # (closed-form counts of grb_tres.cuh)


def flop_model(name):
    p = os.path.join(ROOT, "profiles", "r02_opcount.json")
    if os.path.exists(p):
        t = json.load(open(p)).get(name)
        if t and not t.get("amplitude_as_written_only"):
            return float(t["F_kin"]), float(t["F_born"]), float(t.get("F_node", 0.0)), "op count of the oracle restatement (profiles/r02_opcount.json)"
        if t:   # kinematics counted; the amplitude is restated as explicit FTensor loops: closed-form estimate of the survey
            k, b, n = SURVEY_FLOPS[name]
            return float(t["F_kin"]), float(b), float(n), ("F_kin: op count of the oracle restatement; amplitude: survey's closed-form estimate "
                                                          "(the oracle restates the reference's explicit rank-4 tensor loops, %.0f flop as written)" % t["F_born"])
    k, b, n = SURVEY_FLOPS[name]
    return float(k), float(b), float(n), "survey estimate (SURVEY.md section 8d)"


def sass_mix(kernel_key):
    """Executed fp64 instructions per loop node / per point of a kernel from the committed ncu opcode mix."""
    p = os.path.join(ROOT, "profiles", "r02_sass_mix.json")
    if os.path.exists(p):
        return json.load(open(p)).get(kernel_key)
    return None


GRIDS = os.path.join(ROOT, "tests", "golden", "bench_grids.npz")


def workload_grid(wl, bins=128, dim=6):
    """The converged VEGAS grid of the workload (scripts/make_bench_grids.py), so that every leg of the bench -- device
    iterations, end-to-end call, CPU reference arm -- sees points of the distribution the metric is quoted on."""
    if os.path.exists(GRIDS):
        z = np.load(GRIDS)
        if wl in z.files and z[wl].shape == (bins, dim):
            return z[wl].copy(), "converged VEGAS grid of the workload (tests/golden/bench_grids.npz)"
    return np.tile(((np.arange(bins) + 1) / bins)[:, None], (1, dim)), "uniform grid (no converged grid committed)"


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
def synthetic_points(wl, n, dim, dtape, seed, first=0):
    """Synthetic phase-space points of the workload: the Philox stream of the device path through the workload's converged
    VEGAS grid (tests/philox_ref.py restates the device generator on the host)."""
    from philox_ref import philox_points_np
    xmat, _ = workload_grid(wl, 128, dim)
    u, _, tape = philox_points_np(xmat, first, n, seed, 0, dtape)
    return u, tape


def cpu_reference_sample(wl, points_per_proc, nproc, seed):
    """One bounded sample of the workload on the host cores through the reference's own EventWeight.
    Returns (points, seconds = max over processes of the in-loop time, kind)."""
    import helpers
    card, ref_card, process, pom, ntape = WORKLOADS[wl][:5]
    u_t = [synthetic_points(wl, points_per_proc, 6, ntape, seed, first=k * points_per_proc) for k in range(nproc)]
    if helpers.have_ref():
        td = tempfile.mkdtemp(prefix="grb_ref_")
        procs = []
        for k, (u, tape) in enumerate(u_t):
            fin, fout = os.path.join(td, f"in{k}.bin"), os.path.join(td, f"out{k}.bin")
            with open(fin, "wb") as f:
                f.write(np.array([points_per_proc, 6, ntape], dtype=np.int64).tobytes())
                f.write(np.hstack([u, tape]).tobytes())
            procs.append(subprocess.Popen([helpers.REF_BIN, "eval", ref_card, process, str(pom), fin, fout, "0"],
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
        for f in os.listdir(td):
            os.unlink(os.path.join(td, f))
        os.rmdir(td)
        return points_per_proc * nproc, max(secs), "reference"
    # port: the C restatement, one process per core
    import multiprocessing as mp
    with mp.Pool(nproc) as pool:
        secs = pool.map(_port_worker, [(wl, u, tape) for (u, tape) in u_t])
    return points_per_proc * nproc, max(secs), "port"


def _port_worker(args):
    import helpers
    import graniitti_b200 as g
    wl, u, tape = args
    card, _, process, pom = WORKLOADS[wl][:4]
    msa = np.load(MSA)
    d = g.parse_card(os.path.join(CARDS, card), process, pom)
    eik = helpers.Eikonal(msa[0], msa[1], msa[2]) if pom else None
    helpers.oracle_eval(d, eik, u[:64], tape[:64])
    t0 = time.perf_counter()
    helpers.oracle_eval(d, eik, u, tape)
    return time.perf_counter() - t0


def metric_name(wl):
    return "fp64 screened event weights/sec" if WORKLOADS[wl][3] else "fp64 event weights/sec"


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    wl = args.workload
    nproc = os.cpu_count() or 1
    ppp = args.ref_points or WORKLOADS[wl][6]
    for w in range(args.warmup):
        cpu_reference_sample(wl, max(ppp // 8, 64), nproc, 100 + w)
    tot_pts, tot_s, kind = 0, 0.0, "reference"
    for k in range(args.steps):
        n, s, kind = cpu_reference_sample(wl, ppp, nproc, 1000 + k)
        tot_pts += n
        tot_s += s
    val = tot_pts / tot_s
    line = {"impl": "reference", "metric": metric_name(wl), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[wl][7], "workload_id": wl},
            "run": {"points_per_step": ppp * nproc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": nproc, "kind": kind,
                             "sample": f"{args.steps} x {ppp * nproc} synthetic phase-space points through "
                                       f"MProcess::EventWeight, {nproc} processes, in-loop time"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---- our arm ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import graniitti_b200 as g
    from graniitti_b200 import _abi

    wl = args.workload
    card, ref_card, process, pom, ntape, cpu_pts, _, descr = WORKLOADS[wl]
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E = g.Engine(card=os.path.join(CARDS, card), process=process, pomloop=pom, device=local)
    if E.desc.screening:
        msa = np.load(MSA)
        E.set_eikonal(msa[0], msa[1], msa[2])
    if world > 1 and args.comm == "native":
        # the library's own communicator: rank 0 makes the id, the bytes travel through the process group
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(g.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        E.comm_init(rank, world, bytes(idt.cpu().numpy().tobytes()))
    ext = torch.cuda.ExternalStream(E.stream_ptr(), device=torch.device("cuda", local))
    bins, dim = 128, E.desc.lipsdim
    if args.scaling == "strong":
        first, ncalls = g.local_calls(args.ncalls, rank, world)
        total_calls = args.ncalls
    else:
        first, ncalls = rank * args.ncalls, args.ncalls
        total_calls = args.ncalls * world
    xmat, grid_source = workload_grid(wl, bins, dim)
    nacc = ctypes.sizeof(_abi.Accum) // 8
    seed = 12345678
    state = {"xmat": xmat, "acc": None, "launches": 0, "ms_kernels": 0.0, "ms_screen": 0.0, "n_screen": 0,
             "fid": 0.0, "evals": 0.0, "t_host": 0.0}

    def step(it, timed):
        t_a = time.perf_counter()
        ptr = E.vegas_iterate_async(state["xmat"], first, ncalls, seed, it)
        if world > 1:
            if args.comm == "native":
                E.allreduce()
            else:
                with torch.cuda.stream(ext):
                    g.allreduce_accum(torch.as_tensor(g.DevArray(ptr, nacc + 2 * world), device=torch.device("cuda", local)), rank, world)
        acc = E.vegas_fetch()                      # D2H of the accumulators (41 KB) + stream sync
        t_b = time.perf_counter()
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
            state["t_host"] += time.perf_counter() - t_b   # grid refinement + bookkeeping between two iterations

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

    # ---- e2e: the same metric through the C ABI with HOST buffers (seam A x n) --------------------------------------
    # grb_eval_rng = MProcess::EventWeight(randvec) as the reference calls it: the caller supplies the coordinates, the
    # internal uniforms come from the engine's generator.  grb_eval (coordinates + a caller tape) is timed beside it.
    # One call carries the batch of one step (2^24 points per GPU by default, like the device-timed iteration); the host
    # restatement of the Philox stream makes 2^22 distinct points through the converged grid, tiled to the batch size.
    n_e2e = args.e2e_points if args.e2e_points > 0 else max(1 << 20, min(int(ncalls), 1 << 24))
    n_distinct = min(n_e2e, 1 << 22)
    hu = torch.empty((dim, n_e2e), dtype=torch.float64).pin_memory()
    ht = torch.empty((ntape, n_e2e), dtype=torch.float64).pin_memory()
    hw = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    hf = torch.empty(n_e2e, dtype=torch.uint8).pin_memory()
    pu, ptape = synthetic_points(wl, n_distinct, dim, ntape, 1234, first=rank * n_distinct)
    for o in range(0, n_e2e, n_distinct):
        m = min(n_distinct, n_e2e - o)
        hu.numpy()[:, o:o + m] = pu.T[:, :m]
        ht.numpy()[:, o:o + m] = ptape.T[:, :m]
    del pu, ptape
    e2e_steps = max(3, min(args.steps, 10))
    e2e = {}
    for api in ("eval_rng", "eval"):
        def call(k):
            if api == "eval_rng":
                E.eval_rng_raw(n_e2e, hu.numpy(), dim, seed, k, rank * n_e2e, hw.numpy(), hf.numpy())
            else:
                E.eval_raw(n_e2e, hu.numpy(), dim, ht.numpy(), ntape, hw.numpy(), hf.numpy())
        for k in range(2):
            call(k)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            call(10 + k)                            # H2D + kernels + D2H, synchronous
        e2e[api] = time.perf_counter() - t0
    e2e_launches = E.last_timing()[1] * e2e_steps

    red = torch.tensor([dev_ms, wall_ms, e2e["eval_rng"], e2e["eval"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_rng_s, e2e_tape_s = [float(x) for x in red.cpu()]

    if rank == 0:
        total_points = float(total_calls) * args.steps
        value = total_points / (dev_ms * 1e-3)
        peak = g.measure_dfma_peak(local)
        F_kin, F_born, F_node, f_source = flop_model(wl)
        p_acc = state["fid"] / state["evals"] if state["evals"] else 0.0
        n_items = state["fid"] / world      # accepted points per rank (the counters are all-reduced)
        pts_rank = float(ncalls) * args.steps
        if pom:
            nkt_nodes, nphi_nodes = E.loop_nodes()
            n_nodes = nkt_nodes * nphi_nodes     # the nodes the kernel executes (phi = 2 pi folded onto phi = 0)
            kernel, kernel_ms, kernel_n = "k_screen", state["ms_screen"], max(state["n_screen"], 1)
            flops_kernel = n_items * n_nodes * F_node
            f_point = F_kin + p_acc * (F_born + n_nodes * F_node)
        else:
            n_nodes = 0
            kernel, kernel_ms, kernel_n = "k_points", state["ms_kernels"], max(int(state["launches"]), 1)
            flops_kernel = pts_rank * (F_kin + p_acc * F_born)
            f_point = F_kin + p_acc * F_born
        ach = flops_kernel / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        mix = sass_mix(wl + ":" + kernel)
        roof = {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak if peak else None,
                "peak_source": "self-measured, live: dependent-free DFMA chains on all SMs (grb_measure_dfma_peak); "
                               "MEASURED_PEAKS.json and the profiling guide have no fp64 figure",
                "flops_source": f_source, "F_kin": F_kin, "F_born": F_born, "F_node": F_node,
                "executed_nodes_per_accepted_point": n_nodes, "flops_per_point_model": f_point,
                "whole_path_tflops": f_point * value / 1e12,
                "kernel_share_of_step": kernel_ms / state["ms_kernels"] if state["ms_kernels"] else None,
                "launch_ms_avg": kernel_ms / kernel_n, "traffic": None,
                "frac_note": "frac = algorithmic flops (SURVEY 8d convention: transcendental = 20 flop, op-counted on the oracle) / "
                             "self-measured DFMA peak; the device needs fewer executed flops than the convention counts (table-driven "
                             "log / exp / sincos, merged phases), so read it together with executed_frac (SASS flops / peak) and "
                             "fp64_slot_frac (fp64 pipe slots in use = what is left for this instruction stream)"}
        if mix:   # executed (not algorithmic) fp64 work of the same kernel from the committed ncu opcode mix
            units = n_items * n_nodes if pom else pts_rank
            ex_flops = units * (2 * mix["dfma"] + mix["dmul"] + mix["dadd"])
            slots = units * (mix["dfma"] + mix["dmul"] + mix["dadd"] + mix.get("dsetp", 0) + mix.get("other_fp64", 0))
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6 if clocks else 1965e6
            roof["executed_flops_per_unit"] = 2 * mix["dfma"] + mix["dmul"] + mix["dadd"]
            roof["executed_tflops"] = ex_flops / (kernel_ms * 1e-3) / 1e12
            roof["executed_frac"] = roof["executed_tflops"] / peak if peak else None
            roof["fp64_slot_frac"] = slots / (64.0 * 148 * sm_hz * kernel_ms * 1e-3)
            roof["mix_source"] = mix.get("source")
            if mix.get("dram_bytes_per_unit"):
                roof["traffic"] = mix["dram_bytes_per_unit"] * (n_items if pom else pts_rank) / kernel_n
                roof["traffic_unit"] = "bytes per launch (ncu dram read+write per unit x units per launch)"
        acc = state["acc"]
        sigma = acc.fsum / total_calls
        h2d_rng, h2d_tape = int(dim * 8 * n_e2e), int((dim + ntape) * 8 * n_e2e)
        line = {"metric": metric_name(wl), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": descr, "workload_id": wl,
                           "l2": "inputs are generated on the device per step (Philox counters), nothing is re-read across steps; "
                                 "the per-chunk work list (151 MB) exceeds L2"},
                "run": {"points_per_step_per_gpu": ncalls, "points_per_step": total_calls, "vegas_bins": bins, "dim": dim,
                           "grid": grid_source + "; it keeps adapting during the run",
                           "collective": None if world == 1 else ("grb_allreduce (library communicator, NCCL)" if args.comm == "native"
                                                                  else "torch.distributed all_reduce (NCCL) on the engine stream"),
                           "timing": "CUDA events on the engine stream around K steps, max over ranks",
                           "accepted_fraction": p_acc, "sigma_last_iteration_ub": sigma * 1e6},
                "wall_ms_per_step": wall_ms / args.steps,
                "turnaround_ms_per_step": {"device_kernels": state["ms_kernels"] / args.steps,
                                           "exchange_fetch_sync": (dev_ms - state["ms_kernels"]) / args.steps,
                                           "host_grid_refinement": state["t_host"] / args.steps * 1e3},
                "clocks": clocks,
                "gpu_launches": int(state["launches"]),
                "e2e": {"value": n_e2e * world * e2e_steps / e2e_rng_s, "unit": UNIT,
                        "h2d_bytes_per_step": h2d_rng, "d2h_bytes_per_step": int(9 * n_e2e),
                        "points_per_step_per_gpu": n_e2e, "distinct_points": n_distinct, "steps": e2e_steps, "gpu_launches": int(e2e_launches),
                        "api": "grb_eval_rng (seam A x n: coordinates from pinned host memory, internal draws from the engine's "
                               "Philox stream like the reference's per-thread generator), H2D + D2H inside the timed region",
                        "with_caller_tape": {"value": n_e2e * world * e2e_steps / e2e_tape_s, "api": "grb_eval",
                                             "h2d_bytes_per_step": h2d_tape}},
                "roofline": roof}
        if world == 1 and not args.no_cpu:
            nproc = os.cpu_count() or 1
            n, s, kind = cpu_reference_sample(wl, args.cpu_points or cpu_pts, nproc, 77)
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--comm", default="torch", choices=["torch", "native"])
    ap.add_argument("--ncalls", type=int, default=1 << 24, help="VEGAS points per step per GPU (weak) or in total (strong)")
    ap.add_argument("--e2e-points", type=int, default=0, help="points per e2e call per GPU (0 = the step's own batch, at most 2^24)")
    ap.add_argument("--cpu-points", type=int, default=0, help="cpu_baseline sample: points per host process (0 = per workload)")
    ap.add_argument("--ref-points", type=int, default=0, help="--impl reference: points per host process per step")
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
