"""ctypes binding of include/grb.h (plumbing only)."""
import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("GRB_LIB", os.path.join(_HERE, "libgrb.so"))  # GRB_LIB: experimental builds only
_lib = None


class GrbError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"grb error {rc}: {msg}")
        self.rc = rc


def lib():
    """Load libgrb.so (fails loudly if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: run `python -m graniitti_b200.build` "
                              "(or __graft_entry__.build()); the engine has no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        vp, i64, u64, dbl, cint = C.c_void_p, C.c_int64, C.c_uint64, C.c_double, C.c_int
        L.grb_abi_sizeof.restype = cint
        L.grb_abi_sizeof.argtypes = [cint]
        L.grb_create.restype = cint
        L.grb_create.argtypes = [C.POINTER(_abi.ProcessDesc), cint, C.POINTER(vp)]
        L.grb_create_from_card.restype = cint
        L.grb_create_from_card.argtypes = [C.c_char_p, C.c_char_p, cint, C.c_char_p, cint, C.POINTER(vp)]
        L.grb_parse_card.restype = cint
        L.grb_parse_card.argtypes = [C.c_char_p, C.c_char_p, cint, C.c_char_p, C.POINTER(_abi.ProcessDesc),
                                     C.c_char_p, cint]
        L.grb_get_desc.restype = cint
        L.grb_get_desc.argtypes = [vp, C.POINTER(_abi.ProcessDesc)]
        L.grb_last_error.restype = C.c_char_p
        L.grb_last_error.argtypes = [vp]
        L.grb_destroy.restype = None
        L.grb_destroy.argtypes = [vp]
        L.grb_set_eikonal.restype = cint
        L.grb_set_eikonal.argtypes = [vp, vp, vp, vp, cint]
        L.grb_build_eikonal.restype = cint
        L.grb_build_eikonal.argtypes = [vp, cint, cint, vp, vp, vp, vp, vp, vp, vp]
        L.grb_eikonal_cache_names.restype = cint
        L.grb_eikonal_cache_names.argtypes = [C.POINTER(_abi.ProcessDesc), cint, cint, cint, cint, C.c_char_p,
                                              C.c_char_p, cint]
        L.grb_write_eikonal_file.restype = cint
        L.grb_write_eikonal_file.argtypes = [C.c_char_p, vp, vp, vp, cint]
        L.grb_read_eikonal_file.restype = cint
        L.grb_read_eikonal_file.argtypes = [C.c_char_p, vp, vp, vp, cint]
        L.grb_update_desc.restype = cint
        L.grb_update_desc.argtypes = [vp, C.POINTER(_abi.ProcessDesc)]
        L.grb_set_beam_energies.restype = cint
        L.grb_set_beam_energies.argtypes = [vp, dbl, dbl]
        L.grb_xscan.restype = cint
        L.grb_xscan.argtypes = [vp, vp, cint, C.POINTER(_abi.VegasParam), C.POINTER(_abi.XscanRow)]
        L.grb_eval.restype = cint
        L.grb_eval.argtypes = [vp, i64, vp, cint, vp, cint, vp, vp, vp]
        L.grb_eval_rng.restype = cint
        L.grb_eval_rng.argtypes = [vp, i64, vp, cint, u64, u64, i64, vp, vp, vp]
        L.grb_loop_nodes.restype = cint
        L.grb_loop_nodes.argtypes = [vp, C.POINTER(cint), C.POINTER(cint)]
        L.grb_events_dropped.restype = i64
        L.grb_events_dropped.argtypes = [vp]
        L.grb_event_record.restype = cint
        L.grb_event_record.argtypes = [C.POINTER(_abi.ProcessDesc), vp, cint, cint, vp, C.POINTER(_abi.RecParticle), C.POINTER(cint),
                                       C.POINTER(_abi.RecVertex), C.POINTER(cint)]
        L.grb_write_events.restype = i64
        L.grb_write_events.argtypes = [C.POINTER(_abi.ProcessDesc), C.c_char_p, C.POINTER(_abi.WriterOpts), i64, vp, vp]
        L.grb_write_engine_events.restype = i64
        L.grb_write_engine_events.argtypes = [vp, C.c_char_p, C.POINTER(_abi.WriterOpts)]
        L.grb_vegas_iterate.restype = cint
        L.grb_vegas_iterate.argtypes = [vp, vp, cint, cint, i64, i64, u64, u64, cint, dbl, C.POINTER(_abi.Accum)]
        L.grb_vegas_iterate_async.restype = cint
        L.grb_vegas_iterate_async.argtypes = [vp, vp, cint, cint, i64, i64, u64, u64, cint, dbl, C.POINTER(vp)]
        L.grb_vegas_fetch.restype = cint
        L.grb_vegas_fetch.argtypes = [vp, C.POINTER(_abi.Accum)]
        L.grb_sync.restype = cint
        L.grb_sync.argtypes = [vp]
        L.grb_vegas_optimize_grid.restype = cint
        L.grb_vegas_optimize_grid.argtypes = [vp, vp, cint, cint, dbl]
        L.grb_get_events.restype = i64
        L.grb_get_events.argtypes = [vp, i64, vp, vp, C.POINTER(cint)]
        L.grb_comm_unique_id.restype = cint
        L.grb_comm_unique_id.argtypes = [C.c_char_p]
        L.grb_comm_init.restype = cint
        L.grb_comm_init.argtypes = [vp, cint, cint, C.c_char_p]
        L.grb_allreduce.restype = cint
        L.grb_allreduce.argtypes = [vp]
        L.grb_measure_dfma_peak.restype = cint
        L.grb_measure_dfma_peak.argtypes = [cint, C.POINTER(dbl)]
        L.grb_last_timing.restype = cint
        L.grb_last_timing.argtypes = [vp, C.POINTER(dbl), C.POINTER(cint)]
        L.grb_integrate.restype = cint
        L.grb_integrate.argtypes = [vp, C.POINTER(_abi.VegasParam), cint, cint, _abi.REDUCE_FN, vp,
                                    C.POINTER(_abi.VegasResult)]
        L.grb_sample_flat.restype = cint
        L.grb_sample_flat.argtypes = [vp, C.POINTER(_abi.FlatParam), cint, cint, _abi.REDUCE_FN, vp, C.POINTER(_abi.VegasResult)]
        L.grb_integrate_batch.restype = cint
        L.grb_integrate_batch.argtypes = [vp, C.POINTER(_abi.ProcessDesc), cint, C.POINTER(_abi.VegasParam), cint, cint, _abi.REDUCE_FN, vp,
                                          C.POINTER(_abi.VegasResult)]
        L.grb_local_calls.restype = cint
        L.grb_local_calls.argtypes = [i64, cint, cint, C.POINTER(i64), C.POINTER(i64)]
        L.grb_test_fastmath.restype = cint
        L.grb_test_fastmath.argtypes = [cint, cint, i64, vp, vp]
        L.grb_get_stream.restype = vp
        L.grb_get_stream.argtypes = [vp]
        L.grb_last_timing_kernels.restype = cint
        L.grb_last_timing_kernels.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(cint)]
        _lib = L
    return _lib


def _ptr(x):
    """Raw address of a numpy array / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def parse_card(card, process=None, pomloop=-1, base=None):
    """Host-only: steering card (+ overrides) -> ProcessDesc. Mirrors `gr -i card -p process -l pomloop`."""
    d = _abi.ProcessDesc()
    buf = C.create_string_buffer(2048)
    rc = lib().grb_parse_card(card.encode(), process.encode() if process else None, int(pomloop),
                              base.encode() if base else None, C.byref(d), buf, len(buf))
    if rc != 0:
        raise GrbError(rc, buf.value.decode(errors="replace"))
    return d


def eikonal_cache_names(desc, beam_pdg=(2212, 2212), number_bt=0, number_kt2=0):
    """File names of the reference's eikonal/ table cache for this configuration (MBT_..., MSA_...)."""
    a, b = C.create_string_buffer(128), C.create_string_buffer(128)
    rc = lib().grb_eikonal_cache_names(C.byref(desc), beam_pdg[0], beam_pdg[1], number_bt, number_kt2, a, b, 128)
    if rc != 0:
        raise GrbError(rc, "grb_eikonal_cache_names")
    return a.value.decode(), b.value.decode()


def write_eikonal_file(path, x, re, im):
    x, re, im = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, re, im))
    rc = lib().grb_write_eikonal_file(str(path).encode(), x.ctypes.data, re.ctypes.data, im.ctypes.data, len(x))
    if rc != 0:
        raise GrbError(rc, f"grb_write_eikonal_file({path})")


def read_eikonal_file(path, n_points=10001):
    t = np.zeros((3, n_points))
    rc = lib().grb_read_eikonal_file(str(path).encode(), t[0].ctypes.data, t[1].ctypes.data, t[2].ctypes.data, n_points)
    if rc != 0:
        raise GrbError(rc, f"grb_read_eikonal_file({path})")
    return t


def test_fastmath(kind, x, device=0):
    """Device table-driven log (kind 0) / exp (kind 1) of grb_fastmath.cuh on a host array (unit-test hook)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    rc = lib().grb_test_fastmath(device, kind, x.size, x.ctypes.data, y.ctypes.data)
    if rc != 0:
        raise GrbError(rc, "grb_test_fastmath")
    return y


def optimize_grid(xmat, f2mat, lam=1.5):
    """VEGAS grid refinement (host). xmat [bins][dim] float64 C-order (updated in place),
    f2mat [bins][GRB_MAX_DIM]."""
    bins, dim = xmat.shape
    assert xmat.flags.c_contiguous and f2mat.flags.c_contiguous and f2mat.shape == (bins, _abi.GRB_MAX_DIM) or \
        f2mat.shape[1] == _abi.GRB_MAX_DIM
    rc = lib().grb_vegas_optimize_grid(xmat.ctypes.data, f2mat.ctypes.data, bins, dim, float(lam))
    if rc != 0:
        raise GrbError(rc, "grb_vegas_optimize_grid")
    return xmat


def local_calls(calls, rank, nranks):
    """(first, n): the shard of one iteration that `rank` evaluates (reference MVEGAS.h:59-69)."""
    first, n = C.c_int64(0), C.c_int64(0)
    rc = lib().grb_local_calls(calls, rank, nranks, C.byref(first), C.byref(n))
    if rc != 0:
        raise GrbError(rc, "grb_local_calls")
    return first.value, n.value


def comm_unique_id():
    """128 opaque bytes (ncclUniqueId) made by rank 0 and shipped to the other ranks by any means."""
    buf = C.create_string_buffer(128)
    rc = lib().grb_comm_unique_id(buf)
    if rc != 0:
        raise GrbError(rc, lib().grb_last_error(None).decode(errors="replace"))
    return buf.raw


def event_record(desc, momenta, beam_pdg=(2212, 2212), lifetime_u=None):
    """The reference's event record (MProcess::CommonRecord) of one event from momenta [3 + n_nodes][4]: (particles, vertices)
    as lists of dicts."""
    mom = np.ascontiguousarray(momenta, dtype=np.float64)
    P = (_abi.RecParticle * _abi.GRB_MAX_REC_PARTICLES)()
    V = (_abi.RecVertex * _abi.GRB_MAX_REC_VERTICES)()
    n_p, n_v = C.c_int(0), C.c_int(0)
    lu = np.ascontiguousarray(lifetime_u, dtype=np.float64) if lifetime_u is not None else None
    rc = lib().grb_event_record(C.byref(desc), mom.ctypes.data, beam_pdg[0], beam_pdg[1], lu.ctypes.data if lu is not None else None,
                                P, C.byref(n_p), V, C.byref(n_v))
    if rc != 0:
        raise GrbError(rc, "grb_event_record")
    parts = [dict(id=p.id, pid=p.pid, status=p.status, parent=p.parent, p=(p.px, p.py, p.pz, p.e), m=p.m) for p in P[:n_p.value]]
    verts = [dict(id=v.id, n_in=v.n_in, n_out=v.n_out, inc=list(v.in_[:v.n_in]), pos=(v.x, v.y, v.z, v.t)) for v in V[:n_v.value]]
    return parts, verts


def writer_opts(fmt="hepmc3", sigma_pb=0.0, sigma_err_pb=0.0, weighted=False, nthreads=0, append=False, close_listing=True,
                card_path="", lifetimes=False, number_events=False, first_event=0):
    return _abi.WriterOpts(fmt.encode(), None, None, card_path.encode(), sigma_pb, sigma_err_pb, 0, 0, int(weighted), nthreads,
                           int(append), int(close_listing), int(number_events), int(lifetimes), first_event, 12345)


def write_events(desc, path, momenta, weights=None, **kw):
    """n events as SoA momenta [3 + n_nodes][4][n] (+ weights [2][n]) -> file in the reference's output format."""
    mom = np.ascontiguousarray(momenta, dtype=np.float64)
    n = mom.shape[2]
    w = np.ascontiguousarray(weights, dtype=np.float64) if weights is not None else None
    o = writer_opts(**kw)
    rc = lib().grb_write_events(C.byref(desc), str(path).encode(), C.byref(o), n, mom.ctypes.data, w.ctypes.data if w is not None else None)
    if rc < 0:
        raise GrbError(int(rc), "grb_write_events")
    return int(rc)


def measure_dfma_peak(device=0):
    v = C.c_double(0)
    rc = lib().grb_measure_dfma_peak(device, C.byref(v))
    if rc != 0:
        raise GrbError(rc, "grb_measure_dfma_peak (needs a CUDA device)")
    return v.value


class Engine:
    """One engine per process/GPU rank (reference: one MProcess copy per thread, MGraniitti.cc:278-309)."""

    def __init__(self, desc=None, card=None, process=None, pomloop=-1, base=None, device=0):
        self._h = C.c_void_p()
        L = lib()
        if desc is not None:
            rc = L.grb_create(C.byref(desc), device, C.byref(self._h))
        else:
            rc = L.grb_create_from_card(card.encode(), process.encode() if process else None, int(pomloop),
                                        base.encode() if base else None, device, C.byref(self._h))
        if rc != 0:
            raise GrbError(rc, L.grb_last_error(None).decode(errors="replace"))
        self.desc = _abi.ProcessDesc()
        L.grb_get_desc(self._h, C.byref(self.desc))

    def _check(self, rc):
        if rc != 0:
            raise GrbError(rc, lib().grb_last_error(self._h).decode(errors="replace"))

    def close(self):
        if self._h:
            lib().grb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_eikonal(self, kt2, re, im):
        kt2, re, im = (np.ascontiguousarray(a, dtype=np.float64) for a in (kt2, re, im))
        self._check(lib().grb_set_eikonal(self._h, kt2.ctypes.data, re.ctypes.data, im.ctypes.data, len(kt2)))

    def build_eikonal(self, number_bt=0, number_kt2=0):
        """Build the opacity (MBT) + screening-amplitude (MSA) tables on the device (reference
        MEikonal::S3Constructor) and install the MSA. Returns dict(mbt [3][nbt+1], msa [3][nkt2+1], xs_mb [3])."""
        nb = (number_bt if number_bt > 0 else 10000) + 1
        nk = (number_kt2 if number_kt2 > 0 else 10000) + 1
        mbt, msa, xs = np.zeros((3, nb)), np.zeros((3, nk)), np.zeros(3)
        self._check(lib().grb_build_eikonal(self._h, number_bt, number_kt2, mbt[0].ctypes.data, mbt[1].ctypes.data,
                                            mbt[2].ctypes.data, msa[0].ctypes.data, msa[1].ctypes.data,
                                            msa[2].ctypes.data, xs.ctypes.data))
        return dict(mbt=mbt, msa=msa, xs_mb=xs)

    def eval(self, u, tape, want_dump=False):
        """Host numpy convenience: u [dim][n], tape [dtape][n] (SoA). Returns w, flags, dump."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        tape = np.ascontiguousarray(tape, dtype=np.float64)
        dim, n = u.shape
        dt = tape.shape[0]
        w = np.zeros(n)
        flags = np.zeros(n, dtype=np.uint8)
        dump = np.zeros((_abi.GRB_DUMP_LEN, n)) if want_dump else None
        self._check(lib().grb_eval(self._h, n, u.ctypes.data, dim, tape.ctypes.data if dt else None, dt,
                                   w.ctypes.data, flags.ctypes.data, _ptr(dump)))
        return w, flags, dump

    def eval_raw(self, n, u, dim, tape, dtape, w, flags, dump=None):
        """Pointer-level call (numpy arrays or torch CUDA tensors)."""
        self._check(lib().grb_eval(self._h, n, _ptr(u), dim, _ptr(tape), dtape, _ptr(w), _ptr(flags), _ptr(dump)))

    def eval_rng_raw(self, n, u, dim, seed, iteration, first_index, w, flags, dump=None):
        """Seam A as the reference calls it: coordinates from the caller, internal draws from the engine's Philox stream."""
        self._check(lib().grb_eval_rng(self._h, n, _ptr(u), dim, seed, iteration, first_index, _ptr(w), _ptr(flags), _ptr(dump)))

    def eval_rng(self, u, seed, iteration=0, first_index=0):
        u = np.ascontiguousarray(u, dtype=np.float64)
        dim, n = u.shape
        w = np.zeros(n)
        flags = np.zeros(n, dtype=np.uint8)
        self.eval_rng_raw(n, u, dim, seed, iteration, first_index, w, flags)
        return w, flags

    def loop_nodes(self):
        """(kt nodes, phi nodes) of the screening loop as the kernels execute them."""
        a, b = C.c_int(0), C.c_int(0)
        self._check(lib().grb_loop_nodes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def write_events(self, path, **kw):
        """The events of the last GENERATE iteration -> file (hepmc3 | hepmc2 | hepevt)."""
        o = writer_opts(**kw)
        rc = lib().grb_write_engine_events(self._h, str(path).encode(), C.byref(o))
        if rc < 0:
            self._check(int(rc))
        return int(rc)

    def events_dropped(self):
        return int(lib().grb_events_dropped(self._h))

    def vegas_iterate(self, xmat, first_call, ncalls, seed, iteration, mode=_abi.GRB_MODE_INTEGRATE, maxf=0.0):
        xmat = np.ascontiguousarray(xmat, dtype=np.float64)
        bins, dim = xmat.shape
        acc = _abi.Accum()
        self._check(lib().grb_vegas_iterate(self._h, xmat.ctypes.data, bins, dim, first_call, ncalls, seed,
                                            iteration, mode, maxf, C.byref(acc)))
        return acc

    def vegas_iterate_async(self, xmat, first_call, ncalls, seed, iteration, mode=_abi.GRB_MODE_INTEGRATE, maxf=0.0):
        xmat = np.ascontiguousarray(xmat, dtype=np.float64)
        bins, dim = xmat.shape
        p = C.c_void_p()
        self._check(lib().grb_vegas_iterate_async(self._h, xmat.ctypes.data, bins, dim, first_call, ncalls, seed,
                                                  iteration, mode, maxf, C.byref(p)))
        return p.value

    def vegas_fetch(self):
        acc = _abi.Accum()
        self._check(lib().grb_vegas_fetch(self._h, C.byref(acc)))
        return acc

    def integrate(self, ncall, iters=5, precision=0.01, chi2max=10.0, bins=128, lam=1.5, seed=12345678,
                  burnin=3, rank=0, nranks=1, reduce=None, debug=-1, max_iter=0):
        """Burn-in + integration (reference MGraniitti::Initialize -> SampleVegas). `reduce(device_ptr)` is
        called once per iteration for nranks > 1 (e.g. a torch.distributed all-reduce on the buffer)."""
        par = _abi.VegasParam(bins, iters, burnin, max_iter, ncall, lam, chi2max, precision, seed, debug, 0)
        res = _abi.VegasResult()
        cb = _abi.REDUCE_FN(lambda ctx, ptr: int(reduce(ptr) or 0)) if reduce else _abi.REDUCE_FN()
        self._check(lib().grb_integrate(self._h, C.byref(par), rank, nranks, cb, None, C.byref(res)))
        return res

    def sample_flat(self, precision=0.05, min_events=1e6, batch=0, seed=12345678, rank=0, nranks=1, reduce=None, max_events=0):
        """INTEGRATOR "FLAT" (reference MGraniitti::SampleFlat): plain Monte Carlo in device batches."""
        par = _abi.FlatParam(precision, min_events, batch, max_events, seed)
        res = _abi.VegasResult()
        cb = _abi.REDUCE_FN(lambda ctx, ptr: int(reduce(ptr) or 0)) if reduce else _abi.REDUCE_FN()
        self._check(lib().grb_sample_flat(self._h, C.byref(par), rank, nranks, cb, None, C.byref(res)))
        return res

    def integrate_batch(self, descs, ncall, iters=5, precision=0.01, chi2max=10.0, bins=128, lam=1.5, seed=12345678, burnin=3):
        """K parameter sets (same topology) integrated back to back on this engine (fit loops)."""
        par = _abi.VegasParam(bins, iters, burnin, 0, ncall, lam, chi2max, precision, seed, -1, 0)
        arr = (_abi.ProcessDesc * len(descs))(*descs)
        res = (_abi.VegasResult * len(descs))()
        self._check(lib().grb_integrate_batch(self._h, arr, len(descs), C.byref(par), 0, 1, _abi.REDUCE_FN(), None, res))
        lib().grb_get_desc(self._h, C.byref(self.desc))
        return list(res)

    def update_desc(self, desc):
        """Re-parameterise without re-creating the engine (same decay-tree topology and phase-space class)."""
        self._check(lib().grb_update_desc(self._h, C.byref(desc)))
        lib().grb_get_desc(self._h, C.byref(self.desc))

    def set_beam_energies(self, e1, e2):
        self._check(lib().grb_set_beam_energies(self._h, float(e1), float(e2)))
        lib().grb_get_desc(self._h, C.byref(self.desc))

    def xscan(self, sqrt_s, ncall, iters=5, precision=0.01, chi2max=10.0, bins=128, lam=1.5, seed=12345678, burnin=3):
        """Energy scan (reference src/Program/xscan.cc): rows of (sqrt_s, xs_tot/el/inel [mb], sigma +- err [barn])."""
        e = np.ascontiguousarray(sqrt_s, dtype=np.float64)
        par = _abi.VegasParam(bins, iters, burnin, 0, ncall, lam, chi2max, precision, seed, -1, 0)
        rows = (_abi.XscanRow * len(e))()
        self._check(lib().grb_xscan(self._h, e.ctypes.data, len(e), C.byref(par), rows))
        lib().grb_get_desc(self._h, C.byref(self.desc))
        return list(rows)

    def get_events(self, max_events=1 << 22):
        """(momenta [npart][4][n], weights [2][n]) of the events accepted by the last GENERATE iteration."""
        npart = 3 + self.desc.n_nodes
        mom = np.zeros((npart, 4, max_events))
        wts = np.zeros((2, max_events))
        npp = C.c_int(0)
        n = lib().grb_get_events(self._h, max_events, mom.ctypes.data, wts.ctypes.data, C.byref(npp))
        if n < 0:
            self._check(int(n))
        # the library packs rows with stride n
        mom = mom.reshape(-1)[:npart * 4 * n].reshape(npart, 4, n)
        wts = wts.reshape(-1)[:2 * n].reshape(2, n)
        return mom, wts

    def comm_init(self, rank, nranks, unique_id):
        """Join the library's own NCCL communicator (torch-free hosts): grb_comm_init."""
        self._check(lib().grb_comm_init(self._h, rank, nranks, C.create_string_buffer(bytes(unique_id), 128)))

    def allreduce(self):
        """All-rank sums / maxima of the accumulators of the last vegas_iterate_async, on the engine's stream."""
        self._check(lib().grb_allreduce(self._h))

    def sync(self):
        self._check(lib().grb_sync(self._h))

    def stream_ptr(self):
        return lib().grb_get_stream(self._h)

    def last_timing_kernels(self):
        """(total device ms, summed k_screen device ms, number of k_screen launches) of the last call."""
        tot, scr, n = C.c_double(0), C.c_double(0), C.c_int(0)
        lib().grb_last_timing_kernels(self._h, C.byref(tot), C.byref(scr), C.byref(n))
        return tot.value, scr.value, n.value

    def last_timing(self):
        ms, nl = C.c_double(0), C.c_int(0)
        lib().grb_last_timing(self._h, C.byref(ms), C.byref(nl))
        return ms.value, nl.value
