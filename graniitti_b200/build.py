"""In-tree build of the CUDA library (graniitti_b200/libgrb.so) for sm_100a with nvcc.

One object per translation unit (compiled in parallel, rebuilt only when stale) under graniitti_b200/build/,
then one link.  The built .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgrb.so")
SOURCES = ["grb_consts.cu", "grb_points_regge.cu", "grb_points_tensor.cu", "grb_points_res.cu", "grb_points_rescon.cu",
           "grb_points_yy.cu", "grb_points_tensorvv.cu", "grb_points_tensorff.cu", "grb_points_yyqed.cu", "grb_points_me6.cu", "grb_points_me8.cu", "grb_points_tres.cu", "grb_points_trescon.cu", "grb_kernels_screen.cu", "grb_api.cu", "grb_eikonal.cu", "grb_host.cc", "grb_events.cc"]
# -rdc=true: the __constant__ parameter blocks are defined once (grb_consts.cu) and shared by all kernel TUs
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-rdc=true",
              "-Xcompiler", "-fPIC"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
        os.path.join(HERE, "..", "include", "grb.h")]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, out=None, extra=()):
    """`out`/`extra` build an experimental variant next to the product library (profiling only)."""
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if not (extra and f.startswith("-fmad"))] + list(extra)
    objdir = OBJ if out is None else OBJ + "_" + os.path.basename(out)
    os.makedirs(objdir, exist_ok=True)
    hdrs = _headers()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    lib = out or LIB
    # a library that is newer than every source and header is up to date even when the object directory did not travel
    if not force and not extra and not _newer(lib, [os.path.join(CSRC, s) for s in srcs] + hdrs):
        return lib
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s + ".o")
        if force or _newer(obj, [src] + hdrs):
            jobs.append([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src])
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for rc in ex.map(subprocess.call, jobs):
            if rc != 0:
                raise subprocess.CalledProcessError(rc, "nvcc -c")
    objs = [os.path.join(objdir, s + ".o") for s in srcs]
    if jobs or _newer(lib, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-rdc=true", "-shared", "-Xcompiler", "-fPIC",
                               "-o", lib] + objs + ["-ldl"])
    return lib


if __name__ == "__main__":
    import sys
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
