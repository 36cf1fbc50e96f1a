"""In-tree build of the CUDA library (graniitti_b200/libgrb.so) for sm_100a with nvcc.

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgrb.so")
SOURCES = ["grb_kernels.cu", "grb_api.cu", "grb_eikonal.cu", "grb_host.cc"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "grb.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, out=None, extra=()):
    """`out`/`extra` build an experimental variant next to the product library (profiling only)."""
    if out is None and not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if not (extra and f.startswith("-fmad"))] + list(extra)
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", out or LIB] + [
        os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
