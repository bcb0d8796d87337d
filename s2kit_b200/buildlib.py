"""Builds s2kit_b200/libs2kit_cuda.so in-tree: hand-written CUDA for sm_100a + the C host layer.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libs2kit_cuda.so")

CU = ["kernels_fft.cu", "kernels_legendre.cu", "kernels_table.cu", "kernels_misc.cu", "kernels_pipe.cu", "kernels_uni.cu", "kernels_flow.cu", "kernels_fft16.cu", "plan.cu", "shard.cu", "multi.cu"]
C = ["host_setup.c", "s2kit_compat.c"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# -ffp-contract=off: the libm-based seeds must not be re-associated into FMAs (host_setup.c)
GCC_FLAGS = ["-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-Wall"]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _headers():
    inc = os.path.join(HERE, "..", "include")
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h")]
    return hs


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for f in CU:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            jobs.append((["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj], obj + ".log"))
    for f in C:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            jobs.append((["gcc"] + GCC_FLAGS + ["-c", src, "-o", obj], obj + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            outs = list(ex.map(lambda j: _run(*j), jobs))
        if verbose:
            for o in outs:
                print(o)
    if jobs or not os.path.exists(LIB):
        _run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs +
             ["-lpthread", "-lm"], os.path.join(OBJ, "link.log"))
    # descriptor-only FFTW plan stubs for callers without FFTW (include/s2kit_fftw_shim/fftw3.h)
    shim_src, shim_lib = os.path.join(CSRC, "fftw_shim.c"), os.path.join(HERE, "libs2kit_fftw_shim.so")
    if force or _newer([shim_src] + hdrs, shim_lib):
        _run(["gcc"] + GCC_FLAGS + ["-shared", "-o", shim_lib, shim_src], os.path.join(OBJ, "shim.log"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
