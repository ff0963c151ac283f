"""Builds libhsl_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so is git-ignored
but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhsl_b200.so")
SOURCES = ["hsl_kernels.cu", "hsl_select.cu", "hsl_fall.cu", "hsl_gather.cu", "hsl_capi.cu", "hsl_model_load.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))) + [os.path.join("..", "..", "include", "hsl.h")]  # every header counts as a dependency
# hsl_fall.cu: no FMA contraction.  The sweep kernel is bound by its thread-local working set, not by FP64 issue, and the
# contact set of a step hangs on the sign of depths that are zero up to round-off (the gait puts stance feet exactly at
# touching distance), so it pays to keep the arithmetic identical to the CPU stepper it is compared with.
PER_FILE_FLAGS = {"hsl_fall.cu": ["-fmad=false"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--resource-usage"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    objs = []
    logs = []
    extra = os.environ.get("HSL_NVCC_EXTRA", "").split()
    for s in SOURCES:
        o = os.path.join(LIBDIR, s + ".o")
        cmd = [_nvcc(), "-ccbin", ccbin] + NVCC_FLAGS + PER_FILE_FLAGS.get(s, []) + extra + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs.append(r.stdout)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("nvcc failed on " + s)
        objs.append(o)
    cmd = [_nvcc(), "-ccbin", ccbin, "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
