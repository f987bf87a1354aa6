"""Builds libmvmc.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

    python -m multiview_motion_capture_b200.build          # build if stale
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmvmc.so")
SOURCES = ["affinity.cu", "als.cu", "assign.cu", "ik.cu", "ingest.cu", "matchers.cu", "pipeline.cu"]
# ik.cu: no implicit FMA contraction (every fused operation in the solver is an explicit fma()), so that the CPU build of
# the same sources (tests/emu, g++ -ffp-contract=off) reproduces the GPU's results bit for bit (parity of I4)
# (als.cu also takes -DAL_MAXREG_=112, which caps k_als at 112 registers with no spills to speak of and leaves room for one IK
#  solver warp per SM beside two ALS CTAs: measured 1-2 % slower alone and no better overlapped - 4 691 vs 4 765 frames/s - so off)
PER_FILE_FLAGS = {"ik.cu": ["-fmad=false"]}
# the same source built again with other tile shapes (see als.cu: AL_VARIANT)
VARIANTS = [("als.cu", "als_small.o", ["-DAL_VARIANT=small", "-DAL_FM_=3", "-DAL_THREADS_=128"])]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(ROOT, "include", "mvmc.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    nvcc = _nvcc()
    objs = []
    jobs = []
    for s, name, extra in [(s, s.replace(".cu", ".o"), PER_FILE_FLAGS.get(s, [])) for s in SOURCES] + VARIANTS:
        src = os.path.join(CSRC, s)
        obj = os.path.join(LIBDIR, name)
        objs.append(obj)
        if force or _stale(obj, [src] + headers + [os.path.abspath(__file__)]):
            jobs.append([nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


TORCH_EXT = os.path.join(LIBDIR, "libmvmc_torch.so")


def build_torch_ext(force=False, verbose=True):
    """The PyTorch C++ extension over the C-ABI (csrc/torch_ext.cpp -> lib/libmvmc_torch.so): g++ against the installed
    torch's headers, linked to libmvmc.so next to it ($ORIGIN rpath). Needs no GPU."""
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(CSRC, "torch_ext.cpp")
    if not (force or _stale(TORCH_EXT, [src, os.path.join(ROOT, "include", "mvmc.h"), LIB])):
        return TORCH_EXT
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=libmvmc_torch",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-I" + os.path.join(ROOT, "include"),
           "-I/usr/local/cuda/include"] + ["-isystem" + p for p in ce.include_paths()] + [
           src, "-o", TORCH_EXT, "-L" + LIBDIR, "-lmvmc", "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch",
           "-Wl,-rpath,$ORIGIN", "-Wl,--no-as-needed"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return TORCH_EXT


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv))
    print(build_torch_ext(force="--force" in sys.argv))
