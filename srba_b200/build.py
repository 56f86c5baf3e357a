"""Builds the in-tree shared library srba_b200/libsrba_b200.so for sm_100a (B200):
CUDA kernels (csrc/lm_kernel.cu) + C ABI (csrc/capi.cpp, include/srba_b200.h) + host engine (csrc/engine_capi.cpp).
nvcc cross-compiles without a GPU.  Usage: python -m srba_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsrba_b200.so")
SOURCES = ["lm_kernel.cu", "lm_kernel_cluster.cu", "lm_kernel_wave.cu", "capi.cpp", "engine_capi.cpp"]
HEADERS = ["dev_types.h", "lm_wave.cuh", "host_engine.hpp", "host_pose.hpp", "../../include/srba_b200.h", "../../include/srba_b200_engine.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0", "--shared", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "-Xcompiler", "-pthread", "-cudart", "static"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DSRBA_SOLVER_TIMERS"] if os.environ.get("SRBA_SOLVER_TIMERS") else []   # per-phase clocks inside the solver (tools/stage_breakdown.py)
    extra += os.environ.get("SRBA_B200_NVCC_EXTRA", "").split()   # experiment builds, e.g. "-DSE3_NT=512" (with SRBA_B200_LIB=<other .so> as the output)
    out = os.environ.get("SRBA_B200_LIB", SO)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
