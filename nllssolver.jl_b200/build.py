"""Builds libnlls_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libnlls_b200.so")
SRCS = ["csrc/nlls_b200.cu"]
DEPS = SRCS + sorted("csrc/" + f for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cuh", ".hpp"))) + ["../include/nlls_b200.h"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # residual/assembly kernels are HBM-bound; keeping mul/add unfused tracks the reference's (unfused) Julia arithmetic
    "-shared", "-Xcompiler", "-fPIC",
]
LIBS = ["-ldl", "-lpthread"]   # NCCL is resolved with dlopen at nlls_comm_init; no CUDA math library is linked


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: development variants (A/B timing of compile-time switches), e.g.
    build(force=True, extra_flags=["-DS5_FULL6=0"], out="build/variants/libnlls_bands.so"), loaded with NLLS_B200_LIB=<path>."""
    if out is None and not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = os.path.abspath(out) if out else SO
    os.makedirs(os.path.dirname(target), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + SRCS + LIBS
    subprocess.check_call(cmd, cwd=HERE)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
