"""Builds libnle_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.  No CPU fallback exists."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "dense.cu", "eig.cu", "eig_dc.cu", "eig_topk.cu", "filter_kernels.cu", "cell_kernels.cu", "sinkhorn_cells.cu", "apply_kernels.cu", "nccl_comm.cu", "peaks.cu"]
LIB = os.path.join(HERE, "libnle_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "nle_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libnle_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


HOST_TEST = os.path.join(HERE, "host_mirror_test")


def build_host_test(force=False):
    """C++ host-side test program over include/nle_b200.hpp (g++, links the in-tree libnle_b200.so)."""
    src = os.path.join(HERE, "..", "tests", "cpp", "host_mirror_test.cpp")
    inc = os.path.join(HERE, "..", "include")
    deps = [src, os.path.join(inc, "nle_b200.hpp"), os.path.join(inc, "nle_b200.h"), LIB]
    if not force and os.path.exists(HOST_TEST) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_TEST) for d in deps):
        return HOST_TEST
    build()
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-I", inc, src, "-o", HOST_TEST, "-L", HERE, "-lnle_b200",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building host_mirror_test")
    return HOST_TEST


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    build_host_test(force=True)
    print(LIB)
