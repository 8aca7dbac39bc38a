"""Ahead-of-time build of libgst_cuda.so for sm_100a (no JIT, no run-time kernel sources).

Replaces the reference's run-time OpenCL compilation (gpu/kernel_cache.cpp:146-165): the
kernels are compiled once, here, with nvcc, and the shared library is kept in-tree under
gst_b200/lib/ so it travels with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libgst_cuda.so")
SOURCES = ["gst_kernels.cu", "gst_capi.cu"]
DEPS = SOURCES + ["gst_kernels.cuh", os.path.join("..", "..", "include", "gst_cuda.h")]
# GST_NVCC_EXTRA: extra nvcc flags (e.g. -DSOME_VARIANT) for A/B builds, see scripts/build_var.sh and scripts/ab.sh
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-shared", "-x", "cu",
] + os.environ.get("GST_NVCC_EXTRA", "").split()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    for d in DEPS:
        p = os.path.join(CSRC, d)
        if os.path.exists(p) and os.path.getmtime(p) > built:
            return True
    return False


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into gst_b200/lib/libgst_cuda.so."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I", os.path.join(HERE, "..", "include"), "-o", LIB_PATH + ".tmp"] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgst_cuda.so")
    if verbose:
        sys.stderr.write(res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
