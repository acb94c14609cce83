"""Build the sm_100a extension in-tree: bopy_b200/lib/libbopy_b200.so (C ABI, include/bopy_b200.h).

    python -m bopy_b200.build            # or bopy_b200.build.build()
nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbopy_b200.so")
SOURCES = [os.path.join(CSRC, "bopy_b200.cu")]
HEADERS = [os.path.join(CSRC, f) for f in ("common.cuh", "sweep_kernel.cuh", "aux_kernels.cuh", "fit_kernels.cuh", "probe_kernel.cuh",
                                             "grad_kernel.cuh", "prune_kernels.cuh", "sweep_tc_kernel.cuh", "tc_common.cuh", "minloc_comm.cuh", "small_kernel.cuh", "sweep_group_kernel.cuh", "sweep_warp_kernel.cuh")] + [
    os.path.join(ROOT, "include", "bopy_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libbopy_b200.so")
    return exe


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile (if stale) and return the path of the shared library."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES + ["-ldl"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
