"""Builds torch_random_walk_b200/libtrw_b200.so from csrc/*.cu with nvcc, for sm_100a only.

The library has no PyTorch dependency (plain C ABI, include/trw_b200.h); it is compiled in-tree
so that the shared object travels with the source tree.  `python -m torch_random_walk_b200._build`
rebuilds it; `build()` is what __graft_entry__.build() calls.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtrw_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "--use_fast_math",
    "-Xlinker", "-soname=libtrw_b200.so",
    "-shared",
    "-ldl",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libtrw_b200.so")
    return cand


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(PKG_DIR, "..", "include", "trw_b200.h")]
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB_PATH)
