"""Build recipe for oracle/_ref: the UNMODIFIED reference extension, compiled in place.

TEST INFRASTRUCTURE ONLY.  Compiles the reference's own sources where they lie under
/root/reference (csrc/rw_init.cpp, csrc/cpu/*.cpp, csrc/cuda/*.cu) into
oracle/_ref/torch_rw_native.so.  Nothing is copied into this repository: only the built
shared object lands in oracle/_ref/ (git-ignored, but it travels to the GPU box).

The reference's setup.py cannot be used: it hard-codes -arch=sm_35 (setup.py:42, rejected by
CUDA 12) and takes the ROCm branch when no GPU is visible (setup.py:14-28).  The flags below are
the reference's own (-O2 -fopenmp -DAT_PARALLEL_OPENMP, setup.py:33-35) plus an sm_100 gencode so
that the reference CUDA kernels can also be run on a B200 as a comparison row.

Usage:  python oracle/build_ref.py            (about 4-5 minutes, CPU only)
"""
import glob
import os
import shutil
import sys
import time

REF = os.environ.get("TRW_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(verbose: bool = True) -> str:
    target = os.path.join(OUT, "torch_rw_native.so")
    if not os.path.isdir(os.path.join(REF, "csrc")):
        if os.path.exists(target):
            return target
        raise RuntimeError(f"reference sources not found under {REF} and no prebuilt {target}")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 8))
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load

    srcs = ([os.path.join(REF, "csrc", "rw_init.cpp")]
            + sorted(glob.glob(os.path.join(REF, "csrc", "cpu", "*.cpp")))
            + sorted(glob.glob(os.path.join(REF, "csrc", "cuda", "*.cu"))))
    bdir = os.path.join(OUT, "build")
    os.makedirs(bdir, exist_ok=True)
    t0 = time.time()
    load(name="torch_rw_native", sources=srcs,
         extra_include_paths=[os.path.join(REF, "csrc")],
         extra_cflags=["-O2", "-fopenmp", "-DAT_PARALLEL_OPENMP"],
         extra_cuda_cflags=["-O2", "-gencode=arch=compute_100,code=sm_100"],
         with_cuda=True, build_directory=bdir, verbose=verbose, is_python_module=False)
    shutil.copy2(os.path.join(bdir, "torch_rw_native.so"), target)
    shutil.rmtree(bdir, ignore_errors=True)  # keep only the shared object
    print(f"[oracle/_ref] built {target} in {time.time() - t0:.0f}s", file=sys.stderr)
    return target


if __name__ == "__main__":
    build()
