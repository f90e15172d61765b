"""Builds the optional pybind11 module `torch_rw_native_b200` (csrc/binding/rw_init.cpp): the
reference's extension layout over libtrw_b200.so.  C++ only (no device code: the kernels live in
the C-ABI library), about a minute with the PyTorch headers.  The product path (native.py, ctypes)
does not need it.

    python -m torch_random_walk_b200._build_ext
"""
import ctypes
import importlib.util
import os
import shutil
import sys

from . import _build

PKG_DIR = _build.PKG_DIR
EXT_NAME = "torch_rw_native_b200"
EXT_PATH = os.path.join(PKG_DIR, EXT_NAME + ".so")
SRC = os.path.join(PKG_DIR, "csrc", "binding", "rw_init.cpp")


def needs_build() -> bool:
    return not os.path.exists(EXT_PATH) or os.path.getmtime(EXT_PATH) < max(os.path.getmtime(SRC), os.path.getmtime(
        os.path.join(PKG_DIR, "..", "include", "trw_b200.h")))


def build(force: bool = False, verbose: bool = False) -> str:
    _build.build()
    if not force and not needs_build():
        return EXT_PATH
    from torch.utils.cpp_extension import load

    # the extension names libtrw_b200.so by soname; with the library already mapped, the loader
    # resolves it without any rpath (the in-tree .so may live anywhere)
    ctypes.CDLL(_build.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    bdir = os.path.join(PKG_DIR, "_ext_build")
    os.makedirs(bdir, exist_ok=True)
    load(name=EXT_NAME, sources=[SRC], extra_include_paths=[os.path.join(PKG_DIR, "..", "include")],
         extra_cflags=["-O2"], extra_ldflags=[f"-L{PKG_DIR}", "-ltrw_b200"],
         with_cuda=True, build_directory=bdir, verbose=verbose, is_python_module=False)
    shutil.copy2(os.path.join(bdir, EXT_NAME + ".so"), EXT_PATH)
    shutil.rmtree(bdir, ignore_errors=True)
    return EXT_PATH


def load_module():
    """Imports the built extension (None when it has not been built)."""
    if not os.path.exists(EXT_PATH):
        return None
    import torch  # noqa: F401
    ctypes.CDLL(_build.LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    spec = importlib.util.spec_from_file_location(EXT_NAME, EXT_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
