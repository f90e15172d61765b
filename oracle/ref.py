"""Loader for oracle/_ref/torch_rw_native.so -- the UNMODIFIED reference extension.

TEST INFRASTRUCTURE ONLY.  Built from the sources under /root/reference by oracle/build_ref.py;
used to validate oracle/trw_oracle.c, to generate tests/golden/, and as the `kind: "reference"`
CPU baseline of bench.py.  `native()` returns the pybind11 module whose seven functions take
tensors positionally, exactly as csrc/rw_init.cpp:133-141 defines them.
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_ref", "torch_rw_native.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO_PATH)


def native():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError(f"{SO_PATH} not built; run python oracle/build_ref.py where /root/reference exists")
        import torch  # noqa: F401  (libc10/libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location("torch_rw_native", SO_PATH)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mod = mod
    return _mod
