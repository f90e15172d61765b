"""B200-native random-walk sampler behind torch_rw's API.

    from torch_random_walk_b200 import rw, utils      # or: from torch_rw import rw, utils

`rw` and `utils` mirror /root/reference/torch_rw/{rw,utils}.py; `native` is the
`torch_rw_native` surface over libtrw_b200.so (hand-written sm_100a CUDA, C ABI in
include/trw_b200.h).  Importing `rw`/`native` loads the CUDA library and raises if it is
missing: there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import utils  # noqa: F401


def __getattr__(name):
    if name in ("rw", "native", "dist", "rmat"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
