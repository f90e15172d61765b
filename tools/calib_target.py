"""ncu target for the random-fetch calibration kernels (profiles/): 8-byte dependent random loads
from a 4 GiB table, one launch per load flavour (cache operator / L2 policy), then 32/64/128-byte
granules.  Shows what one random access costs in DRAM traffic on this part.

    python tools/calib_target.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native  # noqa: E402

MODES = {0: "ld.global.nc.L1::no_allocate", 1: "ld.global", 2: "ld.global.cg", 3: "ld.global.cv",
         4: "ld.global.nc + L2 evict_first policy", 5: "ld.global.lu", 6: "ld.global.cs"}


def main():
    lib = native.lib()
    elems = (4 << 30) // 8
    table = torch.randint(0, 1 << 40, (elems,), dtype=torch.int64, device="cuda")
    sink = torch.zeros(8, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(nbytes, label):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.trw_calib_gather(ctypes.c_void_p(table.data_ptr()), elems, 4 << 20, 32, nbytes, 7,
                                  ctypes.c_void_p(sink.data_ptr()), 0, st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ms = e0.elapsed_time(e1)
        print(f"{label}: {ms:.3f} ms, {(4 << 20) * 32 / ms / 1e6:.1f} G fetches/s", flush=True)

    for mode, name in MODES.items():
        native.set_option("calib_mode", mode)
        run(8, f"8B mode {mode} ({name})")
    native.set_option("calib_mode", 0)
    for nbytes in (32, 64, 128):
        run(nbytes, f"{nbytes}B granule")


if __name__ == "__main__":
    main()
