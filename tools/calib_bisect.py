"""Bisects why the walk kernels gather fewer random lines per second than the calibration kernel
(measurement tooling): calibration at walk-like shapes, with an L2-resident side lookup, and the
first-order walk on a degree-regular random graph.  Prints one line per case.

    python tools/calib_bisect.py
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native  # noqa: E402


def main():
    lib = native.lib()
    sink = torch.zeros(8, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def calib(table, n_threads, loads, nbytes, label):
        elems = table.numel()
        for _ in range(2):
            lib.trw_calib_gather(ctypes.c_void_p(table.data_ptr()), elems, n_threads, loads, nbytes, 7,
                                 ctypes.c_void_p(sink.data_ptr()), 0, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            rc = lib.trw_calib_gather(ctypes.c_void_p(table.data_ptr()), elems, n_threads, loads, nbytes, 7,
                                      ctypes.c_void_p(sink.data_ptr()), 0, st)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        ms = e0.elapsed_time(e1) / 3
        print(f"calib {label}: {ms:.3f} ms, {n_threads * loads / ms / 1e6:.1f} G loads/s", flush=True)

    for gib in (4, 8):
        table = torch.randint(0, 1 << 40, (gib * (1 << 30) // 8,), dtype=torch.int64, device="cuda")
        calib(table, 4 << 20, 32, 8, f"{gib}GiB 4Mi x32 8B")
        calib(table, 8868124, 80, 8, f"{gib}GiB 8.87M x80 8B")
        calib(table, 8868124, 80, 32, f"{gib}GiB 8.87M x80 32B")
        for aux in (16, 64):
            native.set_option("calib_aux_mb", aux)
            calib(table, 8868124, 80, 8, f"{gib}GiB 8.87M x80 8B + {aux} MiB L2 side lookup")
        native.set_option("calib_aux_mb", 0)
        del table

    # first-order walk on a degree-regular random graph of c3's size (no hubs)
    n, d, L = 1 << 24, 31, 80
    rp = torch.arange(n + 1, dtype=torch.int64, device="cuda") * d
    ci = torch.randint(0, n, (n * d,), dtype=torch.int64, device="cuda")
    targets = torch.arange(8868124, dtype=torch.int64, device="cuda")
    out = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    for opts in ({}, {"store_mode": 3}, {"records": 0}, {"records": 0, "store_mode": 3}):
        for k, v in opts.items():
            native.set_option(k, v)
        for _ in range(3):
            native.walk(rp, ci, targets, 1.0, 1.0, L, 5, out=out)
        torch.cuda.synchronize()
        b, w = native.last_kernel_ms()
        print(f"uniform walk, regular random graph, {opts}: build {b:.2f} ms walk {w:.3f} ms, "
              f"{targets.numel() * L / w / 1e6:.1f} G steps/s", flush=True)
        for k in opts:
            native.set_option(k, {"store_mode": 0, "records": 1}[k])


if __name__ == "__main__":
    main()
