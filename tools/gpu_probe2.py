"""Second GPU probe (measurement tooling): random-fetch calibration by table size and granule
width, and A/B timings of the rw.walk options on the benchmark graph.  Writes gpurun_out/probe2.json.

    python tools/gpu_probe2.py [--scale 24]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402

DEFAULTS = {"stage_output": 1, "n2v_table": 1, "n2v_speculate": -1, "persist_row_ptr": 0, "row32": 1, "build_mode": 2,
            "n2v_min_ctas": -1, "persist_l2_mb": 64}


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--skip-calib", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe2.json"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    res = {"device": torch.cuda.get_device_name(0)}
    lib = native.lib()

    def save():
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)

    if not args.skip_calib:
        calib = {}
        sink = torch.zeros(8, dtype=torch.int64, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for table_gib in (0.25, 1, 4, 16):
            elems = int(table_gib * (1 << 30) / 8)
            table = torch.randint(0, 1 << 40, (elems,), dtype=torch.int64, device="cuda")
            for nbytes in (8, 32, 64, 128):
                n_threads, loads = 8 << 20, 32
                ms = timed(lambda: lib.trw_calib_gather(ctypes.c_void_p(table.data_ptr()), elems, n_threads, loads, nbytes, 7,
                                                        ctypes.c_void_p(sink.data_ptr()), 0, st), reps=2)
                rate = n_threads * loads / (ms / 1e3)
                calib[f"table{table_gib}GiB_granule{nbytes}B"] = {"ms": ms, "G_fetches_per_s": rate / 1e9,
                                                                  "GBps": rate * max(nbytes, 32) / 1e9}
                print("calib", table_gib, nbytes, f"{rate / 1e9:.1f} G fetches/s", flush=True)
            del table
        res["calib"] = calib
        save()

    t0 = time.time()
    rp, ci = rmat.rmat_csr(args.scale, 16, device="cuda")
    torch.cuda.synchronize()
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    n_walks, L = targets.numel(), 80
    res["graph"] = {"scale": args.scale, "n_nodes": rp.numel() - 1, "nnz": ci.numel(), "max_degree": int(deg.max()),
                    "walks": n_walks, "gen_s": time.time() - t0}
    print(res["graph"], flush=True)
    out = torch.empty((n_walks, L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    native.set_graph_cache(False)  # every variant builds with its own options
    variants = {}

    def run(name, p, q, opts):
        for k, v in {**DEFAULTS, **opts}.items():
            native.set_option(k, v)
        ms = timed(lambda: native.walk(rp, ci, targets, p, q, L, 5, out=out), reps=2, warm=1)
        b, w = native.last_kernel_ms()
        variants[name] = {"call_ms": ms, "build_ms": b, "walk_ms": w, "G_steps_per_s_call": n_walks * L / ms / 1e6,
                          "G_steps_per_s_kernel": n_walks * L / max(w, 1e-9) / 1e6}
        print(name, {k: round(v, 3) for k, v in variants[name].items()}, flush=True)
        for k, v in DEFAULTS.items():
            native.set_option(k, v)
        res["variants"] = variants
        save()

    grid = [
        ("default", {}),
        ("row64", {"row32": 0}),
        ("flat_build", {"build_mode": 0}),
        ("ctas5", {"n2v_min_ctas": 5}),
        ("ctas6", {"n2v_min_ctas": 6}),
        ("spec0", {"n2v_speculate": 0}),
        ("spec1", {"n2v_speculate": 1}),
        ("persist64", {"persist_row_ptr": 1, "persist_l2_mb": 64}),
        ("persist79", {"persist_row_ptr": 1, "persist_l2_mb": 79}),
        ("no_stage", {"stage_output": 0}),
    ]
    for pq_name, p, q in (("c3_p1_q0.5", 1.0, 0.5), ("uniform", 1.0, 1.0), ("c2_p0.5_q2", 0.5, 2.0), ("c5_p0.25_q4", 0.25, 4.0)):
        for vname, opts in grid:
            if pq_name == "uniform" and vname in ("flat_build", "coop_build", "tiles1", "tiles4", "ctas5", "ctas6", "spec0", "spec1"):
                continue
            if pq_name in ("c2_p0.5_q2", "c5_p0.25_q4") and vname in ("flat_build", "coop_build", "tiles1", "tiles4", "persist79", "no_stage"):
                continue
            run(f"{pq_name}/{vname}", p, q, opts)
    save()
    print("probe2 done", flush=True)


if __name__ == "__main__":
    main()
