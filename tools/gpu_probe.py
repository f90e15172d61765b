"""First-contact GPU probe (measurement tooling): device limits, the attainable random-sector
rate (library gather kernels + torch.index_select), kernel-variant timings of rw.walk on the
benchmark graph, and the reference's own CUDA kernels (oracle/_ref, recompiled for sm_100) on the
same inputs.  Writes gpurun_out/probe.json.

    python tools/gpu_probe.py [--scale 24] [--skip-ref]
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
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.json"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    res = {"device": torch.cuda.get_device_name(0)}
    lib = native.lib()
    info = (ctypes.c_int64 * 6)()
    lib.trw_device_info(0, info, 6)
    res["device_info"] = dict(zip(["sms", "l2_bytes", "max_persisting_l2", "max_policy_window", "l2_fetch_granularity",
                                   "max_smem_optin"], list(info)))
    print(res, flush=True)

    def save():
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)

    # ---------------- random gather calibration
    calib = {}
    sink = torch.zeros(8, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for table_gib in (0.125, 4):
        elems = int(table_gib * (1 << 30) / 8)
        table = torch.randint(0, 1 << 40, (elems,), dtype=torch.int64, device="cuda")
        for gran in (64, 32):
            native.set_option("l2_fetch_granularity", gran)
            for nbytes in (8, 32):
                for threads_m in (4, 16):
                    n_threads, loads = threads_m << 20, 64
                    ms = timed(lambda: lib.trw_calib_gather(ctypes.c_void_p(table.data_ptr()), elems, n_threads, loads, nbytes, 7,
                                                            ctypes.c_void_p(sink.data_ptr()), 0, st))
                    rate = n_threads * loads / (ms / 1e3)
                    calib[f"table{table_gib}GiB_gran{gran}_load{nbytes}B_threads{threads_m}M"] = {
                        "ms": ms, "G_loads_per_s": rate / 1e9, "sector_GBps": rate * 32 / 1e9}
                    print("calib", table_gib, gran, nbytes, threads_m, f"{rate / 1e9:.1f} G loads/s", flush=True)
        native.set_option("l2_fetch_granularity", 64)
        idx = torch.randint(0, elems, (1 << 26,), device="cuda")
        ms = timed(lambda: torch.index_select(table, 0, idx))
        calib[f"table{table_gib}GiB_index_select_64M"] = {"ms": ms, "G_loads_per_s": (1 << 26) / (ms / 1e3) / 1e9}
        del table, idx
    res["calib"] = calib
    save()

    # ---------------- the benchmark graph
    t0 = time.time()
    rp, ci = rmat.rmat_csr(args.scale, 16, device="cuda")
    torch.cuda.synchronize()
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    n_walks, L = targets.numel(), 80
    res["graph"] = {"scale": args.scale, "n_nodes": rp.numel() - 1, "nnz": ci.numel(), "max_degree": int(deg.max()),
                    "isolated": int((deg == 0).sum()), "walks": n_walks, "gen_s": time.time() - t0,
                    "deg_ge_16_frac_nnz": float(deg[deg >= 16].sum()) / ci.numel()}
    print(res["graph"], flush=True)
    out = torch.empty((n_walks, L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    variants = {}

    def run(name, p, q, opts):
        for k, v in opts.items():
            native.set_option(k, v)
        try:
            ms = timed(lambda: native.walk(rp, ci, targets, p, q, L, 5, out=out), reps=2, warm=1)
            b, w = native.last_kernel_ms()
            variants[name] = {"call_ms": ms, "build_ms": b, "walk_ms": w, "G_steps_per_s_call": n_walks * L / ms / 1e6,
                              "G_steps_per_s_kernel": n_walks * L / max(w, 1e-9) / 1e6}
            print(name, variants[name], flush=True)
        finally:
            for k in opts:
                native.set_option(k, {"stage_output": 1, "n2v_table": 1, "n2v_speculate": 1, "persist_row_ptr": 0,
                                      "l2_fetch_granularity": 64}[k])
        res["variants"] = variants
        save()

    for pq_name, p, q in (("c3_p1_q0.5", 1.0, 0.5), ("uniform", 1.0, 1.0), ("c2_p0.5_q2", 0.5, 2.0), ("c5_p0.25_q4", 0.25, 4.0)):
        run(f"{pq_name}/default", p, q, {})
        run(f"{pq_name}/no_stage", p, q, {"stage_output": 0})
        run(f"{pq_name}/persist_row_ptr", p, q, {"persist_row_ptr": 1})
        run(f"{pq_name}/fetch32", p, q, {"l2_fetch_granularity": 32})
        if pq_name != "uniform":
            run(f"{pq_name}/no_speculate", p, q, {"n2v_speculate": 0})

    # ---------------- the reference's CUDA kernels on the same graph (sample of walks)
    if not args.skip_ref:
        try:
            from oracle import ref

            rn = ref.native()
            refres = {}
            sample = targets[torch.randperm(n_walks, device="cuda")[: 1 << 20]].contiguous()
            ms = timed(lambda: rn.walk(rp, ci, sample, 1.0, 1.0, L, 5), reps=2)
            refres["uniform_1M_walks"] = {"ms": ms, "G_steps_per_s": sample.numel() * L / ms / 1e6}
            print("ref uniform", refres, flush=True)
            small = sample[: 1 << 14].contiguous()
            ms = timed(lambda: rn.walk(rp, ci, small, 1.0, 0.5, L, 5), reps=1, warm=0)
            refres["p1_q0.5_16k_walks"] = {"ms": ms, "G_steps_per_s": small.numel() * L / ms / 1e6}
            print("ref biased", refres, flush=True)
            res["reference_cuda_sm100"] = refres
        except Exception as exc:  # noqa: BLE001
            res["reference_cuda_sm100"] = {"error": repr(exc)}
    save()
    print("probe done", flush=True)


if __name__ == "__main__":
    main()
