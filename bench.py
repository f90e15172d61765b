#!/usr/bin/env python
"""bench.py -- walk-steps/s of the hot path (rw.walk, node2vec) on synthetic R-MAT graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1|tiny]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's csrc/cpu path on the host cores

A "step" is one rw.walk call over every start node (all nodes with degree > 0), made the way a user
of the drop-in API makes it: the same CSR tensors, resident in HBM, every call.  Default workload (c3)
is BASELINE.json configs[2] -- R-MAT scale 24 (16.8 M nodes, ~2^29 CSR entries), p=1 q=0.5,
walk_length=80 -- the configuration the north-star target is quoted on.

N = 1.  `value` is the drop-in call with the library defaults: native.walk keeps the graph-side
preparation of a graph it has seen before, identified by CONTENT -- every timed call first checksums
row_ptr and col_idx on the device (inside the timed region), then runs the walk kernel on the kept
preparation.  The warm-up calls pay the preparation and its cost is reported (`graph_cache`,
`prepared_handle.prepare_ms`).  `stateless` in the same line is the same loop with the cache off:
everything rebuilt inside every call, like the reference's stateless launcher; `prepared_handle` is
the explicit native.prepare_csr handle (no checksum).  `roofline` is the walk kernel of the timed calls.

N > 1.  One NCCL broadcast replicates the CSR and every rank prepares its replica (dist.ReplicatedCsr),
outside the timed region.  `value` is ONE start-node list sharded block-cyclically over the ranks
("scaling": "strong"): total steps / max-over-ranks time; the shards' digests are all-reduced and
compared with one call over the whole list (`shards_equal_single`).  Contiguous shards and the
replicated list (weak scaling) are reported beside it; at N = 8 configs[4] (c5) is added.

One JSON line goes to stdout (rank 0); progress goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, scale, n_nodes, edge_factor or n_edges, p, q, walk_length)
    "c3": dict(desc="R-MAT scale 24 (16.8M nodes, ~2^29 CSR entries), node2vec p=1 q=0.5 L=80 [BASELINE configs[2]]",
               scale=24, n_nodes=None, edge_factor=16, p=1.0, q=0.5, L=80),
    "c2": dict(desc="ogbn-products-shaped R-MAT (2,449,029 nodes, ~62M CSR entries), node2vec p=0.5 q=2 L=80 [configs[1]]",
               scale=22, n_nodes=2449029, n_edges=34_000_000, p=0.5, q=2.0, L=80),
    "c5": dict(desc="Friendster-shaped R-MAT (65.6M nodes, ~1.8B CSR entries), node2vec p=0.25 q=4 L=40 [configs[4]]",
               scale=26, n_nodes=65608366, n_edges=960_000_000, p=0.25, q=4.0, L=40),
    "c3u": dict(desc="R-MAT scale 24, first-order p=q=1 L=80", scale=24, n_nodes=None, edge_factor=16, p=1.0, q=1.0, L=80),
    "c4": dict(desc="FB15k-237-shaped triples (14,541 entities, 237 relations, 310,116 triples), 10 walks/entity, 40 hops, "
                    "then to_windows_triples(window_size=5) [BASELINE configs[3]]", kind="triples", n_entities=14541,
               n_relations=237, n_triples=310116, walks_per_entity=10, L=40, W=5, p=1.0, q=1.0),
    "tiny": dict(desc="R-MAT scale 16 smoke workload, node2vec p=1 q=0.5 L=80", scale=16, n_nodes=None, edge_factor=16,
                 p=1.0, q=0.5, L=80),
}
BYTES_PER_STEP = {True: 72, False: 104}  # first-order / node2vec algorithmic bytes per step


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line.  Libraries (NCCL's version banner, for one) also write to
# file descriptor 1, so it is pointed at stderr for the whole run and the line goes to the saved one.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples while the timed region runs: one `nvidia-smi -lms 50`
    process started before the region and stopped after it (the recipe of B200_PROFILING.md)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # let the first sample land before the region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is None:
            return
        try:
            time.sleep(0.06)
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=10)
            self.rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass

    def summary(self):
        sm, mx, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


def build_graph(wl, device):
    from torch_random_walk_b200 import rmat

    t0 = time.time()
    row_ptr, col_idx = rmat.rmat_csr(wl["scale"], wl.get("edge_factor", 16), n_nodes=wl.get("n_nodes"), device=device,
                                     n_edges=wl.get("n_edges"))
    if device != "cpu":
        torch.cuda.synchronize()
    log(f"graph: n={row_ptr.numel() - 1:,} nnz={col_idx.numel():,} built in {time.time() - t0:.1f}s")
    return row_ptr, col_idx


def start_nodes(row_ptr):
    deg = row_ptr[1:] - row_ptr[:-1]
    return torch.nonzero(deg > 0).flatten().contiguous()


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own csrc/cpu walk (oracle/_ref) or the oracle port
# ----------------------------------------------------------------------------------------------
def cpu_walk_fn():
    from oracle import ref

    if ref.available():
        native = ref.native()
        return (lambda rp, ci, tg, p, q, L, seed: native.walk(rp, ci, tg, p, q, L, seed)), "reference"
    from oracle import orc

    return (lambda rp, ci, tg, p, q, L, seed: orc.walk(rp, ci, tg, p, q, L, seed)), "port"


def time_cpu(fn, row_ptr, col_idx, targets, wl, budget_s, threads):
    """steps/s of the CPU path on a bounded sample of the start nodes (about budget_s seconds)."""
    torch.set_num_threads(threads)
    n = min(2048, targets.numel())
    t0 = time.perf_counter()
    fn(row_ptr, col_idx, targets[:n].contiguous(), wl["p"], wl["q"], wl["L"], 10)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(targets.numel(), max(n, n * budget_s / dt)))
    t0 = time.perf_counter()
    fn(row_ptr, col_idx, targets[:n].contiguous(), wl["p"], wl["q"], wl["L"], 10)
    dt = time.perf_counter() - t0
    return n * wl["L"] / dt, n, dt


def cpu_sample_targets(targets_cpu, count=200_000):
    g = torch.Generator().manual_seed(1)
    if targets_cpu.numel() <= count:
        return targets_cpu
    return targets_cpu[torch.randperm(targets_cpu.numel(), generator=g)[:count]].contiguous()


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    device = "cuda" if torch.cuda.is_available() else "cpu"
    row_ptr, col_idx = build_graph(wl, device)
    row_ptr, col_idx = row_ptr.cpu(), col_idx.cpu()
    targets = cpu_sample_targets(start_nodes(row_ptr))
    fn, kind = cpu_walk_fn()
    cores = os.cpu_count() or 1
    per_step_budget = max(2.0, min(20.0, 120.0 / max(args.steps + args.warmup, 1)))
    if os.environ.get("TRW_BENCH_CPU_BUDGET_S"):  # tests shrink the CPU sample
        per_step_budget = float(os.environ["TRW_BENCH_CPU_BUDGET_S"])
    results = {}
    for threads in sorted({1, cores}):
        best = 0.0
        for it in range(args.warmup + args.steps):
            sps, n, dt = time_cpu(fn, row_ptr, col_idx, targets, wl, per_step_budget / 2, threads)
            if it >= args.warmup:
                best = max(best, sps)
            log(f"reference cpu threads={threads} iter={it}: {sps:.3e} steps/s ({n} walks, {dt:.2f}s)")
        results[threads] = best
    best_threads = max(results, key=results.get)
    value = results[best_threads]
    line = {
        "impl": "reference", "metric": "walk_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "p": wl["p"], "q": wl["q"], "walk_length": wl["L"], "n_nodes": row_ptr.numel() - 1,
                   "nnz": col_idx.numel()},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": best_threads, "kind": kind,
                         "sample": f"random sample of degree>0 start nodes sized to ~{per_step_budget / 2:.0f}s per timing; "
                                   f"steps/s by thread count: { {k: round(v) for k, v in results.items()} }"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def measure_c4(wl, dev, steps, warmup, e2e_steps):
    """configs[3]: knowledge-graph triple walks + window generation.  A step = rw.walk_triples over
    10 walks per entity followed by rw.to_windows_triples on the result; returns hops/s of the whole
    step, windows/s and the window kernel's HBM-write roofline.  Single GPU (the workload is 145k walks)."""
    from torch_random_walk_b200 import native, rmat, rw

    n_ent, pad = wl["n_entities"], wl["n_entities"] + wl["n_relations"]
    triples = rmat.kg_triples(n_ent, wl["n_relations"], wl["n_triples"], device=dev)
    index, ts = rmat.relation_tail_index(triples, n_ent)
    targets = torch.arange(n_ent, device=dev).repeat_interleave(wl["walks_per_entity"])
    L, W = wl["L"], wl["W"]

    def step(seed):
        walks = rw.walk_triples(ts, index, targets, walk_length=L, padding_idx=pad, seed=seed)
        return walks, rw.to_windows_triples(walks, W, n_ent, pad, ts, seed)

    def timed_loop(fn):
        # Every call returns ~3 GB of fresh tensors.  The previous result is dropped BEFORE the next call so that
        # torch's caching allocator hands the same blocks back; holding two generations made it cudaMalloc/cudaFree
        # inside the loop on some runs (1 ms vs 19 ms per step for identical kernels).
        out_ = None
        for k in range(max(warmup, 3)):
            out_ = None
            out_ = fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            out_ = None
            out_ = fn(100 + k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out_

    native.reset_launch_count()
    step_ms, (walks, outs) = timed_loop(step)                                    # the whole step, K times
    launches = native.launch_count() * steps // (steps + max(warmup, 3))
    outs = None
    walk_ms, walks = timed_loop(lambda sd: rw.walk_triples(ts, index, targets, walk_length=L, padding_idx=pad, seed=sd))
    win_ms, outs = timed_loop(lambda sd: rw.to_windows_triples(walks, W, n_ent, pad, ts, sd))
    hops = targets.numel() * L * steps
    n_win = outs[0].size(0)
    win_bytes = sum(o.numel() for o in outs) * 8
    peak, peak_src = measured_peaks()
    # e2e: pinned CPU tensors in, windows back in pinned host buffers, copies inside the timed region
    ts_h, idx_h, tg_h = ts.cpu().pin_memory(), index.cpu().pin_memory(), targets.cpu().pin_memory()
    host = [torch.empty(o.shape, dtype=torch.int64, pin_memory=True) for o in outs]
    outs = walks = None
    torch.cuda.synchronize()
    t_0 = time.perf_counter()
    ts_d = w_ = o_ = None
    for k in range(e2e_steps + 1):  # the first pass warms the allocator and is not timed
        if k == 1:
            torch.cuda.synchronize()
            t_0 = time.perf_counter()
        ts_d = w_ = o_ = None  # drop the previous pass's tensors before allocating this pass's
        ts_d = ts_h.to(dev, non_blocking=True)
        w_ = rw.walk_triples(ts_d, idx_h.to(dev, non_blocking=True), tg_h.to(dev, non_blocking=True),
                             walk_length=L, padding_idx=pad, seed=k)
        o_ = rw.to_windows_triples(w_, W, n_ent, pad, ts_d, k)
        for h_, x_ in zip(host, o_):
            h_.copy_(x_, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t_0
    res = {"workload": wl["desc"], "walks": targets.numel(), "hops_per_walk": L, "window_size": W,
           "value": hops / (step_ms / 1e3), "unit": "steps/s", "ms_per_step": step_ms / steps, "steps": steps,
           "gpu_launches": launches,
           "windows_per_sec": n_win * steps / (win_ms / 1e3), "walk_hops_per_sec": hops / (walk_ms / 1e3),
           "window_ms": win_ms / steps, "walk_ms": walk_ms / steps,
           "e2e": {"value": targets.numel() * L * e2e_steps / dt, "unit": "steps/s",
                   "h2d_bytes_per_step": int((ts_h.numel() + idx_h.numel() + tg_h.numel()) * 8), "d2h_bytes_per_step": int(win_bytes)},
           "roofline": {"bound": "hbm", "kernel": "windows_kernel<triples>", "achieved": win_bytes * steps / (win_ms / 1e3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": win_bytes * steps / (win_ms / 1e3) / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_window": 504,
                        "note": "timed as the whole to_windows_triples call (one kernel launch + three torch.empty); "
                                "window outputs (%.1f GB) exceed the 126 MB L2" % (win_bytes / 1e9)}}
    del host
    return res


def run_c4(args, wl):
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    with ClockSampler(dev.index) as clocks:
        r = measure_c4(wl, dev, args.steps, args.warmup, args.e2e_steps)
    line = {"metric": "walk_steps_per_sec", "value": r["value"], "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "walks": r["walks"], "hops_per_walk": r["hops_per_walk"], "window_size": r["window_size"],
                       "l2_policy": r["roofline"]["note"]},
            "clocks": clocks.summary(), "gpu_launches": r["gpu_launches"], "e2e": r["e2e"],
            "windows_per_sec": r["windows_per_sec"], "walk_hops_per_sec": r["walk_hops_per_sec"], "roofline": r["roofline"],
            "cpu_baseline": None}
    emit(line)


# ----------------------------------------------------------------------------------------------
def source_sha16():
    """Hash of the sources the walk kernel is compiled from: profiles/traffic_bytes_per_launch.json is only quoted for the
    code it was measured on."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "torch_random_walk_b200", "csrc")
    for f in ("walk_csr.cu", "walk_csr.h", "member_table.cuh", "trw_common.cuh"):
        h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def roofline_of(kernel_ms, steps_per_launch, uniform, name, extra=None):
    peak, peak_src = measured_peaks()
    bps = BYTES_PER_STEP[uniform]
    achieved = steps_per_launch * bps / (kernel_ms / 1e3) / 1e9 if kernel_ms and kernel_ms > 0 else 0.0
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "kernel": name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_step": bps, "steps_per_launch": steps_per_launch,
         "peak_source": peak_src}
    if extra:
        r.update(extra)
    return r


def timed_calls(fn, n, barrier, first_seed):
    """n calls of fn(seed) between two barriers: (total ms by CUDA events, per-call ms)."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    barrier()
    ev[0].record()
    for k in range(n):
        fn(first_seed + k)
        ev[k + 1].record()
    barrier()
    return ev[0].elapsed_time(ev[-1]), [ev[k].elapsed_time(ev[k + 1]) for k in range(n)]


def max_over_ranks(x, dev, world):
    if world == 1:
        return float(x)
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_run(rep, targets, p, q, L, steps, layout, out, barrier, dev, world, seed0):
    """One start-node list sharded over the ranks of `rep` (dist.ReplicatedCsr): ms per step, max over ranks."""
    local, off, blocks, gids = rep.shard(targets, layout)
    o = out[: local.numel()]

    def fn(seed):
        rep.walk_local(local, p, q, L, seed, off, blocks, out=o)

    for k in range(3):
        fn(seed0 + k)
    total_ms, _ = timed_calls(fn, steps, barrier, seed0 + 100)
    return max_over_ranks(total_ms / steps, dev, world), local, gids, o, seed0 + 100 + steps - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("TRW_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary (first-order, c2, c4, c5) measurements")
    ap.add_argument("--stateless", action="store_true", help="graph cache off: the headline call rebuilds the graph-side data")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--option", action="append", default=[], help="library option name=value (experiments)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference" and wl.get("kind") != "triples":
        run_reference_arm(args, wl)
        return
    if wl.get("kind") == "triples":
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "the reference arm is implemented for the CSR walk workloads (c3, c2, c5)"})
            return
        run_c4(args, wl)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the walk has no CPU path (use --impl reference for the CPU arm)")
    import torch.distributed as dist

    from torch_random_walk_b200 import dist as trw_dist
    from torch_random_walk_b200 import native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    overrides = {}
    for kv in args.option:
        k, v = kv.split("=")
        native.set_option(k, int(v))
        overrides[k] = int(v)
    native.set_option("time_kernels", 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    p, q, L = wl["p"], wl["q"], wl["L"]
    uniform = (p == 1.0 and q == 1.0)
    kernel_name = "uniform_walk_kernel" if uniform else "node2vec_walk_kernel"
    W = max(args.warmup, 3)
    line_extra = {}

    # ---- inputs: rank 0 generates; at N > 1 one NCCL broadcast replicates and every rank prepares its replica
    #      (dist.ReplicatedCsr), all outside the timed region
    row_ptr, col_idx = build_graph(wl, dev) if rank == 0 else (None, None)
    rep = None
    if world > 1:
        t0 = time.time()
        rep = trw_dist.ReplicatedCsr(row_ptr, col_idx, src=0, device=dev)
        row_ptr, col_idx = rep.row_ptr, rep.col_idx
        torch.cuda.synchronize()
        log(f"rank {rank}: CSR replicated and prepared in {time.time() - t0:.2f}s")
    targets = start_nodes(row_ptr)
    n_nodes, nnz, n_walks = row_ptr.numel() - 1, col_idx.numel(), targets.numel()
    steps_per_call = n_walks * L
    out = torch.empty((n_walks, L + 1), dtype=torch.int64, device=dev)

    if world == 1:
        # ============================ one GPU: the drop-in call ============================
        cache_on = not args.stateless
        native.set_graph_cache(cache_on)

        def step(seed, cache=None):
            native.walk(row_ptr, col_idx, targets, p, q, L, seed, out=out, cache=cache_on if cache is None else cache)

        # warm-up: the cache grows with use (one-shot, prepared for keeps, triangle Blooms): call until it has settled
        warm = []
        for w in range(12):
            native.reset_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(100 + w)
            e1.record()
            torch.cuda.synchronize()
            st = native.graph_cache_state(dev) if cache_on else {"prepared": False, "blooms": False}
            warm.append({"ms": e0.elapsed_time(e1), "launches": native.launch_count(), "prepared": st["prepared"], "blooms": st["blooms"]})
            settled = (not cache_on) or (st["prepared"] and (st["blooms"] or uniform) and warm[-1]["launches"] <= 3)
            if w + 1 >= W and settled:
                break
        native.reset_launch_count()
        with ClockSampler(local_rank) as clocks:
            total_ms, per_step = timed_calls(step, args.steps, barrier, 1000)
            launches = native.launch_count()
            kms = []
            for k in range(min(args.steps, 3)):  # per-kernel event pairs, read outside the headline loop
                step(2000 + k)
                kms.append(native.last_kernel_ms())
        value = steps_per_call * args.steps / (total_ms / 1e3)
        ms_per_step = total_ms / args.steps
        kernel_ms = sum(k[1] for k in kms) / len(kms)
        build_ms = sum(k[0] for k in kms) / len(kms)
        clk = clocks.summary()
        log(f"drop-in call: {ms_per_step:.2f} ms/step {['%.2f' % x for x in per_step]}; walk kernel {kernel_ms:.2f} ms; warm-up {warm}")
        mode = ("rw.walk on HBM-resident tensors, library defaults: content-validated graph cache (every timed call checksums "
                "row_ptr and col_idx, then runs the walk kernel on the kept preparation)") if cache_on else \
               "rw.walk with the graph cache off: everything rebuilt inside every call"
        roofline = roofline_of(kernel_ms, steps_per_call, uniform, kernel_name, {
            "graph": "kept preparation with triangle Blooms" if cache_on else "per-call preparation",
            "table_build_ms": 0.0 if cache_on else build_ms,
            "whole_call_frac": value * BYTES_PER_STEP[uniform] / 1e9 / measured_peaks()[0],
            "note": "on this part a gather that misses L2 moves a full 128-byte line of HBM traffic (profiles/r01_summary.md): "
                    "measured line traffic, not the 32-byte-sector model behind algorithmic_bytes_per_step, is what the kernel is bound by"})

        # ---- the same loop with the graph cache off: the reference's stateless launcher, call for call
        stateless = None
        if cache_on:
            sl_steps = min(args.steps, 5)
            for k in range(2):
                step(3000 + k, cache=False)
            sl_total, _ = timed_calls(lambda sd: step(sd, cache=False), sl_steps, barrier, 3010)
            sl_build, sl_walk = native.last_kernel_ms()
            sl_ms = sl_total / sl_steps
            stateless = {"value": steps_per_call / (sl_ms / 1e3), "unit": "steps/s", "ms_per_step": sl_ms, "steps": sl_steps,
                         "graph_build_ms": sl_build, "kernel_ms": sl_walk,
                         "roofline_frac_kernel": roofline_of(sl_walk, steps_per_call, uniform, kernel_name)["frac"],
                         "note": "graph cache off: row index + membership table (+ edge records when the walk is long enough) rebuilt "
                                 "inside every call, as the reference's stateless launcher would; no checksum, no Blooms"}
            log(f"stateless: {sl_ms:.2f} ms/step (build {sl_build:.2f}, walk {sl_walk:.2f})")

        # ---- the explicit handle (native.prepare_csr): preparation cost, and the walk kernel without the checksum
        handle = None
        if cache_on:
            native.clear_graph_cache()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g = native.prepare_csr(row_ptr, col_idx, blooms=False)
            torch.cuda.synchronize()
            e0.record()
            g.add_blooms()
            e1.record()
            torch.cuda.synchronize()
            bloom_ms = e0.elapsed_time(e1)
            del g
            e0.record()
            g = native.prepare_csr(row_ptr, col_idx)
            e1.record()
            torch.cuda.synchronize()
            prep_ms = e0.elapsed_time(e1)
            for k in range(2):
                g.walk(targets, p, q, L, 4000 + k, out=out)
            h_total, _ = timed_calls(lambda sd: g.walk(targets, p, q, L, sd, out=out), min(args.steps, 5), barrier, 4010)
            h_ms = h_total / min(args.steps, 5)
            handle = {"value": steps_per_call / (h_ms / 1e3), "unit": "steps/s", "ms_per_step": h_ms, "prepare_ms": prep_ms,
                      "of_which_triangle_blooms_ms": bloom_ms, "info": g.info(),
                      "workspace_bytes": int(g.workspace.numel()),
                      "note": "native.prepare_csr handle: no checksum (the caller promises not to modify the tensors); prepare_ms "
                              "includes the workspace allocation"}
            log(f"explicit handle: {h_ms:.2f} ms/step, prepare {prep_ms:.1f} ms (Blooms {bloom_ms:.1f} ms)")
            del g
        line_extra.update(stateless=stateless, prepared_handle=handle,
                          graph_cache={"enabled": cache_on, "warmup_calls": warm,
                                       "note": "identified by content (sizes + 64-bit checksum computed on the device inside every call); "
                                               "call 1 one-shot, call 2 prepares for keeps, after two more hits the triangle Blooms are added"})
        scaling, parallelism = "strong", "one GPU"
        walks_per_gpu = n_walks
    else:
        # ============================ N GPUs: one start-node list sharded over the ranks ============================
        with ClockSampler(local_rank) as clocks:
            ms_per_step, local, gids, o_local, last_seed = sharded_run(rep, targets, p, q, L, args.steps, "block_cyclic", out, barrier, dev,
                                                                       world, 1000)
        kernel_ms = native.last_kernel_ms()[1]
        launches = args.steps
        value = steps_per_call / (ms_per_step / 1e3)
        clk = clocks.summary()
        walks_per_gpu = local.numel()
        log(f"rank {rank}: sharded (block-cyclic) {ms_per_step:.3f} ms/step, {walks_per_gpu} walks, kernel {kernel_ms:.3f} ms")
        # identity: the shards' digests must add up to the digest of ONE call over the whole list (rank 0 makes it)
        digest = trw_dist.walk_digest(o_local, gids).reshape(1)
        dist.all_reduce(digest, op=dist.ReduceOp.SUM)
        equal = None
        if rank == 0:
            single = rep.graph.walk(targets, p, q, L, last_seed, out=out)
            equal = bool(int(trw_dist.walk_digest(single, torch.arange(n_walks, device=dev))) == int(digest.item()))
            log(f"shards equal the single call (digest of {n_walks} rows): {equal}")
        flag = torch.tensor([1 if equal else 0], dtype=torch.int64, device=dev)
        dist.broadcast(flag, 0)
        # contiguous shards (the simplest drop-in) and the replicated list (weak scaling) beside it
        c_ms, c_local, c_gids, c_out, c_seed = sharded_run(rep, targets, p, q, L, min(args.steps, 5), "contiguous", out, barrier, dev, world, 5000)
        c_digest = trw_dist.walk_digest(c_out, c_gids).reshape(1)
        dist.all_reduce(c_digest, op=dist.ReduceOp.SUM)
        if rank == 0:
            single = rep.graph.walk(targets, p, q, L, c_seed, out=out)
            equal = equal and bool(int(trw_dist.walk_digest(single, torch.arange(n_walks, device=dev))) == int(c_digest.item()))

        def weak(seed):
            rep.graph.walk(targets, p, q, L, seed, walk_id_offset=rank * n_walks, out=out)

        for k in range(2):
            weak(6000 + k)
        w_total, _ = timed_calls(weak, min(args.steps, 5), barrier, 6010)
        w_ms = max_over_ranks(w_total / min(args.steps, 5), dev, world)
        full_kernel_ms = native.last_kernel_ms()[1]
        roofline = roofline_of(full_kernel_ms, steps_per_call, uniform, kernel_name, {
            "graph": "kept preparation with triangle Blooms (dist.ReplicatedCsr)",
            "note": "kernel timed on the full start-node list on this rank (the sharded launches run the same kernel on 1/N of it: "
                    f"{kernel_ms:.3f} ms)"})
        line_extra.update(
            shards_equal_single=equal,
            sharded_contiguous={"value": steps_per_call / (c_ms / 1e3), "unit": "steps/s", "ms_per_step": c_ms, "walks_per_gpu": c_local.numel()},
            replicated_list={"value": world * steps_per_call / (w_ms / 1e3), "unit": "steps/s", "ms_per_step": w_ms, "scaling": "weak",
                             "note": "every rank walks the whole list under its own global walk ids"},
            shard_kernel_ms=kernel_ms)
        scaling = "strong"
        parallelism = (f"CSR replicated x{world} by one NCCL broadcast and prepared per rank (dist.ReplicatedCsr); ONE start-node list "
                       f"sharded block-cyclically (blocks of {trw_dist.DEFAULT_BLOCK} walks), no data-path collective")
        mode = "dist.ReplicatedCsr.walk_local on the rank's shard (explicit kept graph: no per-call checksum)"

    traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            tf = json.load(open(traffic_file))
            if tf.get("source_sha16") == source_sha16():
                roofline["traffic"] = tf.get(args.workload)
                roofline["traffic_source"] = tf.get("how")
            else:
                roofline["traffic_note"] = "profiles/traffic_bytes_per_launch.json was measured on other sources; not quoted"
        except Exception:
            pass

    # ---- e2e: host buffers in, host walks out, through the public host entry (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        try:
            e2e = measure_e2e(native, trw_dist, rep, row_ptr, col_idx, targets, p, q, L, args.e2e_steps, dev, local_rank, rank, world, barrier)
        except Exception as exc:  # noqa: BLE001
            log(f"e2e failed: {exc!r}")
            e2e = {"value": None, "unit": "steps/s", "error": repr(exc)}

    # ---- validity of the last output (sampled: every transition must be an edge) + CPU baseline (rank 0, N=1)
    cpu_baseline = None
    valid = None
    if rank == 0:
        from torch_random_walk_b200 import rmat

        check = native.walk(row_ptr, col_idx, targets[:: max(1, n_walks // 2048)][:2048].contiguous(), p, q, L, 77, cache=False)
        valid = bool(rmat.transitions_are_edges(row_ptr, col_idx, check)) and bool((check[:, 0] == targets[:: max(1, n_walks // 2048)][:2048]).all())
        log(f"sampled validity check: {valid}")
        if world == 1 and not args.no_cpu_baseline:
            rp_c, ci_c = row_ptr.cpu(), col_idx.cpu()
            fn, kind = cpu_walk_fn()
            tg = cpu_sample_targets(targets.cpu())
            res = {}
            for threads in sorted({1, os.cpu_count() or 1}):
                sps, n, dt = time_cpu(fn, rp_c, ci_c, tg, wl, 8.0, threads)
                res[threads] = sps
                log(f"cpu baseline ({kind}) threads={threads}: {sps:.3e} steps/s on {n} walks in {dt:.1f}s")
            best = max(res, key=res.get)
            cpu_baseline = {"value": res[best], "unit": "steps/s", "cores": best, "kind": kind,
                            "host_cores": os.cpu_count(),
                            "sample": f"{kind} csrc/cpu walk on a random sample of degree>0 start nodes, ~8 s per thread "
                                      f"count, same graph/p/q/L; steps/s by threads: { {k: round(v) for k, v in res.items()} }"}
            del rp_c, ci_c

    # ---- secondary lines: the first-order kernel and configs[1] (c2), configs[3] (c4) at N=1; configs[4] (c5) at N=8
    others = None
    if not args.no_extras:
        try:
            others = measure_others(args, native, trw_dist, rep, row_ptr, col_idx, targets, out, L, dev, rank, world, barrier)
        except Exception as exc:  # noqa: BLE001
            log(f"extras failed: {exc!r}")
            others = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "walk_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "p": p, "q": q, "walk_length": L, "n_nodes": n_nodes, "nnz": nnz,
                       "walks": n_walks, "walks_per_gpu": walks_per_gpu, "start_nodes": "all nodes with degree>0, one walk each",
                       "call": mode, "parallelism": parallelism,
                       "l2_policy": "inputs (CSR %.1f GB) and output (%.1f GB) exceed the 126 MB L2; no flush needed"
                                    % ((nnz + n_nodes) * 8 / 1e9, n_walks * (L + 1) * 8 / 1e9),
                       "options": overrides},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "output_valid": valid, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "other_workloads": others,
        }
        line.update(line_extra)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(native, trw_dist, rep, row_ptr, col_idx, targets, p, q, L, e2e_steps, dev, local_rank, rank, world, barrier):
    """The same metric through the host entry: pinned host tensors in, pinned host walks out, copies inside the timed
    region.  N = 1: the whole list through native.walk_host.  N > 1: rank 0 holds the host graph; every rank walks its
    block-cyclic shard of the host start-node list and brings its shard of the walks back to its own host buffer."""
    n_walks = targets.numel()
    steps_per_call = n_walks * L
    if world == 1:
        rp_h, ci_h, tg_h = row_ptr.cpu().pin_memory(), col_idx.cpu().pin_memory(), targets.cpu().pin_memory()
        out_h = torch.empty((n_walks, L + 1), dtype=torch.int64, pin_memory=True)
        # a caller that comes back with the same host arrays finds the device replica kept (content re-checked on every
        # call); like the device-side cache its preparation grows with use, so warm up until it has settled
        fresh_ms = []
        native.set_option("host_keep_graph", 0)
        for k in range(2):
            t0 = time.perf_counter()
            native.walk_host(rp_h, ci_h, tg_h, p, q, L, 5 + k, device=local_rank, out=out_h)
            fresh_ms.append((time.perf_counter() - t0) * 1e3)
        native.set_option("host_keep_graph", 1)
        warm_ms = []
        for k in range(5):
            t0 = time.perf_counter()
            native.walk_host(rp_h, ci_h, tg_h, p, q, L, 10 + k, device=local_rank, out=out_h)
            warm_ms.append((time.perf_counter() - t0) * 1e3)
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            native.walk_host(rp_h, ci_h, tg_h, p, q, L, 3000 + k, device=local_rank, out=out_h)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        graph_bytes = int((rp_h.numel() + ci_h.numel()) * 8)
        info = native.host_replica_info(local_rank)  # of the last timed call
        # h2d / d2h: what crossed PCIe in the last timed call as the library counted its copies -- upwards the start nodes plus
        # the copy engine's share of the content check, downwards the walks at 4 bytes per entry for the chunks that travelled
        # packed and 8 for the others (the split is decided call by call); the tensors themselves are listed beside them
        e2e = {"value": steps_per_call * e2e_steps / dt, "unit": "steps/s", "ms_per_step": dt / e2e_steps * 1e3,
               "h2d_bytes_per_step": info["h2d_bytes"], "d2h_bytes_per_step": info["d2h_bytes"], "steps": e2e_steps,
               "input_tensor_bytes_per_step": int(tg_h.numel() * 8), "result_tensor_bytes_per_step": int(out_h.numel() * 8),
               "graph_bytes_checksummed_per_step": graph_bytes,
               "replica_after_timed_calls": info,  # last_call must read "kept replica validated"
               "warmup_call_ms": warm_ms,
               "fresh_upload": {"value": steps_per_call / (min(fresh_ms) / 1e3), "unit": "steps/s", "ms_per_step": min(fresh_ms),
                                "h2d_bytes_per_step": graph_bytes + int(tg_h.numel() * 8),
                                "note": "option host_keep_graph=0: the graph crosses PCIe and is prepared inside every call"},
               "api": "native.walk_host -> trw_walk_csr_host (pinned host tensors in, pinned host walks out).  The device replica of "
                      "the graph is kept between calls with the same host arrays; every call walks on it at once while row_ptr and "
                      "col_idx are checksummed against it (copy engine re-reading part of the pinned arrays, host threads the rest; a "
                      "mismatch uploads afresh and walks again), so per step only the start nodes go up and the walks come down -- as "
                      "uint32 widened by the host threads for as many chunks as they keep up with, else as int64"}
        # the device path and the host path must agree on the result
        check = native.walk(row_ptr, col_idx, targets[:4096].contiguous(), p, q, L, 3000 + e2e_steps - 1, cache=False)
        assert torch.equal(check.cpu(), out_h[:4096]), "host path and device path disagree"
        return e2e
    import torch.distributed as dist

    tg_h = targets.cpu().pin_memory()
    idx = trw_dist.block_cyclic_indices(n_walks, rank, world)
    local_h = tg_h[idx].contiguous().pin_memory()
    out_h = torch.empty((local_h.numel(), L + 1), dtype=torch.int64, pin_memory=True)
    blocks = (trw_dist.DEFAULT_BLOCK, trw_dist.DEFAULT_BLOCK * world)

    def one(seed):
        rep.walk_local_to_host(local_h, p, q, L, seed, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_h)

    one(5)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        one(3000 + k)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    info = native.host_replica_info(local_rank)  # this rank's last timed call: bytes as the library counted its copies
    moved = torch.tensor([info["h2d_bytes"], info["d2h_bytes"]], dtype=torch.int64, device=dev)
    dist.all_reduce(moved, op=dist.ReduceOp.SUM)
    return {"value": steps_per_call * e2e_steps / dt, "unit": "steps/s", "ms_per_step": dt / e2e_steps * 1e3,
            "h2d_bytes_per_step": int(moved[0].item()), "d2h_bytes_per_step": int(moved[1].item()), "steps": e2e_steps,
            "input_tensor_bytes_per_step": int(n_walks * 8), "result_tensor_bytes_per_step": int(n_walks * (L + 1) * 8),
            "api": "per rank: dist.ReplicatedCsr.walk_local_to_host -> trw_walk_csr_to_host (pinned host start nodes of its shard in, "
                   "walks into its pinned host buffer, chunks walked and copied back in a pipeline); the graph replica was broadcast "
                   "over NVLink once, outside the timed region (bytes are totals over the ranks)"}


def quick_walks(native, rp_, ci_, tg_, p_, q_, L_, out_, peak):
    """Drop-in (validated cache) and stateless timings of one more workload on this GPU."""
    res_ = {}
    native.clear_graph_cache()
    for mode in (True, False):
        for k in range(6 if mode else 2):
            native.walk(rp_, ci_, tg_, p_, q_, L_, 50 + k, out=out_, cache=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3):
            native.walk(rp_, ci_, tg_, p_, q_, L_, 60 + k, out=out_, cache=mode)
        e1.record()
        torch.cuda.synchronize()
        res_[mode] = (e0.elapsed_time(e1) / 3,) + tuple(native.last_kernel_ms())
    native.clear_graph_cache()
    ms, b_, w_ = res_[True]
    first = (p_ == 1.0 and q_ == 1.0)
    return {"steps_per_s": tg_.numel() * L_ / (ms / 1e3), "ms_per_call": ms, "kernel_ms": w_,
            "stateless_ms_per_call": res_[False][0], "stateless_graph_build_ms": res_[False][1], "stateless_kernel_ms": res_[False][2],
            "n_nodes": rp_.numel() - 1, "nnz": ci_.numel(), "walks": tg_.numel(), "p": p_, "q": q_, "walk_length": L_,
            "roofline_frac_kernel": tg_.numel() * L_ * BYTES_PER_STEP[first] / (w_ / 1e3) / 1e9 / peak if w_ > 0 else None}


def measure_others(args, native, trw_dist, rep, row_ptr, col_idx, targets, out, L, dev, rank, world, barrier):
    peak = measured_peaks()[0]
    others = {}
    if world == 1:
        others["c3_first_order_p1_q1"] = quick_walks(native, row_ptr, col_idx, targets, 1.0, 1.0, L, out, peak)
        log(f"extra first-order: {others['c3_first_order_p1_q1']}")
        if args.workload != "c2":
            w2 = WORKLOADS["c2"]
            rp2, ci2 = build_graph(w2, dev)
            tg2 = start_nodes(rp2)
            out2 = out[: tg2.numel()] if (tg2.numel() <= out.size(0) and w2["L"] == L) else \
                torch.empty((tg2.numel(), w2["L"] + 1), dtype=torch.int64, device=dev)
            others["c2_products_shaped_p0.5_q2"] = quick_walks(native, rp2, ci2, tg2, w2["p"], w2["q"], w2["L"], out2, peak)
            del out2
            log(f"extra c2: {others['c2_products_shaped_p0.5_q2']}")
            del rp2, ci2, tg2
        torch.cuda.empty_cache()
        others["c4_triples_and_windows"] = measure_c4(WORKLOADS["c4"], dev, 5, 3, 2)
        log(f"extra c4: {others['c4_triples_and_windows']}")
    elif world >= 8 and args.workload == "c3":
        # configs[4]: Friendster-shaped R-MAT replicated per GPU, node2vec p=0.25 q=4 L=40, one list sharded over the ranks
        import torch.distributed as dist

        w5 = WORKLOADS["c5"]
        rp5, ci5 = build_graph(w5, dev) if rank == 0 else (None, None)
        t0 = time.time()
        rep5 = trw_dist.ReplicatedCsr(rp5, ci5, src=0, device=dev)
        torch.cuda.synchronize()
        prep_s = time.time() - t0
        tg5 = start_nodes(rep5.row_ptr)
        out5 = torch.empty((trw_dist.block_cyclic_count(tg5.numel(), rank, world) + 1, w5["L"] + 1), dtype=torch.int64, device=dev)
        ms5, local5, gids5, o5, _ = sharded_run(rep5, tg5, w5["p"], w5["q"], w5["L"], 3, "block_cyclic", out5, barrier, dev, world, 9000)
        k5 = native.last_kernel_ms()[1]
        others["c5_friendster_shaped_p0.25_q4"] = {
            "value": tg5.numel() * w5["L"] / (ms5 / 1e3), "unit": "steps/s", "ms_per_step": ms5, "n_gpus": world, "scaling": "strong",
            "n_nodes": rep5.row_ptr.numel() - 1, "nnz": rep5.col_idx.numel(), "walks": tg5.numel(), "walks_per_gpu": local5.numel(),
            "p": w5["p"], "q": w5["q"], "walk_length": w5["L"], "shard_kernel_ms": k5, "replicate_and_prepare_s": prep_s,
            "roofline_frac_all_gpus": tg5.numel() * w5["L"] * BYTES_PER_STEP[False] / (ms5 / 1e3) / 1e9 / (peak * world)}
        log(f"extra c5: {others['c5_friendster_shaped_p0.25_q4']}")
        del rep5, rp5, ci5, out5
        dist.barrier()
    return others


if __name__ == "__main__":
    main()
