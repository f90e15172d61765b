#!/usr/bin/env python
"""bench.py -- walk-steps/s of the hot path (rw.walk, node2vec) on synthetic R-MAT graphs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c5|c1|tiny]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's csrc/cpu path on the host cores

A "step" is one rw.walk call over every start node (all nodes with degree > 0), made the way a user
of the drop-in API makes it: the same CSR tensors, resident in HBM, every call.  From the second call
on the library keeps the graph-side preparation (membership table, edge records) of those tensors
(native.walk's graph cache, the library default), so a timed call is the walk kernel and its output
write; the warm-up calls pay the preparation once and its cost is reported (`graph_prepare_ms`).
`stateless` in the same line is the same loop with the cache off: everything rebuilt inside every call,
like the reference's stateless launcher.  Default workload (c3) is
BASELINE.json configs[2] -- R-MAT scale 24 (16.8 M nodes, ~2^29 CSR entries), p=1 q=0.5,
walk_length=80 -- the configuration the north-star target is quoted on.  With N GPUs every rank
holds a replica of the CSR (one NCCL broadcast, outside the timed region) and walks the full
start-node list under its own global walk ids: per-GPU work is fixed ("weak"), value = total steps
of all ranks / max-over-ranks time.

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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": wl["desc"], "p": wl["p"], "q": wl["q"], "walk_length": wl["L"], "n_nodes": row_ptr.numel() - 1,
                   "nnz": col_idx.numel()},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": best_threads, "kind": kind,
                         "sample": f"random sample of degree>0 start nodes sized to ~{per_step_budget / 2:.0f}s per timing; "
                                   f"steps/s by thread count: { {k: round(v) for k, v in results.items()} }"},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_c4(args, wl):
    """configs[3]: knowledge-graph triple walks + window generation.  A step = rw.walk_triples over
    10 walks per entity followed by rw.to_windows_triples on the result; value = hops/s of the whole
    step, windows/s reported beside it.  Single GPU (the workload is 145k walks)."""
    from torch_random_walk_b200 import native, rmat, rw

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if int(os.environ.get("RANK", "0")) != 0:
        return
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
        for k in range(max(args.warmup, 3)):
            out_ = None
            out_ = fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            out_ = None
            out_ = fn(100 + k)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), out_

    native.reset_launch_count()

    with ClockSampler(dev.index) as clocks:
        step_ms, (walks, outs) = timed_loop(step)                                    # the whole step, K times
        outs = None
        walk_ms, walks = timed_loop(lambda sd: rw.walk_triples(ts, index, targets, walk_length=L, padding_idx=pad, seed=sd))
        win_ms, outs = timed_loop(lambda sd: rw.to_windows_triples(walks, W, n_ent, pad, ts, sd))
    # three loops (the step and its two halves), each with its warm-up calls: count the timed step loop's share
    launches = 2 * args.steps
    total_ms = step_ms
    hops = targets.numel() * L * args.steps
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
    for k in range(args.e2e_steps + 1):  # the first pass warms the allocator and is not timed
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
    line = {"metric": "walk_steps_per_sec", "value": hops / (total_ms / 1e3), "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "walks": targets.numel(), "hops_per_walk": L, "window_size": W,
                       "l2_policy": "window outputs (%.1f GB) exceed the 126 MB L2" % (win_bytes / 1e9)},
            "clocks": clocks.summary(), "gpu_launches": launches,
            "e2e": {"value": targets.numel() * L * args.e2e_steps / dt, "unit": "steps/s",
                    "h2d_bytes_per_step": int((ts_h.numel() + idx_h.numel() + tg_h.numel()) * 8), "d2h_bytes_per_step": int(win_bytes)},
            "windows_per_sec": n_win * args.steps / (win_ms / 1e3), "walk_hops_per_sec": hops / (walk_ms / 1e3),
            "roofline": {"bound": "hbm", "kernel": "windows_kernel<triples>", "achieved": win_bytes * args.steps / (win_ms / 1e3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": win_bytes * args.steps / (win_ms / 1e3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_window": 504,
                         "note": "timed as the whole to_windows_triples call (one kernel launch + three torch.empty)"},
            "cpu_baseline": None}
    del host
    emit(line)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("TRW_BENCH_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary (first-order, c2) measurements")
    ap.add_argument("--stateless", action="store_true", help="graph cache off: every call rebuilds the graph-side data")
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

    # ---- inputs: rank 0 generates, one NCCL broadcast replicates (outside the timed region)
    if rank == 0:
        row_ptr, col_idx = build_graph(wl, dev)
    else:
        row_ptr = col_idx = None
    if world > 1:
        t0 = time.time()
        row_ptr, col_idx = trw_dist.replicate_csr(row_ptr, col_idx, src=0, device=dev)
        torch.cuda.synchronize()
        log(f"rank {rank}: CSR replicated in {time.time() - t0:.2f}s")
    targets = start_nodes(row_ptr)
    n_nodes, nnz, n_walks = row_ptr.numel() - 1, col_idx.numel(), targets.numel()
    p, q, L = wl["p"], wl["q"], wl["L"]
    uniform = (p == 1.0 and q == 1.0)
    offset = rank * n_walks  # global walk ids of this rank's replica of the start-node list
    out = torch.empty((n_walks, L + 1), dtype=torch.int64, device=dev)
    steps_per_call = n_walks * L

    cache_on = not args.stateless

    def step(seed, cache=None):
        native.walk(row_ptr, col_idx, targets, p, q, L, seed, walk_id_offset=offset, out=out,
                    cache=cache_on if cache is None else cache)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    native.set_option("time_kernels", 1)
    native.set_graph_cache(cache_on)
    first_ms, prepare_ms = [], None
    for w in range(max(args.warmup, 3)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(100 + w)
        e1.record()
        torch.cuda.synchronize()
        first_ms.append(e0.elapsed_time(e1))
        if w == 1 and cache_on:
            prepare_ms = native.last_kernel_ms()[0]  # the second call with the same tensors prepares the graph for keeps
    barrier()
    native.reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    build_ms, walk_ms = [], []
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev[0].record()
        for k in range(args.steps):
            step(1000 + k)
            ev[k + 1].record()
        barrier()
        launches = native.launch_count()
        # one more, un-timed-by-the-headline call to read the per-kernel event pairs without perturbing the loop
        for k in range(min(args.steps, 3)):
            step(2000 + k)
            b, w_ = native.last_kernel_ms()
            build_ms.append(b)
            walk_ms.append(w_)
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * steps_per_call * args.steps / (total_ms / 1e3)
    ms_per_step = total_ms / args.steps
    clk = clocks.summary()
    log(f"rank {rank}: {ms_per_step:.2f} ms/step, per-step {['%.2f' % x for x in per_step]}, build {build_ms}, walk {walk_ms}")

    # ---- the same loop with the graph cache off (everything rebuilt inside every call), rank 0 only reports it
    stateless = None
    if cache_on:
        sl_steps = min(args.steps, 5)
        step(3000, cache=False)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for k in range(sl_steps):
            step(3001 + k, cache=False)
        s1.record()
        barrier()
        sl_ms = s0.elapsed_time(s1) / sl_steps
        sl_build, sl_walk = native.last_kernel_ms()
        tt = torch.tensor([sl_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sl_ms = float(tt.item())
        stateless = {"value": world * steps_per_call / (sl_ms / 1e3), "unit": "steps/s", "ms_per_step": sl_ms, "steps": sl_steps,
                     "graph_build_ms": sl_build, "kernel_ms": sl_walk,
                     "note": "graph cache off: row index + membership table (+ edge records when the walk is long enough) "
                             "rebuilt inside every call, as the reference's stateless launcher would"}
        log(f"rank {rank}: stateless {sl_ms:.2f} ms/step (build {sl_build:.2f}, walk {sl_walk:.2f})")

    # ---- N > 1: the same start-node list sharded over the ranks (fixed total work)
    strong = None
    if world > 1:
        lo, hi = trw_dist.shard_bounds(n_walks, rank, world)
        shard, out_shard = targets[lo:hi].contiguous(), out[: hi - lo]

        def shard_step(seed):
            native.walk(row_ptr, col_idx, shard, p, q, L, seed, walk_id_offset=lo, out=out_shard, cache=cache_on)

        for k in range(3):
            shard_step(4000 + k)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(args.steps):
            shard_step(4100 + k)
        g1.record()
        barrier()
        tt = torch.tensor([g0.elapsed_time(g1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong_ms = float(tt.item())
        strong = {"value": steps_per_call / (strong_ms / 1e3), "unit": "steps/s", "ms_per_step": strong_ms,
                  "walks_per_gpu": hi - lo, "scaling": "strong",
                  "note": "one copy of the start-node list sharded contiguously over the ranks (global walk ids: the "
                          "concatenated output equals the single-GPU call); max over ranks"}
        log(f"rank {rank}: strong scaling {strong_ms:.2f} ms/step for the sharded list")

    # ---- roofline of the dominant kernel (the walk kernel), timed with its own CUDA event pair
    peak, peak_src = measured_peaks()
    kernel_ms = sum(walk_ms) / max(len(walk_ms), 1)
    bps = BYTES_PER_STEP[uniform]
    achieved = steps_per_call * bps / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "uniform_walk_kernel" if uniform else "node2vec_walk_kernel",
                "kernel_ms": kernel_ms, "table_build_ms": sum(build_ms) / max(len(build_ms), 1),
                "algorithmic_bytes_per_step": bps, "steps_per_launch": steps_per_call, "peak_source": peak_src,
                "whole_call_frac": (value / world) * bps / 1e9 / peak,
                "note": "on this part a gather that misses L2 moves a full 128-byte line of HBM traffic (profiles/r01_summary.md): "
                        "measured line traffic, not the 32-byte-sector model behind algorithmic_bytes_per_step, is what the kernel is bound by"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(args.workload)
        except Exception:
            pass

    # ---- e2e: host buffers in, host walks out, through the public host entry (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        try:
            rp_h, ci_h, tg_h = row_ptr.cpu().pin_memory(), col_idx.cpu().pin_memory(), targets.cpu().pin_memory()
            out_h = torch.empty((n_walks, L + 1), dtype=torch.int64, pin_memory=True)
            native.walk_host(rp_h, ci_h, tg_h, p, q, L, 5, device=local_rank, walk_id_offset=offset, out=out_h)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for k in range(args.e2e_steps):
                native.walk_host(rp_h, ci_h, tg_h, p, q, L, 3000 + k, device=local_rank, walk_id_offset=offset, out=out_h)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": world * steps_per_call * args.e2e_steps / dt, "unit": "steps/s",
                   "h2d_bytes_per_step": int((rp_h.numel() + ci_h.numel() + tg_h.numel()) * 8),
                   "d2h_bytes_per_step": int(out_h.numel() * 8), "steps": args.e2e_steps,
                   "api": "native.walk_host -> trw_walk_csr_host (pinned host tensors in, pinned host walks out); h2d/d2h bytes are "
                          "the caller's int64 tensors -- the library sends col_idx and fetches the walks as uint32 when every id fits "
                          "and it has >= 12 host threads to convert with (option host_compress)"}
            # the device path and the host path must agree on the result
            check = native.walk(row_ptr, col_idx, targets[:4096].contiguous(), p, q, L, 3000 + args.e2e_steps - 1,
                                walk_id_offset=offset)
            assert torch.equal(check.cpu(), out_h[:4096]), "host path and device path disagree"
            del rp_h, ci_h, tg_h, out_h
        except Exception as exc:  # noqa: BLE001
            log(f"e2e failed: {exc!r}")
            e2e = {"value": None, "unit": "steps/s", "error": repr(exc)}

    # ---- validity of the timed output (sampled: every transition must be an edge) + CPU baseline (rank 0, N=1)
    cpu_baseline = None
    valid = None
    if rank == 0:
        from torch_random_walk_b200 import rmat

        sample = out[:: max(1, n_walks // 2048)][:2048]
        valid = bool(rmat.transitions_are_edges(row_ptr, col_idx, sample)) and bool((sample[:, 0] == targets[:: max(1, n_walks // 2048)][:2048]).all())
        log(f"sampled validity check of the timed output: {valid}")
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

    # ---- secondary lines (N=1 only): the first-order kernel on the same graph and BASELINE configs[1] (c2)
    others = None
    if rank == 0 and world == 1 and not args.no_extras:
        others = {}

        def quick(name, rp_, ci_, tg_, p_, q_, L_):
            o_ = out[: tg_.numel()] if (L_ == L and tg_.numel() <= n_walks) else torch.empty((tg_.numel(), L_ + 1), dtype=torch.int64, device=dev)
            res_ = {}
            for mode in ((True, False) if cache_on else (False,)):
                for k in range(3):
                    native.walk(rp_, ci_, tg_, p_, q_, L_, 50 + k, out=o_, cache=mode)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(3):
                    native.walk(rp_, ci_, tg_, p_, q_, L_, 60 + k, out=o_, cache=mode)
                e1.record()
                torch.cuda.synchronize()
                res_[mode] = (e0.elapsed_time(e1) / 3,) + tuple(native.last_kernel_ms())
            ms, b_, w_ = res_[cache_on]
            first = (p_ == 1.0 and q_ == 1.0)
            sps = tg_.numel() * L_ / (ms / 1e3)
            others[name] = {"steps_per_s": sps, "ms_per_call": ms, "kernel_ms": w_, "table_build_ms": b_,
                            "stateless_ms_per_call": res_[False][0] if False in res_ else None,
                            "stateless_graph_build_ms": res_[False][1] if False in res_ else None,
                            "n_nodes": rp_.numel() - 1, "nnz": ci_.numel(), "walks": tg_.numel(), "p": p_, "q": q_, "walk_length": L_,
                            "roofline_frac_kernel": tg_.numel() * L_ * BYTES_PER_STEP[first] / (w_ / 1e3) / 1e9 / peak if w_ > 0 else None}
            log(f"extra {name}: {sps:.3e} steps/s ({ms:.2f} ms/call, kernel {w_:.2f} ms, build {b_:.2f} ms)")

        try:
            quick("c3_first_order_p1_q1", row_ptr, col_idx, targets, 1.0, 1.0, L)
            if args.workload != "c2":
                w2 = WORKLOADS["c2"]
                rp2, ci2 = build_graph(w2, dev)
                quick("c2_products_shaped_p0.5_q2", rp2, ci2, start_nodes(rp2), w2["p"], w2["q"], w2["L"])
                del rp2, ci2
        except Exception as exc:  # noqa: BLE001
            log(f"extras failed: {exc!r}")
            others["error"] = repr(exc)

    if rank == 0:
        line = {
            "metric": "walk_steps_per_sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": wl["desc"], "p": p, "q": q, "walk_length": L, "n_nodes": n_nodes, "nnz": nnz,
                       "walks_per_gpu": n_walks, "start_nodes": "all nodes with degree>0, one walk each per GPU",
                       "parallelism": f"replicated CSR x{world}, start nodes per rank, no data-path collective",
                       "l2_policy": "inputs (CSR %.1f GB) and output (%.1f GB) exceed the 126 MB L2; no flush needed"
                                    % ((nnz + n_nodes) * 8 / 1e9, n_walks * (L + 1) * 8 / 1e9),
                       "graph_cache": cache_on, "options": overrides},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches, "output_valid": valid, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "graph_cache": {"enabled": cache_on, "graph_prepare_ms": prepare_ms, "warmup_call_ms": first_ms,
                            "note": "library default: the second rw.walk call with the same CSR tensors keeps their graph-side "
                                    "preparation; timed calls reuse it (a modified tensor is detected and rebuilt)"},
            "stateless": stateless, "sharded_start_nodes": strong, "other_workloads": others,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
