"""What bounds the end-to-end path when several ranks share one host (measurement tooling; run under torchrun):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_ranks_probe.py [--scale=24]
1. pinned-memory copies with 1, 2, 4, ... ranks active at once: per-rank and aggregate GB/s in each direction -- the
   platform's ceiling for any host path;
2. the sharded end-to-end call (dist.ReplicatedCsr.walk_local_to_host) under different download formats, host thread
   counts and chunk sizes, max over ranks.
Results: gpurun_out/<out>.json (rank 0).
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import dist as trw_dist  # noqa: E402
from torch_random_walk_b200 import native, rmat  # noqa: E402


def arg(name, default):
    for a in sys.argv[1:]:
        if a.startswith(f"--{name}="):
            return a.split("=", 1)[1]
    return default


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    res = {"world": world, "host_cores": os.cpu_count(), "allowed_cpus": len(os.sched_getaffinity(0))}

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- 1. copies, k ranks at once
    n = (1 << 30) // 8
    d_buf = torch.empty(n, dtype=torch.int64, device=dev)
    h_buf = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_buf.zero_()
    copies = {}
    k = 1
    while k <= world:
        for name, (dst, src) in (("d2h", (h_buf, d_buf)), ("h2d", (d_buf, h_buf))):
            barrier()
            t0 = time.perf_counter()
            if rank < k:
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0 if rank < k else 0.0
            dt = max_over_ranks(dt)
            copies[f"{name}_{k}_ranks"] = {"per_rank_GBps": round(4 * n * 8 / dt / 1e9, 1), "aggregate_GBps": round(k * 4 * n * 8 / dt / 1e9, 1)}
            if rank == 0:
                print(name, k, copies[f"{name}_{k}_ranks"], flush=True)
        k *= 2
    res["copies"] = copies
    del d_buf, h_buf

    # ---- 2. the sharded end-to-end call
    scale = int(arg("scale", "24"))
    p, q, L = 1.0, 0.5, 80
    if rank == 0:
        rp, ci = rmat.rmat_csr(scale, 16, device=dev)
    else:
        rp = ci = None
    rep = trw_dist.ReplicatedCsr(rp, ci, src=0, device=dev)
    deg = rep.row_ptr[1:] - rep.row_ptr[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    n_walks = targets.numel()
    tg_h = targets.cpu()
    idx = trw_dist.block_cyclic_indices(n_walks, rank, world)
    local_h = tg_h[idx].contiguous().pin_memory()
    out_h = torch.empty((local_h.numel(), L + 1), dtype=torch.int64, pin_memory=True)
    blocks = (trw_dist.DEFAULT_BLOCK, trw_dist.DEFAULT_BLOCK * world)
    steps = n_walks * L
    runs = {}
    variants = arg("variants", "host_packed_share:0,host_packed_share:-1,host_packed_share:8,"
                               "host_packed_share:0+host_chunk_walks:131072,host_packed_share:-1+host_chunk_walks:131072,"
                               "host_packed_share:8+host_chunk_walks:131072,host_packed_share:-1+host_chunk_walks:131072+host_threads:8,"
                               "host_packed_share:-1+host_chunk_walks:262144")
    for variant in variants.split(","):
        opts = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in variant.split("+")}
        saved = {k_: native.get_option(k_) for k_ in opts}
        try:
            for k_, v_ in opts.items():
                native.set_option(k_, v_)
            for s in range(2):
                rep.walk_local_to_host(local_h, p, q, L, 5 + s, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_h)
            barrier()
            t0 = time.perf_counter()
            for s in range(4):
                rep.walk_local_to_host(local_h, p, q, L, 100 + s, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_h)
            dt = max_over_ranks(time.perf_counter() - t0) / 4
            runs[variant] = {"ms_per_step": round(dt * 1e3, 2), "G_steps_per_s": round(steps / dt / 1e9, 2),
                             "aggregate_d2h_GBps_if_plain": round(n_walks * (L + 1) * 8 / dt / 1e9, 1)}
        except Exception as exc:  # noqa: BLE001
            runs[variant] = {"error": repr(exc)}
        finally:
            for k_, v_ in saved.items():
                native.set_option(k_, v_)
        if rank == 0:
            print(variant, runs[variant], flush=True)
    res["e2e"] = runs
    # the device-side part alone, for scale: the shard's walk without the download
    local_d = local_h.to(dev)
    out_d = torch.empty((local_d.numel(), L + 1), dtype=torch.int64, device=dev)
    for s in range(2):
        rep.walk_local(local_d, p, q, L, 7 + s, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_d)
    barrier()
    t0 = time.perf_counter()
    for s in range(4):
        rep.walk_local(local_d, p, q, L, 200 + s, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_d)
    torch.cuda.synchronize()
    res["device_walk_ms"] = round(max_over_ranks(time.perf_counter() - t0) / 4 * 1e3, 2)
    same = torch.equal(out_d.cpu()[:1000], rep.walk_local_to_host(local_h, p, q, L, 203, rank * trw_dist.DEFAULT_BLOCK, blocks, out=out_h)[:1000])
    res["host_equals_device"] = bool(same)
    if rank == 0:
        print("device walk ms", res["device_walk_ms"], "host == device:", same, flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", arg("out", "e2e_ranks_probe") + ".json"), "w"), indent=1)
    barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
