"""A/B of the L2-resident edge filter (member_table.cuh) on the benchmark graphs (measurement tooling).

    python tools/filter_probe.py [--scale=24] [--pq=1,0.5] [--sizes=0,32,64,96] [--caps=256] [--out=name]
For every filter size (MB) and triangle-Bloom cap: the prepared-graph walk (kernel alone) and its preparation;
with --stateless=1 also the stateless call (build + walk) per filter size.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402


def arg(name, default):
    for a in sys.argv[1:]:
        if a.startswith(f"--{name}="):
            return a.split("=", 1)[1]
    return default


def main():
    scale = int(arg("scale", "24"))
    p, q = map(float, arg("pq", "1,0.5").split(","))
    sizes = [int(x) for x in arg("sizes", "0,32,64,96").split(",")]
    caps = [int(x) for x in arg("caps", "256").split(",")]
    stateless = arg("stateless", "0") == "1"
    extra = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in arg("opts", "").split(",") if kv}
    for k_, v_ in extra.items():
        native.set_option(k_, v_)
    L = int(arg("L", "80"))
    n_nodes = arg("nodes", None)
    n_edges = arg("edges", None)
    rp, ci = rmat.rmat_csr(scale, 16, device="cuda", n_nodes=int(n_nodes) if n_nodes else None,
                           n_edges=int(n_edges) if n_edges else None)
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    print(f"graph n={rp.numel() - 1:,} nnz={ci.numel():,} walks={targets.numel():,} p={p} q={q} L={L}", flush=True)
    out = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    native.set_graph_cache(False)
    res, base = {}, None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for mb in sizes:
        native.set_option("edge_filter_mb", mb)
        if stateless:
            for _ in range(2):
                native.walk(rp, ci, targets, p, q, L, 5, out=out, cache=False)
            torch.cuda.synchronize()
            e0.record()
            for k in range(4):
                native.walk(rp, ci, targets, p, q, L, 6 + k, out=out, cache=False)
            e1.record()
            torch.cuda.synchronize()
            b, w = native.last_kernel_ms()
            row = dict(stateless_call_ms=e0.elapsed_time(e1) / 4, stateless_build_ms=b, stateless_walk_ms=w)
            res[f"stateless_filter{mb}"] = row
            print("stateless", mb, {k: round(v, 3) for k, v in row.items()}, flush=True)
        for cap in caps:
            native.set_option("edge_bloom_cap", cap)
            row = {}
            torch.cuda.synchronize()
            e0.record()
            g = native.prepare_csr(rp, ci)
            e1.record()
            torch.cuda.synchronize()
            row["prepare_ms"] = e0.elapsed_time(e1)
            row.update(g.info())
            for _ in range(2):
                g.walk(targets, p, q, L, 5, out=out)
            torch.cuda.synchronize()
            e0.record()
            for k in range(5):
                g.walk(targets, p, q, L, 6 + k, out=out)
            e1.record()
            torch.cuda.synchronize()
            row["prepared_walk_ms"] = e0.elapsed_time(e1) / 5
            row["prepared_gsteps"] = targets.numel() * L / row["prepared_walk_ms"] / 1e6
            row["frac104"] = row["prepared_gsteps"] * 104 / 6542.4
            check = out[:: max(1, out.size(0) // 4096)].clone()
            if base is None:
                base = check
            row["identical_to_first_variant"] = bool(torch.equal(check, base))
            del g
            res[f"filter{mb}_cap{cap}"] = row
            print(f"filter {mb} MB, cap {cap}:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
    native.set_option("edge_bloom_cap", 256)
    native.set_option("edge_filter_mb", 64)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", arg("out", "filter_probe") + ".json"), "w"), indent=1)


if __name__ == "__main__":
    main()
