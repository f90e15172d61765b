"""A/B of the two node2vec kernel designs on kept graphs (measurement tooling): one thread per walk with O(1)-proposal
rejection (shipped) against one warp per walk with an exact CDF up to 64 neighbours (option n2v_warp, the design the
north star sketches).

    python tools/warp_ab.py
Graphs: a degree-regular random graph (no skew at all: the regime the CDF is made for), the c2-shaped R-MAT and the c3 R-MAT.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402


def regular_graph(n, d, device):
    """Every node gets d//2 random partners; symmetrised and de-duplicated: degrees within a few of d."""
    g = torch.Generator(device=device).manual_seed(3)
    src = torch.arange(n, device=device).repeat_interleave(d // 2)
    dst = torch.randint(0, n, (src.numel(),), generator=g, device=device)
    return rmat.edges_to_csr(src, dst, n)


def run(name, rp, ci, p, q, L, res):
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    out = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    g = native.prepare_csr(rp, ci)
    row = {"n": rp.numel() - 1, "nnz": ci.numel(), "walks": targets.numel(), "max_degree": int(deg.max()), "p": p, "q": q}
    for warp in (0, 1):
        native.set_option("n2v_warp", warp)
        for k in range(2):
            g.walk(targets, p, q, L, 5 + k, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3):
            g.walk(targets, p, q, L, 7 + k, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        row["warp_per_walk" if warp else "thread_per_walk"] = {"ms": ms, "gsteps": targets.numel() * L / ms / 1e6}
    native.set_option("n2v_warp", 0)
    res[name] = row
    print(name, json.dumps(row), flush=True)
    del g


def main():
    native.set_graph_cache(False)
    res = {}
    for d in (16, 32, 64):
        rp, ci = regular_graph(1 << 22, d, "cuda")
        run(f"regular_d{d}_p1_q0.5", rp, ci, 1.0, 0.5, 80, res)
        run(f"regular_d{d}_p0.5_q2", rp, ci, 0.5, 2.0, 80, res)
        del rp, ci
    rp, ci = rmat.rmat_csr(22, 16, n_nodes=2449029, n_edges=34_000_000, device="cuda")
    run("c2_shaped_p0.5_q2", rp, ci, 0.5, 2.0, 80, res)
    del rp, ci
    rp, ci = rmat.rmat_csr(24, 16, device="cuda")
    run("c3_p1_q0.5", rp, ci, 1.0, 0.5, 80, res)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_warp_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
