"""Throughput of rw.walk_edge_list on a BASELINE-shaped graph (measurement tooling).

    python tools/edge_list_probe.py [--walks 16384]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import rmat, rw, utils  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walks", type=int, default=16384)
    args = ap.parse_args()
    rp, ci = rmat.rmat_csr(22, 16, n_nodes=2449029, device="cuda", n_edges=34_000_000)  # the c2 graph
    n = rp.numel() - 1
    deg = rp[1:] - rp[:-1]
    el = torch.stack((torch.repeat_interleave(torch.arange(n, device="cuda"), deg), ci), 1).contiguous()
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nei, el = utils.build_node_edge_index(el, torch.arange(n))
    torch.cuda.synchronize()
    print(f"utils.build_node_edge_index on the device: {el.size(0)} edges, {n} nodes in {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    t0 = time.perf_counter()
    rp2, ci2 = utils.csr_from_edge_index(el, n)
    torch.cuda.synchronize()
    print(f"utils.csr_from_edge_index on the device: {(time.perf_counter() - t0) * 1e3:.1f} ms "
          f"(same CSR as the generator's: {bool(torch.equal(rp2, rp) and torch.equal(ci2, ci))})", flush=True)
    del rp2, ci2
    starts = torch.nonzero(deg > 0).flatten()
    for name, p, q, count in (("first-order", 1.0, 1.0, starts.numel()), ("node2vec p=0.5 q=2", 0.5, 2.0, args.walks),
                              ("node2vec p=1 q=0.5", 1.0, 0.5, args.walks)):
        tg = starts[torch.randperm(starts.numel(), device="cuda")[:count]].contiguous()
        for _ in range(2):
            w = rw.walk_edge_list(el, nei, tg, p, q, 80, 1, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        w = rw.walk_edge_list(el, nei, tg, p, q, 80, 2, n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"edge list {name}: {count} walks x 80 in {ms:.2f} ms = {count * 80 / ms / 1e6:.3f} G steps/s", flush=True)
    print(f"n={n} edges={el.size(0)}")


if __name__ == "__main__":
    main()
