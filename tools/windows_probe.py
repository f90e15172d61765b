"""Throughput of the four window kernels at sizes that fill the GPU (measurement tooling).

    python tools/windows_probe.py [--walks 2000000]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walks", type=int, default=2_000_000)
    ap.add_argument("--option", action="append", default=[], help="library option name=value")
    args = ap.parse_args()
    for kv in args.option:
        native.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        print("option", kv, flush=True)
    n_nodes = 1 << 24
    walks = torch.randint(0, n_nodes, (args.walks, 81), dtype=torch.int64, device="cuda")
    for name, fn in (("to_windows W=5", lambda: native.to_windows(walks, 5, n_nodes, 1)),
                     ("to_windows_cbow W=5", lambda: native.to_windows_cbow(walks, 5, n_nodes, 1)),
                     ("to_windows W=10", lambda: native.to_windows(walks[: args.walks // 2], 10, n_nodes, 1))):
        ms, out = timed(fn)
        nbytes = sum(o.numel() for o in out) * 8
        print(f"{name}: {ms:.3f} ms, {out[0].size(0) / ms / 1e6:.2f} G windows/s, {nbytes / ms / 1e6:.0f} GB/s written "
              f"({nbytes / 1e9:.2f} GB)", flush=True)
        del out
    triples = rmat.kg_triples(14541, 237, 310116, device="cuda")
    _, ts = rmat.relation_tail_index(triples, 14541)
    tw = torch.randint(0, 14541, (args.walks // 4, 81), dtype=torch.int64, device="cuda")
    for name, fn in (("to_windows_triples W=5", lambda: native.to_windows_triples(tw, 5, 14541, 14778, ts, 1)),
                     ("to_windows_triples_cbow W=5", lambda: native.to_windows_triples_cbow(tw, 5, 14541, 14778, ts, 1))):
        ms, out = timed(fn, reps=3)
        nbytes = sum(o.numel() for o in out) * 8
        print(f"{name}: {ms:.3f} ms, {out[0].size(0) / ms / 1e6:.2f} G windows/s, {nbytes / ms / 1e6:.0f} GB/s written "
              f"({nbytes / 1e9:.2f} GB)", flush=True)
        del out


if __name__ == "__main__":
    main()
