"""Short, deterministic run of the hot path for ncu (profiles/): one R-MAT graph, a warm-up and two
timed node2vec walks, one first-order walk, one window generation.  Kept small so that
`ncu --set full` (about 40 replays per kernel) finishes in minutes.

    python tools/profile_target.py [--scale 22] [--p 1.0] [--q 0.5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat, rw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--p", type=float, default=1.0)
    ap.add_argument("--q", type=float, default=0.5)
    ap.add_argument("--walk-length", type=int, default=80)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--option", action="append", default=[])
    ap.add_argument("--prepared", action="store_true", help="walk on a kept graph (native.prepare_csr: edge records + triangle Blooms)")
    ap.add_argument("--no-extras", action="store_true", help="node2vec walks only")
    args = ap.parse_args()
    for kv in args.option:
        k, v = kv.split("=")
        native.set_option(k, int(v))
    rp, ci = rmat.rmat_csr(args.scale, 16, device="cuda")
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    L = args.walk_length
    out = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    native.set_graph_cache(False)
    g = native.prepare_csr(rp, ci) if args.prepared else None
    if g is not None:
        print("prepared:", g.info(), flush=True)
    for k in range(1 + args.reps):
        if g is not None:
            g.walk(targets, args.p, args.q, L, 10 + k, out=out)
        else:
            native.walk(rp, ci, targets, args.p, args.q, L, 10 + k, out=out)
        b, w = native.last_kernel_ms()
        print(f"node2vec p={args.p} q={args.q}: build {b:.3f} ms, walk {w:.3f} ms, "
              f"{targets.numel() * L / w / 1e6:.2f} G steps/s (kernel)", flush=True)
    if args.no_extras:
        torch.cuda.synchronize()
        print(f"n={rp.numel() - 1} nnz={ci.numel()} walks={targets.numel()}")
        return
    native.walk(rp, ci, targets, 1.0, 1.0, L, 10, out=out)
    _, w = native.last_kernel_ms()
    print(f"uniform: walk {w:.3f} ms, {targets.numel() * L / w / 1e6:.2f} G steps/s", flush=True)
    small = out[: 1 << 18]
    rw.to_windows(small.contiguous(), 5, rp.numel() - 1, 1)
    torch.cuda.synchronize()
    print(f"n={rp.numel() - 1} nnz={ci.numel()} walks={targets.numel()}")


if __name__ == "__main__":
    main()
