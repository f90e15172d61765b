"""A/B for SURVEY section 8 (f1), the fused walk -> window pipeline, on the c3 graph (measurement tooling).

separate:  walk kernel (kept graph) writes walks[n, 81]; rw.to_windows(walks, 5) reads them and writes target, pos, neg
fused:     the walk kernel writes target and pos itself (PreparedCsr.walk_windows5); the negatives, which do not depend on
           the walks, still have to be written by someone: their share of to_windows' bytes (32 of 72 per window) is
           charged at to_windows' own rate.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402


def timed(fn, reps=4):
    for _ in range(2):
        out = fn()
    del out
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
        del out
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    p, q, L = 1.0, 0.5, 80
    native.set_graph_cache(False)
    rp, ci = rmat.rmat_csr(scale, 16, device="cuda")
    n = rp.numel() - 1
    targets = torch.nonzero(rp[1:] - rp[:-1] > 0).flatten().contiguous()
    g = native.prepare_csr(rp, ci)
    walks = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    t_walk = timed(lambda: g.walk(targets, p, q, L, 7, out=walks))
    t_win = timed(lambda: native.to_windows(walks, 5, n, 1))
    t_fused = timed(lambda: g.walk_windows5(targets, p, q, L, 7))
    neg_share = 32.0 / 72.0
    res = {"walks": targets.numel(), "walk_length": L, "windows": targets.numel() * (L - 3),
           "separate_ms": {"walk": t_walk, "to_windows": t_win, "total": t_walk + t_win},
           "fused_ms": {"walk_with_target_and_pos": t_fused, "negatives_at_to_windows_rate": neg_share * t_win,
                        "total": t_fused + neg_share * t_win},
           "note": "torch.empty of the outputs is inside every timing (caching allocator: no cudaMalloc after the warm-up)"}
    res["gain"] = 1.0 - res["fused_ms"]["total"] / res["separate_ms"]["total"]
    print(json.dumps(res, indent=1), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_fused_ab.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
