"""The reference's own CUDA kernels (oracle/_ref: unmodified sources recompiled for sm_100) next to
ours on the benchmark graph (measurement tooling; writes gpurun_out/ref_cuda_compare.json).

    python tools/ref_cuda_compare.py [--scale 24]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from torch_random_walk_b200 import native, rmat  # noqa: E402


def timed(fn, reps=3):
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
    scale = 24
    for a in sys.argv[1:]:
        if a.startswith("--scale="):
            scale = int(a.split("=")[1])
    rn = ref.native()
    rp, ci = rmat.rmat_csr(scale, 16, device="cuda")
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    L = 80
    res = {"graph": {"n_nodes": rp.numel() - 1, "nnz": ci.numel(), "walks": targets.numel()}}
    ms = timed(lambda: rn.walk(rp, ci, targets, 1.0, 1.0, L, 5))
    res["reference_cuda_first_order_all_walks"] = {"ms": ms, "G_steps_per_s": targets.numel() * L / ms / 1e6}
    ms = timed(lambda: native.walk(rp, ci, targets, 1.0, 1.0, L, 5))
    res["ours_first_order_all_walks"] = {"ms": ms, "G_steps_per_s": targets.numel() * L / ms / 1e6}
    g = torch.Generator(device="cuda").manual_seed(1)
    small = targets[torch.randperm(targets.numel(), device="cuda", generator=g)[: 1 << 14]].contiguous()
    ms = timed(lambda: rn.walk(rp, ci, small, 1.0, 0.5, L, 5), reps=1)
    res["reference_cuda_node2vec_p1_q0.5_16k_walks"] = {"ms": ms, "G_steps_per_s": small.numel() * L / ms / 1e6}
    ms = timed(lambda: native.walk(rp, ci, targets, 1.0, 0.5, L, 5))
    res["ours_node2vec_p1_q0.5_all_walks"] = {"ms": ms, "G_steps_per_s": targets.numel() * L / ms / 1e6}
    walks = native.walk(rp, ci, targets[: 1 << 20].contiguous(), 1.0, 1.0, L, 5)
    n_win = walks.size(0) * (L + 1 - 5 + 1)
    ms = timed(lambda: rn.to_windows(walks, 5, rp.numel() - 1, 1))
    res["reference_cuda_to_windows_1M_walks"] = {"ms": ms, "G_windows_per_s": n_win / ms / 1e6}
    ms = timed(lambda: native.to_windows(walks, 5, rp.numel() - 1, 1))
    res["ours_to_windows_1M_walks"] = {"ms": ms, "G_windows_per_s": n_win / ms / 1e6}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_compare.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
