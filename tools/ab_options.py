"""A/B of library options on the benchmark graph (measurement tooling).

    python tools/ab_options.py [--scale 24] name=value[,name=value...] ...
Each argument is one variant (comma-separated overrides of the defaults); "default" is always run.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = 24
    for a in sys.argv[1:]:
        if a.startswith("--scale="):
            scale = int(a.split("=")[1])
    p, q = 1.0, 0.5
    for a in sys.argv[1:]:
        if a.startswith("--pq="):
            p, q = map(float, a.split("=")[1].split(","))
    rp, ci = rmat.rmat_csr(scale, 16, device="cuda")
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    L = 80
    out = torch.empty((targets.numel(), L + 1), dtype=torch.int64, device="cuda")
    native.set_option("time_kernels", 1)
    native.set_graph_cache(False)  # every variant must build with its own options
    res = {}
    for variant in ["default"] + args:
        opts = {} if variant == "default" else {kv.split("=")[0]: int(kv.split("=")[1]) for kv in variant.split(",")}
        saved = {k: native.get_option(k) for k in opts}
        for k, v in opts.items():
            native.set_option(k, v)
        for _ in range(2):
            native.walk(rp, ci, targets, p, q, L, 5, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(5):
            native.walk(rp, ci, targets, p, q, L, 6 + k, out=out)
        e1.record()
        torch.cuda.synchronize()
        b, w = native.last_kernel_ms()
        res[variant] = {"call_ms": e0.elapsed_time(e1) / 5, "build_ms": b, "walk_ms": w}
        print(variant, {k: round(v, 3) for k, v in res[variant].items()}, flush=True)
        for k, v in saved.items():
            native.set_option(k, v)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_options.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
