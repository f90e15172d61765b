"""End-to-end host path on the benchmark graph under different wire formats (measurement tooling).

    TRW_HOST_TIMING=1 python tools/host_probe.py [--scale=24] [--variants=compress:1,compress:2,...]
Each variant is a '+'-joined list of option:value pairs applied to trw_walk_csr_host.
"""
import json
import os
import sys
import time

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
    p, q, L = 1.0, 0.5, 80
    rp, ci = rmat.rmat_csr(scale, 16, device="cuda")
    deg = rp[1:] - rp[:-1]
    targets = torch.nonzero(deg > 0).flatten().contiguous()
    rp_h, ci_h, tg_h = rp.cpu().pin_memory(), ci.cpu().pin_memory(), targets.cpu().pin_memory()
    del rp, ci
    out_h = torch.empty((tg_h.numel(), L + 1), dtype=torch.int64, pin_memory=True)
    steps = tg_h.numel() * L
    res = {}
    for variant in arg("variants", "host_compress:1,host_compress:2,host_compress:0").split(","):
        opts = {kv.split(":")[0]: int(kv.split(":")[1]) for kv in variant.split("+")}
        saved = {k: native.get_option(k) for k in opts}
        for k, v in opts.items():
            native.set_option(k, v)
        native.lib().trw_release_cached_buffers()
        ms = []
        for k in range(9):
            t0 = time.perf_counter()
            native.walk_host(rp_h, ci_h, tg_h, p, q, L, 10 + k, device=0, out=out_h)
            ms.append((time.perf_counter() - t0) * 1e3)
        steady = sorted(ms[5:])[len(ms[5:]) // 2]
        res[variant] = {"call_ms": ms, "steady_ms": steady, "steady_gsteps": steps / steady / 1e6}
        print(variant, {"steady_ms": round(steady, 1), "G steps/s": round(steps / steady / 1e6, 2), "calls": [round(x) for x in ms]}, flush=True)
        for k, v in saved.items():
            native.set_option(k, v)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", arg("out", "host_probe") + ".json"), "w"), indent=1)


if __name__ == "__main__":
    main()
