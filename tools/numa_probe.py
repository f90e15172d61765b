"""Host-side topology probe for the end-to-end path (measurement tooling).

    python tools/numa_probe.py [--gb=2]
Prints the box's NUMA layout and where each visible GPU hangs, then times pinned-memory copies in both directions with the
pinned buffer first-touched under the CPU affinity of every NUMA node in turn: a difference between nodes says the host
buffers of a rank should be allocated by a thread bound to its GPU's node.
"""
import glob
import json
import os
import subprocess
import sys

import torch


def arg(name, default):
    for a in sys.argv[1:]:
        if a.startswith(f"--{name}="):
            return a.split("=", 1)[1]
    return default


def cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def main():
    gb = float(arg("gb", "2"))
    res = {"allowed_cpus": len(os.sched_getaffinity(0)), "nodes": {}, "gpus": {}}
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        node = int(path.rsplit("node", 1)[1])
        res["nodes"][node] = cpulist(open(path + "/cpulist").read())
    for d in range(torch.cuda.device_count()):
        bus = torch.cuda.get_device_properties(d).pci_bus_id if hasattr(torch.cuda.get_device_properties(d), "pci_bus_id") else None
        res["gpus"][d] = {"bus": bus}
    try:
        res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=60).stdout
        q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                           timeout=60).stdout
        for line in q.strip().splitlines():
            idx, bus = [x.strip() for x in line.split(",")]
            sysfs = "/sys/bus/pci/devices/" + bus.lower()[4:] + "/numa_node"  # 00000000:1B:00.0 -> 0000:1b:00.0
            res["gpus"][f"smi{idx}"] = {"bus": bus, "numa_node": open(sysfs).read().strip() if os.path.exists(sysfs) else None}
    except Exception as exc:  # noqa: BLE001
        res["topo_error"] = repr(exc)
    print(res.get("topo", ""), flush=True)
    print(json.dumps({k: v for k, v in res.items() if k != "topo"}), flush=True)

    all_cpus = sorted(os.sched_getaffinity(0))
    n = int(gb * (1 << 30)) // 8
    dev = torch.empty(n, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rows = {}
    for node, cpus in list(res["nodes"].items()) + [("any", all_cpus)]:
        use = [c for c in cpus if c in all_cpus]
        if not use:
            rows[str(node)] = "no allowed cpu on this node"
            continue
        os.sched_setaffinity(0, use)
        host = torch.empty(n, dtype=torch.int64, pin_memory=True)
        host.zero_()  # first touch under this affinity
        out = {}
        for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            out[name + "_GBps"] = round(4 * n * 8 / (e0.elapsed_time(e1) / 1e3) / 1e9, 1)
        # both directions at once (the kept-replica check re-reads the host CSR while walks come down)
        s2 = torch.cuda.Stream()
        host2 = torch.empty(n, dtype=torch.int64, pin_memory=True)
        host2.zero_()
        dev2 = torch.empty_like(dev)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(4):
            host.copy_(dev, non_blocking=True)
        with torch.cuda.stream(s2):
            for _ in range(4):
                dev2.copy_(host2, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out["d2h_while_h2d_GBps"] = round(4 * n * 8 / (e0.elapsed_time(e1) / 1e3) / 1e9, 1)
        rows[str(node)] = out
        del host, host2, dev2
        os.sched_setaffinity(0, all_cpus)
        print(node, out, flush=True)
    res["copies"] = rows
    name = arg("out", "numa_probe")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/{name}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
