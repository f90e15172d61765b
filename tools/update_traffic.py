"""Turns an ncu launch list (csv with gpu__time_duration, dram__bytes_read/write, fabric requests) into
profiles/<name>.csv (this library's kernels only) and refreshes profiles/traffic_bytes_per_launch.json, which bench.py
quotes as roofline.traffic ONLY while the hash of the CUDA sources is the one recorded here.

    python tools/update_traffic.py gpurun_out/launches.csv r02_launches_c3_prepared c3
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    src, name, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    hd = rows[h]
    kn, mn, mv, idc = hd.index("Kernel Name"), hd.index("Metric Name"), hd.index("Metric Value"), hd.index("ID")
    agg = {}
    for r in rows[h + 1:]:
        if len(r) > mv and "trw::" in r[kn]:
            agg.setdefault((int(r[idc]), r[kn].split("(")[0].replace("void ", "")), {})[r[mn]] = float(r[mv].replace(",", ""))
    out = os.path.join(ROOT, "profiles", name + ".csv")
    walk = None
    with open(out, "w") as f:
        f.write("id,kernel,time_ms,dram_read_GB,dram_write_GB,l2_to_fabric_requests_M,fabric_requests_G_per_s\n")
        for (i, k), v in sorted(agg.items()):
            ms = v.get("gpu__time_duration.sum", 0.0) / 1e6
            rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
            fab = v.get("lts__t_requests_srcunit_ltcfabric.sum", 0.0)
            f.write(f"{i},\"{k}\",{ms:.4f},{rd / 1e9:.3f},{wr / 1e9:.3f},{fab / 1e6:.2f},{fab / max(ms, 1e-9) / 1e6:.2f}\n")
            if "node2vec_walk_kernel" in k:
                walk = (rd + wr, ms)
    print("wrote", out)
    if walk:
        import bench  # noqa: E402  (source_sha16)

        path = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        data = json.load(open(path)) if os.path.exists(path) else {}
        data.update({workload: walk[0], "source_sha16": bench.source_sha16()})
        data.setdefault("how", "ncu dram__bytes_read.sum + dram__bytes_write.sum of one node2vec_walk_kernel launch per workload key; "
                               "quoted by bench.py only while source_sha16 matches the CUDA sources")
        json.dump(data, open(path, "w"), indent=1)
        print("traffic", workload, walk[0])


if __name__ == "__main__":
    main()
