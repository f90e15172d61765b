"""Small, fast pass over every kernel family and variant for compute-sanitizer (memcheck):

    compute-sanitizer --tool memcheck python tools/sanitizer_target.py
Sizes are tiny on purpose (the tool slows kernels down by one to two orders of magnitude); the
shapes are the awkward ones: row lengths that are not multiples of the staging pieces, strided
outputs, walks not a multiple of the warp size, empty rows, hubs, every (p, q) scheme.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import native, rmat, rw, utils  # noqa: E402


def main():
    native.set_graph_cache(False)
    rp, ci = rmat.rmat_csr(12, 16, device="cuda", seed=2)  # 4096 nodes, hubs >= 2048 neighbours? no: ~1000; plus a star below
    n = rp.numel() - 1
    nodes = torch.arange(n, device="cuda")[: n - 5]  # not a multiple of 32
    laws = ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0), (0.25, 4.0), (2.0, 4.0), (0.25, 0.5), (4.0, 0.25))
    for rec in (1, 0):
        native.set_option("records", rec)
        for p, q in laws:
            for L in (0, 1, 2, 5, 17, 80):
                native.walk(rp, ci, nodes, p, q, L, 3)
    native.set_option("records", -1)
    big = torch.empty((nodes.numel(), 100), dtype=torch.int64, device="cuda")
    for p, q in laws[:3]:
        native.walk(rp, ci, nodes, p, q, 80, 3, out=big[:, 7:88])  # rows that start off any line boundary
    g = native.prepare_csr(rp, ci)  # edge records + triangle Blooms
    for p, q in laws:
        g.walk(nodes[:1000], p, q, 33, 9, walk_id_offset=12345)
        g.walk(nodes[:1000], p, q, 33, 9, walk_id_offset=64, walk_id_blocks=(64, 256))  # block-cyclic walk ids
    g.walk_windows5(nodes[:777], 1.0, 0.5, 17, 9)                      # fused walk -> window prototype
    g.walk_to_host(nodes[:999].cpu(), 1.0, 0.5, 17, 9)                 # download pipeline
    native.set_option("n2v_warp", 1)                                   # warp-per-walk A/B kernel
    g.walk(nodes[:500], 0.5, 2.0, 12, 9)
    native.set_option("n2v_warp", 0)
    for cap, mb in ((8, 0), (1 << 20, 1)):                             # Bloom caps, edge filter
        native.set_option("edge_bloom_cap", cap)
        native.set_option("edge_filter_mb", mb)
        g2 = native.prepare_csr(rp, ci)
        g2.walk(nodes[:1000], 1.0, 0.5, 20, 9)
        del g2
    native.set_option("edge_bloom_cap", 256)
    native.set_option("edge_filter_mb", 0)
    g32 = native.prepare_csr(rp.int(), ci.int())                       # int32 CSR
    g32.walk(nodes[:1000], 0.5, 2.0, 20, 9)
    native.walk(rp.int(), ci.int(), nodes[:1000], 1.0, 0.5, 20, 9)
    native.csr_checksum(rp, ci[1:])                                    # checksum, unaligned start
    native.csr_checksum(rp.int(), ci.int())
    rp_h, ci_h = rp.cpu().pin_memory(), ci.cpu().pin_memory()
    for _ in range(5):                                                 # host path: fresh upload, kept replica, copy-engine check
        native.walk_host(rp_h, ci_h, nodes[:3000].cpu(), 1.0, 0.5, 11, 3, device=0)
    # a star: one hub row of 5000 neighbours (hub-segment builder), leaves of degree 1
    m = 5001
    hub_rp = torch.cat((torch.tensor([0, m - 1]), m - 1 + torch.arange(1, m))).cuda()
    hub_ci = torch.cat((torch.arange(1, m), torch.zeros(m - 1, dtype=torch.int64))).cuda()
    for rec in (1, 0):
        native.set_option("records", rec)
        native.walk(hub_rp, hub_ci, torch.arange(m, device="cuda"), 0.5, 2.0, 9, 1)
        native.walk(hub_rp, hub_ci, torch.arange(m, device="cuda"), 1.0, 0.5, 9, 1)
    native.set_option("records", -1)
    # no edges at all, and ids outside the graph
    z = torch.zeros(11, dtype=torch.int64, device="cuda")
    native.walk(z, torch.empty(0, dtype=torch.int64, device="cuda"), torch.arange(10, device="cuda"), 0.5, 2.0, 7, 1)
    bad = ci.clone()
    bad[::97] = n + 3
    native.set_option("records", 1)
    native.walk(rp, bad, nodes, 1.0, 0.5, 20, 1)
    native.set_option("records", -1)
    # windows: every mode, W = 5 fast paths and odd shapes
    walks = native.walk(rp, ci, nodes, 1.0, 1.0, 80, 3)
    for W in (1, 2, 5, 6, 10):
        rw.to_windows(walks, W, n, 1)
        rw.to_windows_cbow(walks, W, n, 1)
    rw.to_windows(walks[:3].contiguous(), 5, n, 1)
    table = native.negative_table((rp[1:] - rp[:-1]).double(), device="cuda")    # weighted negatives (alias table)
    native.to_windows(walks, 5, n, 1, neg_table=table)
    native.to_windows_cbow(walks, 6, n, 1, neg_table=table)
    triples = rmat.kg_triples(200, 7, 3000, device="cuda")
    index, ts = rmat.relation_tail_index(triples, 200)
    tw = rw.walk_triples(ts, index, torch.arange(200, device="cuda").repeat(3), walk_length=13, padding_idx=207, seed=1)
    for W in (1, 3, 5):
        rw.to_windows_triples(tw, W, 200, 207, ts, 1)
        rw.to_windows_triples_cbow(tw, W, 200, 207, ts, 1)
    native.set_option("win_bulk", 1)                                   # bulk-store stages
    rw.to_windows_triples(tw, 5, 200, 207, ts, 1)
    rw.to_windows_triples_cbow(tw, 5, 200, 207, ts, 1)
    native.set_option("win_bulk", 0)
    # edge-list walks
    el = torch.stack((torch.repeat_interleave(torch.arange(n, device="cuda"), rp[1:] - rp[:-1]), ci), 1).contiguous()
    nei, el = utils.build_node_edge_index(el, torch.arange(n))
    rw.walk_edge_list(el, nei, nodes[:500], 1.0, 1.0, 9, 1, n)
    rw.walk_edge_list(el, nei, nodes[:500], 0.5, 2.0, 9, 1, n)
    torch.cuda.synchronize()
    print("sanitizer target done")


if __name__ == "__main__":
    main()
