"""Generates tests/golden/ref_golden.npz by RUNNING THE UNMODIFIED REFERENCE in this container.

The reference extension is oracle/_ref/torch_rw_native.so (built in place from /root/reference by
oracle/build_ref.py); the reference's Python helpers are imported from
/root/reference/torch_rw/utils.py (with the one-line networkx-3 shim SURVEY.md section 4 names).
Every array stored is either an input we generated from a fixed seed or an output the reference
produced for it; nothing is computed by this repository's own code.  The file pins
oracle/trw_oracle.c (CPU, bit-exact) and the host-side utils mirror, and carries the RNG-free
reference outputs (window positives/targets, dead-end walk rows) the CUDA path must reproduce.

    python tests/golden/make_golden.py          # needs /root/reference; not run on the GPU box
"""
import importlib.util
import os
import sys

import networkx as nx
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

REF_DIR = os.environ.get("TRW_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")


def load_ref_utils():
    if not hasattr(nx, "to_scipy_sparse_matrix"):
        nx.to_scipy_sparse_matrix = nx.to_scipy_sparse_array  # removed in networkx 3 (utils.py:6)
    spec = importlib.util.spec_from_file_location("ref_torch_rw_utils", os.path.join(REF_DIR, "torch_rw", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def toy_graph(directed):
    g = nx.DiGraph() if directed else nx.Graph()
    for a, b in [("A", "B"), ("A", "C"), ("B", "C"), ("B", "D"), ("D", "C"), ("E", "A"), ("E", "D")]:
        g.add_edge(a, b)  # the graph every reference test builds (tests/test_rw.py:31-40)
    return g


def random_csr(rng, n, avg_deg, symmetric=True):
    m = n * avg_deg // (2 if symmetric else 1)
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if symmetric:
        src, dst = np.r_[src, dst], np.r_[dst, src]
    key = np.unique(src.astype(np.int64) * n + dst)
    src, dst = key // n, key % n
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr), dst.astype(np.int64)


def main():
    native = ref.native()
    rutils = load_ref_utils()
    g = {}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731

    # ---------------------------------------------------------------- utils
    for name, graph in (("toy_undirected", toy_graph(False)), ("toy_directed", toy_graph(True)),
                        ("karate", nx.karate_club_graph())):
        rp, ci = rutils.to_csr(graph)
        g[f"utils/{name}/row_ptr"], g[f"utils/{name}/col_idx"] = rp.numpy(), ci.numpy()
        g[f"utils/{name}/nodes"] = rutils.nodes_tensor(graph).numpy()
    for name, graph in (("toy_undirected", toy_graph(False)), ("toy_directed", toy_graph(True))):
        el, mapping = rutils.to_edge_list_indexed(graph)
        g[f"utils/{name}/edge_list"] = el.numpy()
        g[f"utils/{name}/mapping_keys"] = np.array(list(mapping.keys()))
        g[f"utils/{name}/mapping_values"] = np.array(list(mapping.values()), dtype=np.int64)
        nei, el_sorted = rutils.build_node_edge_index(el, torch.unique(el.view(-1)))
        g[f"utils/{name}/node_edge_index"], g[f"utils/{name}/edge_list_sorted"] = nei.numpy(), el_sorted.numpy()
    rng = np.random.default_rng(20261018)
    for case, (n, e) in enumerate(((12, 1), (12, 2), (40, 150), (300, 2000))):
        el = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], 1).astype(np.int64)
        nei, el_sorted = rutils.build_node_edge_index(T(el), torch.arange(n))
        g[f"utils/rand_el{case}/edge_list"] = el
        g[f"utils/rand_el{case}/num_nodes"] = np.array(n)
        g[f"utils/rand_el{case}/node_edge_index"], g[f"utils/rand_el{case}/edge_list_sorted"] = nei.numpy(), el_sorted.numpy()
        tr = np.stack([rng.integers(0, n, e), rng.integers(n, n + 7, e), rng.integers(0, n, e)], 1).astype(np.int64)
        rti, tr_sorted = rutils.build_relation_tail_index(T(tr), torch.arange(n))
        g[f"utils/rand_tr{case}/triples"] = tr
        g[f"utils/rand_tr{case}/relation_tail_index"], g[f"utils/rand_tr{case}/triples_sorted"] = rti.numpy(), tr_sorted.numpy()
    # the reference test's float triples (tests/test_rw_triples.py:26-45)
    A, B, C, D, E, r1, r2, r3 = 0, 1, 2, 3, 4, 5, 6, 7
    toy_triples = torch.Tensor([(A, r1, B), (B, r2, D), (A, r1, C), (C, r2, E), (C, r3, B), (A, r2, D), (D, r3, A), (D, r2, C)])
    ents = torch.Tensor(list(set(toy_triples[:, 0].tolist() + toy_triples[:, 2].tolist()))).to(int)
    rti, tr_sorted = rutils.build_relation_tail_index(toy_triples, ents)
    g["utils/toy_triples/triples"] = toy_triples.numpy()
    g["utils/toy_triples/entities"] = ents.numpy()
    g["utils/toy_triples/relation_tail_index"], g["utils/toy_triples/triples_sorted"] = rti.numpy(), tr_sorted.numpy()

    # ---------------------------------------------------------------- CSR walks (CPU reference)
    rp, ci = T(g["utils/toy_undirected/row_ptr"]), T(g["utils/toy_undirected/col_idx"])
    nodes = T(g["utils/toy_undirected/nodes"])
    for tag, (p, q) in (("uniform", (1.0, 1.0)), ("biased", (0.7, 0.5))):
        g[f"walk/toy_{tag}"] = native.walk(rp, ci, nodes, p, q, 6, 10).numpy()  # tests/test_rw.py:46,114
    krp, kci = T(g["utils/karate/row_ptr"]), T(g["utils/karate/col_idx"])
    knodes = T(g["utils/karate/nodes"]).repeat_interleave(10)
    g["walk/karate_uniform_L80"] = native.walk(krp, kci, knodes, 1.0, 1.0, 80, 10).numpy()  # BASELINE config 1
    g["walk/karate_p0.5_q2_L80"] = native.walk(krp, kci, knodes, 0.5, 2.0, 80, 10).numpy()
    for case, (n, deg, p, q, L, seed) in enumerate(((60, 6, 1.0, 1.0, 12, 3), (60, 6, 0.25, 4.0, 12, 4),
                                                    (200, 10, 1.0, 0.5, 20, 5), (200, 10, 2.0, 0.5, 9, 6))):
        rp_, ci_ = random_csr(rng, n, deg)
        starts = np.flatnonzero(np.diff(rp_) > 0).astype(np.int64)  # the CPU reference divides by zero on degree 0
        g[f"walk/rand{case}/row_ptr"], g[f"walk/rand{case}/col_idx"], g[f"walk/rand{case}/targets"] = rp_, ci_, starts
        g[f"walk/rand{case}/params"] = np.array([p, q, L, seed])
        g[f"walk/rand{case}/walks"] = native.walk(T(rp_), T(ci_), T(starts), p, q, L, seed).numpy()

    # ---------------------------------------------------------------- edge-list walks
    for name in ("toy_directed", "toy_undirected"):
        el_sorted, nei = T(g[f"utils/{name}/edge_list_sorted"]), T(g[f"utils/{name}/node_edge_index"])
        targets = T(g[f"utils/{name}/mapping_values"])
        pad = int(sorted(targets.tolist())[-1] + 1)  # tests/test_rw_edge_list.py:40
        for tag, (p, q) in (("uniform", (1.0, 1.0)), ("biased", (0.7, 0.5))):
            for restart in (True, False):
                g[f"walk_el/{name}_{tag}_restart{int(restart)}"] = native.walk_edge_list(
                    el_sorted, nei, targets, p, q, 6, 10, pad, restart).numpy()
    for case, (n, e, p, q, L, seed, restart) in enumerate(((30, 60, 1.0, 1.0, 10, 7, True), (30, 60, 1.0, 1.0, 10, 8, False),
                                                           (30, 90, 0.7, 0.5, 10, 9, True), (50, 200, 2.0, 0.25, 8, 11, False))):
        el = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)], 1).astype(np.int64)
        # one spare index row for the padding id keeps the reference's is_neighbor(., padding) read in bounds
        nei, el_sorted = rutils.build_node_edge_index(T(el), torch.arange(n + 1))
        targets = torch.arange(n)
        g[f"walk_el/rand{case}/edge_list_sorted"], g[f"walk_el/rand{case}/node_edge_index"] = el_sorted.numpy(), nei.numpy()
        g[f"walk_el/rand{case}/params"] = np.array([p, q, L, seed, n, int(restart)])
        g[f"walk_el/rand{case}/walks"] = native.walk_edge_list(el_sorted, nei, targets, p, q, L, seed, n, restart).numpy()

    # ---------------------------------------------------------------- triple walks
    rti, trs = T(g["utils/toy_triples/relation_tail_index"]), T(g["utils/toy_triples/triples_sorted"])
    tt = T(g["utils/toy_triples/entities"]).repeat_interleave(2, 0)
    g["walk_tr/toy"] = native.walk_triples(trs, rti, tt, 6, r3 + 1, False, 10).numpy()  # tests/test_rw_triples.py:61-68
    for case in range(2):
        trs, rti = T(g[f"utils/rand_tr{case + 2}/triples_sorted"]), T(g[f"utils/rand_tr{case + 2}/relation_tail_index"])
        n = rti.size(0)
        g[f"walk_tr/rand{case}/walks"] = native.walk_triples(trs, rti, torch.arange(n), 7, n + 7, False, 21 + case).numpy()

    # ---------------------------------------------------------------- windows
    torch.manual_seed(20)  # tests/test_windows.py:6-7
    walks = torch.randint(low=0, high=30, size=(3, 10))
    g["win/test_walks"] = walks.numpy()
    for k, a in enumerate(native.to_windows(walks, 5, 30, 20)):
        g[f"win/test_skipgram/{k}"] = a.numpy()
    for k, a in enumerate(native.to_windows_cbow(walks, 5, 30, 20)):
        g[f"win/test_cbow/{k}"] = a.numpy()
    torch.manual_seed(20)  # tests/test_windows.py:124-127
    twalks = torch.randint(low=0, high=30, size=(3, 21))
    triples = torch.randint(low=0, high=30, size=(10, 3))
    g["win/test_twalks"], g["win/test_triples"] = twalks.numpy(), triples.numpy()
    for k, a in enumerate(native.to_windows_triples(twalks, 4, 30, -1, triples, 20)):
        g[f"win/test_triples_sg/{k}"] = a.numpy()
    for k, a in enumerate(native.to_windows_triples_cbow(twalks, 4, 30, -1, triples, 20)):
        g[f"win/test_triples_cbow/{k}"] = a.numpy()
    shapes = []
    for case, (n, wl, W) in enumerate(((4, 10, 1), (5, 11, 2), (3, 12, 3), (7, 13, 4), (2, 20, 7), (6, 21, 5), (3, 81, 5), (9, 5, 5))):
        w = torch.from_numpy(rng.integers(0, 50, (n, wl)).astype(np.int64))
        tri = torch.from_numpy(rng.integers(0, 50, (17, 3)).astype(np.int64))
        g[f"win/rand{case}/walks"], g[f"win/rand{case}/triples"] = w.numpy(), tri.numpy()
        shapes.append((n, wl, W))
        for k, a in enumerate(native.to_windows(w, W, 50, case)):
            g[f"win/rand{case}/skipgram/{k}"] = a.numpy()
        for k, a in enumerate(native.to_windows_cbow(w, W, 50, case)):
            g[f"win/rand{case}/cbow/{k}"] = a.numpy()
        for k, a in enumerate(native.to_windows_triples(w, W, 50, 77, tri, case)):
            g[f"win/rand{case}/triples_sg/{k}"] = a.numpy()
        for k, a in enumerate(native.to_windows_triples_cbow(w, W, 50, 77, tri, case)):
            g[f"win/rand{case}/triples_cbow/{k}"] = a.numpy()
    g["win/rand_shapes"] = np.array(shapes)
    np.savez_compressed(OUT, **g)
    print(f"wrote {OUT}: {len(g)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
