"""Pins oracle/trw_oracle.c: bit-exact against (a) the known-answer vectors hard-coded in the
reference's own CPU tests and (b) outputs of the unmodified reference built in oracle/_ref
(tests/golden/ref_golden.npz).  CPU only."""
import numpy as np
import torch

from torch_random_walk_b200 import utils


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def I(rows):
    return torch.tensor(rows, dtype=torch.int64)


# ---- known answers transcribed from the reference's tests (file:line in each test) ----------

def _toy_csr(golden):
    return T(golden["utils/toy_undirected/row_ptr"]), T(golden["utils/toy_undirected/col_idx"]), torch.arange(5)


def test_uniform_walk_known_answer(orc, golden):
    # /root/reference/tests/test_rw.py:46-55
    rp, ci, nodes = _toy_csr(golden)
    walks = orc.walk(rp, ci, nodes, 1.0, 1.0, 6, 10)
    expected = I([[0, 2, 1, 3, 4, 0, 4], [1, 3, 2, 3, 4, 3, 4], [2, 0, 1, 3, 2, 0, 2], [3, 4, 0, 1, 2, 1, 2],
                  [4, 0, 4, 0, 2, 1, 0]])
    assert torch.equal(walks, expected)


def test_biased_walk_known_answer(orc, golden):
    # /root/reference/tests/test_rw.py:114-122
    rp, ci, nodes = _toy_csr(golden)
    walks = orc.walk(rp, ci, nodes, 0.7, 0.5, 6, 10)
    expected = I([[0, 2, 3, 4, 3, 4, 3], [1, 2, 1, 2, 1, 0, 4], [2, 0, 2, 3, 4, 3, 2], [3, 2, 0, 4, 3, 4, 3],
                  [4, 0, 4, 0, 2, 3, 4]])
    assert torch.equal(walks, expected)


def _toy_edge_list(directed):
    from helpers import toy_graph

    el, mapping = utils.to_edge_list_indexed(toy_graph(directed))
    targets = torch.tensor(list(mapping.values()), dtype=torch.int64)
    nei, el_sorted = utils.build_node_edge_index(el, torch.unique(el.view(-1)))
    pad = sorted(targets.tolist())[-1] + 1
    return el_sorted, nei, targets, pad


def test_edge_list_uniform_known_answers(orc):
    el, nei, targets, pad = _toy_edge_list(directed=True)
    # /root/reference/tests/test_rw_edge_list.py:31-35
    assert torch.equal(nei, I([[0, 1], [2, 3], [-1, -1], [4, 4], [5, 6]]))
    # :43-60 (restart) and :95-112 (no restart)
    walks = orc.walk_edge_list(el, nei, targets, 1.0, 1.0, 6, 10, pad)
    assert torch.equal(walks, I([[0, 2, 5, 0, 1, 2, 5], [1, 3, 2, 5, 1, 2, 5], [2, 5, 2, 5, 2, 5, 2],
                                 [3, 2, 5, 3, 2, 5, 3], [4, 3, 2, 5, 4, 3, 2]]))
    walks = orc.walk_edge_list(el, nei, targets, 1.0, 1.0, 6, 10, pad, restart=False)
    assert torch.equal(walks, I([[0, 2, 5, 5, 5, 5, 5], [1, 2, 5, 5, 5, 5, 5], [2, 5, 5, 5, 5, 5, 5],
                                 [3, 2, 5, 5, 5, 5, 5], [4, 0, 2, 5, 5, 5, 5]]))


def test_triple_walk_known_answer(orc, golden):
    # /root/reference/tests/test_rw_triples.py:47-81
    rti = T(golden["utils/toy_triples/relation_tail_index"])
    assert torch.equal(rti, I([[0, 2], [3, 3], [4, 5], [6, 7], [-1, -1]]))
    trs = T(golden["utils/toy_triples/triples_sorted"])
    targets = T(golden["utils/toy_triples/entities"]).repeat_interleave(2, 0)
    walks = orc.walk_triples(trs, rti, targets, 6, 8, 10, restart=False)
    expected = I([[0, 5, 2, 6, 4, 8, 8, 8, 8, 8, 8, 8, 8], [0, 6, 3, 6, 2, 6, 4, 8, 8, 8, 8, 8, 8],
                  [1, 6, 3, 6, 2, 7, 1, 6, 3, 6, 2, 7, 1], [1, 6, 3, 6, 2, 7, 1, 6, 3, 6, 2, 6, 4],
                  [2, 7, 1, 6, 3, 7, 0, 5, 2, 6, 4, 8, 8], [2, 6, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8],
                  [3, 6, 2, 6, 4, 8, 8, 8, 8, 8, 8, 8, 8], [3, 7, 0, 5, 2, 7, 1, 6, 3, 6, 2, 6, 4],
                  [4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8], [4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8]])
    assert torch.equal(walks, expected)


def test_windows_known_answers(orc):
    # /root/reference/tests/test_windows.py:4-31 and :34-55
    torch.manual_seed(20)
    walks = torch.randint(low=0, high=30, size=(3, 10))
    target, pos, neg = orc.to_windows(walks, 5, 30, 20)
    assert target.size(0) == 6 * 3
    assert torch.equal(target[:6], I([27, 13, 24, 20, 13, 6]))
    pos_expected = I([[11, 10, 13, 24], [10, 27, 24, 20], [27, 13, 20, 13], [13, 24, 13, 6], [24, 20, 6, 27],
                      [20, 13, 27, 0]])
    assert torch.equal(pos[:6], pos_expected)
    assert torch.equal(neg[:6], I([[1, 18, 17, 9], [26, 1, 22, 11], [10, 1, 20, 4], [17, 9, 14, 9],
                                   [25, 17, 29, 29], [15, 16, 11, 11]]))
    pos_nodes, neg_nodes, windows = orc.to_windows_cbow(walks, 5, 30, 20)
    assert torch.equal(pos_nodes[:6], I([27, 13, 24, 20, 13, 6]))
    assert torch.equal(neg_nodes[:6], I([1, 18, 17, 9, 26, 1]))
    assert torch.equal(windows[:6], pos_expected)


def test_triple_windows_known_answers(orc):
    # /root/reference/tests/test_windows.py:122-180
    torch.manual_seed(20)
    walks = torch.randint(low=0, high=30, size=(3, 21))
    triples = torch.randint(low=0, high=30, size=(10, 3))
    target, pos, neg = orc.to_windows_triples(walks, 4, 30, -1, triples, 20)
    assert torch.equal(target[:2], I([[11, 10, 27], [27, 13, 24]]))
    assert torch.equal(pos[:2], I([[[-1, -1, 11], [-1, -1, -1], [-1, -1, -1], [-1, -1, -1], [27, 13, 24], [24, 20, 13],
                                    [13, 6, 27], [27, 0, 7]],
                                   [[10, 10, 27], [-1, -1, 11], [-1, -1, -1], [-1, -1, -1], [24, 20, 13], [13, 6, 27],
                                    [27, 0, 7], [7, 14, 20]]]))
    assert torch.equal(neg[:2], I([[[18, 5, 19], [7, 25, 24], [10, 4, 14], [16, 24, 21], [20, 23, 10], [18, 5, 19],
                                    [20, 5, 14], [18, 5, 19]],
                                   [[29, 9, 17], [18, 5, 19], [29, 9, 17], [1, 8, 6], [10, 4, 14], [16, 24, 21],
                                    [1, 8, 6], [16, 24, 21]]]))


# ---- outputs of the unmodified reference on generated inputs ---------------------------------

def test_csr_walks_match_reference(orc, golden):
    rp, ci, nodes = _toy_csr(golden)
    assert np.array_equal(orc.walk(rp, ci, nodes, 1.0, 1.0, 6, 10).numpy(), golden["walk/toy_uniform"])
    assert np.array_equal(orc.walk(rp, ci, nodes, 0.7, 0.5, 6, 10).numpy(), golden["walk/toy_biased"])
    krp, kci = T(golden["utils/karate/row_ptr"]), T(golden["utils/karate/col_idx"])
    knodes = T(golden["utils/karate/nodes"]).repeat_interleave(10)
    assert np.array_equal(orc.walk(krp, kci, knodes, 1.0, 1.0, 80, 10).numpy(), golden["walk/karate_uniform_L80"])
    assert np.array_equal(orc.walk(krp, kci, knodes, 0.5, 2.0, 80, 10).numpy(), golden["walk/karate_p0.5_q2_L80"])
    for case in range(4):
        p, q, L, seed = golden[f"walk/rand{case}/params"]
        got = orc.walk(T(golden[f"walk/rand{case}/row_ptr"]), T(golden[f"walk/rand{case}/col_idx"]),
                       T(golden[f"walk/rand{case}/targets"]), float(p), float(q), int(L), int(seed))
        assert np.array_equal(got.numpy(), golden[f"walk/rand{case}/walks"]), case


def test_edge_list_walks_match_reference(orc, golden):
    for case in range(4):
        p, q, L, seed, pad, restart = golden[f"walk_el/rand{case}/params"]
        el, nei = T(golden[f"walk_el/rand{case}/edge_list_sorted"]), T(golden[f"walk_el/rand{case}/node_edge_index"])
        got = orc.walk_edge_list(el, nei, torch.arange(int(pad)), float(p), float(q), int(L), int(seed), int(pad),
                                 bool(restart))
        assert np.array_equal(got.numpy(), golden[f"walk_el/rand{case}/walks"]), case
    for name in ("toy_directed", "toy_undirected"):
        el, nei = T(golden[f"utils/{name}/edge_list_sorted"]), T(golden[f"utils/{name}/node_edge_index"])
        targets = T(golden[f"utils/{name}/mapping_values"])
        pad = int(sorted(targets.tolist())[-1] + 1)
        for restart in (True, False):
            got = orc.walk_edge_list(el, nei, targets, 1.0, 1.0, 6, 10, pad, restart)
            assert np.array_equal(got.numpy(), golden[f"walk_el/{name}_uniform_restart{int(restart)}"]), (name, restart)


def test_edge_list_biased_toy_matches_reference_where_defined(orc, golden):
    """The reference's second-order edge-list walk reads row `padding_idx` of node_edge_index when
    the previous node is padding (csrc/cpu/rw_cpu_edge_list.cpp:221 with t == pad) -- one row past
    the tensor in its own tests (pad == N).  The oracle treats that row as empty.  Undirected toy
    graph: no dead ends, never pads, so the comparison is exact."""
    name = "toy_undirected"
    el, nei = T(golden[f"utils/{name}/edge_list_sorted"]), T(golden[f"utils/{name}/node_edge_index"])
    targets = T(golden[f"utils/{name}/mapping_values"])
    pad = int(sorted(targets.tolist())[-1] + 1)
    for restart in (True, False):
        got = orc.walk_edge_list(el, nei, targets, 0.7, 0.5, 6, 10, pad, restart)
        assert np.array_equal(got.numpy(), golden[f"walk_el/{name}_biased_restart{int(restart)}"])


def test_triple_walks_match_reference(orc, golden):
    for case in range(2):
        trs, rti = T(golden[f"utils/rand_tr{case + 2}/triples_sorted"]), T(golden[f"utils/rand_tr{case + 2}/relation_tail_index"])
        n = rti.size(0)
        got = orc.walk_triples(trs, rti, torch.arange(n), 7, n + 7, 21 + case, restart=False)
        assert np.array_equal(got.numpy(), golden[f"walk_tr/rand{case}/walks"])


def test_windows_match_reference(orc, golden):
    walks = T(golden["win/test_walks"])
    for k, a in enumerate(orc.to_windows(walks, 5, 30, 20)):
        assert np.array_equal(a.numpy(), golden[f"win/test_skipgram/{k}"])
    for k, a in enumerate(orc.to_windows_cbow(walks, 5, 30, 20)):
        assert np.array_equal(a.numpy(), golden[f"win/test_cbow/{k}"])
    twalks, triples = T(golden["win/test_twalks"]), T(golden["win/test_triples"])
    for k, a in enumerate(orc.to_windows_triples(twalks, 4, 30, -1, triples, 20)):
        assert np.array_equal(a.numpy(), golden[f"win/test_triples_sg/{k}"])
    for k, a in enumerate(orc.to_windows_triples_cbow(twalks, 4, 30, -1, triples, 20)):
        assert np.array_equal(a.numpy(), golden[f"win/test_triples_cbow/{k}"])
    for case, (n, wl, W) in enumerate(golden["win/rand_shapes"].tolist()):
        w, tri = T(golden[f"win/rand{case}/walks"]), T(golden[f"win/rand{case}/triples"])
        for k, a in enumerate(orc.to_windows(w, W, 50, case)):
            assert np.array_equal(a.numpy(), golden[f"win/rand{case}/skipgram/{k}"]), (case, k)
        for k, a in enumerate(orc.to_windows_cbow(w, W, 50, case)):
            assert np.array_equal(a.numpy(), golden[f"win/rand{case}/cbow/{k}"]), (case, k)
        for k, a in enumerate(orc.to_windows_triples(w, W, 50, 77, tri, case)):
            assert np.array_equal(a.numpy(), golden[f"win/rand{case}/triples_sg/{k}"]), (case, k)
        for k, a in enumerate(orc.to_windows_triples_cbow(w, W, 50, 77, tri, case)):
            assert np.array_equal(a.numpy(), golden[f"win/rand{case}/triples_cbow/{k}"]), (case, k)


def test_edge_list_biased_known_answers(orc):
    # /root/reference/tests/test_rw_edge_list.py:363-381 (restart), :414-433 (no restart), :577-597 (undirected)
    el, nei, targets, pad = _toy_edge_list(directed=True)
    walks = orc.walk_edge_list(el, nei, targets, 0.7, 0.2, 6, 20, pad)
    assert torch.equal(walks, I([[0, 2, 0, 1, 3, 2, 0], [1, 3, 2, 1, 3, 2, 1], [2, 5, 2, 5, 2, 5, 2],
                                 [3, 2, 3, 2, 3, 2, 3], [4, 0, 1, 3, 2, 4, 0]]))
    walks = orc.walk_edge_list(el, nei, targets, 0.7, 0.2, 6, 20, pad, restart=False)
    assert torch.equal(walks, I([[0, 2, 5, 5, 5, 5, 5], [1, 3, 2, 5, 5, 5, 5], [2, 5, 5, 5, 5, 5, 5],
                                 [3, 2, 5, 5, 5, 5, 5], [4, 0, 2, 5, 5, 5, 5]]))
    el, nei, targets, pad = _toy_edge_list(directed=False)
    # :246-250
    assert torch.equal(nei, I([[0, 2], [3, 5], [6, 8], [9, 11], [12, 13]]))
    # :254-275 (uniform) and :577-597 (biased)
    walks = orc.walk_edge_list(el, nei, targets, 1.0, 1.0, 6, 10, pad)
    assert torch.equal(walks, I([[0, 2, 0, 4, 3, 4, 3], [1, 0, 2, 1, 0, 4, 3], [2, 3, 4, 0, 2, 3, 1],
                                 [4, 3, 4, 0, 2, 0, 2], [3, 1, 0, 2, 0, 2, 3]]))
    walks = orc.walk_edge_list(el, nei, targets, 0.7, 0.2, 6, 20, pad)
    assert torch.equal(walks, I([[0, 2, 3, 4, 3, 2, 0], [1, 3, 2, 0, 4, 3, 2], [2, 0, 4, 3, 1, 0, 4],
                                 [4, 3, 1, 0, 4, 3, 4], [3, 4, 0, 1, 0, 4, 3]]))


def test_triple_cbow_windows_known_answers(orc):
    # /root/reference/tests/test_windows.py:243-285
    torch.manual_seed(20)
    walks = torch.randint(low=0, high=30, size=(3, 21))
    triples = torch.randint(low=0, high=30, size=(10, 3))
    pos_triples, neg_triples, pos_windows = orc.to_windows_triples_cbow(walks, 4, 30, -1, triples, 20)
    assert torch.equal(pos_triples[:2], I([[11, 10, 27], [27, 13, 24]]))
    assert torch.equal(pos_windows[:2, :2], I([[[-1, -1, 11], [-1, -1, -1]], [[10, 10, 27], [-1, -1, 11]]]))
    assert torch.equal(neg_triples[:2], I([[18, 5, 19], [7, 25, 24]]))
    assert torch.equal(pos_windows[:2, 4:], I([[[27, 13, 24], [24, 20, 13], [13, 6, 27], [27, 0, 7]],
                                               [[24, 20, 13], [13, 6, 27], [27, 0, 7], [7, 14, 20]]]))
