"""Host-side graph preparation (torch_random_walk_b200.utils) against the reference's own
torch_rw/utils.py outputs (tests/golden/ref_golden.npz) and its tests' known answers.  CPU only."""
import networkx as nx
import numpy as np
import pytest
import torch

from helpers import toy_graph
from torch_random_walk_b200 import utils


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name,graph", [("toy_undirected", toy_graph(False)), ("toy_directed", toy_graph(True)),
                                        ("karate", nx.karate_club_graph())])
def test_to_csr_and_nodes_tensor(golden, name, graph):
    row_ptr, col_idx = utils.to_csr(graph)
    nodes = utils.nodes_tensor(graph)
    for t in (row_ptr, col_idx, nodes):
        assert t.dtype == torch.int64 and t.is_contiguous() and t.device.type == "cpu"
    assert np.array_equal(row_ptr.numpy(), golden[f"utils/{name}/row_ptr"])
    assert np.array_equal(col_idx.numpy(), golden[f"utils/{name}/col_idx"])
    assert np.array_equal(nodes.numpy(), golden[f"utils/{name}/nodes"])


def test_karate_is_baseline_config_1(golden):
    # BASELINE.json configs[0]: 34 nodes, 156 CSR entries
    assert golden["utils/karate/row_ptr"].shape == (35,) and golden["utils/karate/col_idx"].shape == (156,)


def test_to_csr_exact_above_2_pow_24():
    # the reference's float32 round trip (torch_rw/utils.py:7-8) loses ids above 2^24; ours must not
    g = nx.Graph()
    g.add_nodes_from(range(3))
    g.add_edge(0, 2)
    row_ptr, col_idx = utils.to_csr(g)
    assert row_ptr.tolist() == [0, 1, 1, 2] and col_idx.tolist() == [2, 0]


@pytest.mark.parametrize("name,directed", [("toy_undirected", False), ("toy_directed", True)])
def test_edge_list_indexed_and_node_edge_index(golden, name, directed):
    el, mapping = utils.to_edge_list_indexed(toy_graph(directed))
    assert np.array_equal(el.numpy(), golden[f"utils/{name}/edge_list"])
    assert list(mapping.keys()) == golden[f"utils/{name}/mapping_keys"].tolist()
    assert list(mapping.values()) == golden[f"utils/{name}/mapping_values"].tolist()
    nei, el_sorted = utils.build_node_edge_index(el, torch.unique(el.view(-1)))
    assert np.array_equal(nei.numpy(), golden[f"utils/{name}/node_edge_index"])
    assert np.array_equal(el_sorted.numpy(), golden[f"utils/{name}/edge_list_sorted"])


def test_node_edge_index_known_answers():
    # /root/reference/tests/test_rw_edge_list.py:31-35 and :246-250
    el, _ = utils.to_edge_list_indexed(toy_graph(True))
    nei, _ = utils.build_node_edge_index(el, torch.unique(el.view(-1)))
    assert nei.tolist() == [[0, 1], [2, 3], [-1, -1], [4, 4], [5, 6]]
    el, _ = utils.to_edge_list_indexed(toy_graph(False))
    nei, _ = utils.build_node_edge_index(el, torch.unique(el.view(-1)))
    assert nei.tolist() == [[0, 2], [3, 5], [6, 8], [9, 11], [12, 13]]


@pytest.mark.parametrize("case", range(4))
def test_random_indices_match_reference(golden, case):
    el = T(golden[f"utils/rand_el{case}/edge_list"])
    n = int(golden[f"utils/rand_el{case}/num_nodes"])
    nei, el_sorted = utils.build_node_edge_index(el, torch.arange(n))
    assert np.array_equal(nei.numpy(), golden[f"utils/rand_el{case}/node_edge_index"])
    assert np.array_equal(el_sorted.numpy(), golden[f"utils/rand_el{case}/edge_list_sorted"])
    tr = T(golden[f"utils/rand_tr{case}/triples"])
    rti, tr_sorted = utils.build_relation_tail_index(tr, torch.arange(n))
    assert np.array_equal(rti.numpy(), golden[f"utils/rand_tr{case}/relation_tail_index"])
    assert np.array_equal(tr_sorted.numpy(), golden[f"utils/rand_tr{case}/triples_sorted"])


def test_relation_tail_index_known_answer(golden):
    # /root/reference/tests/test_rw_triples.py:26-53 (float32 triples in, int64 out)
    triples = T(golden["utils/toy_triples/triples"])
    assert triples.dtype == torch.float32
    rti, tr_sorted = utils.build_relation_tail_index(triples, T(golden["utils/toy_triples/entities"]))
    assert rti.tolist() == [[0, 2], [3, 3], [4, 5], [6, 7], [-1, -1]]
    assert tr_sorted.dtype == torch.int64
    assert np.array_equal(tr_sorted.numpy(), golden["utils/toy_triples/triples_sorted"])


def test_empty_edge_list_raises_like_reference():
    with pytest.raises(IndexError):
        utils.build_node_edge_index(torch.zeros((0, 2), dtype=torch.int64), torch.arange(3))


# ---- vectorised / device-side graph preparation --------------------

def _same_rows_per_head(a, b):
    """Two row lists sorted by head hold the same rows per head (order inside a head may differ)."""
    key = lambda t: sorted(map(tuple, t.tolist()))  # noqa: E731
    return a.shape == b.shape and torch.equal(a[:, 0], b[:, 0]) and key(a) == key(b)


@pytest.mark.parametrize("case", range(4))
def test_torch_index_builders_match_reference(golden, case):
    from torch_random_walk_b200.utils import _sorted_rows_and_ranges_torch

    n = int(golden[f"utils/rand_el{case}/num_nodes"])
    index, rows = _sorted_rows_and_ranges_torch(T(golden[f"utils/rand_el{case}/edge_list"]), n)
    assert np.array_equal(index.numpy(), golden[f"utils/rand_el{case}/node_edge_index"])
    assert _same_rows_per_head(rows, T(golden[f"utils/rand_el{case}/edge_list_sorted"]))
    index, rows = _sorted_rows_and_ranges_torch(T(golden[f"utils/rand_tr{case}/triples"]), n)
    assert np.array_equal(index.numpy(), golden[f"utils/rand_tr{case}/relation_tail_index"])
    assert _same_rows_per_head(rows, T(golden[f"utils/rand_tr{case}/triples_sorted"]))


def test_torch_index_builder_known_answers(golden):
    from torch_random_walk_b200.utils import _sorted_rows_and_ranges_torch

    # /root/reference/tests/test_rw_edge_list.py:31-35, tests/test_rw_triples.py:47-51
    el, _ = utils.to_edge_list_indexed(toy_graph(True))
    index, _ = _sorted_rows_and_ranges_torch(el, 5)
    assert index.tolist() == [[0, 1], [2, 3], [-1, -1], [4, 4], [5, 6]]
    index, rows = _sorted_rows_and_ranges_torch(T(golden["utils/toy_triples/triples"]), 5)
    assert index.tolist() == [[0, 2], [3, 3], [4, 5], [6, 7], [-1, -1]]
    assert rows.dtype == torch.int64
    with pytest.raises(IndexError):
        _sorted_rows_and_ranges_torch(torch.zeros((0, 2), dtype=torch.int64), 3)


@pytest.mark.parametrize("name,graph,symmetric", [("toy_undirected", toy_graph(False), True), ("toy_directed", toy_graph(True), False),
                                                   ("karate", nx.karate_club_graph(), True)])
def test_csr_from_edge_index_equals_to_csr(golden, name, graph, symmetric):
    nodes = list(graph.nodes())
    pos = {v: i for i, v in enumerate(nodes)}
    edges = torch.tensor([[pos[a], pos[b]] for a, b in graph.edges()], dtype=torch.int64)
    for layout in (edges, edges.t().contiguous()):
        row_ptr, col_idx = utils.csr_from_edge_index(layout, len(nodes), symmetric=symmetric)
        assert np.array_equal(row_ptr.numpy(), golden[f"utils/{name}/row_ptr"])
        assert np.array_equal(col_idx.numpy(), golden[f"utils/{name}/col_idx"])
    dup = torch.cat((edges, edges[:3]))  # duplicate edges are merged, as scipy's canonical CSR does
    row_ptr, col_idx = utils.csr_from_edge_index(dup, len(nodes), symmetric=symmetric)
    assert np.array_equal(col_idx.numpy(), golden[f"utils/{name}/col_idx"])
    with pytest.raises(IndexError):
        utils.csr_from_edge_index(torch.tensor([[0, len(nodes)]]), len(nodes))


def test_csr_files_round_trip_and_compact(tmp_path):
    """f4 of SURVEY section 8: a CSR graph kept on disk ('.npz' / '.pt'), int32 where the values allow it, and an
    edge-list file turned into the same CSR that utils.csr_from_edge_index gives."""
    rng = np.random.default_rng(5)
    edges = torch.from_numpy(rng.integers(0, 500, (4000, 2)))
    row_ptr, col_idx = utils.csr_from_edge_index(edges, 500, symmetric=True)
    for ext in (".npz", ".pt"):
        path = str(tmp_path / ("graph" + ext))
        utils.save_csr(path, row_ptr, col_idx)
        rp, ci = utils.load_csr(path)
        assert rp.dtype == torch.int32 and ci.dtype == torch.int32  # compact by default
        assert torch.equal(rp.long(), row_ptr) and torch.equal(ci.long(), col_idx)
        rp64, ci64 = utils.load_csr(path, dtype=torch.int64)
        assert rp64.dtype == torch.int64 and torch.equal(rp64, row_ptr) and torch.equal(ci64, col_idx)
        utils.save_csr(path, row_ptr, col_idx, compact=False)
        assert utils.load_csr(path)[1].dtype == torch.int64
    big = col_idx.clone()
    big[3] = 2 ** 31 + 5
    assert utils.compact_csr(row_ptr, big)[1].dtype == torch.int64 and utils.compact_csr(row_ptr, big)[0].dtype == torch.int32
    np.save(str(tmp_path / "edges.npy"), edges.numpy())
    torch.save({"edge_index": edges.t().contiguous()}, str(tmp_path / "edges.pt"))
    for name in ("edges.npy", "edges.pt"):
        rp, ci = utils.load_edge_index(str(tmp_path / name), num_nodes=500, symmetric=True)
        assert torch.equal(rp, row_ptr) and torch.equal(ci, col_idx)
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, row_ptr=np.array([0, 3]), col_idx=np.array([1]))
    try:
        utils.load_csr(bad)
        raise AssertionError("a malformed file must be refused")
    except ValueError:
        pass
