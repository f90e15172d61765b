"""Edge-list and triple walks on the GPU against the reference's conventions: RNG-free rows are
bit-exact with the reference's goldens, the rest is checked structurally and statistically against
the oracle (for the second-order edge-list walk the oracle IS the specification: its acceptance
rule deviates from analytic node2vec)."""
import numpy as np
import pytest
import torch

from helpers import chi2_pvalue, second_order_counts, toy_graph, two_sample_chi2
from torch_random_walk_b200 import utils

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rw():
    from torch_random_walk_b200 import rw as _rw

    return _rw


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _toy(directed):
    el, mapping = utils.to_edge_list_indexed(toy_graph(directed))
    targets = torch.tensor(list(mapping.values()), dtype=torch.int64)
    nei, el_sorted = utils.build_node_edge_index(el, torch.unique(el.view(-1)))
    return el_sorted, nei, targets, sorted(targets.tolist())[-1] + 1


def _check_edge_list_walks(walks, el, nei, targets, pad, restart):
    w = walks.cpu().numpy()
    el, nei = el.numpy(), nei.numpy()
    succ = {}
    for h, t in el.tolist():
        succ.setdefault(h, set()).add(t)
    assert np.array_equal(w[:, 0], targets.numpy())
    for row in w:
        jump = row[0] if restart else pad
        for a, b in zip(row[:-1], row[1:]):
            if a == pad:
                assert b == jump
            elif a not in succ:
                assert b == pad
            else:
                assert b in succ[a]


def test_uniform_edge_list_known_rows(rw):
    # /root/reference/tests/test_rw_edge_list.py:115-223: rows of nodes 2 (no out-edge) and 3 (one out-edge) are RNG-free
    el, nei, targets, pad = _toy(directed=True)
    walks = rw.walk_edge_list(edge_list_indexed=el.cuda(), node_edge_index=nei.cuda(), target_nodes=targets.cuda(),
                              p=1.0, q=1.0, walk_length=6, seed=10, padding_idx=pad)
    assert walks.shape == (5, 7)
    assert walks[2].tolist() == [2, 5, 2, 5, 2, 5, 2] and walks[3].tolist() == [3, 2, 5, 3, 2, 5, 3]
    _check_edge_list_walks(walks, el, nei, targets, pad, True)
    walks = rw.walk_edge_list(el.cuda(), nei.cuda(), targets.cuda(), 1.0, 1.0, 6, 10, pad, restart=False)
    assert walks[2].tolist() == [2, 5, 5, 5, 5, 5, 5] and walks[3].tolist() == [3, 2, 5, 5, 5, 5, 5]
    _check_edge_list_walks(walks, el, nei, targets, pad, False)
    for row in walks.tolist():  # no restart: once padded, padded for ever
        if pad in row:
            k = row.index(pad)
            assert all(x == pad for x in row[k:])


@pytest.mark.parametrize("restart", [True, False])
def test_biased_edge_list_known_rows_and_structure(rw, restart):
    # :436-546: with restart the biased walk emits the start node instead of the padding
    el, nei, targets, pad = _toy(directed=True)
    walks = rw.walk_edge_list(el.cuda(), nei.cuda(), targets.cuda(), 0.7, 0.2, 6, 20, pad, restart=restart)
    if restart:
        assert walks[2].tolist() == [2, 5, 2, 5, 2, 5, 2] and walks[3].tolist() == [3, 2, 3, 2, 3, 2, 3]
    else:
        assert walks[2].tolist() == [2, 5, 5, 5, 5, 5, 5] and walks[3].tolist() == [3, 2, 5, 5, 5, 5, 5]


def test_undirected_edge_list_structure(rw):
    el, nei, targets, pad = _toy(directed=False)
    for p, q in ((1.0, 1.0), (0.7, 0.2)):
        walks = rw.walk_edge_list(el.cuda(), nei.cuda(), targets.cuda(), p, q, 6, 10, pad)
        assert int(walks.max()) < pad
        _check_edge_list_walks(walks, el, nei, targets, pad, True)


@pytest.mark.parametrize("p,q,restart", [(0.7, 0.2, True), (0.5, 2.0, True), (2.0, 0.5, False)])
def test_biased_edge_list_statistics_match_oracle(rw, orc, p, q, restart):
    g = torch.Generator().manual_seed(5)
    n = 24
    el = torch.stack((torch.randint(0, n - 2, (120,), generator=g), torch.randint(0, n, (120,), generator=g)), 1)
    nei, el_sorted = utils.build_node_edge_index(el, torch.arange(n + 1))  # nodes n-2, n-1: dead ends; row n: padding
    pad = n
    targets = torch.arange(n).repeat_interleave(4000)
    L = 12
    got = rw.walk_edge_list(el_sorted.cuda(), nei.cuda(), targets.cuda(), p, q, L, 3, pad, restart=restart)
    ref = orc.walk_edge_list(el_sorted, nei, targets, p, q, L, 3, pad, restart)
    assert got.shape == ref.shape
    a, b = second_order_counts(got, n + 1), second_order_counts(ref, n + 1)
    chi2, dof = two_sample_chi2(a, b, n + 1)
    assert dof > 50
    assert chi2_pvalue(chi2, dof) > 0.01, (chi2, dof)
    # first-order marginals as well
    f = lambda w: dict(zip(*np.unique(w.cpu().numpy()[:, 1].astype(np.int64) + (n + 1) * w.cpu().numpy()[:, 0], return_counts=True)))  # noqa: E731
    fa, fb = f(got), f(ref)
    assert set(fa) == set(fb)


@pytest.mark.parametrize("duplicates,sort_tails", [(False, True), (False, False), (True, False)])
def test_edge_list_table_path_equals_the_reference_scan(rw, duplicates, sort_tails):
    """The second-order edge-list walk answers "x in adj(t)" from the hashed table when it has a
    workspace; the reference scans [first, last) -- the last out-edge excluded.  Same draws, so the
    walks must be bit-identical, on long rows (hashed), short rows (packed), unsorted tails and
    tails stored twice (where "x is the last tail" has to count the earlier copy)."""
    from helpers import random_csr
    from torch_random_walk_b200 import native

    rp, ci = random_csr(31, 1500, 30, sort_rows=sort_tails)
    rp_np, ci_np = rp.numpy(), ci.numpy().copy()
    if duplicates:  # the last tail of every other row also appears in the row's first slot, or not at all elsewhere
        for v in range(0, 1500, 2):
            if rp_np[v + 1] - rp_np[v] >= 3:
                ci_np[rp_np[v]] = ci_np[rp_np[v + 1] - 1]
    n = 1500
    heads = np.repeat(np.arange(n), np.diff(rp_np))
    el = T(np.stack((heads, ci_np), 1)).cuda()
    nei, el = utils.build_node_edge_index(el, torch.arange(n))
    nodes = torch.arange(n, device="cuda")
    for p, q, restart in ((0.5, 2.0, True), (1.0, 0.5, False), (0.25, 4.0, True), (2.0, 0.5, True)):
        a = rw.walk_edge_list(el, nei, nodes, p, q, 30, 4, n, restart=restart)
        native.set_option("el_table", 0)
        try:
            b = rw.walk_edge_list(el, nei, nodes, p, q, 30, 4, n, restart=restart)
        finally:
            native.set_option("el_table", 1)
        assert torch.equal(a, b), (p, q, restart)
    # an index that does not describe the edge list must not be trusted: the walk falls back to the scan
    bad = nei.clone()
    rows_with_edges = torch.nonzero(bad[:, 0] >= 0).flatten()
    bad[rows_with_edges[5], 1] -= 1  # node loses its last edge in the index only
    a = rw.walk_edge_list(el, bad, nodes, 0.5, 2.0, 20, 4, n)
    native.set_option("el_table", 0)
    try:
        b = rw.walk_edge_list(el, bad, nodes, 0.5, 2.0, 20, 4, n)
    finally:
        native.set_option("el_table", 1)
    assert torch.equal(a, b)


def test_edge_list_matches_reference_golden_dead_ends(rw, golden):
    # rows that never branch are RNG-free: compare them with the reference's own output
    for case in range(2):
        p, q, L, seed, pad, restart = golden[f"walk_el/rand{case}/params"]
        el, nei = T(golden[f"walk_el/rand{case}/edge_list_sorted"]), T(golden[f"walk_el/rand{case}/node_edge_index"])
        ref = golden[f"walk_el/rand{case}/walks"]
        got = rw.walk_edge_list(el.cuda(), nei.cuda(), torch.arange(int(pad)).cuda(), float(p), float(q), int(L),
                                int(seed), int(pad), bool(restart)).cpu().numpy()
        count = (nei[:, 1] - nei[:, 0] + 1).numpy()
        count[nei[:, 0].numpy() < 0] = 0
        checked = 0
        for row_ref, row_got in zip(ref, got):
            # deterministic as long as every visited node has at most one out-edge
            det = True
            for k, v in enumerate(row_ref):
                if not det:
                    break
                assert row_got[k] == v
                checked += 1
                if v != int(pad) and count[v] > 1:
                    det = False
        assert checked > len(ref)


def test_triple_walk_known_rows_and_structure(rw, golden):
    # /root/reference/tests/test_rw_triples.py:84-159; entity 4 has no outgoing triple -> all padding
    rti, trs = T(golden["utils/toy_triples/relation_tail_index"]), T(golden["utils/toy_triples/triples_sorted"])
    targets = T(golden["utils/toy_triples/entities"]).repeat_interleave(2, 0)
    walks = rw.walk_triples(triples_indexed=trs.cuda(), relation_tail_index=rti.cuda(), target_nodes=targets.cuda(),
                            walk_length=6, seed=10, padding_idx=8, restart=False)
    assert walks.shape == (10, 13)
    assert walks[8].tolist() == [4] + [8] * 12 and walks[9].tolist() == [4] + [8] * 12
    triple_set = set(map(tuple, trs.tolist()))
    for row in walks.tolist():
        for k in range(0, 12, 2):
            h, r, t = row[k], row[k + 1], row[k + 2]
            if h == 8 or not any(x[0] == h for x in triple_set):
                assert (r, t) == (8, 8)
            else:
                assert (h, r, t) in triple_set


def test_triple_walk_statistics_match_oracle(rw, orc, golden):
    trs, rti = T(golden["utils/rand_tr3/triples_sorted"]), T(golden["utils/rand_tr3/relation_tail_index"])
    n = rti.size(0)
    pad = n + 7
    targets = torch.arange(n).repeat_interleave(300)
    got = rw.walk_triples(trs.cuda(), rti.cuda(), targets.cuda(), 5, pad, 9, restart=False).cpu().numpy()
    ref = orc.walk_triples(trs, rti, targets, 5, pad, 9, restart=False).numpy()
    # first hop: (head, rel, tail) frequencies
    key = lambda w: dict(zip(*np.unique((w[:, 0] * (pad + 1) + w[:, 1]) * (pad + 1) + w[:, 2], return_counts=True)))  # noqa: E731
    a, b = key(got), key(ref)
    assert set(a) == set(b)
    cells = sorted(a)
    ca, cb = np.array([a[c] for c in cells], float), np.array([b[c] for c in cells], float)
    chi2 = float(((ca - cb) ** 2 / (ca + cb)).sum())
    assert chi2_pvalue(chi2, len(cells) - n) > 0.01
    # padding propagates
    for row in got[:2000]:
        if pad in row[1:]:
            k = list(row[1:]).index(pad) + 1
            assert all(x == pad for x in row[k:])


def test_c4_shape_triple_walks_and_windows_full_size(rw, orc):
    """BASELINE.json configs[3]: FB15k-237-shaped triples (14,541 entities, 237 relations, 310,116
    triples), 10 walks per entity, 40 hops, then to_windows_triples(window_size=5).  Full size through
    size-independent properties; a slice of the windows bit-exact against the oracle."""
    from torch_random_walk_b200 import rmat

    n_ent, n_rel, n_tr = 14541, 237, 310116
    triples = rmat.kg_triples(n_ent, n_rel, n_tr, seed=7, device="cuda")
    index, ts = rmat.relation_tail_index(triples, n_ent)
    pad = n_ent + n_rel
    targets = torch.arange(n_ent, device="cuda").repeat_interleave(10)
    L, W = 40, 5
    walks = rw.walk_triples(ts, index, targets, walk_length=L, padding_idx=pad, seed=10)
    assert walks.shape == (n_ent * 10, 2 * L + 1)
    w = walks.cpu()
    assert torch.equal(w[:, 0], targets.cpu())
    # every (head, rel, tail) hop is a row of `triples`, or padding after a dead end, and padding absorbs
    key = lambda h, r, t: (h * (pad + 1) + r) * (pad + 1) + t  # noqa: E731
    valid = torch.unique(key(ts[:, 0], ts[:, 1], ts[:, 2])).cpu()
    heads, rels, tails = w[:, 0:-2:2], w[:, 1::2], w[:, 2::2]
    hop_key = key(heads, rels, tails).reshape(-1)
    is_pad = ((rels == pad) & (tails == pad)).reshape(-1)
    has_out = (index[:, 0] >= 0).cpu()
    head_flat = heads.reshape(-1)
    head_dead = (head_flat == pad) | ~has_out[head_flat.clamp(max=n_ent - 1)]
    assert bool((torch.isin(hop_key, valid) | is_pad).all())
    assert bool((is_pad == head_dead).all())  # padding exactly where the reference pads (rw_cuda_triples.cu:23-43)
    # windows at full size: shapes, targets are the walk's own triples, negatives are rows of `triples`
    tt, tp, tn = rw.to_windows_triples(walks, W, n_ent, pad, ts, 3)
    K = n_ent * 10 * L
    assert tt.shape == (K, 3) and tp.shape == (K, 2 * W, 3) and tn.shape == (K, 2 * W, 3)
    assert torch.equal(tt.view(n_ent * 10, L, 3)[:, :, 0], walks[:, 0:-2:2])
    assert torch.equal(tt.view(n_ent * 10, L, 3)[:, :, 2], walks[:, 2::2])
    assert bool(torch.isin(key(tn[..., 0], tn[..., 1], tn[..., 2]).reshape(-1), valid.cuda()).all())
    # right-hand windows of hop j are the targets of hops j+1.. (size-independent shift property)
    right = tp.view(n_ent * 10, L, 2 * W, 3)[:, :-1, W, :]
    assert torch.equal(right, tt.view(n_ent * 10, L, 3)[:, 1:, :])
    # a slice bit-exact against the oracle (positives and targets are RNG-free)
    sl = walks[:2000].cpu()
    o_t, o_p, _ = orc.to_windows_triples(sl, W, n_ent, pad, ts.cpu(), 3)
    assert torch.equal(tt[: 2000 * L].cpu(), o_t) and torch.equal(tp[: 2000 * L].cpu(), o_p)


def test_device_side_index_builders_feed_the_walks(rw, golden):
    """utils.build_node_edge_index / build_relation_tail_index / csr_from_edge_index on CUDA tensors
    (SURVEY.md section 8f rank 2): same index as the host path, and the walks accept the result."""
    el = T(golden["utils/rand_el3/edge_list"])
    n = int(golden["utils/rand_el3/num_nodes"])
    idx_gpu, rows_gpu = utils.build_node_edge_index(el.cuda(), torch.arange(n))
    assert idx_gpu.is_cuda and np.array_equal(idx_gpu.cpu().numpy(), golden["utils/rand_el3/node_edge_index"])
    walks = rw.walk_edge_list(rows_gpu, idx_gpu, torch.arange(n, device="cuda"), 1.0, 1.0, 8, 3, n)
    _check_edge_list_walks(walks, rows_gpu.cpu(), idx_gpu.cpu(), torch.arange(n), n, True)
    tr = T(golden["utils/rand_tr3/triples"])
    rti_gpu, trs_gpu = utils.build_relation_tail_index(tr.cuda(), torch.arange(n))
    assert np.array_equal(rti_gpu.cpu().numpy(), golden["utils/rand_tr3/relation_tail_index"])
    w = rw.walk_triples(trs_gpu, rti_gpu, torch.arange(n, device="cuda"), 4, n + 7, 1)
    assert w.shape == (n, 9)
    rp, ci = utils.csr_from_edge_index(el.cuda(), n, symmetric=True)
    w = rw.walk(rp, ci, torch.arange(n, device="cuda"), 0.5, 2.0, 10, 1)
    from helpers import check_walks_follow_edges

    check_walks_follow_edges(w, rp, ci, torch.arange(n))
