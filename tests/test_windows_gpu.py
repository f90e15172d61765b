"""Window generation on the GPU: positives/targets bit-exact against the oracle and the reference's
goldens (they are RNG-free); negatives checked for the reference's support and constraints."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rw():
    from torch_random_walk_b200 import rw as _rw

    return _rw


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def I(rows):
    return torch.tensor(rows, dtype=torch.int64)


def test_known_answers_of_the_reference_gpu_tests(rw):
    # /root/reference/tests/test_windows.py:58-119 (positives identical to the CPU tests :14-20, :44-51)
    torch.manual_seed(20)
    walks = torch.randint(low=0, high=30, size=(3, 10)).cuda()
    target, pos, neg = rw.to_windows(walks=walks, window_size=5, num_nodes=30, seed=20)
    assert target.size(0) == 6 * 3
    pos_expected = I([[11, 10, 13, 24], [10, 27, 24, 20], [27, 13, 20, 13], [13, 24, 13, 6], [24, 20, 6, 27],
                      [20, 13, 27, 0]]).cuda()
    assert torch.equal(target[:6], I([27, 13, 24, 20, 13, 6]).cuda())
    assert torch.equal(pos[:6], pos_expected)
    assert neg.shape == pos.shape and int(neg.min()) >= 0 and int(neg.max()) < 30
    pos_nodes, neg_nodes, windows = rw.to_windows_cbow(walks=walks, window_size=5, num_nodes=30, seed=20)
    assert torch.equal(pos_nodes[:6], I([27, 13, 24, 20, 13, 6]).cuda())
    assert torch.equal(windows[:6], pos_expected)
    assert bool((neg_nodes != pos_nodes).all())
    # :183-240 and :288-329 (triples; [10,10,27] pins the reference's head-slot quirk)
    torch.manual_seed(20)
    twalks = torch.randint(low=0, high=30, size=(3, 21)).cuda()
    triples = torch.randint(low=0, high=30, size=(10, 3)).cuda()
    tt, tp, tn = rw.to_windows_triples(walks=twalks, window_size=4, num_nodes=30, padding_idx=-1, triples=triples, seed=20)
    exp_pos = I([[[-1, -1, 11], [-1, -1, -1], [-1, -1, -1], [-1, -1, -1], [27, 13, 24], [24, 20, 13], [13, 6, 27],
                  [27, 0, 7]],
                 [[10, 10, 27], [-1, -1, 11], [-1, -1, -1], [-1, -1, -1], [24, 20, 13], [13, 6, 27], [27, 0, 7],
                  [7, 14, 20]]]).cuda()
    assert torch.equal(tt[:2], I([[11, 10, 27], [27, 13, 24]]).cuda())
    assert torch.equal(tp[:2], exp_pos)
    ct, cn, cp = rw.to_windows_triples_cbow(walks=twalks, window_size=4, num_nodes=30, padding_idx=-1, triples=triples,
                                            seed=20)
    assert torch.equal(ct[:2], I([[11, 10, 27], [27, 13, 24]]).cuda())
    assert torch.equal(cp[:2], exp_pos)


def _rows_of(triples, x):
    key = lambda t: (t[..., 0] * 1_000_003 + t[..., 1]) * 1_000_003 + t[..., 2]  # noqa: E731
    return torch.isin(key(x.reshape(-1, 3)), key(triples))


def _check_all_variants(rw, orc, walks, triples, W, num_nodes, pad, seed):
    wc, tc = walks.cuda(), triples.cuda()
    o_t, o_p, o_n = orc.to_windows(walks, W, num_nodes, seed)
    g_t, g_p, g_n = rw.to_windows(wc, W, num_nodes, seed)
    assert torch.equal(g_t.cpu(), o_t) and torch.equal(g_p.cpu(), o_p)
    assert g_n.shape == o_n.shape
    if g_n.numel():
        assert int(g_n.min()) >= 0 and int(g_n.max()) < num_nodes
    c_p, c_n, c_w = rw.to_windows_cbow(wc, W, num_nodes, seed)
    oc_p, oc_n, oc_w = orc.to_windows_cbow(walks, W, num_nodes, seed)
    assert torch.equal(c_p.cpu(), oc_p) and torch.equal(c_w.cpu(), oc_w) and c_n.shape == oc_n.shape
    if c_n.numel():
        assert int(c_n.min()) >= 0 and int(c_n.max()) < num_nodes and bool((c_n != c_p).all())
    t_t, t_p, t_n = rw.to_windows_triples(wc, W, num_nodes, pad, tc, seed)
    ot_t, ot_p, ot_n = orc.to_windows_triples(walks, W, num_nodes, pad, triples, seed)
    assert torch.equal(t_t.cpu(), ot_t) and torch.equal(t_p.cpu(), ot_p) and t_n.shape == ot_n.shape
    if t_n.numel():
        assert bool(_rows_of(tc, t_n).all())
    b_t, b_n, b_p = rw.to_windows_triples_cbow(wc, W, num_nodes, pad, tc, seed)
    ob_t, ob_n, ob_p = orc.to_windows_triples_cbow(walks, W, num_nodes, pad, triples, seed)
    assert torch.equal(b_t.cpu(), ob_t) and torch.equal(b_p.cpu(), ob_p) and b_n.shape == ob_n.shape
    if b_n.numel():
        assert bool(_rows_of(tc, b_n).all())
        assert bool((b_n != b_t).any(dim=1).all())  # never the positive triple itself
    for t in (g_t, g_p, g_n, c_p, c_n, c_w, t_t, t_p, t_n, b_t, b_n, b_p):
        assert t.dtype == torch.int64 and t.is_cuda and t.is_contiguous()


def test_positives_match_reference_goldens(rw, orc, golden):
    for case, (n, wl, W) in enumerate(golden["win/rand_shapes"].tolist()):
        walks, tri = T(golden[f"win/rand{case}/walks"]), T(golden[f"win/rand{case}/triples"])
        wc, tc = walks.cuda(), tri.cuda()
        outs = {"skipgram": rw.to_windows(wc, W, 50, case), "cbow": rw.to_windows_cbow(wc, W, 50, case),
                "triples_sg": rw.to_windows_triples(wc, W, 50, 77, tc, case),
                "triples_cbow": rw.to_windows_triples_cbow(wc, W, 50, 77, tc, case)}
        rng_free = {"skipgram": (0, 1), "cbow": (0, 2), "triples_sg": (0, 1), "triples_cbow": (0, 2)}
        for name, tensors in outs.items():
            for k, t in enumerate(tensors):
                ref = golden[f"win/rand{case}/{name}/{k}"]
                assert tuple(t.shape) == ref.shape, (case, name, k)
                if k in rng_free[name]:
                    assert np.array_equal(t.cpu().numpy(), ref), (case, name, k)


@pytest.mark.parametrize("n,wl,W", [(1, 5, 5), (2, 7, 1), (33, 81, 5), (1000, 81, 5), (257, 41, 10), (5, 3001, 7),
                                    (3, 30001, 3), (64, 13, 6), (511, 21, 4), (100, 41, 11), (70, 61, 12), (129, 81, 9), (77, 9, 2)])
def test_all_variants_against_oracle(rw, orc, n, wl, W):
    g = torch.Generator().manual_seed(n * 1000 + wl)
    walks = torch.randint(0, 500, (n, wl), generator=g)
    triples = torch.randint(0, 500, (97, 3), generator=g)
    _check_all_variants(rw, orc, walks, triples, W, 500, 777, seed=wl)


def test_deterministic_and_seeded(rw):
    walks = torch.randint(0, 100, (300, 41), device="cuda")
    a = rw.to_windows(walks, 5, 100, 1)
    b = rw.to_windows(walks, 5, 100, 1)
    c = rw.to_windows(walks, 5, 100, 2)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(a[1], c[1]) and not torch.equal(a[2], c[2])
    # negatives are roughly uniform over [0, num_nodes)
    hist = torch.bincount(a[2].flatten(), minlength=100).float()
    assert float(hist.min()) > 0.8 * float(hist.mean()) and float(hist.max()) < 1.2 * float(hist.mean())


def test_row_fast_paths_equal_the_element_path(rw):
    """W=5 rows and skip-gram negatives are written as whole 32-byte rows when the output is 32-byte
    aligned; an output shifted by one element takes the per-element path.  Same values either way."""
    import ctypes

    from torch_random_walk_b200 import native

    lib = native.lib()
    n, wl, W, nodes = 1234, 81, 5, 1 << 20
    walks = torch.randint(0, nodes, (n, wl), device="cuda")
    k = n * (wl - W + 1)
    fast = rw.to_windows(walks, W, nodes, 9)
    fast_cbow = rw.to_windows_cbow(walks, W, nodes, 9)
    # pure-torch restatement of the positives (windows_cuda.cu:40-55): the window without its middle element
    idx = torch.arange(wl - W + 1, device="cuda")[:, None] + torch.tensor([0, 1, 3, 4], device="cuda")[None, :]
    assert torch.equal(fast[1], walks[:, idx].reshape(k, W - 1)) and torch.equal(fast_cbow[2], fast[1])
    assert torch.equal(fast[0], walks[:, 2:wl - 2].reshape(k))
    bufs = [torch.empty(k * (W - 1) + 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    tgt = torch.empty(k, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.trw_windows(ctypes.c_void_p(walks.data_ptr()), n, wl, W, nodes, 9, ctypes.c_void_p(tgt.data_ptr()),
                         ctypes.c_void_p(bufs[0].data_ptr() + 8), ctypes.c_void_p(bufs[1].data_ptr() + 8), 0, st)
    assert rc == 0
    assert torch.equal(bufs[0][1:1 + k * (W - 1)].view(k, W - 1), fast[1])
    assert torch.equal(bufs[1][1:1 + k * (W - 1)].view(k, W - 1), fast[2])
    assert int(fast[2].min()) >= 0 and int(fast[2].max()) < nodes


def test_bulk_store_path_of_the_triple_windows_equals_plain_stores(rw):
    """Option win_bulk hands every warp's staged rows to the copy engine (cp.async.bulk) instead of storing them
    from the lanes; the tensors must be the same, for full tiles, a ragged last tile and odd row counts."""
    from torch_random_walk_b200 import native

    triples = torch.randint(0, 5000, (20011, 3), device="cuda")
    for n, wl, W in ((1001, 81, 5), (37, 7, 2), (5, 3, 1), (64, 21, 3)):
        walks = torch.randint(0, 5000, (n, wl), device="cuda")
        plain = rw.to_windows_triples(walks, W, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, W, 5000, 4999, triples, 3)
        native.set_option("win_bulk", 1)
        try:
            bulk = rw.to_windows_triples(walks, W, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, W, 5000, 4999, triples, 3)
        finally:
            native.set_option("win_bulk", 0)
        for a, b in zip(plain, bulk):
            assert torch.equal(a, b), (n, wl, W)


def test_direct_positive_windows_of_the_triple_kernels_equal_the_staged_ones(rw):
    """Option win_direct_pos (the default) writes the positive windows of the triple kernels straight from the walk tile,
    one fixed 16-byte piece slot per lane; off, they go through the per-warp stage like the other rows.  Same tensors for every
    window size the direct form takes (3W <= 32) and beyond, full and ragged tiles, walks shorter than a window."""
    from torch_random_walk_b200 import native

    triples = torch.randint(0, 5000, (20011, 3), device="cuda")
    for n, wl, W in ((1001, 81, 5), (37, 7, 2), (5, 3, 1), (64, 21, 3), (333, 41, 10), (50, 41, 11), (129, 5, 4), (2, 81, 7), (4097, 13, 6)):
        walks = torch.randint(0, 5000, (n, wl), device="cuda")
        assert native.get_option("win_direct_pos") == 1
        direct = rw.to_windows_triples(walks, W, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, W, 5000, 4999, triples, 3)
        native.set_option("win_direct_pos", 0)
        try:
            staged = rw.to_windows_triples(walks, W, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, W, 5000, 4999, triples, 3)
        finally:
            native.set_option("win_direct_pos", 1)
        for a, b in zip(direct, staged):
            assert torch.equal(a, b), (n, wl, W)


def test_compact_triple_table_gives_the_same_negatives(rw):
    """The negative rows come from a 16-byte uint32 copy of `triples` made inside the call; with it switched off
    (option win_table16), and for a table with an id that does not fit 32 bits, the outputs must be the same."""
    from torch_random_walk_b200 import native

    walks = torch.randint(0, 5000, (777, 41), device="cuda")
    for triples in (torch.randint(0, 5000, (30011, 3), device="cuda"),
                    torch.cat((torch.randint(0, 5000, (999, 3), device="cuda"), torch.tensor([[1, (1 << 33) + 2, 3]], device="cuda")))):
        packed = rw.to_windows_triples(walks, 5, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, 5, 5000, 4999, triples, 3)
        native.set_option("win_table16", 0)
        try:
            plain = rw.to_windows_triples(walks, 5, 5000, 4999, triples, 3) + rw.to_windows_triples_cbow(walks, 5, 5000, 4999, triples, 3)
        finally:
            native.set_option("win_table16", 1)
        for a, b in zip(packed, plain):
            assert torch.equal(a, b)
        # every negative row is a row of `triples`
        neg = packed[2].reshape(-1, 3)[:5000]
        keys = set(map(tuple, triples.tolist()))
        assert all(tuple(r) in keys for r in neg.tolist())


def test_window_outputs_stay_inside_their_tensors(rw):
    """Guard bands around every output of the four window kernels (see the walk test of the same name)."""
    import ctypes

    from torch_random_walk_b200 import native

    lib = native.lib()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    guard, sentinel = 1024, -0x0123456789ABCDEF
    walks = torch.randint(0, 300, (131, 41), device="cuda")
    triples = torch.randint(0, 300, (500, 3), device="cuda")

    def padded(numel):
        return torch.full((guard + numel + guard,), sentinel, dtype=torch.int64, device="cuda")

    def ptr(t):
        return ctypes.c_void_p(t.data_ptr() + guard * 8)

    def intact(*ts):
        torch.cuda.synchronize()
        return all(bool((t[:guard] == sentinel).all()) and bool((t[-guard:] == sentinel).all()) for t in ts)

    for W in (1, 4, 5, 7):
        k = 131 * (41 - W + 1)
        a, b, c = padded(k), padded(k * (W - 1)), padded(k * (W - 1))
        assert lib.trw_windows(ctypes.c_void_p(walks.data_ptr()), 131, 41, W, 300, 3, ptr(a), ptr(b), ptr(c), 0, st) == 0
        assert intact(a, b, c)
        ref = rw.to_windows(walks, W, 300, 3)
        assert torch.equal(b[guard:guard + k * (W - 1)].view(k, W - 1), ref[1])
        a, b, c = padded(k), padded(k), padded(k * (W - 1))
        assert lib.trw_windows_cbow(ctypes.c_void_p(walks.data_ptr()), 131, 41, W, 300, 3, ptr(a), ptr(b), ptr(c), 0, st) == 0
        assert intact(a, b, c)
    for W in (1, 3, 5):
        k = 131 * 20
        a, b, c = padded(k * 3), padded(k * 2 * W * 3), padded(k * 2 * W * 3)
        assert lib.trw_windows_triples(ctypes.c_void_p(walks.data_ptr()), 131, 41, W, 300, 999, ctypes.c_void_p(triples.data_ptr()),
                                       500, 3, ptr(a), ptr(b), ptr(c), 0, st) == 0
        assert intact(a, b, c)
        ref = rw.to_windows_triples(walks, W, 300, 999, triples, 3)
        assert torch.equal(b[guard:guard + k * 2 * W * 3].view(k, 2 * W, 3), ref[1])
        assert torch.equal(c[guard:guard + k * 2 * W * 3].view(k, 2 * W, 3), ref[2])
        a, b, c = padded(k * 3), padded(k * 3), padded(k * 2 * W * 3)
        assert lib.trw_windows_triples_cbow(ctypes.c_void_p(walks.data_ptr()), 131, 41, W, 300, 999,
                                            ctypes.c_void_p(triples.data_ptr()), 500, 3, ptr(a), ptr(b), ptr(c), 0, st) == 0
        assert intact(a, b, c)


def test_cbow_single_node_and_degenerate_triples(rw):
    walks = torch.zeros((4, 9), dtype=torch.int64, device="cuda")
    pos, neg, win = rw.to_windows_cbow(walks, 3, 1, 5)  # only node 0 exists: 101 redraws, then the same node
    assert torch.equal(neg, torch.zeros_like(neg)) and torch.equal(pos, neg)
    triples = torch.zeros((1, 3), dtype=torch.int64, device="cuda")
    pt, nt, pw = rw.to_windows_triples_cbow(walks, 2, 1, 9, triples, 5)
    assert torch.equal(nt, torch.zeros_like(nt))


def test_non_contiguous_walks_raise_like_reference(rw):
    walks = torch.randint(0, 10, (6, 20), device="cuda")[:, ::2]
    with pytest.raises(RuntimeError, match="contigous"):  # csrc/cuda/utils.cuh:9 spelling
        rw.to_windows(walks, 3, 10, 1)
    empty = torch.empty((0, 11), dtype=torch.int64, device="cuda")
    t, p, n = rw.to_windows(empty, 5, 10, 1)
    assert t.shape == (0,) and p.shape == (0, 4) and n.shape == (0, 4)


def test_negatives_from_an_alias_table_follow_the_weights(rw):
    """Extension: with neg_table= the negatives of to_windows / to_windows_cbow are drawn with
    P(v) ~ degree[v]**0.75 instead of uniformly.  Targets and positives are untouched, the draws follow the distribution
    (chi-square), nodes of weight 0 never appear, the CBOW form still avoids its positive node, and the 16-byte pair path
    writes what the per-element path writes."""
    import ctypes

    from scipy import stats

    from torch_random_walk_b200 import native

    nodes, n, wl, W = 600, 3000, 41, 5
    g = torch.Generator().manual_seed(11)
    degree = torch.randint(1, 400, (nodes,), generator=g).double()
    degree[::7] = 0.0
    table = native.negative_table(degree, power=0.75, device="cuda")
    walks = torch.randint(0, nodes, (n, wl), generator=g).cuda()
    plain = rw.to_windows(walks, W, nodes, 4)
    tgt, pos, neg = native.to_windows(walks, W, nodes, 4, neg_table=table)
    assert torch.equal(tgt, plain[0]) and torch.equal(pos, plain[1]) and neg.shape == plain[2].shape
    assert torch.equal(neg, native.to_windows(walks, W, nodes, 4, neg_table=table)[2])            # deterministic
    assert not torch.equal(neg, native.to_windows(walks, W, nodes, 5, neg_table=table)[2])        # seeded
    want = degree.numpy() ** 0.75
    want[degree.numpy() == 0] = 0.0
    want = want / want.sum()

    def follows(sample, from_table=True):
        counts = np.bincount(sample.cpu().numpy().ravel(), minlength=nodes).astype(np.float64)
        if from_table:
            assert counts[want == 0].sum() == 0  # a node of weight 0 is never drawn
        keep = want > 0
        return stats.chisquare(counts[keep], want[keep] * counts[keep].sum()).pvalue

    assert follows(neg) > 1e-3
    assert follows(plain[2], from_table=False) < 1e-6  # (the uniform negatives do not: the test can tell the two apart)
    pos_nodes, neg_nodes, windows = native.to_windows_cbow(walks, W, nodes, 4, neg_table=table)
    plain_cbow = rw.to_windows_cbow(walks, W, nodes, 4)
    assert torch.equal(pos_nodes, plain_cbow[0]) and torch.equal(windows, plain_cbow[2])
    assert not bool((neg_nodes == pos_nodes).any())
    assert follows(neg_nodes) > 1e-4  # (conditioned on != the positive node: a small, second-order distortion)
    # element path (output shifted by one element) == pair path
    lib = native.lib()
    k = n * (wl - W + 1)
    bufs = [torch.empty(k * (W - 1) + 8, dtype=torch.int64, device="cuda") for _ in range(2)]
    out_t = torch.empty(k, dtype=torch.int64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.trw_windows_alias(ctypes.c_void_p(walks.data_ptr()), n, wl, W, nodes, 4, ctypes.c_void_p(table.data_ptr()),
                               ctypes.c_void_p(out_t.data_ptr()), ctypes.c_void_p(bufs[0].data_ptr() + 8),
                               ctypes.c_void_p(bufs[1].data_ptr() + 8), 0, st)
    assert rc == 0, lib.trw_last_error()
    assert torch.equal(bufs[1][1:1 + k * (W - 1)].view(k, W - 1), neg)
    with pytest.raises(RuntimeError):
        native.to_windows(walks, W, nodes + 1, 4, neg_table=table)
