"""Parity of the CUDA CSR walk (rw.walk -> trw_walk_csr) with the reference's algorithm.

The reference draws from a different RNG, so parity is what BASELINE.json's north_star defines:
structure (column 0, every transition an edge, the dead-end convention), reproducibility, and
second-order transition statistics against BOTH the analytic node2vec law and the oracle's own
empirical frequencies (chi-square p > 0.01, pooled total variation < 1e-2 at 1e7 samples)."""
import numpy as np
import pytest
import torch

from helpers import (check_walks_follow_edges, chi2_and_tv, chi2_pvalue, node2vec_probs, random_csr,
                     second_order_counts, two_sample_chi2)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rw():
    from torch_random_walk_b200 import rw as _rw

    return _rw


@pytest.fixture(scope="module")
def native():
    from torch_random_walk_b200 import native as _n

    return _n


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def cuda(*ts):
    return [t.cuda() for t in ts]


def test_native_library_is_the_one_loaded(native):
    maps = open("/proc/self/maps").read()
    assert "libtrw_b200.so" in maps
    assert native.lib().trw_device_check(0) == 0


def test_toy_graph_shapes_and_structure(rw, golden):
    # the call of /root/reference/tests/test_rw.py:79 and :146 (uniform / biased on GPU)
    rp, ci, nodes = cuda(T(golden["utils/toy_undirected/row_ptr"]), T(golden["utils/toy_undirected/col_idx"]),
                         torch.arange(5))
    for p, q in ((1.0, 1.0), (0.7, 0.5)):
        walks = rw.walk(row_ptr=rp, col_idx=ci, target_nodes=nodes, p=p, q=q, walk_length=6, seed=10)
        assert walks.shape == (5, 7) and walks.dtype == torch.int64 and walks.is_cuda and walks.is_contiguous()
        check_walks_follow_edges(walks, rp, ci, nodes)


def test_karate_config1(rw, golden):
    # BASELINE.json configs[0]: karate club, p=q=1, walk_length=80, 10 walks per node
    rp, ci = cuda(T(golden["utils/karate/row_ptr"]), T(golden["utils/karate/col_idx"]))
    nodes = T(golden["utils/karate/nodes"]).repeat_interleave(10).cuda()
    walks = rw.walk(rp, ci, nodes, 1.0, 1.0, 80, 10)
    assert walks.shape == golden["walk/karate_uniform_L80"].shape == (340, 81)
    check_walks_follow_edges(walks, rp, ci, nodes)


@pytest.mark.parametrize("p,q", [(1.0, 1.0), (0.5, 2.0), (1.0, 0.5), (0.25, 4.0), (2.0, 1.0)])
def test_reproducible_and_seed_sensitive(rw, p, q):
    rp, ci = cuda(*random_csr(1, 3000, 24))
    nodes = torch.arange(3000, device="cuda")
    a = rw.walk(rp, ci, nodes, p, q, 40, 123)
    b = rw.walk(rp, ci, nodes, p, q, 40, 123)
    c = rw.walk(rp, ci, nodes, p, q, 40, 124)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    check_walks_follow_edges(a, rp, ci, nodes)


@pytest.mark.parametrize("p,q", [(1.0, 1.0), (0.5, 2.0)])
def test_shards_are_bit_identical_to_one_call(native, p, q):
    # what makes 1/2/4/8-GPU output identical: Philox is keyed by the GLOBAL walk id
    rp, ci = cuda(*random_csr(2, 5000, 20))
    nodes = torch.randint(0, 5000, (20001,), device="cuda")
    full = native.walk(rp, ci, nodes, p, q, 30, 7)
    for world in (2, 3, 8):
        from torch_random_walk_b200.dist import shard_bounds

        parts = []
        for r in range(world):
            lo, hi = shard_bounds(nodes.numel(), r, world)
            parts.append(native.walk(rp, ci, nodes[lo:hi].contiguous(), p, q, 30, 7, walk_id_offset=lo))
        assert torch.equal(torch.cat(parts), full), world


DEFAULTS = {"n2v_table": 1, "n2v_speculate": -1, "stage_output": 1, "row32": 1, "build_mode": 2, "n2v_min_ctas": -1, "n2v_fold": 1,
            "records": -1, "n2v_slots": 8, "edge_filter_mb": 32}


def test_kernel_variants_agree_bit_for_bit(native):
    """Hashed membership table vs the reference's linear scan, cooperative vs flat table build,
    uint32 vs int64 row index, speculative row fetch on/off, staged vs plain stores and every
    register budget must not change a single entry: same draws, same decisions."""
    rp, ci = random_csr(3, 4000, 60)  # mean degree 60: most rows use the table (>= 12 neighbours)
    rp, ci = cuda(rp, ci)
    nodes = torch.arange(4000, device="cuda")
    variants = [{}, {"records": 1}, {"records": 0}, {"records": 0, "n2v_table": 0}, {"records": 1, "n2v_table": 0},
                {"records": 1, "n2v_min_ctas": 4}, {"records": 1, "n2v_slots": 16}, {"n2v_min_ctas": 5}, {"smem_carveout_kb": 64}, {"n2v_table": 0}, {"n2v_speculate": 0}, {"n2v_speculate": 1}, {"stage_output": 0}, {"row32": 0},
                {"build_mode": 0}, {"n2v_min_ctas": 5}, {"n2v_min_ctas": 6}, {"edge_filter_mb": 0}, {"edge_filter_mb": 1},
                {"edge_filter_mb": 0, "records": 1}, {"edge_filter_mb": 1, "records": 0},
                {"n2v_table": 0, "row32": 0, "stage_output": 0}, {"build_mode": 0, "row32": 0, "n2v_min_ctas": 6}]
    base = None
    try:
        for opts in variants:
            for k, v in {**DEFAULTS, **opts}.items():
                native.set_option(k, v)
            w = native.walk(rp, ci, nodes, 0.5, 2.0, 25, 99)
            w2 = native.walk(rp, ci, nodes, 1.0, 0.5, 25, 99)
            u = native.walk(rp, ci, nodes, 1.0, 1.0, 25, 99)
            if base is None:
                base = (w, w2, u)
            assert torch.equal(w, base[0]) and torch.equal(w2, base[1]) and torch.equal(u, base[2]), opts
    finally:
        for k, v in DEFAULTS.items():
            native.set_option(k, v)


def test_table_build_on_skewed_graph_matches_scan(native):
    """The three table builds (tiled shared-memory, cooperative chunks, flat) against each other and
    against the linear scan, on a graph with hubs, many empty rows and many tiles."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(17, 16, device="cuda", seed=5)
    deg = rp[1:] - rp[:-1]
    nodes = torch.nonzero(deg > 0).flatten()[:50000].contiguous()
    assert int(deg.max()) >= 2048  # hub rows take the global-CAS path of the tiled build
    a = native.walk(rp, ci, nodes, 0.5, 2.0, 12, 1)
    try:
        # a filter so small that it is saturated ("maybe" for nearly every pair), and none at all
        native.set_option("edge_filter_mb", 1)
        assert torch.equal(a, native.walk(rp, ci, nodes, 0.5, 2.0, 12, 1))
        native.set_option("edge_filter_mb", 0)
        assert torch.equal(a, native.walk(rp, ci, nodes, 0.5, 2.0, 12, 1))
        native.set_option("edge_filter_mb", 64)
        native.set_option("build_mode", 0)
        b = native.walk(rp, ci, nodes, 0.5, 2.0, 12, 1)
        native.set_option("n2v_table", 0)
        c = native.walk(rp, ci, nodes[:3000].contiguous(), 0.5, 2.0, 12, 1)
    finally:
        native.set_option("build_mode", 2)
        native.set_option("n2v_table", 1)
        native.set_option("edge_filter_mb", 64)
    assert torch.equal(a, b)
    assert torch.equal(a[:3000], c)


def test_edge_records_match_row_index_path(native):
    """Edge records (the proposal gather returns the neighbour's row span) vs row-index lookups, on a
    graph with hubs, empty rows and neighbour ids outside [0, n): bit-identical walks, all three laws."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(16, 16, device="cuda", seed=11)
    ci = ci.clone()
    n = rp.numel() - 1
    ci[::9973] = n + 5          # ids outside the graph: nodes without out-edges (the walk stays there)
    ci[5::19997] = (1 << 40) + 3  # and one that needs the high word of the record
    nodes = torch.arange(n, device="cuda")
    ci[7::15013] = (1 << 32) + 17  # aliases node 17 in a 32-bit slot: the uint32 table must not answer for it
    d17 = min(3, int(rp[18] - rp[17]))  # and a row that holds such ids next to real ones
    ci[int(rp[17]):int(rp[17]) + d17] = torch.tensor([(1 << 32) + 17, n + 5, (1 << 40) + 3], device="cuda")[:d17]
    nodes = torch.arange(n, device="cuda")
    laws = ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0), (2.0, 1.0))
    res = {}
    try:
        for rec in (1, 0):
            native.set_option("records", rec)
            res[rec] = [native.walk(rp, ci, nodes, p_, q_, 20, 7) for p_, q_ in laws]
        # the reference's answer: a linear scan of the int64 adjacency (csrc/cuda/rw_cuda.cu:33-57), no table, no filter
        native.set_option("n2v_table", 0)
        sub = nodes[:6000].contiguous()
        scan = [native.walk(rp, ci, sub, p_, q_, 20, 7) for p_, q_ in laws]
    finally:
        native.set_option("records", -1)
        native.set_option("n2v_table", 1)
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)
    for a, b in zip(res[1], scan):
        assert torch.equal(a[:6000], b)
    assert bool((res[1][1] == (1 << 40) + 3).any())


def test_prepared_graph_and_graph_cache(native, rw):
    """A graph prepared once (explicit handle, or rw.walk's cache from the second call on) must give
    the walks of the stateless call, for every law, on any stream, and must notice in-place edits."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(15, 16, device="cuda", seed=3)
    n = rp.numel() - 1
    nodes = torch.arange(n, device="cuda")
    laws = ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0), (0.25, 4.0), (2.0, 1.0))
    base = [native.walk(rp, ci, nodes, p_, q_, 30, 11, cache=False) for p_, q_ in laws]
    g = native.prepare_csr(rp, ci)
    for (p_, q_), b in zip(laws, base):
        assert torch.equal(g.walk(nodes, p_, q_, 30, 11), b)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        w = g.walk(nodes, 1.0, 0.5, 30, 11)
    side.synchronize()
    assert torch.equal(w, base[1])
    # sharded calls through the handle keep global walk ids
    half = n // 2
    parts = [g.walk(nodes[:half], 0.5, 2.0, 30, 11), g.walk(nodes[half:], 0.5, 2.0, 30, 11, walk_id_offset=half)]
    assert torch.equal(torch.cat(parts), base[2])
    del g

    native.set_graph_cache(True)
    try:
        launches, states = [], []
        for k in range(6):
            native.reset_launch_count()
            assert torch.equal(rw.walk(rp, ci, nodes, 1.0, 0.5, 30, 11), base[1])
            launches.append(native.launch_count())
            states.append(native.graph_cache_state(rp.device))
        # call 1 one-shot (checksum + build + walk), call 2 prepares for keeps, the call after two hits adds the
        # triangle Blooms, every later call is checksum + walk kernel
        assert launches[0] > 3 and launches[1] > launches[0]
        assert [st["prepared"] for st in states] == [False, True, True, True, True, True]
        assert [st["blooms"] for st in states] == [False, False, False, True, True, True]
        # (a cached call = two checksum launches, one per array, + the walk kernel)
        assert launches[2] == 3 and launches[3] == 4 and launches[4] == 3 and launches[5] == 3
        assert torch.equal(rw.walk(rp, ci, nodes, 0.25, 4.0, 30, 11), base[3])  # same graph, other law: still cached
        # the cache goes by content: a re-created tensor with the same bytes hits ...
        native.reset_launch_count()
        assert torch.equal(rw.walk(rp.clone(), ci.clone(), nodes, 0.5, 2.0, 30, 11), base[2])
        assert native.launch_count() == 3
        # ... and any write is seen, whether torch counts it (in-place op) or not (.data, a view made earlier)
        ci2 = ci.clone()
        for _ in range(4):
            ref = rw.walk(rp, ci2, nodes, 1.0, 0.5, 30, 11)
        assert native.graph_cache_state(rp.device)["blooms"]
        version = ci2._version
        ci2.data[rp[5]:rp[6]] = ci2[rp[5]]  # all neighbours of node 5 become the same node; no version bump
        assert ci2._version == version
        edited = rw.walk(rp, ci2, nodes, 1.0, 0.5, 30, 11)
        assert torch.equal(edited, native.walk(rp, ci2, nodes, 1.0, 0.5, 30, 11, cache=False))
        assert not torch.equal(edited, ref)
        assert not native.graph_cache_state(rp.device)["prepared"]  # the stale preparation was dropped, not used
        # the cache holds no reference to the caller's tensors
        import weakref
        probe = ci2.clone()
        for _ in range(3):
            rw.walk(rp, probe, nodes, 1.0, 0.5, 30, 11)
        assert native.graph_cache_state(rp.device)["prepared"]
        alive = weakref.ref(probe)
        del probe
        assert alive() is None
    finally:
        native.set_graph_cache(False)


def test_int32_csr_gives_the_int64_walks(native, rw):
    """f4 of SURVEY section 8: row_ptr / col_idx may be int32 (each on its own).  Same values, same walks -- one-shot,
    through a prepared graph with triangle Blooms, through the content cache, for every law."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(15, 16, device="cuda", seed=13)
    n = rp.numel() - 1
    nodes = torch.arange(n, device="cuda")
    laws = ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0), (0.25, 0.5), (2.0, 1.0))
    base = [native.walk(rp, ci, nodes, p_, q_, 30, 11, cache=False) for p_, q_ in laws]
    for rp_t, ci_t in ((rp.int(), ci.int()), (rp, ci.int()), (rp.int(), ci)):
        for (p_, q_), b in zip(laws, base):
            w = native.walk(rp_t, ci_t, nodes, p_, q_, 30, 11, cache=False)
            assert w.dtype == torch.int64 and torch.equal(w, b), (rp_t.dtype, ci_t.dtype, p_, q_)
        g = native.prepare_csr(rp_t, ci_t)
        assert g.symmetric
        for (p_, q_), b in zip(laws, base):
            assert torch.equal(g.walk(nodes, p_, q_, 30, 11), b)
        del g
        assert native.csr_checksum(rp_t, ci_t) == native.csr_checksum(rp, ci)  # a checksum of the values, not of the bytes
    native.set_graph_cache(True)
    try:
        rp32, ci32 = rp.int(), ci.int()
        for _ in range(5):
            assert torch.equal(rw.walk(rp32, ci32, nodes, 1.0, 0.5, 30, 11), base[1])
        assert native.graph_cache_state(rp.device)["blooms"]
        assert torch.equal(rw.walk(rp, ci, nodes, 1.0, 0.5, 30, 11), base[1])  # int64 arrays: another key, same walks
    finally:
        native.set_graph_cache(False)
    with pytest.raises(RuntimeError):
        native.walk(rp.short(), ci, nodes, 1.0, 0.5, 30, 11)
    with pytest.raises(RuntimeError):
        native.walk(rp, ci, nodes.int(), 1.0, 0.5, 30, 11)


def test_checksum_sees_every_element(native):
    rp, ci = cuda(*random_csr(8, 3000, 20))
    base = native.csr_checksum(rp, ci)
    assert base == native.csr_checksum(rp.clone(), ci.clone())
    for arr in (rp, ci):
        for pos in (0, 1, arr.numel() // 2, arr.numel() - 1):
            keep = int(arr[pos])
            arr[pos] = keep + 1
            assert native.csr_checksum(rp, ci) != base
            arr[pos] = keep
    swapped = ci.clone()
    i = int(rp[7])
    if swapped[i] != swapped[i + 1]:
        swapped[i], swapped[i + 1] = ci[i + 1], ci[i]
        assert native.csr_checksum(rp, swapped) != base  # position-sensitive
    assert native.csr_checksum(rp, ci[1:]) != base        # unaligned start, shorter array
    assert native.csr_checksum(rp, ci) == base


def _clustered_csr(seed, n, communities, avg_deg):
    """A graph with many triangles: nodes fall into communities and mostly link inside them."""
    rng = np.random.default_rng(seed)
    m = n * avg_deg // 2
    src = rng.integers(0, n, m)
    inside = rng.random(m) < 0.8
    size = n // communities
    dst = np.where(inside, (src // size) * size + rng.integers(0, size, m), rng.integers(0, n, m)) % n
    keep = src != dst
    src, dst = np.r_[src[keep], dst[keep]], np.r_[dst[keep], src[keep]]
    key = np.unique(src.astype(np.int64) * n + dst)
    src, dst = key // n, key % n
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return torch.from_numpy(np.cumsum(row_ptr)), torch.from_numpy(dst.astype(np.int64))


@pytest.mark.parametrize("kind", ["rmat", "clustered", "directed", "duplicates", "out_of_graph_ids", "self_loops"])
def test_triangle_blooms_and_edge_filter_do_not_change_a_walk(native, kind):
    """A kept graph carries triangle Blooms in its edge records and an L2-resident edge filter
    (member_table.cuh).  Both are one-sided short cuts in front of the membership table: whatever the
    graph (symmetric or not, with or without triangles, duplicate or out-of-graph entries), every walk
    must equal the stateless call's, which has neither, for every law and every cap."""
    from torch_random_walk_b200 import rmat

    if kind == "rmat":
        rp, ci = rmat.rmat_csr(15, 16, device="cuda", seed=21)
    elif kind == "clustered":
        rp, ci = cuda(*_clustered_csr(5, 6000, 60, 40))
    elif kind == "directed":
        rp, ci = cuda(*random_csr(6, 4000, 30, symmetric=False))
    elif kind == "duplicates":
        rp, ci = random_csr(7, 3000, 40, sort_rows=False)
        rp_np, ci_np = rp.numpy(), ci.numpy().copy()
        for v in range(0, 3000, 2):
            if rp_np[v + 1] - rp_np[v] >= 2:
                ci_np[rp_np[v] + 1] = ci_np[rp_np[v]]
        rp, ci = cuda(rp, T(ci_np))
    elif kind == "self_loops":
        # a symmetric graph in which every third node also lists itself (a self-loop is its own mirror entry)
        n_ = 5000
        rng = np.random.default_rng(8)
        src, dst = rng.integers(0, n_, 60000), rng.integers(0, n_, 60000)
        loops = np.arange(0, n_, 3)
        src, dst = np.r_[src, dst, loops], np.r_[dst, src, loops]
        key = np.unique(src.astype(np.int64) * n_ + dst)
        rp_np = np.zeros(n_ + 1, dtype=np.int64)
        np.add.at(rp_np, key // n_ + 1, 1)
        rp, ci = cuda(T(np.cumsum(rp_np)), T(key % n_))
    else:
        rp, ci = rmat.rmat_csr(14, 16, device="cuda", seed=4)
        ci = ci.clone()
        ci[::577] = (1 << 32) + 9
        ci[3::1013] = rp.numel() + 7
    n = rp.numel() - 1
    nodes = torch.arange(n, device="cuda")
    laws = ((1.0, 0.5), (0.5, 2.0), (0.25, 4.0), (0.4, 0.8), (2.0, 3.0))
    saved = {k: native.get_option(k) for k in ("edge_bloom_cap", "edge_filter_mb")}
    try:
        native.set_option("edge_filter_mb", 0)
        base = [native.walk(rp, ci, nodes, p_, q_, 40, 3, cache=False) for p_, q_ in laws]
        # cap 8 and 2: most edges join two rows longer than the cap, so their questions go to the hub-pair filter (1 MB: crowded)
        for cap, filter_mb in ((256, 64), (8, 64), (2, 1), (1 << 20, 0), (0, 64), (256, 1), (8, 0)):
            native.set_option("edge_bloom_cap", cap)
            native.set_option("edge_filter_mb", filter_mb)
            g = native.prepare_csr(rp, ci)
            assert (g.info()["edge_filter_bits"] > 0) == (cap > 0 and filter_mb > 0)
            for (p_, q_), b in zip(laws, base):
                assert torch.equal(g.walk(nodes, p_, q_, 40, 3), b), (kind, cap, filter_mb, p_, q_)
            if cap == 8 and filter_mb == 64:  # the same through kernels without edge records: the filter must stay out of it
                native.set_option("records", 0)
                try:
                    for (p_, q_), b in zip(laws[:2], base[:2]):
                        assert torch.equal(g.walk(nodes, p_, q_, 40, 3), b), (kind, "records off", p_, q_)
                finally:
                    native.set_option("records", -1)
            if cap == 256 and filter_mb == 64:
                assert g.symmetric == (kind in ("rmat", "clustered", "self_loops"))
            del g
    finally:
        for k, v in saved.items():
            native.set_option(k, v)


def test_unsorted_rows_and_duplicate_edges(native):
    # the API does not promise sorted or duplicate-free rows; membership must not depend on either
    rp, ci = random_csr(4, 1500, 40, sort_rows=False)
    rp_np, ci_np = rp.numpy(), ci.numpy().copy()
    # duplicate the first neighbour of every row with >= 2 entries into its second slot
    for v in range(1500):
        if rp_np[v + 1] - rp_np[v] >= 2:
            ci_np[rp_np[v] + 1] = ci_np[rp_np[v]]
    rp, ci = cuda(rp, T(ci_np))
    nodes = torch.arange(1500, device="cuda")
    a = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
    native.set_option("n2v_table", 0)
    try:
        b = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
    finally:
        native.set_option("n2v_table", 1)
    assert torch.equal(a, b)
    check_walks_follow_edges(a, rp, ci, nodes)
    # folding and the mixture need duplicate-free rows: the prepare step must have switched them off here
    native.set_option("n2v_fold", 0)
    native.set_option("n2v_mix", 0)
    try:
        c = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
    finally:
        native.set_option("n2v_fold", 1)
        native.set_option("n2v_mix", 1)
    assert torch.equal(a, c)


def test_folding_changes_the_draws_not_the_law(native, golden):
    """With strictly increasing rows the kernel picks the tightest sampling scheme the law allows:
    the two-sided mixture (q > 1, p <= q), else return-edge folding (1/p > max(1, 1/q)), else plain
    rejection.  Other draws, same distribution (the statistical tests below run all three);
    switching the faster schemes off restores the plain-rejection stream bit for bit."""
    rp, ci = cuda(*random_csr(6, 2000, 20))
    nodes = torch.arange(2000, device="cuda")
    mixed = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
    native.set_option("n2v_mix", 0)
    try:
        folded = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
        native.set_option("n2v_fold", 0)
        plain = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
        plain2 = native.walk(rp, ci, nodes, 0.25, 4.0, 30, 5)
    finally:
        native.set_option("n2v_fold", 1)
        native.set_option("n2v_mix", 1)
    assert torch.equal(plain, plain2)
    assert not torch.equal(folded, plain) and not torch.equal(mixed, plain) and not torch.equal(mixed, folded)
    for w in (mixed, folded, plain):
        check_walks_follow_edges(w, rp, ci, nodes)
    # p >= min(1, q): nothing to fold, the option is inert
    a = native.walk(rp, ci, nodes, 1.0, 0.5, 30, 5)
    native.set_option("n2v_fold", 0)
    try:
        b = native.walk(rp, ci, nodes, 1.0, 0.5, 30, 5)
    finally:
        native.set_option("n2v_fold", 1)
    assert torch.equal(a, b)


def test_dead_end_and_isolated_nodes_stay(rw):
    # convention of csrc/cuda/rw_cuda.cu:25-30: no out-edge -> the walk stays on the node
    # 0 -> 1 -> 2 (sink), 3 isolated, 4 -> 0
    rp = torch.tensor([0, 1, 2, 2, 2, 3], dtype=torch.int64).cuda()
    ci = torch.tensor([1, 2, 0], dtype=torch.int64).cuda()
    nodes = torch.arange(5, device="cuda")
    for p, q in ((1.0, 1.0), (0.5, 2.0)):
        w = rw.walk(rp, ci, nodes, p, q, 6, 3).cpu()
        assert w[0].tolist() == [0, 1, 2, 2, 2, 2, 2]
        assert w[2].tolist() == [2] * 7
        assert w[3].tolist() == [3] * 7
        assert w[4].tolist() == [4, 0, 1, 2, 2, 2, 2]


def test_edge_cases(rw, native):
    rp, ci = cuda(*random_csr(5, 100, 6))
    empty = torch.empty(0, dtype=torch.int64, device="cuda")
    assert rw.walk(rp, ci, empty, 1.0, 1.0, 5, 1).shape == (0, 6)
    assert rw.walk(rp, ci, empty, 0.5, 2.0, 5, 1).shape == (0, 6)
    nodes = torch.arange(100, device="cuda")
    for p, q in ((1.0, 1.0), (0.5, 2.0)):
        for L in (0, 1, 2, 3, 4, 5):
            w = rw.walk(rp, ci, nodes, p, q, L, 1)
            assert w.shape == (100, L + 1)
            check_walks_follow_edges(w, rp, ci, nodes)
    # output rows that do not start on a sector boundary: write into a strided, offset view
    big = torch.full((100, 13), -7, dtype=torch.int64, device="cuda")
    out = big[:, 1:10]
    native.walk(rp, ci, nodes, 0.5, 2.0, 8, 11, out=out)
    ref = native.walk(rp, ci, nodes, 0.5, 2.0, 8, 11)
    assert torch.equal(out, ref)
    assert (big[:, 0] == -7).all() and (big[:, 10:] == -7).all()
    # dtypes: the reference's accessors raise on anything but int64; int32 CSR arrays are accepted here as an extension
    # (test_int32_csr_gives_the_int64_walks), everything else still raises the reference's way
    for bad_rp, bad_ci, bad_nodes in ((rp.short(), ci, nodes), (rp, ci.double(), nodes), (rp, ci, nodes.int())):
        with pytest.raises(RuntimeError, match="Long"):
            rw.walk(bad_rp, bad_ci, bad_nodes, 1.0, 1.0, 3, 1)
    for bad in (torch.empty((100, 8), dtype=torch.int64, device="cuda"), torch.empty((99, 9), dtype=torch.int64, device="cuda"),
                torch.empty((100, 9), dtype=torch.int32, device="cuda"), torch.empty((100, 9), dtype=torch.int64),
                torch.empty((100, 18), dtype=torch.int64, device="cuda")[:, ::2]):
        with pytest.raises(RuntimeError, match="out must"):
            native.walk(rp, ci, nodes, 0.5, 2.0, 8, 11, out=bad)
    # non-contiguous inputs are accepted (the reference reads through strided accessors)
    w = rw.walk(rp, ci, torch.arange(100, device="cuda")[::2], 1.0, 1.0, 4, 1)
    check_walks_follow_edges(w, rp, ci, torch.arange(100, device="cuda")[::2])


def test_degenerate_graphs(rw):
    # no edges at all: every walk stays on its start node, for every (p, q)
    rp = torch.zeros(11, dtype=torch.int64, device="cuda")
    ci = torch.empty(0, dtype=torch.int64, device="cuda")
    nodes = torch.arange(10, device="cuda")
    for p, q in ((1.0, 1.0), (0.5, 2.0), (1.0, 0.5), (0.25, 4.0)):
        assert torch.equal(rw.walk(rp, ci, nodes, p, q, 7, 1), nodes[:, None].expand(-1, 8))
    # a single self-loop and a 2-cycle; long walks; one walk
    rp = torch.tensor([0, 1, 2, 3], dtype=torch.int64, device="cuda")
    ci = torch.tensor([0, 2, 1], dtype=torch.int64, device="cuda")
    for p, q in ((1.0, 1.0), (0.25, 4.0), (4.0, 0.25)):
        w = rw.walk(rp, ci, torch.tensor([1], device="cuda"), p, q, 1000, 3).cpu()
        assert w.shape == (1, 1001) and w[0, ::2].eq(1).all() and w[0, 1::2].eq(2).all()
        w = rw.walk(rp, ci, torch.tensor([0], device="cuda"), p, q, 33, 3)
        assert bool((w == 0).all())
    # a star: the hub row (>= 2048 neighbours) goes through the hub-segment builder
    n = 5000
    hub_rp = torch.cat((torch.tensor([0, n - 1]), n - 1 + torch.arange(1, n))).cuda()
    hub_ci = torch.cat((torch.arange(1, n), torch.zeros(n - 1, dtype=torch.int64))).cuda()
    w = rw.walk(hub_rp, hub_ci, torch.arange(n, device="cuda"), 0.5, 2.0, 9, 1).cpu()
    assert bool((w[1:, 1] == 0).all()) and bool((w[0, 1] > 0).all())
    assert bool((w[:, 2:][w[:, 1:-1] == 0] > 0).all()) and bool((w[:, 2:][w[:, 1:-1] > 0] == 0).all())


def test_first_order_transitions_are_uniform(rw, golden):
    rp, ci = T(golden["utils/karate/row_ptr"]), T(golden["utils/karate/col_idx"])
    n = 34
    nodes = torch.arange(n).repeat_interleave(3000)
    walks = rw.walk(rp.cuda(), ci.cuda(), nodes.cuda(), 1.0, 1.0, 80, 17).cpu().numpy()
    a, b = walks[:, :-1].ravel(), walks[:, 1:].ravel()
    deg = np.diff(rp.numpy())
    keys, counts = np.unique(a.astype(np.int64) * n + b, return_counts=True)
    visits = np.bincount(a, minlength=n)
    expected = visits[keys // n] / deg[keys // n]
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    dof = len(keys) - n
    assert len(keys) == 156  # every edge of the karate club is used
    assert chi2_pvalue(chi2, dof) > 0.01, (chi2, dof)


# (p, q, options): every sampling scheme of the node2vec kernel -- two-sided mixture (q > 1, p <= q), return-edge
# folding (1/p the strict maximum), plain rejection -- on the laws that select it, and the slower schemes forced
# on the same laws by switching the faster ones off.
SCHEMES = [(0.5, 2.0, {}), (0.25, 4.0, {}), (2.0, 4.0, {}), (0.5, 2.0, {"n2v_mix": 0}), (0.25, 4.0, {"n2v_mix": 0}),
           (0.25, 4.0, {"n2v_mix": 0, "n2v_fold": 0}), (1.0, 0.5, {}), (0.25, 0.5, {}), (0.5, 0.25, {}), (4.0, 2.0, {})]


@pytest.mark.parametrize("p,q,opts", SCHEMES)
def test_second_order_statistics_match_analytic_and_oracle(rw, native, orc, golden, p, q, opts):
    """North-star criterion: chi-square p > 0.01 and pooled TV < 1e-2 at 1e7 samples, against the
    analytic node2vec probabilities and against the reference algorithm's own empirical counts."""
    rp, ci = T(golden["utils/karate/row_ptr"]), T(golden["utils/karate/col_idx"])
    n = 34
    table = node2vec_probs(rp, ci, p, q)
    L = 100
    reps = 3000  # 34 * 3000 walks * 99 second-order transitions = 1.0e7 samples
    nodes = torch.arange(n).repeat_interleave(reps)
    for k, v in opts.items():
        native.set_option(k, v)
    try:
        walks = rw.walk(rp.cuda(), ci.cuda(), nodes.cuda(), p, q, L, 2024)
    finally:
        for k in opts:
            native.set_option(k, 1)
    check_walks_follow_edges(walks, rp, ci, nodes)
    got = second_order_counts(walks, n)
    assert sum(got.values()) >= 10_000_000
    chi2, dof, tv = chi2_and_tv(got, table, n)
    assert chi2_pvalue(chi2, dof) > 0.01, ("analytic", chi2, dof)
    assert tv < 1e-2, tv
    # the oracle (reference algorithm, glibc rand) on a third of the samples
    ref_walks = orc.walk(rp, ci, torch.arange(n).repeat_interleave(1000), p, q, L, 7)
    ref_counts = second_order_counts(ref_walks, n)
    chi2_o, dof_o, tv_o = chi2_and_tv(ref_counts, table, n)
    assert tv_o < 1e-2  # the oracle itself is a faithful node2vec sampler
    chi2_2, dof_2 = two_sample_chi2(got, ref_counts, n)
    assert chi2_pvalue(chi2_2, dof_2) > 0.01, ("vs oracle", chi2_2, dof_2)


@pytest.mark.parametrize("p,q", [(0.5, 2.0), (1.0, 0.5), (0.25, 4.0), (2.0, 1.0)])
def test_warp_per_walk_exact_cdf_kernel_samples_the_same_law(native, orc, golden, p, q):
    """The A/B kernel of option n2v_warp (one warp per walk, exact CDF by prefix sum up to 64 neighbours, lock-step
    rejection above): other draws, same law -- analytic chi-square / TV on karate (every row takes the CDF), and
    acceptance classes against the analytic law and the oracle on a graph whose rows take the rejection branch."""
    rp, ci = T(golden["utils/karate/row_ptr"]), T(golden["utils/karate/col_idx"])
    n = 34
    nodes = torch.arange(n).repeat_interleave(3000)
    native.set_option("n2v_warp", 1)
    try:
        g = native.prepare_csr(rp.cuda(), ci.cuda())
        walks = g.walk(nodes.cuda(), p, q, 100, 77)
        del g
        rp2, ci2 = random_csr(21, 400, 130)  # rows of ~110 neighbours: the rejection branch
        assert int((rp2[1:] - rp2[:-1]).min()) > 64
        g2 = native.prepare_csr(rp2.cuda(), ci2.cuda())
        walks2 = g2.walk(torch.arange(400).repeat_interleave(250).cuda(), p, q, 100, 78)
        native.set_option("n2v_warp", 0)
        shipped2 = g2.walk(torch.arange(400).repeat_interleave(250).cuda(), p, q, 100, 79)  # the thread-per-walk kernel, other seed
        del g2
    finally:
        native.set_option("n2v_warp", 0)
    check_walks_follow_edges(walks, rp, ci, nodes)
    got = second_order_counts(walks, n)
    chi2, dof, tv = chi2_and_tv(got, node2vec_probs(rp, ci, p, q), n)
    assert sum(got.values()) >= 10_000_000
    assert chi2_pvalue(chi2, dof) > 0.01, (chi2, dof)
    assert tv < 1e-2, tv
    got2 = second_order_counts(walks2, 400)
    # 44 k contexts share 1e7 samples, so the per-context class TV is sampling noise (~0.03) for any exact sampler: the
    # warp kernel must sit where the shipped kernel sits, and be homogeneous with it and with the oracle
    ship2 = second_order_counts(shipped2, 400)
    assert abs(_class_tv(got2, rp2, ci2, p, q, 400) - _class_tv(ship2, rp2, ci2, p, q, 400)) < 2e-3
    ref2 = second_order_counts(orc.walk(rp2, ci2, torch.arange(400).repeat_interleave(25), p, q, 100, 3), 400)
    # homogeneity per context on the three acceptance classes (a per-neighbour table would have ~2 samples per cell here)
    for other, what in ((ship2, "shipped kernel"), (ref2, "oracle")):
        chi2_s, dof_s = _two_sample_class_chi2(got2, other, rp2, ci2, 400)
        assert dof_s > 10_000 and chi2_pvalue(chi2_s, dof_s) > 0.01, (what, chi2_s, dof_s)


def test_fused_walk_to_windows_prototype_equals_walk_then_to_windows(native, rw):
    """The A/B kernel of SURVEY section 8 (f1): skip-gram targets and positive windows of width 5 written by the walk
    kernel itself must be, bit for bit, what rw.to_windows makes of the walks of the same call."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(14, 16, device="cuda", seed=6)
    n = rp.numel() - 1
    nodes = torch.arange(n, device="cuda")[: n - 5]  # a ragged last CTA and warp
    g = native.prepare_csr(rp, ci)
    for L in (80, 7, 4):
        for p_, q_ in ((1.0, 0.5), (2.0, 0.5)):
            walks = g.walk(nodes, p_, q_, L, 3)
            target, pos, _ = rw.to_windows(walks, 5, n, 1)
            f_target, f_pos = g.walk_windows5(nodes, p_, q_, L, 3)
            assert torch.equal(f_target, target) and torch.equal(f_pos, pos), (L, p_, q_)
    with pytest.raises(RuntimeError):
        g.walk_windows5(nodes, 0.5, 2.0, 80, 3)  # the mixture laws are not part of the prototype


def _two_sample_class_chi2(counts_a, counts_b, row_ptr, col_idx, n, min_total=20):
    """Homogeneity of two sets of (t, v, x) counts on the acceptance classes (return / common neighbour / far) of
    every (t, v) context: sum of the 2 x 3 contingency chi-squares.  Returns (chi2, dof)."""
    from collections import defaultdict

    rp, ci = row_ptr.numpy(), col_idx.numpy()
    adj = [set(ci[rp[i]:rp[i + 1]].tolist()) for i in range(n)]
    ctx = defaultdict(lambda: [[0, 0], [0, 0], [0, 0]])
    for which, counts in enumerate((counts_a, counts_b)):
        for key, c in counts.items():
            tv_, x = divmod(key, n)
            t, _ = divmod(tv_, n)
            ctx[tv_][0 if x == t else (1 if x in adj[t] else 2)][which] += c
    chi2, dof = 0.0, 0
    for cells in ctx.values():
        rows = [ab for ab in cells if ab[0] + ab[1] >= min_total]
        na, nb = sum(a for a, _ in rows), sum(b for _, b in rows)
        if len(rows) < 2 or na == 0 or nb == 0:
            continue
        for a, b in rows:
            ea, eb = (a + b) * na / (na + nb), (a + b) * nb / (na + nb)
            chi2 += (a - ea) ** 2 / ea + (b - eb) ** 2 / eb
        dof += len(rows) - 1
    return chi2, dof


def _class_tv(counts, row_ptr, col_idx, p, q, n):
    """Count-weighted mean TV over the three acceptance classes (return / common neighbour / far)
    per (t,v) context: the quantity the rejection rule controls, and far less noisy than the
    per-neighbour TV when a context has ~30 outcomes."""
    from collections import defaultdict

    rp, ci = row_ptr.numpy(), col_idx.numpy()
    adj = [set(ci[rp[i]:rp[i + 1]].tolist()) for i in range(n)]
    ctx = defaultdict(lambda: [0, 0, 0])
    for key, c in counts.items():
        tv_, x = divmod(key, n)
        t, v = divmod(tv_, n)
        ctx[(t, v)][0 if x == t else (1 if x in adj[t] else 2)] += c
    tv_sum, total = 0.0, 0
    for (t, v), obs in ctx.items():
        nb = ci[rp[v]:rp[v + 1]].tolist()
        w = [sum(1.0 / p for x in nb if x == t), sum(1.0 for x in nb if x != t and x in adj[t]),
             sum(1.0 / q for x in nb if x != t and x not in adj[t])]
        m, z = sum(obs), sum(w)
        tv_sum += 0.5 * sum(abs(o / m - e / z) for o, e in zip(obs, w)) * m
        total += m
    return tv_sum / total


@pytest.mark.parametrize("p,q,opts", [(0.5, 2.0, {}), (0.5, 2.0, {"records": 1}), (0.5, 2.0, {"n2v_mix": 0}), (2.0, 4.0, {}),
                                      (1.0, 0.5, {"records": 1})])
def test_second_order_statistics_with_table_rows(rw, native, orc, p, q, opts):
    """Same criterion on a graph whose rows are long enough (>= 12) to go through the hashed
    membership table, with triangles so that all three acceptance classes occur and with rows of
    different lengths so that the mixture proposes common neighbours from both sides.  With ~30
    outcomes per context the per-neighbour TV at 1e7 samples is dominated by sampling noise
    (~0.02), so the TV bound is applied to the acceptance classes; chi-square stays per neighbour."""
    rp, ci = random_csr(11, 60, 40)
    n = 60
    assert int((rp[1:] - rp[:-1]).min()) >= 16 and int((rp[1:] - rp[:-1]).max()) > int((rp[1:] - rp[:-1]).min())
    table = node2vec_probs(rp, ci, p, q)
    nodes = torch.arange(n).repeat_interleave(3000)
    for k, v in opts.items():
        native.set_option(k, v)
    try:
        walks = rw.walk(rp.cuda(), ci.cuda(), nodes.cuda(), p, q, 100, 5)
    finally:
        for k in opts:
            native.set_option(k, {"records": -1}.get(k, 1))
    got = second_order_counts(walks, n)
    chi2, dof, tv = chi2_and_tv(got, table, n)
    assert chi2_pvalue(chi2, dof) > 0.01, (chi2, dof)
    assert tv < 4e-2
    assert _class_tv(got, rp, ci, p, q, n) < 1e-2
    ref_counts = second_order_counts(orc.walk(rp, ci, torch.arange(n).repeat_interleave(300), p, q, 100, 3), n)
    chi2_2, dof_2 = two_sample_chi2(got, ref_counts, n)
    assert chi2_pvalue(chi2_2, dof_2) > 0.01, (chi2_2, dof_2)


@pytest.mark.parametrize("p,q,opts", [(0.5, 2.0, {}), (0.5, 2.0, {"records": 1}), (0.25, 0.5, {}), (1.0, 0.5, {"records": 1})])
def test_second_order_statistics_on_a_directed_graph(rw, native, orc, p, q, opts):
    """A directed graph: t -> v does not imply v -> t, so the return edge and the common-neighbour
    class are decided by adj(v) and adj(t) separately -- the mixture's third term (t accepted iff t is
    a neighbour of v) and its t-side proposals (accepted iff in adj(v)) have to get this right."""
    rp, ci = random_csr(17, 50, 24, symmetric=False)
    n = 50
    deg = (rp[1:] - rp[:-1])
    assert int(deg.min()) >= 12  # no dead ends, every row through the hashed table
    table = node2vec_probs(rp, ci, p, q)
    nodes = torch.arange(n).repeat_interleave(3000)
    for k, v in opts.items():
        native.set_option(k, v)
    try:
        walks = rw.walk(rp.cuda(), ci.cuda(), nodes.cuda(), p, q, 80, 23)
    finally:
        for k in opts:
            native.set_option(k, {"records": -1}.get(k, 1))
    check_walks_follow_edges(walks, rp, ci, nodes)
    got = second_order_counts(walks, n)
    chi2, dof, tv = chi2_and_tv(got, table, n)
    assert chi2_pvalue(chi2, dof) > 0.01, (chi2, dof)
    assert _class_tv(got, rp, ci, p, q, n) < 1e-2
    ref_counts = second_order_counts(orc.walk(rp, ci, torch.arange(n).repeat_interleave(300), p, q, 80, 3), n)
    chi2_2, dof_2 = two_sample_chi2(got, ref_counts, n)
    assert chi2_pvalue(chi2_2, dof_2) > 0.01, (chi2_2, dof_2)


def test_large_graph_structure_and_hubs(rw):
    """A skewed graph (R-MAT scale 16) at a size where hubs, the table build's tile logic and many
    CTAs are all exercised; size-independent property: every transition is an edge."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(16, 16, device="cuda", seed=3)
    deg = rp[1:] - rp[:-1]
    assert int(deg.max()) > 1000 and int((deg == 0).sum()) > 0
    nodes = torch.nonzero(deg > 0).flatten()
    for p, q in ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0)):
        w = rw.walk(rp, ci, nodes, p, q, 20, 1)
        check_walks_follow_edges(w, rp, ci, nodes)
    # isolated start nodes stay put
    iso = torch.nonzero(deg == 0).flatten()[:100]
    w = rw.walk(rp, ci, iso, 1.0, 0.5, 5, 1)
    assert torch.equal(w, iso[:, None].expand(-1, 6))


@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_baseline_sizes_through_size_independent_properties(native, workload):
    """BASELINE.json configs[1] (2.45 M nodes, 66 M CSR entries, p=0.5 q=2) and configs[2] (R-MAT scale 24,
    521 M entries, p=1 q=0.5) at full size, L=80, one walk per degree>0 node.  The oracle cannot run these in
    seconds, so the checks are the properties that hold at any size: column 0 = start nodes, every transition
    an edge, reruns identical, any sharding of the start nodes concatenates to the single call, and the
    prepared graph (edge records, kept table) gives the walks of the stateless call."""
    from torch_random_walk_b200 import rmat

    # the generator parameters bench.py uses for these configs
    wl = {"c2": dict(scale=22, n_nodes=2449029, n_edges=34_000_000, p=0.5, q=2.0, L=80),
          "c3": dict(scale=24, n_nodes=None, edge_factor=16, p=1.0, q=0.5, L=80)}[workload]
    free, _ = torch.cuda.mem_get_info()
    if workload == "c3" and free < 60 << 30:
        pytest.skip("needs ~40 GB of free device memory")
    rp, ci = rmat.rmat_csr(wl["scale"], wl.get("edge_factor", 16), n_nodes=wl.get("n_nodes"), device="cuda",
                           n_edges=wl.get("n_edges"))
    nodes = torch.nonzero(rp[1:] - rp[:-1] > 0).flatten().contiguous()
    p, q, L = wl["p"], wl["q"], wl["L"]
    n = nodes.numel()
    one_shot = native.walk(rp, ci, nodes, p, q, L, 77, cache=False)
    assert one_shot.shape == (n, L + 1) and torch.equal(one_shot[:, 0], nodes)
    sample = one_shot[:: max(1, n // 200_000)]
    assert rmat.transitions_are_edges(rp, ci, sample)
    # checksums instead of a second full copy: a position-weighted sum that any changed entry changes
    weights = torch.arange(1, L + 2, device="cuda", dtype=torch.int64)

    def digest(w):
        return (w * weights).sum(1)

    want = digest(one_shot)
    del one_shot, sample
    graph = native.prepare_csr(rp, ci)
    assert torch.equal(digest(graph.walk(nodes, p, q, L, 77)), want)
    cuts = [0, n // 3, n // 3 + 12345, n]
    parts = [digest(graph.walk(nodes[a:b].contiguous(), p, q, L, 77, walk_id_offset=a)) for a, b in zip(cuts, cuts[1:])]
    assert torch.equal(torch.cat(parts), want)
    uni = graph.walk(nodes, 1.0, 1.0, L, 5)
    assert rmat.transitions_are_edges(rp, ci, uni[:: max(1, n // 200_000)])
    assert not torch.equal(digest(graph.walk(nodes, p, q, L, 78)), want)  # another seed, another set of walks


def test_guard_bands_around_outputs_and_workspace(native):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are looked for directly: the
    workspace (row index, table, records, work lists) and the walk output get sentinel-filled guard
    bands on both sides, the library is called through the C ABI with pointers into the middle, and
    the bands must come back untouched.  Shapes: hubs, empty rows, a walk count that is no multiple
    of the warp size, a row length that is no multiple of the 16-element staging pieces."""
    import ctypes

    from torch_random_walk_b200 import rmat

    lib = native.lib()
    rp, ci = rmat.rmat_csr(14, 16, device="cuda", seed=9)
    n, nnz = rp.numel() - 1, ci.numel()
    nodes = torch.arange(n - 7, device="cuda")
    nw, L, guard = nodes.numel(), 37, 4096
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    sentinel = -0x0123456789ABCDEF
    for p_, q_, rec in ((1.0, 1.0, 1), (1.0, 0.5, 1), (0.5, 2.0, 1), (0.25, 4.0, 0), (1.0, 0.5, 0)):
        native.set_option("records", rec)
        try:
            need = lib.trw_walk_csr_workspace_bytes(n, nnz, p_, q_)
            ws_words = (need + 7) // 8
            ws = torch.full((guard + ws_words + guard,), sentinel, dtype=torch.int64, device="cuda")
            out = torch.full((guard + nw * (L + 1) + guard,), sentinel, dtype=torch.int64, device="cuda")
            base = ws.data_ptr() + guard * 8
            assert base % 256 == 0
            rc = lib.trw_walk_csr(ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()), n, nnz,
                                  ctypes.c_void_p(nodes.data_ptr()), nw, 0, p_, q_, L, 5,
                                  ctypes.c_void_p(out.data_ptr() + guard * 8), L + 1, ctypes.c_void_p(base), need, 0, st)
            assert rc == 0, lib.trw_last_error()
            torch.cuda.synchronize()
            for t in (ws, out):
                assert bool((t[:guard] == sentinel).all()) and bool((t[-guard:] == sentinel).all()), (p_, q_, rec)
            walks = out[guard:guard + nw * (L + 1)].view(nw, L + 1)
            assert torch.equal(walks, native.walk(rp, ci, nodes, p_, q_, L, 5, cache=False))
        finally:
            native.set_option("records", -1)


def test_guard_bands_around_a_kept_graph(native):
    """The same for a kept graph called through the C ABI: the workspace of trw_csr_graph_prepare_ex (table, edge
    records with their triangle Blooms, edge filter, work lists, flag cells), the walk output in contiguous and
    block-cyclic numbering, the fused window outputs and the checksum cell sit between sentinel bands."""
    import ctypes

    from torch_random_walk_b200 import rmat

    lib = native.lib()
    rp, ci = rmat.rmat_csr(14, 16, device="cuda", seed=10)
    n, nnz = rp.numel() - 1, ci.numel()
    nodes = torch.arange(n - 3, device="cuda")
    nw, L, guard = nodes.numel(), 23, 4096
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    sentinel = -0x0123456789ABCDEF

    def banded(words):
        t = torch.full((guard + words + guard,), sentinel, dtype=torch.int64, device="cuda")
        return t, t.data_ptr() + guard * 8

    def intact(*tensors):
        torch.cuda.synchronize()
        return all(bool((t[:guard] == sentinel).all()) and bool((t[-guard:] == sentinel).all()) for t in tensors)

    try:
        for filter_mb in (0, 1):
            native.set_option("edge_filter_mb", filter_mb)
            need = lib.trw_csr_graph_workspace_bytes(n, nnz)
            ws, ws_ptr = banded((need + 7) // 8)
            assert ws_ptr % 256 == 0
            handle = ctypes.c_void_p()
            rc = lib.trw_csr_graph_prepare_ex(ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()), n, nnz,
                                              ctypes.c_void_p(ws_ptr), need, 0, st, 8, ctypes.byref(handle))
            assert rc == 0, lib.trw_last_error()
            assert lib.trw_csr_graph_add_blooms(handle, None, None, 64, st) == 0  # already there at cap 8: a no-op
            for p_, q_ in ((1.0, 0.5), (0.5, 2.0), (0.25, 0.5), (1.0, 1.0)):
                out, out_ptr = banded(nw * (L + 1))
                for ids in ((0, 0, 0), (128, 128, 512)):
                    rc = lib.trw_walk_csr_prepared_at(handle, ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()),
                                                      ctypes.c_void_p(nodes.data_ptr()), nw, ids[0], ids[1], ids[2], p_, q_, L, 5,
                                                      ctypes.c_void_p(out_ptr), L + 1, st)
                    assert rc == 0, lib.trw_last_error()
                    assert intact(ws, out), (filter_mb, p_, q_, ids)
                    if ids == (0, 0, 0):
                        walks = out[guard:guard + nw * (L + 1)].view(nw, L + 1)
                        assert torch.equal(walks, native.walk(rp, ci, nodes, p_, q_, L, 5, cache=False))
            tgt, tgt_ptr = banded(nw * (L - 3))
            pos, pos_ptr = banded(nw * (L - 3) * 4)
            rc = lib.trw_walk_csr_prepared_windows5(handle, ctypes.c_void_p(nodes.data_ptr()), nw, 0, 1.0, 0.5, L, 5,
                                                    ctypes.c_void_p(tgt_ptr), ctypes.c_void_p(pos_ptr), st)
            assert rc == 0, lib.trw_last_error()
            assert intact(ws, tgt, pos)
            cell, cell_ptr = banded(1)
            assert lib.trw_csr_checksum(ctypes.c_void_p(rp.data_ptr()), ctypes.c_void_p(ci.data_ptr()), n, nnz, ctypes.c_void_p(cell_ptr), 0, st) == 0
            assert intact(cell) and int(cell[guard]) == native.csr_checksum(rp, ci) - (1 << 64 if native.csr_checksum(rp, ci) >= 1 << 63 else 0)
            lib.trw_csr_graph_destroy(handle)
    finally:
        native.set_option("edge_filter_mb", 0)


def test_walk_host_matches_device_path(native):
    rp, ci = random_csr(8, 3000, 20)
    nodes = torch.randint(0, 3000, (10000,))
    native.set_option("host_chunk_walks", 3000)  # force several pipelined chunks
    try:
        for p, q in ((1.0, 1.0), (0.5, 2.0)):
            host = native.walk_host(rp, ci, nodes, p, q, 16, 42, device=0)
            dev = native.walk(rp.cuda(), ci.cuda(), nodes.cuda(), p, q, 16, 42)
            assert host.device.type == "cpu" and torch.equal(host, dev.cpu())
    finally:
        native.set_option("host_chunk_walks", 1 << 20)


def test_walk_host_keeps_the_replica_and_notices_changes(native):
    """The host entry keeps the device replica of the graph for a caller that comes back with the same arrays, grows
    its preparation with use, and re-checks the arrays' content on every call: an in-place change of the host graph
    must give the walks of the changed graph, never those of the stale replica."""
    from torch_random_walk_b200 import rmat

    rp, ci = rmat.rmat_csr(13, 16, seed=9)
    rp, ci = rp.pin_memory(), ci.pin_memory()
    n = rp.numel() - 1
    nodes = torch.arange(n)
    laws = ((1.0, 0.5), (0.5, 2.0), (1.0, 1.0))
    expect = {law: native.walk(rp.cuda(), ci.cuda(), nodes.cuda(), law[0], law[1], 20, 3, cache=False).cpu() for law in laws}
    saved = {k: native.get_option(k) for k in ("host_threads", "host_compress", "host_chunk_walks", "host_check_dma", "host_packed_share", "host_sum_piece")}
    try:
        native.lib().trw_release_cached_buffers()
        native.set_option("host_sum_piece", 10007)  # ~25 pieces for the copy engine and the host threads to share
        native.set_option("host_chunk_walks", 1024)  # eight chunks: the ring of three buffers goes round
        # dma: eighths of the content check the copy engine may take; share: chunks of 8 that travel packed (-1: adaptive)
        for threads, compress, dma, share in ((2, 1, 8, -1), (12, 1, 8, -1), (12, 2, 0, -1), (12, 1, 3, 5), (5, 1, 5, 8), (1, 1, 4, -1), (3, 0, 4, -1)):
            native.set_option("host_threads", threads)
            native.set_option("host_compress", compress)
            native.set_option("host_check_dma", dma)
            native.set_option("host_packed_share", share)
            for rounds in range(6):  # fresh, full preparation, ..., triangle Blooms, steady
                for law in laws:
                    assert torch.equal(native.walk_host(rp, ci, nodes, law[0], law[1], 20, 3, device=0), expect[law]), (threads, compress, rounds, law)
            info = native.host_replica_info(0)  # a content check that failed would show as a fresh upload on every call
            assert info["held"] and info["level"] == 2 and info["last_call"] == "kept replica validated", (threads, compress, dma, share, info)
            # the bytes the call moved: walks at 4 or 8 bytes per entry; start nodes plus at most the copy engine's share of the check
            assert n * 21 * (4 if share == 8 else 8 if compress == 0 or share == 0 else 4) <= info["d2h_bytes"] <= n * 21 * 8 + 64, info
            assert n * 8 <= info["h2d_bytes"] <= n * 8 + (rp.numel() + ci.numel()) * 8 * dma // 8 + 8 * 10007 * 2, info
        assert native.csr_checksum_host(rp, ci) == native.csr_checksum(rp.cuda(), ci.cuda())
        # pageable arrays: kept as well, always summed by the host threads
        rp_p, ci_p = rp.clone(), ci.clone()
        assert not rp_p.is_pinned()
        for rounds in range(4):
            assert torch.equal(native.walk_host(rp_p, ci_p, nodes, 1.0, 0.5, 20, 3, device=0), expect[(1.0, 0.5)])
        ci_p[int(rp[9])] = ci_p[int(rp[9]) + 1]
        assert native.host_replica_info(0)["last_call"] == "kept replica validated"
        assert torch.equal(native.walk_host(rp_p, ci_p, nodes, 1.0, 0.5, 20, 3, device=0),
                           native.walk(rp_p.cuda(), ci_p.cuda(), nodes.cuda(), 1.0, 0.5, 20, 3, cache=False).cpu())
        assert native.host_replica_info(0)["last_call"] == "kept replica found changed"
        native.set_option("host_check_dma", 4)
        for rounds in range(3):
            assert torch.equal(native.walk_host(rp, ci, nodes, 1.0, 0.5, 20, 3, device=0), expect[(1.0, 0.5)])
        # change the graph under the same pointers
        lo, hi = int(rp[3]), int(rp[4])
        ci[lo:hi] = ci[lo]
        changed = native.walk(rp.cuda(), ci.cuda(), nodes.cuda(), 1.0, 0.5, 20, 3, cache=False).cpu()
        assert not torch.equal(changed, expect[(1.0, 0.5)])
        assert torch.equal(native.walk_host(rp, ci, nodes, 1.0, 0.5, 20, 3, device=0), changed)
        assert native.host_replica_info(0)["last_call"] == "kept replica found changed"
        assert torch.equal(native.walk_host(rp, ci, nodes, 1.0, 0.5, 20, 3, device=0), changed)
        assert native.host_replica_info(0)["last_call"] == "kept replica validated"
        rp2 = rp.clone().pin_memory()  # same content at another address: a fresh upload, same walks
        assert torch.equal(native.walk_host(rp2, ci, nodes, 1.0, 0.5, 20, 3, device=0), changed)
        assert native.host_replica_info(0)["last_call"] == "fresh upload"
        with pytest.raises(RuntimeError):
            native.walk_host(rp, ci, nodes, 1.0, 0.5, 20, 3, device=0, out=torch.empty((n, 20), dtype=torch.int64))
    finally:
        for k, v in saved.items():
            native.set_option(k, v)
        native.lib().trw_release_cached_buffers()


def test_prepared_graph_walks_to_host(native):
    from torch_random_walk_b200 import rmat
    from torch_random_walk_b200.dist import block_cyclic_indices

    rp, ci = rmat.rmat_csr(13, 16, device="cuda", seed=2)
    n = rp.numel() - 1
    nodes = torch.arange(n)
    g = native.prepare_csr(rp, ci)
    full = g.walk(nodes.cuda(), 1.0, 0.5, 25, 8).cpu()
    saved = {k: native.get_option(k) for k in ("host_threads", "host_compress", "host_chunk_walks")}
    try:
        native.set_option("host_chunk_walks", 1024)
        for threads, compress in ((1, 1), (12, 1), (12, 2), (12, 0)):
            native.set_option("host_threads", threads)
            native.set_option("host_compress", compress)
            assert torch.equal(g.walk_to_host(nodes, 1.0, 0.5, 25, 8), full)
            # a block-cyclic shard (rank 1 of 3, blocks of 256 walks) in several pipeline chunks
            idx = block_cyclic_indices(n, 1, 3, 256)
            part = g.walk_to_host(nodes[idx].contiguous(), 1.0, 0.5, 25, 8, walk_id_offset=256, walk_id_blocks=(256, 768))
            assert torch.equal(part, full[idx])
    finally:
        for k, v in saved.items():
            native.set_option(k, v)


def test_walk_host_wire_formats_agree(native):
    """The host path sends col_idx and fetches the walks as uint32 when it has the threads for it; the
    int64 copies, the compressed path over several upload and download chunks, pageable and pinned
    buffers, and the fallback for an id that needs more than 32 bits must all give the device path's walks."""
    rp, ci = random_csr(9, 20000, 16)  # ~320 k CSR entries: five upload chunks of 65,536
    nodes = torch.randint(0, 20000, (25000,))
    dev = native.walk(rp.cuda(), ci.cuda(), nodes.cuda(), 1.0, 0.5, 21, 7, cache=False).cpu()
    saved = {k: native.get_option(k) for k in ("host_chunk_walks", "host_up_chunk", "host_threads", "host_compress")}
    try:
        native.set_option("host_chunk_walks", 4000)
        native.set_option("host_up_chunk", 1 << 16)
        native.set_option("host_threads", 12)  # enough to switch the compression on whatever the machine has
        for compress in (1, 0):
            native.set_option("host_compress", compress)
            assert torch.equal(native.walk_host(rp, ci, nodes, 1.0, 0.5, 21, 7, device=0), dev), compress
        native.set_option("host_compress", 1)
        pinned_out = torch.empty((25000, 22), dtype=torch.int64, pin_memory=True)
        pageable_out = torch.empty((25000, 22), dtype=torch.int64)
        for out in (pinned_out, pageable_out):
            native.walk_host(rp.pin_memory(), ci.pin_memory(), nodes, 1.0, 0.5, 21, 7, device=0, out=out)
            assert torch.equal(out, dev)
        # one neighbour id beyond 32 bits in the last upload chunk: the call must notice and use plain copies
        wide = ci.clone()
        wide[-3] = (1 << 33) + 5
        ref = native.walk(rp.cuda(), wide.cuda(), nodes.cuda(), 1.0, 0.5, 21, 7, cache=False).cpu()
        assert torch.equal(native.walk_host(rp, wide, nodes, 1.0, 0.5, 21, 7, device=0), ref)
        # and a start node beyond 32 bits (outside the graph: the walk stays there)
        far = nodes.clone()
        far[17] = (1 << 35) + 1
        ref = native.walk(rp.cuda(), ci.cuda(), far.cuda(), 1.0, 1.0, 21, 7, cache=False).cpu()
        got = native.walk_host(rp, ci, far, 1.0, 1.0, 21, 7, device=0)
        assert torch.equal(got, ref) and bool((got[17] == (1 << 35) + 1).all())
    finally:
        for k, v in saved.items():
            native.set_option(k, v)
