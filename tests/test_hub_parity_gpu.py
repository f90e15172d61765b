"""Hub-scale parity of the CSR node2vec walk with the reference's algorithm.

The chi-square / TV tests of test_walk_gpu.py run on graphs of 34-60 nodes, whose rows take the packed
short-row path or a single small table segment.  Here the graph has a hub of ~108,000 neighbours (a
segmented table built by build_hub_kernel, mixture thresholds at degree 1e5), a second hub of ~60,000
that only half of the nodes see, and cliques of eight among the leaves, so every step class occurs
(return, common neighbour, far) with hub rows on either side of the membership question.  A context
(t, v) at a hub has 1e5 outcomes, so statistics are taken per acceptance CLASS -- the quantity the
rejection rule of csrc/cuda/rw_cuda.cu:146-179 controls -- over context types that share their
analytic class probabilities by construction, plus the uniformity of the far class inside a hub row.
The oracle (the reference's CPU algorithm, linear scan and all) is compared on the same tables."""
import numpy as np
import pytest
import torch

from helpers import check_walks_follow_edges, chi2_pvalue

pytestmark = pytest.mark.gpu

N_LEAF = 120_000
HUB0, HUB1 = N_LEAF, N_LEAF + 1
N = N_LEAF + 2
PERIOD = 40  # lcm of the clique size (8) and the two hub patterns (10, 2)


def _kind(ids):
    """Context type of a node: hubs are their own kind, leaves fall into PERIOD kinds."""
    return np.where(ids >= N_LEAF, PERIOD + (ids - N_LEAF), ids % PERIOD)


def _hub_graph():
    leaf = np.arange(N_LEAF, dtype=np.int64)
    # cliques of eight consecutive leaves
    base = (leaf // 8) * 8
    src = np.repeat(leaf, 8)
    dst = (np.repeat(base, 8) + np.tile(np.arange(8), N_LEAF))
    keep = src != dst
    src, dst = src[keep], dst[keep]
    h0 = leaf[leaf % 10 != 0]  # hub 0 sees nine leaves in ten
    h1 = leaf[leaf % 2 == 0]   # hub 1 sees every other leaf
    src = np.r_[src, h0, np.full(h0.size, HUB0), h1, np.full(h1.size, HUB1), [HUB0, HUB1]]
    dst = np.r_[dst, np.full(h0.size, HUB0), h0, np.full(h1.size, HUB1), h1, [HUB1, HUB0]]
    key = np.unique(src * N + dst)
    rows, cols = key // N, key % N
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(row_ptr, rows + 1, 1)
    return np.cumsum(row_ptr), cols, key


def _class_table(walks, key):
    """[type(t), type(v), class] counts over all (t, v, x) triples; class 0 return, 1 common neighbour, 2 far."""
    w = walks.numpy() if isinstance(walks, torch.Tensor) else walks
    t, v, x = w[:, :-2].ravel(), w[:, 1:-1].ravel(), w[:, 2:].ravel()
    q = t * N + x
    pos = np.searchsorted(key, q)
    member = key[np.minimum(pos, key.size - 1)] == q
    cls = np.where(x == t, 0, np.where(member, 1, 2))
    flat = (_kind(t) * (PERIOD + 2) + _kind(v)) * 3 + cls
    return np.bincount(flat, minlength=(PERIOD + 2) ** 2 * 3).reshape(PERIOD + 2, PERIOD + 2, 3), (t, v, x, cls)


def _expected_classes(rp, ci, adj_sets, t, v, p, q):
    nb = ci[rp[v]:rp[v + 1]]
    st = adj_sets(t)
    w = np.zeros(3)
    for x in nb.tolist():
        if x == t:
            w[0] += 1.0 / p
        elif x in st:
            w[1] += 1.0
        else:
            w[2] += 1.0 / q
    return w / w.sum()


@pytest.fixture(scope="module")
def hub_graph():
    rp, ci, key = _hub_graph()
    cache = {}

    def adj_sets(node):
        if node not in cache:
            cache[node] = set(ci[rp[node]:rp[node + 1]].tolist())
        return cache[node]

    return rp, ci, key, adj_sets


LAWS = [(1.0, 0.5, "plain rejection"), (0.25, 0.5, "return-edge folding"), (0.5, 2.0, "two-sided mixture")]


@pytest.mark.parametrize("p,q,scheme", LAWS)
def test_hub_rows_follow_the_node2vec_law_and_the_oracle(hub_graph, orc, p, q, scheme):
    from torch_random_walk_b200 import native

    rp, ci, key, adj_sets = hub_graph
    deg = np.diff(rp)
    assert deg[HUB0] >= 100_000 and deg[HUB1] >= 50_000 and deg[:N_LEAF].max() < 12
    rp_t, ci_t = torch.from_numpy(rp), torch.from_numpy(ci)
    rng = np.random.default_rng(1)
    # start everywhere: 84 steps per walk, ~1.0e7 second-order transitions
    nodes = torch.from_numpy(np.r_[np.arange(N), rng.integers(0, N, 1000)])
    L = 84
    walks = native.walk(rp_t.cuda(), ci_t.cuda(), nodes.cuda(), p, q, L, 99, cache=False)
    check_walks_follow_edges(walks[:2000], rp_t, ci_t, nodes[:2000])
    # the kept graph (edge records, triangle Blooms, symmetric-graph short cuts) must not change an entry
    g = native.prepare_csr(rp_t.cuda(), ci_t.cuda())
    assert g.symmetric and torch.equal(g.walk(nodes.cuda(), p, q, L, 99), walks)
    del g
    got, (t, v, x, cls) = _class_table(walks.cpu(), key)
    assert got.sum() >= 10_000_000

    # analytic class probabilities per context type (all contexts of a type share them by construction: checked on three)
    chi2, dof, tv_sum, total = 0.0, 0, 0.0, 0
    types = np.argwhere(got.sum(2) >= 2000)
    tk, vk = _kind(t), _kind(v)
    for a, b in types:
        sel = np.flatnonzero((tk == a) & (vk == b))
        reps = sel[rng.integers(0, sel.size, 3)]
        probs = [_expected_classes(rp, ci, adj_sets, int(t[i]), int(v[i]), p, q) for i in reps]
        assert np.allclose(probs[0], probs[1]) and np.allclose(probs[0], probs[2]), (a, b)
        obs = got[a, b].astype(float)
        m = obs.sum()
        exp = probs[0] * m
        cells = exp >= 5.0
        if cells.sum() >= 2:
            chi2 += float((((obs - exp) ** 2)[cells] / exp[cells]).sum())
            dof += int(cells.sum()) - 1
        tv_sum += 0.5 * float(np.abs(obs / m - probs[0]).sum()) * m
        total += m
    assert total >= 0.98 * got.sum()  # nearly every sample sits in a tested type
    assert dof >= 100
    assert chi2_pvalue(chi2, dof) > 0.01, (scheme, chi2, dof)
    assert tv_sum / total < 1e-2, (scheme, tv_sum / total)

    # inside the far class of a hub row every far neighbour is equally likely: residues of x modulo a prime
    for hub in (HUB0, HUB1):
        sel = (v == hub) & (cls == 2)
        nb = ci[rp[hub]:rp[hub + 1]]
        expect = np.bincount(nb % 61, minlength=61).astype(float)
        obs = np.bincount(x[sel] % 61, minlength=61).astype(float)
        assert obs.sum() >= 200_000
        expect *= obs.sum() / expect.sum()
        c2 = float(((obs - expect) ** 2 / expect).sum())
        assert chi2_pvalue(c2, 60) > 0.001, (scheme, hub, c2)

    # the reference's algorithm on the same graph (linear scans of the hub rows: a few seconds for 3e5 samples)
    ref_nodes = torch.from_numpy(np.r_[np.arange(0, N_LEAF, 40), [HUB0, HUB1] * 50].astype(np.int64))
    ref_walks = orc.walk(rp_t, ci_t, ref_nodes, p, q, 100, 5)
    ref, _ = _class_table(ref_walks, key)
    assert ref.sum() >= 300_000
    c2, d2 = 0.0, 0
    for a, b in np.argwhere((got.sum(2) >= 2000) & (ref.sum(2) >= 200)):
        ga, rb = got[a, b].astype(float), ref[a, b].astype(float)
        keep = (ga + rb) >= 10
        if keep.sum() < 2:
            continue
        na, nb_ = ga[keep].sum(), rb[keep].sum()
        ea, eb = (ga + rb)[keep] * na / (na + nb_), (ga + rb)[keep] * nb_ / (na + nb_)
        c2 += float(((ga[keep] - ea) ** 2 / ea + (rb[keep] - eb) ** 2 / eb).sum())
        d2 += int(keep.sum()) - 1
    assert d2 >= 50
    assert chi2_pvalue(c2, d2) > 0.01, (scheme, "vs oracle", c2, d2)
