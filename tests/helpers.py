"""Shared builders for the tests: the reference's toy graph, random CSR graphs, analytic node2vec."""
import numpy as np
import torch


def toy_edges():
    # the 5-node / 7-edge graph every reference test builds (tests/test_rw.py:31-40)
    return [("A", "B"), ("A", "C"), ("B", "C"), ("B", "D"), ("D", "C"), ("E", "A"), ("E", "D")]


def toy_graph(directed=False):
    import networkx as nx

    g = nx.DiGraph() if directed else nx.Graph()
    for a, b in toy_edges():
        g.add_edge(a, b)
    return g


def random_csr(seed, n, avg_deg, symmetric=True, sort_rows=True):
    rng = np.random.default_rng(seed)
    m = n * avg_deg // (2 if symmetric else 1)
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if symmetric:
        src, dst = np.r_[src, dst], np.r_[dst, src]
    key = np.unique(src.astype(np.int64) * n + dst)
    src, dst = key // n, key % n
    if not sort_rows:  # shuffle inside rows
        order = np.lexsort((rng.random(len(src)), src))
        src, dst = src[order], dst[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return torch.from_numpy(np.cumsum(row_ptr)), torch.from_numpy(dst.astype(np.int64))


def edge_set(row_ptr, col_idx):
    rp, ci = row_ptr.cpu().numpy(), col_idx.cpu().numpy()
    n = len(rp) - 1
    src = np.repeat(np.arange(n), np.diff(rp))
    return set(zip(src.tolist(), ci.tolist())), n


def check_walks_follow_edges(walks, row_ptr, col_idx, targets):
    """Structural contract of rw.walk: column 0 = targets, every consecutive pair is an edge, or a
    stay on a node without out-edges (csrc/cuda/rw_cuda.cu:25-30)."""
    w = walks.cpu().numpy()
    rp, ci = row_ptr.cpu().numpy(), col_idx.cpu().numpy()
    n = len(rp) - 1
    assert np.array_equal(w[:, 0], targets.cpu().numpy())
    assert w.min() >= 0 and w.max() < n
    a, b = w[:, :-1].ravel(), w[:, 1:].ravel()
    deg = np.diff(rp)
    key = np.unique(np.repeat(np.arange(n), deg).astype(np.int64) * n + ci)
    is_edge = np.isin(a.astype(np.int64) * n + b, key)
    stay = (deg[a] == 0) & (a == b)
    assert np.all(is_edge | stay), f"{(~(is_edge | stay)).sum()} transitions are not edges"


def node2vec_probs(row_ptr, col_idx, p, q):
    """Analytic second-order transition table: {(t, v): (neighbours of v, probabilities)}.
    P(x | t, v) proportional to mult_v(x) * alpha, alpha = 1/p (x == t), 1 (x in adj(t)), 1/q else
    (csrc/cuda/rw_cuda.cu:156-175)."""
    rp, ci = row_ptr.cpu().numpy(), col_idx.cpu().numpy()
    n = len(rp) - 1
    adj = [ci[rp[i]:rp[i + 1]] for i in range(n)]
    sets = [set(a.tolist()) for a in adj]
    table = {}
    for t in range(n):
        for v in sets[t]:
            nb = adj[v]
            if len(nb) == 0:
                continue
            w = np.array([1.0 / p if x == t else (1.0 if x in sets[t] else 1.0 / q) for x in nb])
            table[(t, int(v))] = (nb, w / w.sum())
    return table


def second_order_counts(walks, n):
    """Counts of (t, v, x) over all consecutive triples of the walks, as a dict keyed by t*n*n+v*n+x."""
    w = walks.cpu().numpy().astype(np.int64)
    t, v, x = w[:, :-2].ravel(), w[:, 1:-1].ravel(), w[:, 2:].ravel()
    keys, counts = np.unique((t * n + v) * n + x, return_counts=True)
    return dict(zip(keys.tolist(), counts.tolist()))


def chi2_and_tv(counts, table, n, min_expected=5.0):
    """Pearson chi-square (pooled over all (t,v) contexts, cells with expectation >= min_expected),
    its degrees of freedom, and the count-weighted mean total variation distance."""
    from collections import defaultdict

    ctx = defaultdict(dict)
    for key, c in counts.items():
        tv_, x = divmod(key, n)
        t, v = divmod(tv_, n)
        ctx[(t, v)][x] = c
    chi2, dof, tv_sum, total = 0.0, 0, 0.0, 0
    for (t, v), obs in ctx.items():
        nb, pr = table[(t, v)]
        # merge duplicate neighbour ids
        probs = defaultdict(float)
        for x, pp in zip(nb.tolist(), pr.tolist()):
            probs[x] += pp
        m = sum(obs.values())
        assert set(obs) <= set(probs), "walk took a transition that is not an edge"
        tv = 0.5 * sum(abs(obs.get(x, 0) / m - pp) for x, pp in probs.items())
        tv_sum += tv * m
        total += m
        cells = [(obs.get(x, 0), m * pp) for x, pp in probs.items() if m * pp >= min_expected]
        if len(cells) >= 2:
            chi2 += sum((o - e) ** 2 / e for o, e in cells)
            dof += len(cells) - 1
    return chi2, dof, tv_sum / max(total, 1)


def two_sample_chi2(counts_a, counts_b, n, min_total=10):
    """Homogeneity test of two sets of (t,v,x) counts, conditioned on the (t,v) context: sum over
    contexts of the 2 x k contingency chi-square.  Returns (chi2, dof)."""
    from collections import defaultdict

    ctx = defaultdict(lambda: defaultdict(lambda: [0, 0]))
    for which, counts in enumerate((counts_a, counts_b)):
        for key, c in counts.items():
            tv_, x = divmod(key, n)
            ctx[tv_][x][which] += c
    chi2, dof = 0.0, 0
    for cells in ctx.values():
        rows = [ab for ab in cells.values() if ab[0] + ab[1] >= min_total]
        na, nb = sum(a for a, _ in rows), sum(b for _, b in rows)
        if len(rows) < 2 or na == 0 or nb == 0:
            continue
        for a, b in rows:
            ea, eb = (a + b) * na / (na + nb), (a + b) * nb / (na + nb)
            chi2 += (a - ea) ** 2 / ea + (b - eb) ** 2 / eb
        dof += len(rows) - 1
    return chi2, dof


def chi2_pvalue(chi2, dof):
    from scipy.stats import chi2 as chi2_dist

    return float(chi2_dist.sf(chi2, dof)) if dof > 0 else 1.0
