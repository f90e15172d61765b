"""Synthetic inputs for the benchmarks: R-MAT graphs in CSR form and FB15k-237-shaped triples.

Measurement plumbing, written with torch ops so it runs on the GPU that is
about to be measured (or on the CPU for tests).  Not part of the reference's surface.
"""
import torch


def rmat_edges(scale, n_edges, a=0.57, b=0.19, c=0.19, seed=20261018, device="cpu", chunk=1 << 26):
    """Directed R-MAT edge list (src, dst), ids in [0, 2^scale): one quadrant choice per bit level."""
    gen = torch.Generator(device=device).manual_seed(int(seed))
    src = torch.empty(n_edges, dtype=torch.int64, device=device)
    dst = torch.empty(n_edges, dtype=torch.int64, device=device)
    for lo in range(0, n_edges, chunk):
        m = min(chunk, n_edges - lo)
        s = torch.zeros(m, dtype=torch.int64, device=device)
        d = torch.zeros(m, dtype=torch.int64, device=device)
        for _ in range(scale):
            r = torch.rand(m, generator=gen, device=device)
            s_bit = r >= (a + b)                       # quadrants c, d
            d_bit = ((r >= a) & (r < a + b)) | (r >= a + b + c)  # quadrants b, d
            s = s * 2 + s_bit
            d = d * 2 + d_bit
        src[lo:lo + m] = s
        dst[lo:lo + m] = d
    return src, dst


def edges_to_csr(src, dst, n_nodes, symmetric=True):
    """(src, dst) -> (row_ptr[n+1], col_idx[nnz]) int64: optional symmetrisation, self-loops and
    duplicate edges removed, rows sorted -- what utils.to_csr produces for an undirected graph."""
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = src * n_nodes + dst
    if symmetric:
        key = torch.cat((key, dst * n_nodes + src))
    del src, dst
    key = torch.unique(key)  # sorted
    rows = torch.div(key, n_nodes, rounding_mode="floor")
    col_idx = (key - rows * n_nodes).contiguous()
    del key
    counts = torch.bincount(rows, minlength=n_nodes)
    row_ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=col_idx.device)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    return row_ptr.contiguous(), col_idx


def rmat_csr(scale, edge_factor=16, n_nodes=None, seed=20261018, device="cpu", n_edges=None):
    """Symmetrised, de-duplicated R-MAT CSR.  `n_nodes` (<= 2^scale) folds ids with a modulo, for
    shapes that are not a power of two (e.g. the 2,449,029-node ogbn-products shape)."""
    full = 1 << scale
    n = full if n_nodes is None else int(n_nodes)
    m = int(n_edges) if n_edges is not None else n * edge_factor
    src, dst = rmat_edges(scale, m, seed=seed, device=device)
    if n != full:
        src, dst = src % n, dst % n
    return edges_to_csr(src, dst, n)


def kg_triples(n_entities=14541, n_relations=237, n_triples=310116, seed=20261018, device="cpu", zipf=True):
    """FB15k-237-shaped triples [T,3] = (head, relation, tail); relation ids follow the entity ids
    (the shared id space of the reference's tests, tests/test_rw_triples.py:14-24)."""
    gen = torch.Generator(device=device).manual_seed(int(seed))

    def draw(n, count):
        u = torch.rand(count, generator=gen, device=device)
        if zipf:  # heavy-tailed entity popularity
            u = u ** 2.5
        return (u * n).long().clamp_(max=n - 1)

    heads = draw(n_entities, n_triples)
    tails = draw(n_entities, n_triples)
    rels = torch.randint(0, n_relations, (n_triples,), generator=gen, device=device) + n_entities
    return torch.stack((heads, rels, tails), 1).contiguous()


def relation_tail_index(triples, n_entities):
    """Vectorised equivalent of utils.build_relation_tail_index for large inputs (stable sort by
    head): returns (relation_tail_index[n_entities,2], triples sorted by head)."""
    order = torch.sort(triples[:, 0], stable=True).indices
    ts = triples[order].contiguous()
    heads = ts[:, 0]
    counts = torch.bincount(heads, minlength=n_entities)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    index = torch.stack((starts, ends - 1), 1)
    index[counts == 0] = -1
    return index.contiguous(), ts


def transitions_are_edges(row_ptr, col_idx, walks):
    """True when every consecutive pair of every walk row is a CSR edge (rows must be sorted), or a
    stay on a node without out-edges.  Vectorised per-row bisection; used to validate benchmark output."""
    a, b = walks[:, :-1].reshape(-1), walks[:, 1:].reshape(-1)
    lo, hi = row_ptr[a].clone(), row_ptr[a + 1].clone()
    empty = lo == hi
    steps = max(1, int(torch.log2((hi - lo).max().float() + 1).ceil().item()) + 1)
    for _ in range(steps):  # lower_bound of b in col_idx[lo:hi)
        mid = (lo + hi) // 2
        go_right = (col_idx[mid.clamp(max=col_idx.numel() - 1)] < b) & (lo < hi)
        lo = torch.where(go_right, mid + 1, lo)
        hi = torch.where(go_right | (lo >= hi), hi, mid)
    found = (lo < row_ptr[a + 1]) & (col_idx[lo.clamp(max=col_idx.numel() - 1)] == b)
    return bool((found | (empty & (a == b))).all())
