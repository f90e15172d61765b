"""Graph -> tensor preparation; drop-in for the reference's torch_rw/utils.py.

Same function names, argument meaning and outputs (int64, contiguous CPU tensors) as
/root/reference/torch_rw/utils.py:5-120, without its O(n^2) Python loops, its float32 round
trip (utils.py:7-8 is wrong above 2^24) and its dependency on an API networkx 3 removed
(`nx.to_scipy_sparse_matrix`, utils.py:6).  These run on the host, like the reference's: they
feed the CUDA walk kernels, they are not part of them.
"""
import numpy as np
import torch


def to_csr(graph):
    """networkx graph -> (row_ptr[n+1], col_idx[nnz]).  Reference: torch_rw/utils.py:5-9.

    Node id = position in graph.nodes() order; rows are sorted and duplicate edges merged
    (scipy canonical CSR, what `nx.to_scipy_sparse_matrix(graph, format='csr')` produced);
    edge weights are dropped.
    """
    import networkx as nx

    csr = nx.to_scipy_sparse_array(graph, format="csr")
    csr.sum_duplicates()
    csr.sort_indices()
    row_ptr = torch.from_numpy(np.asarray(csr.indptr, dtype=np.int64)).contiguous()
    col_idx = torch.from_numpy(np.asarray(csr.indices, dtype=np.int64)).contiguous()
    return row_ptr, col_idx


def nodes_tensor(graph):
    """arange(number_of_nodes) as int64.  Reference: torch_rw/utils.py:11-18 (which finds each
    node's position with list.index, i.e. its own position, in O(n^2))."""
    return torch.arange(graph.number_of_nodes(), dtype=torch.int64).contiguous()


def to_edge_list_indexed(graph):
    """networkx graph -> (edge_list_indexed[E,2], node_index_mapping).  Reference:
    torch_rw/utils.py:21-56.

    A node's id is its rank among sorted(graph.nodes()).  The mapping dict holds only nodes that
    occur in an edge, in order of first occurrence (head before tail) -- the reference's tests
    take `list(node_idx_map.values())` as start nodes, so the order is part of the contract.
    Undirected graphs get the reversed edges appended (utils.py:52-54).
    """
    import networkx as nx

    edges = list(graph.edges())
    rank = {node: i for i, node in enumerate(sorted(graph.nodes()))}
    node_index_mapping = {}
    flat = np.empty((len(edges), 2), dtype=np.int64)
    for i, (head, tail) in enumerate(edges):
        h = node_index_mapping.get(head)
        if h is None:
            h = node_index_mapping[head] = rank[head]
        t = node_index_mapping.get(tail)
        if t is None:
            t = node_index_mapping[tail] = rank[tail]
        flat[i, 0] = h
        flat[i, 1] = t
    edge_list_indexed = torch.from_numpy(flat).contiguous()
    if not nx.is_directed(graph):
        edge_list_indexed = torch.cat((edge_list_indexed, torch.fliplr(edge_list_indexed)), dim=0)
    return edge_list_indexed, node_index_mapping


def _sort_rows_by_head(rows: np.ndarray) -> np.ndarray:
    # The reference sorts with pandas `sort_values(by="head")` (utils.py:62, 96): a single-key
    # sort with the default kind, i.e. numpy's argsort(kind="quicksort") on the head column.
    # Using the same call keeps the (unstable) order of rows that share a head identical.
    order = np.argsort(rows[:, 0], kind="quicksort")
    return rows[order]


def _head_ranges(heads: np.ndarray, num_nodes: int) -> torch.Tensor:
    """Inclusive [first,last] row range per head id, [-1,-1] for ids without rows.

    Vectorised restatement of the loop at torch_rw/utils.py:74-87 / 106-118, including its
    single-row corner: with exactly one row the loop only ever writes column 0, leaving [0,-1].
    """
    num_rows = heads.shape[0]
    if num_rows == 0:
        # the reference reads row 0 unconditionally (utils.py:73 / 105)
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
    if heads.min() < 0 or heads.max() >= num_nodes:
        raise IndexError(f"head id out of range for an index with {num_nodes} rows")
    index = np.full((num_nodes, 2), -1, dtype=np.int64)
    starts = np.flatnonzero(np.r_[True, heads[1:] != heads[:-1]])
    ends = np.r_[starts[1:] - 1, num_rows - 1]
    index[heads[starts], 0] = starts
    index[heads[starts], 1] = ends
    if num_rows == 1:
        index[heads[0], 1] = -1
    return torch.from_numpy(index).contiguous()


def _sorted_rows_and_ranges_torch(rows: torch.Tensor, num_nodes: int):
    """Device-side (or large-input) twin of _sort_rows_by_head + _head_ranges: one stable sort by
    head, then segment boundaries with bincount/cumsum -- no Python loop, runs where `rows` lives.
    Rows that share a head keep their input order (the reference's pandas sort leaves that order
    unspecified), everything else is identical, including the single-row corner ([0,-1])."""
    rows = rows.to(torch.int64)
    if rows.size(0) == 0:
        raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
    heads = rows[:, 0]
    if int(heads.min()) < 0 or int(heads.max()) >= num_nodes:
        raise IndexError(f"head id out of range for an index with {num_nodes} rows")
    order = torch.sort(heads, stable=True).indices
    rows = rows[order].contiguous()
    counts = torch.bincount(rows[:, 0], minlength=num_nodes)
    ends = torch.cumsum(counts, 0)
    index = torch.stack((ends - counts, ends - 1), 1)
    index[counts == 0] = -1
    if rows.size(0) == 1:
        index[rows[0, 0], 1] = -1
    return index.contiguous(), rows


def build_node_edge_index(edge_list_indexed, nodes_tensor):
    """(edge_list_indexed[E,2], node ids) -> (node_edge_index[N,2], edge list sorted by head).
    Reference: torch_rw/utils.py:58-89.  N = number of distinct ids in nodes_tensor.
    CUDA inputs are processed on the device (stable sort) and CUDA tensors are returned."""
    if torch.is_tensor(edge_list_indexed) and edge_list_indexed.is_cuda:
        num_nodes = int(torch.unique(torch.as_tensor(nodes_tensor)).numel())
        return _sorted_rows_and_ranges_torch(edge_list_indexed.reshape(-1, 2), num_nodes)
    rows = np.ascontiguousarray(torch.as_tensor(edge_list_indexed).cpu().numpy()).astype(np.int64, copy=False)
    rows = _sort_rows_by_head(rows.reshape(-1, 2))
    num_nodes = int(torch.unique(torch.as_tensor(nodes_tensor)).numel())
    node_edge_index = _head_ranges(rows[:, 0], num_nodes)
    return node_edge_index, torch.from_numpy(np.ascontiguousarray(rows)).contiguous()


def build_relation_tail_index(triples_indexed_tensor, all_entities_tensor):
    """(triples[T,3], entity ids) -> (relation_tail_index[N,2], triples sorted by head).
    Reference: torch_rw/utils.py:91-120.  N = len(all_entities_tensor) (not de-duplicated, as in
    the reference); float inputs are truncated to int64 after sorting, like `.to(int)` there.
    CUDA inputs are processed on the device (stable sort) and CUDA tensors are returned."""
    if torch.is_tensor(triples_indexed_tensor) and triples_indexed_tensor.is_cuda:
        return _sorted_rows_and_ranges_torch(triples_indexed_tensor.reshape(-1, 3), int(torch.as_tensor(all_entities_tensor).numel()))
    raw = np.ascontiguousarray(torch.as_tensor(triples_indexed_tensor).cpu().numpy()).reshape(-1, 3)
    rows = _sort_rows_by_head(raw).astype(np.int64)
    num_nodes = int(torch.as_tensor(all_entities_tensor).numel())
    relation_tail_index = _head_ranges(rows[:, 0], num_nodes)
    return relation_tail_index, torch.from_numpy(np.ascontiguousarray(rows)).contiguous()


def csr_from_edge_index(edge_index, num_nodes, symmetric=False):
    """Tensor edge list -> (row_ptr[n+1], col_idx[nnz]) with the properties of `to_csr` output:
    rows sorted, duplicate edges merged, int64, contiguous, on the device of `edge_index`
    ([2,E] or [E,2]).  `symmetric=True` adds the reversed edges (an undirected graph, as
    `to_csr(nx.Graph)` yields).  Vectorised: one sort (torch.unique) and a bincount."""
    e = torch.as_tensor(edge_index).to(torch.int64)
    if e.dim() != 2 or (e.size(0) != 2 and e.size(1) != 2):
        raise ValueError("edge_index must be [2,E] or [E,2]")
    src, dst = (e[0], e[1]) if e.size(0) == 2 and e.size(1) != 2 else (e[:, 0], e[:, 1])
    if src.numel() and (int(torch.min(torch.minimum(src, dst))) < 0 or int(torch.max(torch.maximum(src, dst))) >= num_nodes):
        raise IndexError(f"node id out of range for a graph with {num_nodes} nodes")
    key = src * num_nodes + dst
    if symmetric:
        key = torch.cat((key, dst * num_nodes + src))
    key = torch.unique(key)
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    col_idx = (key - rows * num_nodes).contiguous()
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=col_idx.device)
    torch.cumsum(torch.bincount(rows, minlength=num_nodes), 0, out=row_ptr[1:])
    return row_ptr.contiguous(), col_idx


# ---------------------------------------------------------------------------------------------
# Interchange formats (the reference has none: its only inputs are networkx
# objects).  A graph prepared once can be kept on disk as two arrays and handed to rw.walk as they
# are -- int32 arrays included, which halve the file, the PCIe transfer and the graph in HBM.
# ---------------------------------------------------------------------------------------------
def compact_csr(row_ptr, col_idx):
    """(row_ptr, col_idx) in the narrowest integer type rw.walk takes: int32 where every value fits
    (node ids < 2^31 for col_idx, nnz < 2^31 for row_ptr), int64 otherwise.  Values are unchanged."""
    def narrow(t):
        t = torch.as_tensor(t)
        if t.numel() == 0 or (int(t.max()) < 2 ** 31 and int(t.min()) >= -(2 ** 31)):
            return t.to(torch.int32).contiguous()
        return t.to(torch.int64).contiguous()

    return narrow(row_ptr), narrow(col_idx)


def save_csr(path, row_ptr, col_idx, compact=True):
    """Writes a CSR graph to `path`: '.npz' (numpy, arrays `row_ptr` / `col_idx`) or '.pt' (torch.save of
    a dict with the same keys).  compact=True stores int32 where the values allow it."""
    if compact:
        row_ptr, col_idx = compact_csr(row_ptr, col_idx)
    row_ptr, col_idx = torch.as_tensor(row_ptr).cpu().contiguous(), torch.as_tensor(col_idx).cpu().contiguous()
    if str(path).endswith(".npz"):
        np.savez(path, row_ptr=row_ptr.numpy(), col_idx=col_idx.numpy())
    elif str(path).endswith(".pt"):
        torch.save({"row_ptr": row_ptr, "col_idx": col_idx}, path)
    else:
        raise ValueError("save_csr writes '.npz' or '.pt'")


def load_csr(path, device=None, dtype=None):
    """Reads what save_csr wrote -> (row_ptr, col_idx) on `device` (default CPU).  dtype=None keeps the
    stored integer types (rw.walk takes int32 and int64); torch.int64 gives the reference's dtype."""
    if str(path).endswith(".npz"):
        with np.load(path) as z:
            row_ptr, col_idx = torch.from_numpy(z["row_ptr"]), torch.from_numpy(z["col_idx"])
    elif str(path).endswith(".pt"):
        d = torch.load(path, map_location="cpu")
        row_ptr, col_idx = d["row_ptr"], d["col_idx"]
    else:
        raise ValueError("load_csr reads '.npz' or '.pt'")
    if row_ptr.dim() != 1 or col_idx.dim() != 1 or row_ptr.numel() == 0 or int(row_ptr[-1]) != col_idx.numel():
        raise ValueError("not a CSR graph: row_ptr[-1] must equal len(col_idx)")
    if dtype is not None:
        row_ptr, col_idx = row_ptr.to(dtype), col_idx.to(dtype)
    if device is not None:
        row_ptr, col_idx = row_ptr.to(device), col_idx.to(device)
    return row_ptr.contiguous(), col_idx.contiguous()


def load_edge_index(path, num_nodes=None, symmetric=False, device=None):
    """Edge list file -> CSR: '.npy' (an integer array [E,2] or [2,E]) or '.pt' (a tensor of that shape,
    or a dict with key 'edge_index').  num_nodes defaults to max id + 1.  Rows sorted, duplicates
    merged (csr_from_edge_index); symmetric=True adds the reversed edges."""
    if str(path).endswith(".npy"):
        e = torch.from_numpy(np.load(path))
    elif str(path).endswith(".pt"):
        e = torch.load(path, map_location="cpu")
        if isinstance(e, dict):
            e = e["edge_index"]
    else:
        raise ValueError("load_edge_index reads '.npy' or '.pt'")
    e = torch.as_tensor(e).to(torch.int64)
    if device is not None:
        e = e.to(device)
    if num_nodes is None:
        num_nodes = int(e.max()) + 1 if e.numel() else 0
    return csr_from_edge_index(e, num_nodes, symmetric=symmetric)
