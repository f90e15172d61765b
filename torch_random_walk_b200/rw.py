"""Walk / window API -- drop-in for the reference's torch_rw/rw.py (same names, keyword names,
argument order and return values; /root/reference/torch_rw/rw.py:3-39).  Every function forwards
positionally to the native surface, exactly as the reference forwards to torch_rw_native."""
from . import native as torch_rw_native


def walk(row_ptr, col_idx, target_nodes, p, q, walk_length, seed):
    return torch_rw_native.walk(row_ptr, col_idx, target_nodes, p, q, walk_length, seed)


def walk_edge_list(edge_list_indexed, node_edge_index, target_nodes, p, q, walk_length, seed, padding_idx,
                   restart=True):
    return torch_rw_native.walk_edge_list(edge_list_indexed, node_edge_index, target_nodes, p, q, walk_length, seed,
                                          padding_idx, restart)


def walk_triples(triples_indexed, relation_tail_index, target_nodes, walk_length, padding_idx, seed, restart=True):
    # the native order is (..., walk_length, padding_idx, restart, seed): torch_rw/rw.py:18-26
    return torch_rw_native.walk_triples(triples_indexed, relation_tail_index, target_nodes, walk_length, padding_idx,
                                        restart, seed)


def to_windows(walks, window_size, num_nodes, seed):
    return torch_rw_native.to_windows(walks, window_size, num_nodes, seed)


def to_windows_cbow(walks, window_size, num_nodes, seed):
    return torch_rw_native.to_windows_cbow(walks, window_size, num_nodes, seed)


def to_windows_triples(walks, window_size, num_nodes, padding_idx, triples, seed):
    return torch_rw_native.to_windows_triples(walks, window_size, num_nodes, padding_idx, triples, seed)


def to_windows_triples_cbow(walks, window_size, num_nodes, padding_idx, triples, seed):
    return torch_rw_native.to_windows_triples_cbow(walks, window_size, num_nodes, padding_idx, triples, seed)
