"""ctypes front end of oracle/trw_oracle.c, mirroring torch_rw.rw's signatures on CPU tensors.

TEST INFRASTRUCTURE ONLY (see oracle/trw_oracle.c).  Each function follows the reference entry
point named in its docstring and returns torch.int64 CPU tensors of the reference's shapes.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtrw_oracle.so")
_lib = None

_I64P = ctypes.c_void_p
_i64 = ctypes.c_int64


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "trw_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-Wall", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _c(t):
    t = torch.as_tensor(t)
    assert t.dtype == torch.int64 and t.device.type == "cpu", "oracle takes int64 CPU tensors"
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def walk(row_ptr, col_idx, target_nodes, p, q, walk_length, seed):
    """rw.walk on CPU tensors: csrc/cpu/rw_cpu.cpp:203-226."""
    row_ptr, col_idx, target_nodes = _c(row_ptr), _c(col_idx), _c(target_nodes)
    n = target_nodes.numel()
    out = torch.empty((n, walk_length + 1), dtype=torch.int64)
    lib().orc_walk_csr(_p(row_ptr), _p(col_idx), _i64(col_idx.numel()), _p(target_nodes), _i64(n),
                       ctypes.c_double(p), ctypes.c_double(q), ctypes.c_int(walk_length),
                       ctypes.c_int(seed), _p(out))
    return out


def walk_edge_list(edge_list_indexed, node_edge_index, target_nodes, p, q, walk_length, seed,
                   padding_idx, restart=True):
    """rw.walk_edge_list on CPU tensors: csrc/cpu/rw_cpu_edge_list.cpp:240-266."""
    el, nei, tg = _c(edge_list_indexed), _c(node_edge_index), _c(target_nodes)
    n = tg.numel()
    out = torch.empty((n, walk_length + 1), dtype=torch.int64)
    lib().orc_walk_edge_list(_p(el), _p(nei), _i64(nei.size(0)), _p(tg), _i64(n),
                             ctypes.c_double(p), ctypes.c_double(q), ctypes.c_int(walk_length),
                             ctypes.c_int(seed), _i64(padding_idx), ctypes.c_int(bool(restart)),
                             _p(out))
    return out


def walk_triples(triples_indexed, relation_tail_index, target_nodes, walk_length, padding_idx,
                 seed, restart=True):
    """rw.walk_triples on CPU tensors: csrc/cpu/rw_cpu_triples.cpp:105-127."""
    tr, rti, tg = _c(triples_indexed), _c(relation_tail_index), _c(target_nodes)
    n = tg.numel()
    out = torch.empty((n, 2 * walk_length + 1), dtype=torch.int64)
    lib().orc_walk_triples(_p(tr), _p(rti), _p(tg), _i64(n), ctypes.c_int(walk_length),
                           _i64(padding_idx), ctypes.c_int(bool(restart)), ctypes.c_int(seed),
                           _p(out))
    return out


def to_windows(walks, window_size, num_nodes, seed):
    """rw.to_windows on CPU tensors: csrc/cpu/windows_cpu.cpp:5-77."""
    walks = _c(walks)
    n, wl = walks.shape
    k = (wl - window_size + 1) * n
    target = torch.empty((k,), dtype=torch.int64)
    pos = torch.empty((k, window_size - 1), dtype=torch.int64)
    neg = torch.empty((k, window_size - 1), dtype=torch.int64)
    lib().orc_windows(_p(walks), _i64(n), _i64(wl), ctypes.c_int(window_size), _i64(num_nodes),
                      ctypes.c_int(seed), _p(target), _p(pos), _p(neg))
    return target, pos, neg


def to_windows_cbow(walks, window_size, num_nodes, seed):
    """rw.to_windows_cbow on CPU tensors: csrc/cpu/windows_cpu.cpp:80-159."""
    walks = _c(walks)
    n, wl = walks.shape
    k = (wl - window_size + 1) * n
    pos_nodes = torch.empty((k,), dtype=torch.int64)
    neg_nodes = torch.empty((k,), dtype=torch.int64)
    windows = torch.empty((k, window_size - 1), dtype=torch.int64)
    lib().orc_windows_cbow(_p(walks), _i64(n), _i64(wl), ctypes.c_int(window_size),
                           _i64(num_nodes), ctypes.c_int(seed), _p(pos_nodes), _p(neg_nodes),
                           _p(windows))
    return pos_nodes, neg_nodes, windows


def to_windows_triples(walks, window_size, num_nodes, padding_idx, triples, seed):
    """rw.to_windows_triples on CPU tensors: csrc/cpu/windows_cpu.cpp:161-310."""
    walks, triples = _c(walks), _c(triples)
    n, wl = walks.shape
    k = ((wl - 1) // 2) * n
    target = torch.empty((k, 3), dtype=torch.int64)
    pos = torch.empty((k, window_size * 2, 3), dtype=torch.int64)
    neg = torch.empty((k, window_size * 2, 3), dtype=torch.int64)
    lib().orc_windows_triples(_p(walks), _i64(n), _i64(wl), ctypes.c_int(window_size),
                              _i64(num_nodes), _i64(padding_idx), _p(triples),
                              _i64(triples.size(0)), ctypes.c_int(seed), _p(target), _p(pos),
                              _p(neg))
    return target, pos, neg


def to_windows_triples_cbow(walks, window_size, num_nodes, padding_idx, triples, seed):
    """rw.to_windows_triples_cbow on CPU tensors: csrc/cpu/windows_cpu.cpp:312-475."""
    walks, triples = _c(walks), _c(triples)
    n, wl = walks.shape
    k = ((wl - 1) // 2) * n
    pos_triples = torch.empty((k, 3), dtype=torch.int64)
    neg_triples = torch.empty((k, 3), dtype=torch.int64)
    pos_windows = torch.empty((k, window_size * 2, 3), dtype=torch.int64)
    lib().orc_windows_triples_cbow(_p(walks), _i64(n), _i64(wl), ctypes.c_int(window_size),
                                   _i64(num_nodes), _i64(padding_idx), _p(triples),
                                   _i64(triples.size(0)), ctypes.c_int(seed), _p(pos_triples),
                                   _p(neg_triples), _p(pos_windows))
    return pos_triples, neg_triples, pos_windows
