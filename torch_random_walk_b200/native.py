"""The `torch_rw_native` surface over libtrw_b200.so (ctypes; C ABI in include/trw_b200.h).

Stands where the reference's pybind11 module stood (/root/reference/csrc/rw_init.cpp:133-141):
the same seven functions, positional arguments in the same order, same return values.  What the
reference's launchers did around their kernels is done here -- device checks with the reference's
messages (csrc/cuda/utils.cuh:7-9), `torch.empty` outputs on the inputs' device, launch on the
current stream, no host synchronisation -- and nothing else: all arithmetic is in the CUDA
library.  There is NO CPU path: CPU tensors raise, and a missing/unloadable library raises at
import of this module.
"""
import ctypes
import os
import threading

import torch

from . import _build

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_c_ptr = ctypes.c_void_p
_c_size = ctypes.c_size_t


class _GraphView(ctypes.Structure):
    """trw_csr_graph_view (include/trw_b200.h)."""
    _fields_ = [("graph", ctypes.c_void_p), ("row_ptr", ctypes.c_void_p), ("col_idx", ctypes.c_void_p), ("ready_stream", ctypes.c_void_p)]


def _load():
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if os.environ.get("TRW_NO_AUTOBUILD"):
            raise ImportError(f"{path} is missing; run `python -m torch_random_walk_b200._build`")
        _build.build()
    lib = ctypes.CDLL(path)
    lib.trw_last_error.restype = ctypes.c_char_p
    lib.trw_launch_count.restype = _c_i64
    lib.trw_get_option.restype = _c_i64
    lib.trw_get_option.argtypes = [ctypes.c_char_p]
    lib.trw_set_option.argtypes = [ctypes.c_char_p, _c_i64]
    lib.trw_walk_csr_workspace_bytes.restype = _c_size
    lib.trw_walk_csr_workspace_bytes.argtypes = [_c_i64, _c_i64, _c_dbl, _c_dbl]
    lib.trw_walk_csr_workspace_bytes_for.restype = _c_size
    lib.trw_walk_csr_workspace_bytes_for.argtypes = [_c_i64, _c_i64, _c_dbl, _c_dbl, _c_i64, _c_int]
    lib.trw_csr_graph_workspace_bytes.restype = _c_size
    lib.trw_csr_graph_workspace_bytes.argtypes = [_c_i64, _c_i64]
    lib.trw_csr_graph_prepare.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr,
                                          ctypes.POINTER(_c_ptr)]
    lib.trw_csr_graph_prepare_ex.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr, _c_i64,
                                             ctypes.POINTER(_c_ptr)]
    lib.trw_walk_csr_prepared_at.argtypes = [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                             _c_i64, _c_ptr, _c_i64, _c_ptr]
    lib.trw_csr_graph_add_blooms.argtypes = [_c_ptr, _c_ptr, _c_ptr, _c_i64, _c_ptr]
    lib.trw_csr_checksum.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_int, _c_ptr]
    lib.trw_csr_checksum_typed.argtypes = [_c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_i64, _c_ptr, _c_int, _c_ptr]
    lib.trw_csr_graph_prepare_typed.argtypes = [_c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr, _c_i64,
                                                ctypes.POINTER(_c_ptr)]
    lib.trw_walk_csr_typed.argtypes = [_c_ptr, _c_int, _c_ptr, _c_int, _c_i64, _c_i64, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                       _c_i64, _c_ptr, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr]
    lib.trw_walk_csr_to_host.argtypes = [ctypes.POINTER(_GraphView), _c_ptr, _c_i64, _c_i64, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                         _c_i64, _c_ptr]
    lib.trw_walk_csr_prepared.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int, _c_i64, _c_ptr, _c_i64,
                                          _c_ptr]
    lib.trw_walk_csr_prepared_windows5.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int, _c_i64, _c_ptr, _c_ptr, _c_ptr]
    lib.trw_csr_checksum_host.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_int, _c_int, ctypes.POINTER(ctypes.c_uint64)]
    lib.trw_host_replica_info.argtypes = [_c_int, ctypes.POINTER(_c_i64), _c_int]
    lib.trw_csr_graph_destroy.argtypes = [_c_ptr]
    lib.trw_csr_graph_destroy.restype = None
    lib.trw_alias_table_build.argtypes = [_c_ptr, _c_i64, _c_dbl, _c_ptr]
    lib.trw_windows_alias.argtypes = [_c_ptr, _c_i64, _c_i64, _c_int, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]
    lib.trw_windows_cbow_alias.argtypes = [_c_ptr, _c_i64, _c_i64, _c_int, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]
    lib.trw_reset_options.restype = None
    lib.trw_csr_graph_info.argtypes = [_c_ptr, _c_ptr, ctypes.POINTER(_c_i64), _c_int]
    lib.trw_walk_csr.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                 _c_i64, _c_ptr, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr]
    lib.trw_walk_csr_host.argtypes = [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                      _c_i64, _c_ptr, _c_int]
    lib.trw_walk_edge_list.argtypes = [_c_ptr, _c_i64, _c_ptr, _c_i64, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                       _c_i64, _c_i64, _c_int, _c_ptr, _c_i64, _c_int, _c_ptr]
    lib.trw_walk_edge_list_workspace_bytes.restype = _c_size
    lib.trw_walk_edge_list_workspace_bytes.argtypes = [_c_i64, _c_i64, _c_dbl, _c_dbl]
    lib.trw_walk_edge_list_ws.argtypes = [_c_ptr, _c_i64, _c_ptr, _c_i64, _c_ptr, _c_i64, _c_i64, _c_dbl, _c_dbl, _c_int,
                                          _c_i64, _c_i64, _c_int, _c_ptr, _c_i64, _c_ptr, _c_size, _c_int, _c_ptr]
    lib.trw_walk_triples.argtypes = [_c_ptr, _c_i64, _c_ptr, _c_i64, _c_ptr, _c_i64, _c_i64, _c_int, _c_i64, _c_int,
                                     _c_i64, _c_ptr, _c_i64, _c_int, _c_ptr]
    win = [_c_ptr, _c_i64, _c_i64, _c_int, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]
    lib.trw_windows.argtypes = win
    lib.trw_windows_cbow.argtypes = win
    wint = [_c_ptr, _c_i64, _c_i64, _c_int, _c_i64, _c_i64, _c_ptr, _c_i64, _c_i64, _c_ptr, _c_ptr, _c_ptr, _c_int, _c_ptr]
    lib.trw_windows_triples.argtypes = wint
    lib.trw_windows_triples_cbow.argtypes = wint
    wint_ws = wint[:-2] + [_c_ptr, _c_size] + wint[-2:]
    lib.trw_windows_triples_ws.argtypes = wint_ws
    lib.trw_windows_triples_cbow_ws.argtypes = wint_ws
    lib.trw_windows_triples_workspace_bytes.restype = _c_size
    lib.trw_windows_triples_workspace_bytes.argtypes = [_c_i64]
    lib.trw_calib_gather.argtypes = [_c_ptr, _c_i64, _c_i64, _c_int, _c_int, _c_i64, _c_ptr, _c_int, _c_ptr]
    lib.trw_device_check.argtypes = [_c_int]
    if lib.trw_abi_version() != 3:
        raise ImportError("libtrw_b200.so ABI version mismatch; rebuild with python -m torch_random_walk_b200._build")
    return lib


_lib = _load()
LIB_PATH = _build.LIB_PATH


def lib():
    """The loaded ctypes library (tests check its exported symbols against include/trw_b200.h)."""
    return _lib


def _check(status):
    if status != 0:
        raise RuntimeError(_lib.trw_last_error().decode("utf-8", "replace") or f"libtrw_b200 status {status}")


def _require_cuda(t, name, int32_too=False):
    # message convention of the reference's CHECK_CUDA (csrc/cuda/utils.cuh:7)
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"(*{name}) must be a CUDA tensor")
    if int32_too and t.dtype == torch.int32:
        return  # extension: row_ptr / col_idx of the CSR walk may be int32 (the reference raises, csrc/cuda/rw_cuda.cu:206-209)
    if t.dtype != torch.int64:
        # the reference's packed_accessor64<int64_t> raises the same way
        raise RuntimeError(f"expected scalar type Long but found {str(t.dtype).replace('torch.', '')} ({name})")


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t.numel() else 0)


def set_option(name, value):
    _check(_lib.trw_set_option(name.encode(), int(value)))


def reset_options():
    """Every option of the calling thread back to the shipped default (trw_reset_options)."""
    _lib.trw_reset_options()


def get_option(name):
    return int(_lib.trw_get_option(name.encode()))


def last_kernel_ms():
    """(table build ms, walk kernel ms) of the last rw.walk call; needs set_option("time_kernels", 1)."""
    b, w = ctypes.c_float(0), ctypes.c_float(0)
    _check(_lib.trw_last_kernel_ms(ctypes.byref(b), ctypes.byref(w)))
    return b.value, w.value


def launch_count():
    return int(_lib.trw_launch_count())


def reset_launch_count():
    _lib.trw_reset_launch_count()


def _check_out(out, n, wl, dev):
    """`out=` is an extension of this binding (the reference always allocates): refuse anything the kernels would
    write outside of."""
    if out is None:
        return
    if not (isinstance(out, torch.Tensor) and out.is_cuda and out.device == dev and out.dtype == torch.int64):
        raise RuntimeError("out must be an int64 CUDA tensor on the graph's device")
    if out.dim() != 2 or out.size(0) != n or out.size(1) != wl or (n and wl > 1 and out.stride(1) != 1) or \
            (n > 1 and out.stride(0) < wl):
        raise RuntimeError(f"out must have shape ({n}, {wl}) with unit column stride and non-overlapping rows")


class PreparedCsr:
    """A CSR graph prepared once for any number of walks (trw_csr_graph_prepare): the uint32 row
    index, the membership table, the duplicate-edge check, the edge records and (blooms=True) their
    triangle Blooms live in a workspace tensor this object owns (24 bytes per CSR entry).  With
    hold=True (the explicit prepare_csr handle) it keeps `row_ptr` / `column_idx` alive and they must
    not be modified while it is in use; with hold=False every walk names the arrays it runs on, which
    the caller has verified to hold the prepared content (the graph cache of `walk` does, by
    checksum).  Walks through it are bit-identical to the one-shot call."""

    def __init__(self, row_ptr, column_idx, hold=True, blooms=True):
        _require_cuda(row_ptr, "row_ptr", int32_too=True)
        _require_cuda(column_idx, "column_idx", int32_too=True)
        row_ptr, column_idx = row_ptr.contiguous(), column_idx.contiguous()
        self.row_ptr, self.column_idx = (row_ptr, column_idx) if hold else (None, None)
        self.device = row_ptr.device
        self.n_nodes, self.nnz = max(row_ptr.numel() - 1, 0), column_idx.numel()
        self._handle = _c_ptr()
        self._destroy = _lib.trw_csr_graph_destroy  # bound now: module globals may be gone at interpreter exit
        # add_blooms rewrites what the handle points to: launches through one handle are issued one at a time, so that a
        # walk is always stream-ordered after the pass whose flags its plan reads
        self._lock = threading.Lock()
        self.has_blooms = bool(blooms)
        with torch.cuda.device(self.device):
            need = _lib.trw_csr_graph_workspace_bytes(self.n_nodes, self.nnz)
            self.workspace = torch.empty((max(need, 1),), dtype=torch.uint8, device=self.device)
            self.dtypes = (row_ptr.dtype, column_idx.dtype)
            _check(_lib.trw_csr_graph_prepare_typed(_ptr(row_ptr), row_ptr.element_size(), _ptr(column_idx), column_idx.element_size(),
                                                    self.n_nodes, self.nnz, _ptr(self.workspace) if need else None, need,
                                                    self.device.index, _stream(self.device), -1 if blooms else 0,
                                                    ctypes.byref(self._handle)))
            self._ready = torch.cuda.Event()
            self._ready.record(torch.cuda.current_stream(self.device))
            self._stream_id = torch.cuda.current_stream(self.device).cuda_stream

    def _order_after_preparation(self, stream):
        if stream.cuda_stream != self._stream_id:  # prepared on another stream: order after it, tell the allocator
            stream.wait_event(self._ready)
            self.workspace.record_stream(stream)

    def add_blooms(self, row_ptr=None, column_idx=None, cap=0):
        """Adds the triangle Blooms to a graph prepared without them (trw_csr_graph_add_blooms): one pass,
        ~0.3 s on a 0.5 G-entry R-MAT at the default cap, after which q != 1 walks probe memory far less."""
        ci = self.column_idx if column_idx is None else column_idx
        with self._lock, torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            self._order_after_preparation(stream)
            _check(_lib.trw_csr_graph_add_blooms(self._handle, None, _ptr(ci) if ci is not None else None, int(cap),
                                                 ctypes.c_void_p(stream.cuda_stream)))
            self._ready = torch.cuda.Event()
            self._ready.record(stream)
            self._stream_id = stream.cuda_stream
            self.has_blooms = True

    def walk(self, target_nodes, p, q, walk_length, seed, walk_id_offset=0, out=None, csr=None, walk_id_blocks=None):
        """`csr=(row_ptr, column_idx)`: the arrays this walk reads (required when the graph does not hold its own).
        `walk_id_blocks=(block, stride)`: global ids of a block-cyclic shard (see dist.block_cyclic_shard)."""
        id_block, id_stride = (0, 0) if walk_id_blocks is None else (int(walk_id_blocks[0]), int(walk_id_blocks[1]))
        _require_cuda(target_nodes, "target_nodes")
        dev = self.device
        if target_nodes.device != dev:
            raise RuntimeError(f"target_nodes is on {target_nodes.device}, the prepared graph on {dev}")
        rp, ci = (self.row_ptr, self.column_idx) if csr is None else csr
        if rp is None:
            raise RuntimeError("this prepared graph does not hold its CSR arrays: pass csr=(row_ptr, column_idx)")
        if (rp.dtype, ci.dtype) != self.dtypes:
            raise RuntimeError("csr arrays must have the dtypes the graph was prepared with")
        target_nodes = target_nodes.contiguous()
        n, wl = target_nodes.size(0), int(walk_length) + 1
        _check_out(out, n, wl, dev)
        with self._lock, torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            self._order_after_preparation(stream)
            walks = torch.empty((n, wl), dtype=torch.int64, device=dev) if out is None else out
            _check(_lib.trw_walk_csr_prepared_at(self._handle, _ptr(rp), _ptr(ci), _ptr(target_nodes), n, int(walk_id_offset),
                                                 id_block, id_stride, float(p), float(q), int(walk_length), int(seed),
                                                 _ptr(walks), walks.stride(0) if n else wl, _stream(dev)))
        return walks

    def walk_to_host(self, target_nodes, p, q, walk_length, seed, walk_id_offset=0, walk_id_blocks=None, out=None, csr=None):
        """The walk for start nodes and walks that live in HOST memory (trw_walk_csr_to_host): the graph stays on the
        device, chunks are walked and copied back in a pipeline.  Blocks until `out` is complete."""
        rp, ci = (self.row_ptr, self.column_idx) if csr is None else csr
        if rp is None:
            raise RuntimeError("this prepared graph does not hold its CSR arrays: pass csr=(row_ptr, column_idx)")
        target_nodes = _host_int64(target_nodes, "target_nodes")
        n, wl = target_nodes.size(0), int(walk_length) + 1
        out = _host_out(out, n, wl)
        id_block, id_stride = (0, 0) if walk_id_blocks is None else (int(walk_id_blocks[0]), int(walk_id_blocks[1]))
        with self._lock:
            view = _GraphView(self._handle, rp.data_ptr(), ci.data_ptr() if ci.numel() else None, self._stream_id)
            _check(_lib.trw_walk_csr_to_host(ctypes.byref(view), _ptr(target_nodes), n, int(walk_id_offset), id_block, id_stride,
                                             float(p), float(q), int(walk_length), int(seed), _ptr(out)))
        return out

    def walk_windows5(self, target_nodes, p, q, walk_length, seed, walk_id_offset=0):
        """A/B of the fused walk -> window pipeline (trw_walk_csr_prepared_windows5): (target_nodes[K], pos_windows[K, 4]) of
        rw.to_windows(rw.walk(...), 5, ...) without the walks being written.  Measurement only."""
        _require_cuda(target_nodes, "target_nodes")
        target_nodes = target_nodes.contiguous()
        n, per_walk = target_nodes.size(0), int(walk_length) - 3
        with self._lock, torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            self._order_after_preparation(stream)
            tgt = torch.empty((n * per_walk,), dtype=torch.int64, device=self.device)
            pos = torch.empty((n * per_walk, 4), dtype=torch.int64, device=self.device)
            _check(_lib.trw_walk_csr_prepared_windows5(self._handle, _ptr(target_nodes), n, int(walk_id_offset), float(p), float(q),
                                                       int(walk_length), int(seed), _ptr(tgt), _ptr(pos), _stream(self.device)))
        return tgt, pos

    def info(self):
        """What the preparation built (trw_csr_graph_info; waits for the preparing stream)."""
        vals = (_c_i64 * 6)()
        with torch.cuda.device(self.device):
            _check(_lib.trw_csr_graph_info(self._handle, ctypes.c_void_p(self._stream_id), vals, 6))
        keys = ("table", "records", "edge_filter_bits", "triangle_blooms", "symmetric", "table_overflowed")
        return dict(zip(keys, (int(v) for v in vals)))

    @property
    def symmetric(self):
        """True when the triangle-Bloom pass proved that every stored (t -> v) has its (v -> t)."""
        return self.info()["symmetric"] == 1

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        destroy = getattr(self, "_destroy", None)
        if h and destroy is not None:
            destroy(h)


def prepare_csr(row_ptr, column_idx, blooms=True):
    """Explicit form of what rw.walk's graph cache does: prepare once, then `.walk(...)` many times.  The
    handle keeps the two tensors alive; they must not be modified while it is in use (no checksum here)."""
    return PreparedCsr(row_ptr, column_idx, hold=True, blooms=blooms)


def csr_checksum(row_ptr, column_idx):
    """64-bit content checksum of a CSR graph on its device (trw_csr_checksum; one streaming pass, then a
    host wait for 8 bytes).  Equal sizes and checksums identify the graph a kept preparation belongs to."""
    dev = row_ptr.device
    with _checksum_lock:  # the result cells are shared by the callers of a device: one checksum at a time
        return _csr_checksum_locked(row_ptr, column_idx, dev)


def _csr_checksum_locked(row_ptr, column_idx, dev):
    bufs = _checksum_bufs.get(dev.index)
    if bufs is None:
        bufs = (torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64).pin_memory(), torch.cuda.Event())
        _checksum_bufs[dev.index] = bufs
    d_buf, h_buf, done = bufs
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _check(_lib.trw_csr_checksum_typed(_ptr(row_ptr), row_ptr.element_size(), _ptr(column_idx), column_idx.element_size(),
                                           max(row_ptr.numel() - 1, 0), column_idx.numel(), _ptr(d_buf), dev.index,
                                           ctypes.c_void_p(stream.cuda_stream)))
        h_buf.copy_(d_buf, non_blocking=True)
        done.record(stream)
        done.synchronize()
    return int(h_buf.item())


# Graph cache of the drop-in `walk`.  A user of the reference calls rw.walk(row_ptr, col_idx, ...) once
# per epoch on the same graph; everything the walk derives from the graph is the same every time, so the
# last graph per device is kept -- identified by CONTENT, not by tensor identity: every cached call first
# checksums row_ptr and col_idx on the device (one streaming pass at HBM speed, 0.7 ms for the 4.3 GB of a
# 0.5 G-entry graph, then a host wait for 8 bytes) and reuses a preparation only for equal sizes and equal
# checksum.  Writes through `.data`, raw pointers, DLPack or another library are therefore seen like any
# other, a re-created tensor with the same content hits, and the cache holds no reference to the caller's
# tensors (they are named afresh by every call).  The preparation grows with use:
#   call 1 on a graph   one-shot, everything built inside the call (the reference's stateless launcher)
#   call 2              the graph is prepared for keeps (row index, table, edge records: ~10 ms on c3)
#   call 2 + _BLOOM_AFTER_HITS   the triangle Blooms are added (~0.3 s on c3), later calls are the walk kernel alone
# It costs trw_csr_graph_workspace_bytes (24 bytes per CSR entry) of device memory per device; when that
# cannot be had the call falls back to the one-shot path.  clear_graph_cache(), set_graph_cache(False) or
# TRW_GRAPH_CACHE=0 switch it off; the one-shot path needs no checksum.
_graph_cache = {}   # device index -> dict(key=(n_nodes, nnz, checksum), graph=PreparedCsr, hits=int)
_seen_once = {}     # device index -> key of the last one-shot graph (a second call with it prepares the graph for keeps)
_no_room = {}       # device index -> key whose preparation ran out of memory (not retried)
_checksum_bufs = {}
_checksum_lock = threading.Lock()
_cache_lock = threading.RLock()
_graph_cache_on = os.environ.get("TRW_GRAPH_CACHE", "1") != "0"
_BLOOM_AFTER_HITS = int(os.environ.get("TRW_BLOOM_AFTER_HITS", "2"))


def set_graph_cache(enabled):
    global _graph_cache_on
    _graph_cache_on = bool(enabled)
    if not _graph_cache_on:
        clear_graph_cache()


def graph_cache_enabled():
    return _graph_cache_on


def clear_graph_cache():
    with _cache_lock:
        _graph_cache.clear()
        _seen_once.clear()
        _no_room.clear()


def graph_cache_state(device=None):
    """For measurements and tests: what the cache holds for `device` (default: the current one)."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    with _cache_lock:
        e = _graph_cache.get(idx)
        if e is None:
            return {"prepared": False, "hits": 0, "blooms": False}
        return {"prepared": True, "hits": e["hits"], "blooms": e["graph"].has_blooms, "key": e["key"]}


def _cached_walk(row_ptr, column_idx, target_nodes, p, q, walk_length, seed, walk_id_offset, out):
    """The cached form of `walk`, or None when this call has to take the one-shot path."""
    dev = row_ptr.device
    key = (max(row_ptr.numel() - 1, 0), column_idx.numel(), row_ptr.dtype, column_idx.dtype, csr_checksum(row_ptr, column_idx))
    with _cache_lock:
        entry = _graph_cache.get(dev.index)
        if entry is not None and entry["key"] != key:
            del _graph_cache[dev.index]  # another graph (or this one, modified): free the workspace before anything else is allocated
            entry = None
        if entry is None:
            if _seen_once.get(dev.index) != key or _no_room.get(dev.index) == key:
                _seen_once[dev.index] = key
                return None
            try:
                graph = PreparedCsr(row_ptr, column_idx, hold=False, blooms=False)
            except torch.cuda.OutOfMemoryError:
                _no_room[dev.index] = key  # 24 bytes per entry do not fit beside the caller's data: stay one-shot
                return None
            entry = {"key": key, "graph": graph, "hits": 0}
            _graph_cache[dev.index] = entry
        else:
            entry["hits"] += 1
            if entry["hits"] >= _BLOOM_AFTER_HITS and not entry["graph"].has_blooms and not (p == 1.0 and q == 1.0):
                entry["graph"].add_blooms(column_idx=column_idx)
        graph = entry["graph"]
    return graph.walk(target_nodes, p, q, walk_length, seed, walk_id_offset=walk_id_offset, out=out, csr=(row_ptr, column_idx))


def walk(row_ptr, column_idx, target_nodes, p, q, walk_length, seed, walk_id_offset=0, out=None, cache=None):
    """csrc/rw_init.cpp:11-25 -> csrc/cuda/rw_cuda.cu:186-248.  Returns walks[n, walk_length+1] on
    row_ptr's device.  `walk_id_offset` / `out` are extensions for sharded callers; `cache` overrides
    the graph cache for this call (None: the module setting).  With the cache, repeated calls on a graph
    of the same content reuse its preparation (see the comment above _graph_cache); cached and one-shot
    calls return identical walks.  Extension: row_ptr and column_idx may each be int32 (half the graph in
    HBM; the reference raises on anything but int64) -- start nodes and walks stay int64 and the walks
    equal those of the int64 call on the same values."""
    _require_cuda(row_ptr, "row_ptr", int32_too=True)
    _require_cuda(column_idx, "column_idx", int32_too=True)
    _require_cuda(target_nodes, "target_nodes")
    dev = row_ptr.device
    if column_idx.device != dev or target_nodes.device != dev:
        raise RuntimeError(f"row_ptr, column_idx and target_nodes must be on one device (got {dev}, {column_idx.device}, "
                           f"{target_nodes.device})")
    use_cache = _graph_cache_on if cache is None else bool(cache)
    n_nodes, nnz = max(row_ptr.numel() - 1, 0), column_idx.numel()
    row_ptr, column_idx, target_nodes = row_ptr.contiguous(), column_idx.contiguous(), target_nodes.contiguous()
    if use_cache and nnz > 0 and target_nodes.size(0) > 0:
        walks = _cached_walk(row_ptr, column_idx, target_nodes, p, q, walk_length, seed, walk_id_offset, out)
        if walks is not None:
            return walks
    n, wl = target_nodes.size(0), int(walk_length) + 1
    _check_out(out, n, wl, dev)
    with torch.cuda.device(dev):
        walks = torch.empty((n, wl), dtype=torch.int64, device=dev) if out is None else out
        need = _lib.trw_walk_csr_workspace_bytes_for(n_nodes, nnz, float(p), float(q), n, int(walk_length)) if n else 0
        ws = torch.empty((need,), dtype=torch.uint8, device=dev) if need else None
        _check(_lib.trw_walk_csr_typed(_ptr(row_ptr), row_ptr.element_size(), _ptr(column_idx), column_idx.element_size(), n_nodes,
                                       nnz, _ptr(target_nodes), n, int(walk_id_offset), float(p), float(q), int(walk_length),
                                       int(seed), _ptr(walks), walks.stride(0) if n else wl, _ptr(ws) if ws is not None else None,
                                       need, dev.index, _stream(dev)))
    return walks


def walk_edge_list(edge_list_indexed, node_edges_idx, target_nodes, p, q, walk_length, seed, padding_idx, restart,
                   walk_id_offset=0):
    """csrc/rw_init.cpp:27-45 -> csrc/cuda/rw_cuda_edge_list.cu:243-308.  Output on
    node_edges_idx's device (rw_cuda_edge_list.cu:260)."""
    _require_cuda(edge_list_indexed, "edge_list_indexed")
    _require_cuda(node_edges_idx, "node_edges_idx")
    _require_cuda(target_nodes, "target_nodes")
    dev = node_edges_idx.device
    el, nei, tg = edge_list_indexed.contiguous(), node_edges_idx.contiguous(), target_nodes.contiguous()
    n, wl = tg.size(0), int(walk_length) + 1
    with torch.cuda.device(dev):
        walks = torch.empty((n, wl), dtype=torch.int64, device=dev)
        need = _lib.trw_walk_edge_list_workspace_bytes(el.size(0), nei.size(0), float(p), float(q)) if n else 0
        ws = torch.empty((need,), dtype=torch.uint8, device=dev) if need else None
        _check(_lib.trw_walk_edge_list_ws(_ptr(el), el.size(0), _ptr(nei), nei.size(0), _ptr(tg), n, int(walk_id_offset),
                                          float(p), float(q), int(walk_length), int(seed), int(padding_idx),
                                          1 if restart else 0, _ptr(walks), wl, _ptr(ws) if ws is not None else None, need,
                                          dev.index, _stream(dev)))
    return walks


def walk_triples(triples_indexed, relation_tail_index, target_nodes, walk_length, padding_idx, restart, seed,
                 walk_id_offset=0):
    """csrc/rw_init.cpp:47-75 -> csrc/cuda/rw_cuda_triples.cu:103-169.  Output [n, 2*walk_length+1]
    on target_nodes' device (rw_cuda_triples.cu:118)."""
    _require_cuda(triples_indexed, "triples_indexed")
    _require_cuda(relation_tail_index, "relation_tail_index")
    _require_cuda(target_nodes, "target_nodes")
    dev = target_nodes.device
    tr, rti, tg = triples_indexed.contiguous(), relation_tail_index.contiguous(), target_nodes.contiguous()
    n, wl = tg.size(0), 2 * int(walk_length) + 1
    with torch.cuda.device(dev):
        walks = torch.empty((n, wl), dtype=torch.int64, device=dev)
        _check(_lib.trw_walk_triples(_ptr(tr), tr.size(0), _ptr(rti), rti.size(0), _ptr(tg), n, int(walk_id_offset),
                                     int(walk_length), int(padding_idx), 1 if restart else 0, int(seed), _ptr(walks),
                                     wl, dev.index, _stream(dev)))
    return walks


def _require_walks(walks):
    _require_cuda(walks, "walks")
    if walks.dim() != 2:
        raise RuntimeError("walks must be a 2-d tensor")
    if not walks.is_contiguous():
        # CHECK_CONTIGUOUS, csrc/cuda/utils.cuh:9 (spelling kept)
        raise RuntimeError("walks must be a contigous tensor")


def _node_windows(fn, walks, window_size, num_nodes, seed, cbow, neg_table=None):
    _require_walks(walks)
    dev = walks.device
    n, wl = walks.shape
    w = int(window_size)
    k = (wl - w + 1) * n
    if neg_table is not None:
        if not (isinstance(neg_table, torch.Tensor) and neg_table.is_cuda and neg_table.device == dev and neg_table.dtype == torch.int64 and
                neg_table.dim() == 1 and neg_table.is_contiguous() and neg_table.numel() == int(num_nodes)):
            raise RuntimeError("neg_table must be the tensor native.negative_table made for num_nodes nodes, on the walks' device")
        fn = _lib.trw_windows_cbow_alias if cbow else _lib.trw_windows_alias
    with torch.cuda.device(dev):
        first = torch.empty((k,), dtype=torch.int64, device=dev)
        win_a = torch.empty((k, w - 1), dtype=torch.int64, device=dev)
        other = torch.empty((k,) if cbow else (k, w - 1), dtype=torch.int64, device=dev)
        outs = (first, other, win_a) if cbow else (first, win_a, other)
        table = () if neg_table is None else (_ptr(neg_table),)
        _check(fn(_ptr(walks), n, wl, w, int(num_nodes), int(seed), *table, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]),
                  dev.index, _stream(dev)))
    return outs


def negative_table(weights, power=0.75, device=None):
    """Alias table for drawing window negatives with P(v) ~ weights[v]**power instead of uniformly (trw_alias_table_build; SURVEY
    section 8 f3 -- word2vec's unigram^0.75 with weights = node degrees).  Built on the host in O(n); returns an int64 tensor of
    num_nodes cells on `device` (default: the current CUDA device) for the `neg_table=` argument of to_windows / to_windows_cbow.
    An extension: the reference's negatives are uniform, and so are ours without it."""
    w = torch.as_tensor(weights).detach().to("cpu", torch.float64).contiguous()
    if w.dim() != 1 or w.numel() == 0:
        raise RuntimeError("weights must be a non-empty 1-D tensor (one weight per node)")
    table = torch.empty(w.numel(), dtype=torch.int64)
    _check(_lib.trw_alias_table_build(_ptr(w), w.numel(), float(power), _ptr(table)))
    return table.to(torch.device("cuda", torch.cuda.current_device()) if device is None else device)


def to_windows(walks, window_size, num_nodes, seed, neg_table=None):
    """csrc/rw_init.cpp:77-88 -> (target_nodes[K], pos_windows[K,W-1], neg_windows[K,W-1]).  neg_table (extension): see
    negative_table."""
    return _node_windows(_lib.trw_windows, walks, window_size, num_nodes, seed, cbow=False, neg_table=neg_table)


def to_windows_cbow(walks, window_size, num_nodes, seed, neg_table=None):
    """csrc/rw_init.cpp:90-101 -> (pos_nodes[K], neg_nodes[K], windows[K,W-1]).  neg_table (extension): see negative_table."""
    return _node_windows(_lib.trw_windows_cbow, walks, window_size, num_nodes, seed, cbow=True, neg_table=neg_table)


def _triple_windows(fn, walks, window_size, num_nodes, padding_idx, triples, seed, cbow):
    _require_walks(walks)
    _require_cuda(triples, "triples")
    dev = walks.device
    triples = triples.contiguous()
    n, wl = walks.shape
    w = int(window_size)
    k = ((wl - 1) // 2) * n
    with torch.cuda.device(dev):
        first = torch.empty((k, 3), dtype=torch.int64, device=dev)
        win_a = torch.empty((k, 2 * w, 3), dtype=torch.int64, device=dev)
        other = torch.empty((k, 3) if cbow else (k, 2 * w, 3), dtype=torch.int64, device=dev)
        outs = (first, other, win_a) if cbow else (first, win_a, other)
        need = _lib.trw_windows_triples_workspace_bytes(triples.size(0))
        ws = torch.empty((need,), dtype=torch.uint8, device=dev) if need else None
        _check(fn(_ptr(walks), n, wl, w, int(num_nodes), int(padding_idx), _ptr(triples), triples.size(0), int(seed),
                  _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(ws) if ws is not None else None, need, dev.index, _stream(dev)))
    return outs


def to_windows_triples(walks, window_size, num_nodes, padding_idx, triples, seed):
    """csrc/rw_init.cpp:103-116 -> (target_triples[K,3], pos_windows[K,2W,3], neg_windows[K,2W,3])."""
    return _triple_windows(_lib.trw_windows_triples_ws, walks, window_size, num_nodes, padding_idx, triples, seed, False)


def to_windows_triples_cbow(walks, window_size, num_nodes, padding_idx, triples, seed):
    """csrc/rw_init.cpp:118-131 -> (pos_triples[K,3], neg_triples[K,3], pos_windows[K,2W,3])."""
    return _triple_windows(_lib.trw_windows_triples_cbow_ws, walks, window_size, num_nodes, padding_idx, triples, seed, True)


def _host_int64(t, name):
    if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.int64:
        raise RuntimeError(f"(*{name}) must be an int64 CPU tensor for the host path")
    return t.contiguous()


def _host_out(out, n, wl):
    """`out=` of the host entries: the C side writes n*wl int64 values through a raw pointer, so anything but a
    contiguous int64 CPU tensor of exactly that shape is refused here."""
    if out is None:
        return torch.empty((n, wl), dtype=torch.int64, pin_memory=True)
    if not (isinstance(out, torch.Tensor) and not out.is_cuda and out.dtype == torch.int64 and out.dim() == 2 and
            out.size(0) == n and out.size(1) == wl and out.is_contiguous()):
        raise RuntimeError(f"out must be a contiguous int64 CPU tensor of shape ({n}, {wl})")
    return out


def csr_checksum_host(row_ptr, column_idx, threads=0, simd=True):
    """The checksum of csr_checksum for int64 CSR arrays in host memory (trw_csr_checksum_host): what the host path compares
    with its kept device replica.  `simd=False` keeps to the scalar loop (same value)."""
    row_ptr, column_idx = _host_int64(row_ptr, "row_ptr"), _host_int64(column_idx, "column_idx")
    out = ctypes.c_uint64(0)
    _check(_lib.trw_csr_checksum_host(_ptr(row_ptr), _ptr(column_idx), max(row_ptr.numel() - 1, 0), column_idx.numel(), int(threads),
                                      1 if simd else 0, ctypes.byref(out)))
    v = int(out.value)
    return v - (1 << 64) if v >= (1 << 63) else v  # as csr_checksum reports it (an int64 cell)


def host_replica_info(device=0):
    """What trw_walk_csr_host keeps on `device` (trw_host_replica_info)."""
    out = (ctypes.c_int64 * 6)()
    _check(_lib.trw_host_replica_info(int(device), out, 6))
    return {"held": bool(out[0]), "level": int(out[1]), "hits": int(out[2]),
            "last_call": {0: "none", 1: "kept replica validated", 2: "fresh upload", 3: "kept replica found changed"}[int(out[3])],
            "h2d_bytes": int(out[4]), "d2h_bytes": int(out[5])}  # what the last host-path call moved over PCIe


def walk_host(row_ptr, column_idx, target_nodes, p, q, walk_length, seed, device=0, walk_id_offset=0, out=None):
    """End-to-end entry for callers holding CPU tensors (trw_walk_csr_host): the graph and start
    nodes are copied to `device`, walked in chunks, and the walks are streamed back into a (pinned)
    CPU tensor.  A caller that comes back with the same arrays finds the device replica kept (its content
    is re-checked by checksum on every call).  Still the CUDA path -- there is no CPU implementation to
    fall back to."""
    row_ptr, column_idx = _host_int64(row_ptr, "row_ptr"), _host_int64(column_idx, "column_idx")
    target_nodes = _host_int64(target_nodes, "target_nodes")
    n, wl = target_nodes.size(0), int(walk_length) + 1
    out = _host_out(out, n, wl)
    _check(_lib.trw_walk_csr_host(_ptr(row_ptr), _ptr(column_idx), max(row_ptr.numel() - 1, 0), column_idx.numel(),
                                  _ptr(target_nodes), n, int(walk_id_offset), float(p), float(q), int(walk_length),
                                  int(seed), _ptr(out), int(device)))
    return out
