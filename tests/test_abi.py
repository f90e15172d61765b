"""The drop-in boundary without a GPU: libtrw_b200.so loads, exports every symbol that
include/trw_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "trw_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from torch_random_walk_b200 import native

    lib = native.lib()
    names = declared_symbols()
    assert len(names) >= 16
    for name in names:
        assert getattr(lib, name) is not None, name
    assert lib.trw_abi_version() == 3


def test_library_has_no_torch_or_oracle_dependency():
    from torch_random_walk_b200 import native

    out = os.popen(f"ldd {native.LIB_PATH}").read()
    assert "torch" not in out and "oracle" not in out and "c10" not in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torch_random_walk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} mentions the oracle"


def test_reference_signatures_are_kept():
    import inspect

    from torch_random_walk_b200 import rw

    # /root/reference/torch_rw/rw.py:3-39
    expect = {
        "walk": ["row_ptr", "col_idx", "target_nodes", "p", "q", "walk_length", "seed"],
        "walk_edge_list": ["edge_list_indexed", "node_edge_index", "target_nodes", "p", "q", "walk_length", "seed",
                           "padding_idx", "restart"],
        "walk_triples": ["triples_indexed", "relation_tail_index", "target_nodes", "walk_length", "padding_idx", "seed",
                         "restart"],
        "to_windows": ["walks", "window_size", "num_nodes", "seed"],
        "to_windows_cbow": ["walks", "window_size", "num_nodes", "seed"],
        "to_windows_triples": ["walks", "window_size", "num_nodes", "padding_idx", "triples", "seed"],
        "to_windows_triples_cbow": ["walks", "window_size", "num_nodes", "padding_idx", "triples", "seed"],
    }
    for name, params in expect.items():
        assert list(inspect.signature(getattr(rw, name)).parameters) == params, name
    assert inspect.signature(rw.walk_edge_list).parameters["restart"].default is True
    assert inspect.signature(rw.walk_triples).parameters["restart"].default is True
    import torch_rw.rw as shim  # the reference's import path
    import torch_rw_native

    assert shim.walk is rw.walk and callable(torch_rw_native.to_windows_triples_cbow)


def test_cpu_tensors_raise_like_check_cuda():
    from torch_random_walk_b200 import rw

    z = torch.zeros(3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):  # csrc/cuda/utils.cuh:7
        rw.walk(z, z, z, 1.0, 1.0, 3, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        rw.to_windows(torch.zeros((2, 5), dtype=torch.int64), 3, 10, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        rw.walk_triples(z.view(1, 3), z.view(1, 3)[:, :2], z, 2, 9, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_means_error_not_fallback():
    from torch_random_walk_b200 import native

    lib = native.lib()
    assert lib.trw_device_check(0) == -2  # TRW_ERR_DEVICE
    assert b"no CPU fallback" in lib.trw_last_error()
    buf = (ctypes.c_int64 * 8)()
    rc = lib.trw_walk_csr(buf, buf, 1, 1, buf, 1, 0, 1.0, 1.0, 2, 1, buf, 3, None, 0, 0, None)
    assert rc == -2
    rc = lib.trw_windows(buf, 1, 4, 2, 5, 1, buf, buf, buf, 0, None)
    assert rc == -2


def test_argument_validation_precedes_device_use():
    from torch_random_walk_b200 import native

    lib = native.lib()
    buf = (ctypes.c_int64 * 8)()
    assert lib.trw_walk_csr(buf, buf, 1, 1, buf, -1, 0, 1.0, 1.0, 2, 1, buf, 3, None, 0, 0, None) == -1
    assert lib.trw_walk_csr(buf, buf, 1, 1, buf, 1, 0, 1.0, 1.0, 2, 1, buf, 2, None, 0, 0, None) == -1  # stride < L+1
    assert lib.trw_set_option(b"no_such_option", 1) == -1
    assert lib.trw_get_option(b"stage_output") in (0, 1)
    short_ws = lib.trw_walk_csr_workspace_bytes_for(10, 100, 1.0, 1.0, 5, 10)  # short walk: only the uint32 row index
    assert 11 * 4 <= short_ws <= 1024
    # long walks add the 16-byte edge records; the unconditional bound covers them
    assert lib.trw_walk_csr_workspace_bytes_for(10, 100, 1.0, 1.0, 50, 10) >= short_ws + 100 * 16
    assert lib.trw_walk_csr_workspace_bytes(10, 100, 1.0, 1.0) >= short_ws + 100 * 16
    assert lib.trw_walk_csr_workspace_bytes_for(10, 100, 0.5, 2.0, 5, 10) >= 100 * 8 + short_ws
    assert lib.trw_csr_graph_workspace_bytes(10, 100) >= 100 * 24
    handle = ctypes.c_void_p()
    assert lib.trw_csr_graph_prepare(buf, buf, 1, 1, None, 0, 0, None, None) == -1  # null out_graph
    assert lib.trw_walk_csr_prepared(None, buf, 1, 0, 1.0, 1.0, 2, 1, buf, 3, None) == -1  # null graph
    lib.trw_csr_graph_destroy(None)


def test_options_belong_to_the_calling_thread():
    """trw_set_option changes the knobs of the calling thread only: another thread keeps the shipped defaults."""
    import ctypes
    import threading

    from torch_random_walk_b200 import native

    lib = native.lib()
    default = native.get_option("edge_bloom_cap")
    seen = {}

    def other():
        seen["before"] = native.get_option("edge_bloom_cap")
        native.set_option("edge_bloom_cap", default + 7)
        seen["own"] = native.get_option("edge_bloom_cap")

    native.set_option("edge_bloom_cap", default + 1)
    try:
        t = threading.Thread(target=other)
        t.start()
        t.join()
        assert seen == {"before": default, "own": default + 7}
        assert native.get_option("edge_bloom_cap") == default + 1
    finally:
        native.set_option("edge_bloom_cap", default)
    assert lib.trw_set_option(b"no_such_option", ctypes.c_int64(1)) != 0


def test_reset_options_restores_the_shipped_defaults():
    from torch_random_walk_b200 import native

    before = {k: native.get_option(k) for k in ("edge_bloom_cap", "edge_filter_mb", "host_chunk_walks", "records", "win_direct_pos")}
    for k in before:
        native.set_option(k, 5)
    native.reset_options()
    assert {k: native.get_option(k) for k in before} == before
    assert before["edge_bloom_cap"] == 256 and before["edge_filter_mb"] == 32 and before["records"] == -1


def test_host_checksum_forms_agree():
    """trw_csr_checksum_host (what the host path compares with its kept device replica): the AVX-512 loop, the scalar
    loop and a numpy restatement of the definition give one value, for every length around the vector width and any
    thread count.  Pure host code: no device is touched."""
    import numpy as np

    from torch_random_walk_b200 import native

    def mix(z):
        z = z ^ (z >> np.uint64(30))
        z = z * np.uint64(0xBF58476D1CE4E5B9)
        z = z ^ (z >> np.uint64(27))
        z = z * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    def restated(rp, ci):
        total = np.uint64(0)
        with np.errstate(over="ignore"):
            for arr, golden in ((ci, 0x9E3779B97F4A7C15), (rp, 0xD6E8FEB86659FD93)):
                pos = np.arange(1, arr.size + 1, dtype=np.uint64) * np.uint64(golden)
                total = total + mix(arr.astype(np.uint64) + pos).sum(dtype=np.uint64)
        v = int(total)
        return v - (1 << 64) if v >= (1 << 63) else v

    gen = torch.Generator().manual_seed(5)
    for nnz in (0, 1, 7, 8, 15, 16, 17, 31, 33, 100, 4099):
        ci = torch.randint(-5, 1 << 40, (nnz,), generator=gen, dtype=torch.int64)
        rp = torch.randint(0, 1 << 33, (nnz % 13 + 1,), generator=gen, dtype=torch.int64)
        want = restated(rp.numpy(), ci.numpy())
        for threads in (1, 3, 16):
            assert native.csr_checksum_host(rp, ci, threads=threads, simd=True) == want, (nnz, threads)
            assert native.csr_checksum_host(rp, ci, threads=threads, simd=False) == want, (nnz, threads)
    ci = torch.arange(100, dtype=torch.int64)
    rp = torch.tensor([0, 100])
    a = native.csr_checksum_host(rp, ci)
    ci[50], ci[51] = ci[51].item(), ci[50].item()  # position-sensitive: a swap is seen
    assert native.csr_checksum_host(rp, ci) != a


def test_alias_table_reproduces_the_weights():
    """trw_alias_table_build (host code): the table's cells must add up to P(v) ~ weights[v]**power -- a node's mass is
    what its own cell keeps plus what the cells aliased to it give away -- for skewed, flat, sparse and single-node inputs."""
    import numpy as np

    from torch_random_walk_b200 import native

    lib = native.lib()
    rng = np.random.default_rng(3)
    cases = [(rng.pareto(1.2, 5000) + 1.0, 0.75), (np.ones(17), 0.75), (np.r_[np.zeros(50), rng.integers(1, 9, 200)].astype(float), 0.75),
             (np.array([7.0]), 0.75), (rng.integers(0, 1000, 4096).astype(float), 1.0), (np.r_[1e6, np.ones(999)], 0.5)]
    for weights, power in cases:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        table = np.zeros(w.size, dtype=np.uint64)
        rc = lib.trw_alias_table_build(w.ctypes.data_as(ctypes.c_void_p), w.size, float(power), table.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0, native.lib().trw_last_error()
        keep = (table & np.uint64(0xFFFFFFFF)).astype(np.float64) / 2.0**32
        alias = (table >> np.uint64(32)).astype(np.int64)
        assert alias.min() >= 0 and alias.max() < w.size
        mass = keep.copy()
        np.add.at(mass, alias, 1.0 - keep)
        want = np.where(w > 0, w**power, 0.0)
        want = want / want.sum() * w.size
        assert np.allclose(mass, want, rtol=0, atol=1e-6 * max(1.0, want.max())), (w.size, power, np.abs(mass - want).max())
        assert (keep[w == 0] < 1e-9).all()  # a node of weight 0 is never drawn: its own cell keeps nothing, nothing is aliased to it
        assert not np.isin(alias[keep < 1.0 - 1e-9], np.nonzero(w == 0)[0]).any()
    bad = np.array([1.0, -1.0])
    out = np.zeros(2, dtype=np.uint64)
    assert lib.trw_alias_table_build(bad.ctypes.data_as(ctypes.c_void_p), 2, 0.75, out.ctypes.data_as(ctypes.c_void_p)) != 0
    zero = np.zeros(4)
    assert lib.trw_alias_table_build(zero.ctypes.data_as(ctypes.c_void_p), 4, 0.75, out.ctypes.data_as(ctypes.c_void_p)) != 0
