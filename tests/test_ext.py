"""The pybind11 successor of the reference's csrc/rw_init.cpp (torch_random_walk_b200/csrc/binding):
same module surface as `torch_rw_native`, and -- on a GPU -- the same results as the ctypes path."""
import pytest
import torch

from torch_random_walk_b200 import _build_ext


@pytest.fixture(scope="module")
def ext():
    mod = _build_ext.load_module()
    if mod is None:
        pytest.skip("torch_rw_native_b200.so not built (python -m torch_random_walk_b200._build_ext)")
    return mod


def test_module_surface_and_cpu_error(ext):
    # /root/reference/csrc/rw_init.cpp:133-141
    for name in ("walk", "walk_edge_list", "walk_triples", "to_windows", "to_windows_cbow", "to_windows_triples",
                 "to_windows_triples_cbow"):
        assert callable(getattr(ext, name))
    z = torch.zeros(3, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ext.walk(z, z, z, 1.0, 1.0, 3, 1)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ext.to_windows(torch.zeros((2, 5), dtype=torch.int64), 3, 10, 1)


@pytest.mark.gpu
def test_extension_matches_ctypes_binding(ext):
    from helpers import random_csr
    from torch_random_walk_b200 import native, utils

    rp, ci = [t.cuda() for t in random_csr(21, 2000, 24)]
    nodes = torch.arange(2000, device="cuda")
    for p, q in ((1.0, 1.0), (0.5, 2.0), (1.0, 0.5)):
        assert torch.equal(ext.walk(rp, ci, nodes, p, q, 20, 9), native.walk(rp, ci, nodes, p, q, 20, 9))
    walks = native.walk(rp, ci, nodes, 1.0, 1.0, 20, 9)
    for a, b in zip(ext.to_windows(walks, 5, 2000, 3), native.to_windows(walks, 5, 2000, 3)):
        assert torch.equal(a, b)
    for a, b in zip(ext.to_windows_cbow(walks, 5, 2000, 3), native.to_windows_cbow(walks, 5, 2000, 3)):
        assert torch.equal(a, b)
    g = torch.Generator().manual_seed(1)
    el = torch.stack((torch.randint(0, 50, (300,), generator=g), torch.randint(0, 52, (300,), generator=g)), 1)
    nei, els = utils.build_node_edge_index(el, torch.arange(53))
    t = torch.arange(52, device="cuda")
    assert torch.equal(ext.walk_edge_list(els.cuda(), nei.cuda(), t, 0.7, 0.2, 9, 4, 52, True),
                       native.walk_edge_list(els.cuda(), nei.cuda(), t, 0.7, 0.2, 9, 4, 52, True))
    tr = torch.stack((el[:, 0], torch.randint(60, 65, (300,), generator=g), el[:, 1].clamp(max=49)), 1)
    rti, trs = utils.build_relation_tail_index(tr, torch.arange(50))
    t = torch.arange(50, device="cuda")
    wt = ext.walk_triples(trs.cuda(), rti.cuda(), t, 6, 70, False, 5)
    assert torch.equal(wt, native.walk_triples(trs.cuda(), rti.cuda(), t, 6, 70, False, 5))
    for a, b in zip(ext.to_windows_triples(wt, 3, 50, 70, trs.cuda(), 2), native.to_windows_triples(wt, 3, 50, 70, trs.cuda(), 2)):
        assert torch.equal(a, b)
    for a, b in zip(ext.to_windows_triples_cbow(wt, 3, 50, 70, trs.cuda(), 2),
                    native.to_windows_triples_cbow(wt, 3, 50, 70, trs.cuda(), 2)):
        assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="contigous"):
        ext.to_windows(walks[:, ::2], 3, 2000, 1)
