"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): sharding arithmetic, the CSR
broadcast and the gather that reassembles walk shards in the caller's order."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_random_walk_b200 import dist as trw_dist


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [trw_dist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_block_cyclic_shards_partition_the_list():
    for n in (0, 1, 4095, 4096, 4097, 100003):
        for world in (1, 2, 3, 8):
            for block in (1, 7, 4096):
                parts = [trw_dist.block_cyclic_indices(n, r, world, block) for r in range(world)]
                assert [p.numel() for p in parts] == [trw_dist.block_cyclic_count(n, r, world, block) for r in range(world)]
                allidx = torch.cat(parts)
                assert allidx.numel() == n and torch.equal(torch.sort(allidx).values, torch.arange(n))
                sizes = [p.numel() for p in parts]
                assert max(sizes) - min(sizes) <= block
                # the (offset, block, stride) triple the kernels use reproduces the index list
                for r, p in enumerate(parts):
                    i = torch.arange(p.numel())
                    assert torch.equal(p, r * block + (i // block) * (block * world) + i % block)


def test_walk_digest_adds_up_over_any_partition():
    g = torch.Generator().manual_seed(3)
    walks = torch.randint(0, 1 << 40, (5000, 17), generator=g)
    ids = torch.arange(5000)
    whole = trw_dist.walk_digest(walks, ids)
    for world, block in ((2, 64), (3, 4096), (8, 5)):
        parts = [trw_dist.block_cyclic_indices(5000, r, world, block) for r in range(world)]
        total = sum(int(trw_dist.walk_digest(walks[p], p)) for p in parts)
        assert (total - int(whole)) % (1 << 64) == 0
    perm = torch.randperm(5000, generator=g)
    assert int(trw_dist.walk_digest(walks[perm], ids[perm])) == int(whole)
    changed = walks.clone()
    changed[123, 5] += 1
    assert int(trw_dist.walk_digest(changed, ids)) != int(whole)
    assert int(trw_dist.walk_digest(walks, ids + 1)) != int(whole)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if rank == 0:
            row_ptr = torch.tensor([0, 2, 3, 5], dtype=torch.int64)
            col_idx = torch.tensor([1, 2, 0, 0, 1], dtype=torch.int64)
        else:
            row_ptr = col_idx = None
        row_ptr, col_idx = trw_dist.replicate_csr(row_ptr, col_idx, src=0, device=torch.device("cpu"))
        targets = torch.arange(11, dtype=torch.int64)
        local, offset = trw_dist.shard_targets(targets)
        # stand-in for the walk: a row that is a pure function of the GLOBAL walk id, as the kernels' Philox is
        gid = torch.arange(offset, offset + local.numel())
        fake = torch.stack((local, gid * 7 + 1, gid * gid), 1)
        full = trw_dist.gather_walks(fake, targets.numel())
        # the same through a block-cyclic split (blocks of two walks dealt round-robin)
        local_bc, (off, blk, stride) = trw_dist.shard_targets_block_cyclic(targets, block=2)
        i = torch.arange(local_bc.numel())
        gid_bc = off + (i // blk) * stride + i % blk
        fake_bc = torch.stack((local_bc, gid_bc * 7 + 1, gid_bc * gid_bc), 1)
        full_bc = trw_dist.gather_walks_block_cyclic(fake_bc, targets.numel(), block=2)
        assert torch.equal(full_bc, full)
        ret[rank] = (row_ptr.tolist(), col_idx.tolist(), offset, full.tolist())
    finally:
        dist.destroy_process_group()


def test_replicate_shard_gather_world2():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        out = dict(ret)
    g = torch.arange(11)
    expected = torch.stack((g, g * 7 + 1, g * g), 1).tolist()
    for rank in range(world):
        rp, ci, offset, full = out[rank]
        assert rp == [0, 2, 3, 5] and ci == [1, 2, 0, 0, 1]
        assert full == expected
    assert out[0][2] == 0 and out[1][2] == 6
