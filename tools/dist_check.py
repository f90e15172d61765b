"""Multi-GPU check over NCCL (run under torchrun): CSR replicated by broadcast, start nodes sharded,
walks gathered -- the result must equal the single-GPU call bit for bit (global walk ids).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch_random_walk_b200 import dist as trw_dist  # noqa: E402
from torch_random_walk_b200 import native, rmat  # noqa: E402


def c_abi_collectives(rp, ci, nodes, dev, rank, world):
    """trw_replicate_csr / trw_gather_walks (include/trw_b200.h) over a communicator of our own, as a binder without
    Python would use them: ncclCommInitRank through ctypes on the NCCL the process has loaded."""
    import ctypes

    lib = native.lib()
    if not lib.trw_nccl_available():
        print("NCCL entry points not found: C-ABI collectives skipped", flush=True)
        return True
    nccl_path = [ln.split()[-1] for ln in open("/proc/self/maps") if "libnccl" in ln][0]
    nccl = ctypes.CDLL(nccl_path)

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_ubyte * 128)]

    uid = UniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    box = [ctypes.string_at(ctypes.byref(uid), 128) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctypes.memmove(ctypes.byref(uid), box[0], 128)
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    nccl.ncclCommInitRank.restype = ctypes.c_int
    rc = nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank -> {rc}"
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.trw_replicate_csr.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]
    lib.trw_gather_walks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]
    ok = True
    for rp_t, ci_t in ((rp, ci), (rp.int(), ci.int())):
        rp2 = rp_t.clone() if rank == 0 else torch.zeros_like(rp_t)
        ci2 = ci_t.clone() if rank == 0 else torch.zeros_like(ci_t)
        rc = lib.trw_replicate_csr(comm, 0, rp2.data_ptr(), rp2.element_size(), rp2.numel() - 1, ci2.data_ptr(), ci2.element_size(),
                                   ci2.numel(), stream)
        torch.cuda.synchronize()
        ok = ok and rc == 0 and torch.equal(rp2, rp_t) and torch.equal(ci2, ci_t)
    lo, hi = trw_dist.shard_bounds(nodes.numel(), rank, world)
    part = native.walk(rp, ci, nodes[lo:hi].contiguous(), 1.0, 0.5, 20, 5, walk_id_offset=lo, cache=False)
    out = torch.empty((nodes.numel(), 21), dtype=torch.int64, device=dev)
    rows = (ctypes.c_int64 * world)(*[trw_dist.shard_bounds(nodes.numel(), r, world)[1] - trw_dist.shard_bounds(nodes.numel(), r, world)[0]
                                      for r in range(world)])
    rc = lib.trw_gather_walks(comm, rank, world, part.data_ptr(), 21, out.data_ptr(), rows, stream)
    torch.cuda.synchronize()
    single = native.walk(rp, ci, nodes, 1.0, 0.5, 20, 5, cache=False)
    ok = ok and rc == 0 and torch.equal(out, single)
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    if rank == 0:
        print(f"C-ABI collectives (trw_replicate_csr int64 + int32, trw_gather_walks) over our own communicator: {ok}", flush=True)
    return ok


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        rp, ci = rmat.rmat_csr(18, 16, device=dev, seed=4)
    else:
        rp = ci = None
    rp, ci = trw_dist.replicate_csr(rp, ci, src=0, device=dev)
    nodes = torch.nonzero(rp[1:] - rp[:-1] > 0).flatten().contiguous()
    ok = True
    for p, q in ((1.0, 1.0), (1.0, 0.5), (0.5, 2.0)):
        for _ in range(3):  # the third call goes through the kept graph of the cache
            gathered = trw_dist.walk_sharded(rp, ci, nodes, p, q, 40, 123, gather=True)
        single = native.walk(rp, ci, nodes, p, q, 40, 123, cache=False)
        same = torch.equal(gathered, single)
        ok = ok and same
        if rank == 0:
            print(f"p={p} q={q}: world={world} walks={nodes.numel()} gathered == single-GPU: {same}", flush=True)
    # block-cyclic shards through the replicated-graph handle: the shards' digests add up to the single call's
    rep = trw_dist.ReplicatedCsr(rp if rank == 0 else None, ci if rank == 0 else None, src=0, device=dev)
    for p, q in ((1.0, 0.5), (0.5, 2.0)):
        local_nodes, off, blocks, gids = rep.shard(nodes, "block_cyclic", 1024)
        part = rep.walk_local(local_nodes, p, q, 40, 123, off, blocks)
        digest = trw_dist.walk_digest(part, gids).reshape(1)
        dist.all_reduce(digest, op=dist.ReduceOp.SUM)
        single = native.walk(rp, ci, nodes, p, q, 40, 123, cache=False)
        same = int(trw_dist.walk_digest(single, torch.arange(nodes.numel(), device=dev))) == int(digest.item())
        same = same and torch.equal(part, single[gids])
        host_part = rep.walk_local_to_host(local_nodes.cpu(), p, q, 40, 123, off, blocks)
        same = same and torch.equal(host_part, part.cpu())
        ok = ok and same
        if rank == 0:
            print(f"p={p} q={q}: block-cyclic shards (device and host path) == single-GPU: {same}", flush=True)
    ok = ok and c_abi_collectives(rp, ci, nodes, dev, rank, world)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        raise SystemExit("multi-GPU result differs from the single-GPU call")
    if rank == 0:
        print("dist check ok")


if __name__ == "__main__":
    main()
