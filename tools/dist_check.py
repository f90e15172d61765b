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
