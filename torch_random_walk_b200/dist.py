"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on the GPU box, gloo in
the CPU tests).  The reference has no distributed code; walks are
independent, so the only collective is the one-off replication of the CSR, and start nodes are
sharded with no traffic while walking.

Each shard passes its global walk offset to the kernels, which key Philox by the *global* walk id:
the concatenation of all shards is bit-identical to a single-GPU call for every world size.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world_size: int):
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_targets(target_nodes: torch.Tensor, rank=None, world_size=None):
    """-> (this rank's slice of target_nodes, global index of its first walk)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(target_nodes.size(0), rank, world_size)
    return target_nodes[lo:hi], lo


def replicate_csr(row_ptr, col_idx, src=0, device=None, group=None):
    """Broadcast the CSR from rank `src` to every rank: one size exchange, then one broadcast each
    for row_ptr and col_idx (ncclBroadcast over NVLink with the nccl backend).  Ranks other than
    `src` may pass None."""
    rank = dist.get_rank(group)
    if device is None:
        device = row_ptr.device if row_ptr is not None else torch.device("cpu")
    sizes = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        sizes[0], sizes[1] = row_ptr.numel(), col_idx.numel()
    dist.broadcast(sizes, src, group=group)
    n_rp, n_ci = int(sizes[0]), int(sizes[1])
    if rank != src:
        row_ptr = torch.empty(n_rp, dtype=torch.int64, device=device)
        col_idx = torch.empty(n_ci, dtype=torch.int64, device=device)
    else:
        row_ptr, col_idx = row_ptr.to(device).contiguous(), col_idx.to(device).contiguous()
    dist.broadcast(row_ptr, src, group=group)
    dist.broadcast(col_idx, src, group=group)
    return row_ptr, col_idx


def gather_walks(local_walks: torch.Tensor, n_total: int, group=None):
    """All-gather the per-rank walk shards (contiguous sharding) into the caller's order:
    [n_total, row_len] on every rank."""
    world = dist.get_world_size(group)
    row_len = local_walks.size(1)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = local_walks
    if buf.size(0) != pad:  # equalise shard sizes for all_gather_into_tensor
        buf = torch.cat((buf, buf.new_zeros((pad - buf.size(0), row_len))))
    out = torch.empty((world * pad, row_len), dtype=local_walks.dtype, device=local_walks.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    return torch.cat([out[r * pad: r * pad + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def walk_sharded(row_ptr, col_idx, target_nodes, p, q, walk_length, seed, gather=False):
    """rw.walk over all ranks: every rank holds the (replicated) CSR and the full target list,
    walks its contiguous shard with global walk ids, and optionally gathers the result."""
    from . import native

    local, offset = shard_targets(target_nodes)
    walks = native.walk(row_ptr, col_idx, local.contiguous(), p, q, walk_length, seed, walk_id_offset=offset)
    if gather:
        return gather_walks(walks, target_nodes.size(0))
    return walks
