"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on the GPU box, gloo in
the CPU tests).  The reference has no distributed code; walks are
independent, so the only collective is the one-off replication of the CSR, and start nodes are
sharded with no traffic while walking.

Each shard passes its global walk ids to the kernels, which key Philox by the *global* walk id:
the union of all shards is bit-identical to a single-GPU call for every world size and for both
shard layouts -- contiguous slices (the simplest drop-in) and block-cyclic (every world-th block of
4096 walks: start nodes sorted by id are sorted by degree on R-MAT-like graphs, and interleaving
blocks evens the ranks out).
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


DEFAULT_BLOCK = 4096


def block_cyclic_count(n_items: int, rank: int, world_size: int, block: int = DEFAULT_BLOCK) -> int:
    """Number of items rank `rank` owns when blocks of `block` consecutive items are dealt round-robin."""
    n_items, block = int(n_items), int(block)
    full_rounds, rest = divmod(n_items, block * world_size)
    return full_rounds * block + min(max(rest - rank * block, 0), block)


def block_cyclic_indices(n_items: int, rank: int, world_size: int, block: int = DEFAULT_BLOCK, device=None):
    """Global indices of the items rank `rank` owns, in its local order: local item i is global
    rank*block + (i // block) * (block*world_size) + i % block."""
    m = block_cyclic_count(n_items, rank, world_size, block)
    i = torch.arange(m, dtype=torch.int64, device=device)
    blk = torch.div(i, block, rounding_mode="floor")
    return rank * block + blk * (block * world_size) + (i - blk * block)


def shard_targets_block_cyclic(target_nodes: torch.Tensor, rank=None, world_size=None, block: int = DEFAULT_BLOCK):
    """-> (this rank's start nodes, (walk_id_offset, walk_id_block, walk_id_stride)) for a block-cyclic shard:
    the triple is what the walk needs to number its local walks globally."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    idx = block_cyclic_indices(target_nodes.size(0), rank, world_size, block, device=target_nodes.device)
    return target_nodes[idx].contiguous(), (rank * block, block, block * world_size)


def walk_digest(walks: torch.Tensor, global_ids: torch.Tensor) -> torch.Tensor:
    """Order- and partition-independent digest of walk rows: a wrapping int64 sum over rows of a hash of
    (row content, global walk id).  The digests of the shards of any partition add up to the digest of the
    single call, which is how bench.py checks on real multi-GPU runs that sharding changed nothing."""
    cols = walks.size(1)
    w = (torch.arange(1, cols + 1, dtype=torch.int64, device=walks.device) * 0x9E3779B1 + 0x7F4A7C15) | 1
    h = (walks * w).sum(1) + global_ids.to(walks.device) * 0x632BE5AB
    h = (h ^ (h >> 29)) * 0x2545F4914F6CDD1D
    h = h ^ (h >> 32)
    return h.sum()


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


def gather_walks_block_cyclic(local_walks: torch.Tensor, n_total: int, block: int = DEFAULT_BLOCK, group=None):
    """All-gather the shards of a block-cyclic split (shard_targets_block_cyclic) back into the caller's order:
    [n_total, row_len] on every rank."""
    world = dist.get_world_size(group)
    row_len = local_walks.size(1)
    counts = [block_cyclic_count(n_total, r, world, block) for r in range(world)]
    pad = max(counts) if counts else 0
    buf = local_walks
    if buf.size(0) != pad:  # equalise shard sizes for all_gather_into_tensor
        buf = torch.cat((buf, buf.new_zeros((pad - buf.size(0), row_len))))
    flat = torch.empty((world * pad, row_len), dtype=local_walks.dtype, device=local_walks.device)
    dist.all_gather_into_tensor(flat, buf.contiguous(), group=group)
    out = torch.empty((n_total, row_len), dtype=local_walks.dtype, device=local_walks.device)
    for r in range(world):
        out[block_cyclic_indices(n_total, r, world, block, device=local_walks.device)] = flat[r * pad: r * pad + counts[r]]
    return out


def walk_sharded(row_ptr, col_idx, target_nodes, p, q, walk_length, seed, gather=False):
    """rw.walk over all ranks: every rank holds the (replicated) CSR and the full target list,
    walks its contiguous shard with global walk ids, and optionally gathers the result."""
    from . import native

    local, offset = shard_targets(target_nodes)
    walks = native.walk(row_ptr, col_idx, local.contiguous(), p, q, walk_length, seed, walk_id_offset=offset)
    if gather:
        return gather_walks(walks, target_nodes.size(0))
    return walks


class ReplicatedCsr:
    """The north star's multi-GPU form: one NCCL broadcast puts the CSR on every rank, each rank prepares it
    once (native.prepare_csr: table, edge records, triangle Blooms) and then walks its shard of any start-node
    list with global walk ids -- no traffic while walking.  The replica belongs to this object (the ranks
    other than `src` never had another copy), so it is kept without per-call validation; `src`'s own
    tensors must not be modified while it lives."""

    def __init__(self, row_ptr, col_idx, src=0, device=None, group=None, blooms=True):
        from . import native

        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.row_ptr, self.col_idx = replicate_csr(row_ptr, col_idx, src=src, device=device, group=group)
        self.graph = native.prepare_csr(self.row_ptr, self.col_idx, blooms=blooms)

    def shard(self, target_nodes, layout="block_cyclic", block=DEFAULT_BLOCK):
        """-> (local start nodes, walk_id_offset, walk_id_blocks or None, global ids of the local walks)."""
        n = target_nodes.size(0)
        if layout == "contiguous":
            lo, hi = shard_bounds(n, self.rank, self.world)
            return target_nodes[lo:hi].contiguous(), lo, None, torch.arange(lo, hi, device=target_nodes.device)
        if layout != "block_cyclic":
            raise ValueError("layout must be 'contiguous' or 'block_cyclic'")
        local, (off, blk, stride) = shard_targets_block_cyclic(target_nodes, self.rank, self.world, block)
        return local, off, (blk, stride), block_cyclic_indices(n, self.rank, self.world, block, device=target_nodes.device)

    def walk(self, target_nodes, p, q, walk_length, seed, layout="block_cyclic", block=DEFAULT_BLOCK, out=None):
        """Walks this rank's shard of `target_nodes` (the full list, identical on every rank)."""
        local, off, blocks, _ = self.shard(target_nodes, layout, block)
        return self.graph.walk(local, p, q, walk_length, seed, walk_id_offset=off, walk_id_blocks=blocks, out=out)

    def gather(self, local_walks, n_total, layout="block_cyclic", block=DEFAULT_BLOCK):
        """The ranks' shards of one `walk` call back in the caller's order, on every rank."""
        if layout == "contiguous":
            return gather_walks(local_walks, n_total, group=self.group)
        return gather_walks_block_cyclic(local_walks, n_total, block, group=self.group)

    def walk_local_to_host(self, local_targets_host, p, q, walk_length, seed, walk_id_offset, walk_id_blocks=None, out=None):
        """The rank's shard for start nodes and walks in HOST memory (native.PreparedCsr.walk_to_host)."""
        return self.graph.walk_to_host(local_targets_host, p, q, walk_length, seed, walk_id_offset=walk_id_offset,
                                       walk_id_blocks=walk_id_blocks, out=out)

    def walk_local(self, local_targets, p, q, walk_length, seed, walk_id_offset, walk_id_blocks=None, out=None):
        """The same for a shard the caller has already cut (bench.py keeps the shard across calls)."""
        return self.graph.walk(local_targets, p, q, walk_length, seed, walk_id_offset=walk_id_offset,
                               walk_id_blocks=walk_id_blocks, out=out)
