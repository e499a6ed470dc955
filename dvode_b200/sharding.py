"""Multi-GPU sharding of a batch: contiguous index blocks, no data-path collective.

The systems of a batch are independent: rank r of G owns systems
[r*N/G, (r+1)*N/G) and integrates them with its own handle on its own GPU.  The only exchange
is the final gather of solutions and statistics to rank 0.

On GPUs that gather is the C ABI's `bvode_gather_device` (csrc/bvode_multi.cu): one pack kernel per
rank and one NCCL group of sends / receives straight into the root's arrays.  `make_comm` builds
its communicator; torch.distributed is only the plumbing that carries the 128-byte NCCL id from
rank 0 to the other ranks.  `gather_to_rank0` is the host-logic twin over torch.distributed
point-to-point operations (gloo in the CPU tests): the same block arithmetic, receives in place.
"""


def shard_range(nsys_total, rank, world):
    """[first, last) of the contiguous block owned by `rank`; blocks differ by at most one
    (the same arithmetic as bvode_shard_range in the C ABI)."""
    base, rem = divmod(int(nsys_total), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def make_comm(dist, rank, world, device):
    """A dvode_b200.Comm (NCCL communicator behind the C ABI) for this process group: rank 0
    creates the unique id, torch.distributed broadcasts it."""
    from . import api
    box = [api.Comm.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return api.Comm(world, rank, device, box[0])


def gather_to_rank0(local, dist, rank, world, nsys_total=None):
    """Gather per-rank arrays (first dimension = this rank's block of systems, sizes given by
    shard_range) to rank 0 in system order.  `local` is a torch tensor; returns the
    [nsys_total, ...] tensor on rank 0 and None elsewhere.  No size exchange (the blocks follow from
    nsys_total), no padding, no concatenation: rank 0 receives every block in its place."""
    import torch
    if world == 1:
        return local
    if nsys_total is None:  # equal blocks
        nsys_total = local.shape[0] * world
    first, last = shard_range(nsys_total, rank, world)
    assert last - first == local.shape[0], "local block does not match shard_range(nsys_total)"
    local = local.contiguous()
    if rank != 0:
        if local.shape[0] > 0:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, 0)]):
                req.wait()
        return None
    out = torch.empty((nsys_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    out[first:last].copy_(local)
    ops = []
    for r in range(1, world):
        f, l = shard_range(nsys_total, r, world)
        if l > f:
            ops.append(dist.P2POp(dist.irecv, out[f:l], r))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out
