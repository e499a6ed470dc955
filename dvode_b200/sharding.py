"""Multi-GPU sharding of a batch: contiguous index blocks, no data-path collective.

The systems of a batch are independent: rank r of G owns systems
[r*N/G, (r+1)*N/G) and integrates them with its own handle on its own GPU.  The only exchange
is the final gather of solutions and statistics to rank 0 (torch.distributed: NCCL on GPUs,
gloo in the CPU tests).
"""
import numpy as np


def shard_range(nsys_total, rank, world):
    """[first, last) of the contiguous block owned by `rank`; blocks differ by at most one."""
    base, rem = divmod(int(nsys_total), int(world))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def gather_to_rank0(local, dist, rank, world, device=None):
    """Gather per-rank arrays (first dimension = this rank's systems, possibly ragged by one)
    to rank 0, concatenated in system order.  `local` is a torch tensor; returns a tensor on
    rank 0 and None elsewhere."""
    import torch
    if world == 1:
        return local
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, torch.zeros((nmax - local.shape[0],) + tuple(local.shape[1:]),
                                            dtype=local.dtype, device=local.device)])
    pad = pad.contiguous()
    outs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, outs, dst=0)
    if rank != 0:
        return None
    return torch.cat([o[:c] for o, c in zip(outs, counts)])
