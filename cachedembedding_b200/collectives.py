"""The pooled-embedding exchanges of the parallel bags (upstream colossalai/nn/_ops/_utils.py: dual_all_to_all,
dual_all_to_all_tablewise).  One all-to-all forward, the mirror all-to-all backward.

NCCL over NVLink is used when the group's backend is nccl; for gloo (CPU tests of the host logic) the same exchange
is done with point-to-point sends, since gloo has no list all_to_all.
"""
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def split_sizes(total: int, world_size: int) -> List[int]:
    """torch.tensor_split sizes: the first `total % W` parts get one extra."""
    return [total // world_size + int(i < total % world_size) for i in range(world_size)]


def get_partition(embedding_dim: int, rank: int, world_size: int):
    """(start, end, divides_evenly): the columns of `rank` when D columns are dealt out by the torch.tensor_split rule
    (what the reference's recsys/utils/misc.py:138-154 computes for the column-wise bag)."""
    if world_size > 1 and embedding_dim < world_size:
        raise AssertionError(f"cannot split {embedding_dim} embedding columns over {world_size} ranks")
    widths = split_sizes(embedding_dim, world_size)
    start = sum(widths[:rank])
    return start, start + widths[rank], embedding_dim % world_size == 0


def exchange(recv: List[torch.Tensor], send: List[torch.Tensor], group=None):
    """recv[r] <- what rank r sends to me; send[r] -> rank r."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all(recv, send, group=group)
        return
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    recv[rank].copy_(send[rank])
    ops = []
    for r in range(world):
        if r == rank:
            continue
        peer = dist.get_global_rank(group, r) if group is not None else r
        ops.append(dist.P2POp(dist.isend, send[r], peer, group))
        ops.append(dist.P2POp(dist.irecv, recv[r], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _all_to_all(x: torch.Tensor, group, scatter_dim: int, gather_dim: int,
                gather_sizes: Optional[Sequence[int]] = None) -> torch.Tensor:
    world = dist.get_world_size(group)
    if world == 1:
        return x
    rank = dist.get_rank(group)
    scatter_dim = scatter_dim % x.dim()
    gather_dim = gather_dim % x.dim()
    send = [c.contiguous() for c in torch.tensor_split(x, world, scatter_dim)]
    recv = []
    for r in range(world):
        shape = list(send[rank].shape)
        if gather_sizes is not None:
            shape[gather_dim] = gather_sizes[r]
        recv.append(torch.empty(shape, dtype=x.dtype, device=x.device))
    exchange(recv, send, group)
    return torch.cat(recv, dim=gather_dim).contiguous()


class _DualAllToAll(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, group, scatter_dim, gather_dim, fwd_gather_sizes, bwd_gather_sizes):
        ctx.group, ctx.scatter_dim, ctx.gather_dim = group, scatter_dim, gather_dim
        ctx.bwd_gather_sizes = bwd_gather_sizes
        return _all_to_all(x, group, scatter_dim, gather_dim, fwd_gather_sizes)

    @staticmethod
    def backward(ctx, grad):
        return (_all_to_all(grad.contiguous(), ctx.group, ctx.gather_dim, ctx.scatter_dim, ctx.bwd_gather_sizes),
                None, None, None, None, None)


def dual_all_to_all(x, group, scatter_dim: int, gather_dim: int, fwd_gather_sizes=None, bwd_gather_sizes=None):
    """Scatter `x` along scatter_dim, gather along gather_dim; backward is the same exchange with the dims swapped.

    *_gather_sizes: size along the gather dim of the piece each rank contributes, when they differ (column split of a
    D that W does not divide; batch split of a B that W does not divide)."""
    return _DualAllToAll.apply(x, group, scatter_dim, gather_dim, fwd_gather_sizes, bwd_gather_sizes)


class _DualAllToAllTablewise(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, group, scatter_strides, dim_per_rank):
        ctx.group, ctx.scatter_strides, ctx.dim_per_rank = group, list(scatter_strides), list(dim_per_rank)
        world = dist.get_world_size(group)
        if world == 1:
            return x
        rank = dist.get_rank(group)
        send = [c.contiguous() for c in x.split(ctx.scatter_strides, 0)]
        recv = [torch.empty(ctx.scatter_strides[rank], ctx.dim_per_rank[r], dtype=x.dtype, device=x.device)
                for r in range(world)]
        exchange(recv, send, group)
        return torch.cat(recv, 1).contiguous()

    @staticmethod
    def backward(ctx, grad):
        group = ctx.group
        world = dist.get_world_size(group)
        if world == 1:
            return grad, None, None, None
        rank = dist.get_rank(group)
        send = [c.contiguous() for c in grad.split(ctx.dim_per_rank, 1)]
        recv = [torch.empty(ctx.scatter_strides[r], ctx.dim_per_rank[rank], dtype=grad.dtype, device=grad.device)
                for r in range(world)]
        exchange(recv, send, group)
        return torch.cat(recv, 0).contiguous(), None, None, None


def dual_all_to_all_tablewise(x, group, scatter_strides, dim_per_rank):
    """(B, F_loc*D) on every rank -> (B_rank, sum_r F_r*D), features in rank-major order (A.6)."""
    return _DualAllToAllTablewise.apply(x, group, scatter_strides, dim_per_rank)
