"""Sparse-input exchange of the distributed data loader: every rank ends up with the KeyedJaggedTensor of the GLOBAL
batch (all-gather of the ranks' jagged inputs, key-major).

Replaces /root/reference/recsys/datasets/utils.py:8-54 (`KJTAllToAll.all_to_all`), which
runs TWO blocking collectives (lengths at :31, values at :41) and synchronises the host once per rank and once per
(rank, key) in between (`.item()` at :36, `.cpu().tolist()` at :34).  Here:

  * ONE collective: every rank contributes one packed int64 buffer [count | lengths (F * B_loc) | values, padded to
    `capacity`]; buffers have the same size on every rank, so it is a plain all-gather;
  * the merge (values of key f = rank 0's values of f, then rank 1's, ...) is computed on the device from the gathered
    lengths with cumulative sums and ONE gather -- no per-rank or per-key host round trip;
  * `capacity` (ids per rank per batch) is the only size the host has to know.  With pooling factor 1 (Criteo: one id
    per feature per sample, recsys/datasets/criteo.py:129-130) it is F * B_loc and the exchange never waits for the GPU;
    for ragged inputs the total is read back once per call (one `.item()`), and a rank that exceeds `capacity` is an
    error raised from that read-back.

`all_to_all(kjt)` accepts any object with keys() / values() / lengths() / stride() (torchrec's KeyedJaggedTensor or the
shim's) and returns (keys, values, lengths, stride) -- or, when `kjt_factory` is given (e.g.
`KeyedJaggedTensor.from_lengths_sync`), whatever the factory builds from keys / values / lengths.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def merge_gathered(packed: torch.Tensor, num_keys: int, local_batch: int, capacity: int, total: Optional[int] = None):
    """packed: int64 [W, 1 + F * B_loc + capacity] as gathered.  Returns (values, lengths) of the global batch, key-major:
    lengths[f, r * B_loc + b] = rank r's length of (f, b); values of key f are the ranks' values of f in rank order.
    `total` = number of values over all ranks if the caller knows it (no host synchronisation then)."""
    W = packed.shape[0]
    F, B = num_keys, local_batch
    counts = packed[:, 0]                                              # [W]
    lengths = packed[:, 1:1 + F * B].view(W, F, B)                     # rank-major as gathered
    values = packed[:, 1 + F * B:]                                     # [W, capacity]
    per_rank_key = lengths.sum(2)                                      # [W, F]
    # source start of (r, f) inside rank r's values; destination start of (f, r) inside the merged values
    src_start = torch.cumsum(per_rank_key, 1) - per_rank_key           # [W, F]
    by_key = per_rank_key.t().contiguous()                             # [F, W]
    dst_start = (torch.cumsum(by_key.view(-1), 0) - by_key.view(-1)).view(F, W)
    if total is None:
        total = int(counts.sum().item())                               # the one host read-back of the ragged path
        if int(counts.max().item()) > capacity:
            raise RuntimeError(f"a rank sent {int(counts.max())} ids, more than the exchange capacity {capacity}")
    # one gather: for every output position, which (rank, position) it comes from
    seg_len = by_key.view(-1)                                          # segments in output order (f major, then r)
    seg_id = torch.repeat_interleave(torch.arange(F * W, device=packed.device), seg_len, output_size=total)
    within = torch.arange(total, device=packed.device) - dst_start.view(-1)[seg_id]
    f_of, r_of = seg_id // W, seg_id % W
    src = src_start[r_of, f_of] + within
    merged = values[r_of, src]
    merged_lengths = lengths.permute(1, 0, 2).reshape(-1)              # [F, W, B] -> key-major, rank-major, sample
    return merged, merged_lengths


class FusedKJTAllToAll:
    """Drop-in for the reference's KJTAllToAll (same constructor, same `all_to_all` entry point)."""

    def __init__(self, group=None, capacity: Optional[int] = None, fixed_lengths: bool = False,
                 kjt_factory: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.capacity = capacity
        self.fixed_lengths = fixed_lengths          # every rank sends exactly `capacity` ids: nothing is read back
        self.kjt_factory = kjt_factory

    @torch.no_grad()
    def all_to_all(self, kjt):
        if self.world_size == 1:
            return kjt
        values, lengths = kjt.values(), kjt.lengths()
        keys, batch = list(kjt.keys()), int(kjt.stride())
        F = len(keys)
        n = values.numel()
        cap = self.capacity if self.capacity is not None else n
        if self.fixed_lengths and n != cap:
            raise ValueError(f"fixed_lengths exchange expects {cap} ids per rank, got {n}")
        if n > cap:
            raise ValueError(f"{n} ids exceed the exchange capacity {cap}")
        packed = torch.zeros(1 + F * batch + cap, dtype=torch.int64, device=values.device)
        packed[0] = n
        packed[1:1 + F * batch] = lengths.view(-1)
        packed[1 + F * batch:1 + F * batch + n] = values
        gathered = [torch.empty_like(packed) for _ in range(self.world_size)]
        dist.all_gather(gathered, packed, group=self.group)            # the one collective
        total = cap * self.world_size if self.fixed_lengths else None
        merged, merged_lengths = merge_gathered(torch.stack(gathered), F, batch, cap, total)
        merged = merged.to(values.dtype)
        merged_lengths = merged_lengths.to(lengths.dtype)
        if self.kjt_factory is not None:
            return self.kjt_factory(keys=keys, values=merged, lengths=merged_lengths)
        return keys, merged, merged_lengths, batch * self.world_size
