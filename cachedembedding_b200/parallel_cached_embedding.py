"""ParallelCachedEmbeddingBag: column-wise sharded cached embedding bag.

Drop-in for ``colossalai.nn.parallel.layers.ParallelCachedEmbeddingBag`` as constructed at
/root/reference/recsys/models/dlrm.py:70-81: rank r holds columns ``get_partition(D, r, W)`` of every row, runs the
cache + bag kernels on ALL ids of the global batch, then one all-to-all turns (B, F, D/W) into (B/W, F, D).
"""
from typing import Optional

import torch
import torch.distributed as dist

from .cached_embedding import CachedEmbeddingBag
from .collectives import dual_all_to_all, get_partition, split_sizes
from .evict_strategy import EvictionStrategy


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class ParallelCachedEmbeddingBag(CachedEmbeddingBag):

    def __init__(self,
                 num_embeddings,
                 embedding_dim,
                 padding_idx=None,
                 max_norm=None,
                 norm_type=2.,
                 scale_grad_by_freq=False,
                 sparse=False,
                 _weight=None,
                 mode='mean',
                 include_last_offset=False,
                 dtype=None,
                 device=None,
                 cache_ratio=0.01,
                 ids_freq_mapping=None,
                 warmup_ratio=0.7,
                 buffer_size=50_000,
                 pin_weight=False,
                 evict_strategy: EvictionStrategy = EvictionStrategy.DATASET,
                 process_group=None,
                 **kwargs):
        self.process_group = process_group
        self.rank, self.world_size = _world(process_group)
        self.partition_start_index, self.partition_end_index, divisible = get_partition(
            embedding_dim, self.rank, self.world_size)
        self.embedding_dim_per_partition = self.partition_end_index - self.partition_start_index
        self.full_embedding_dim = embedding_dim
        super().__init__(num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                         _weight, mode, include_last_offset, dtype, device, cache_ratio, ids_freq_mapping,
                         warmup_ratio, buffer_size, pin_weight, evict_strategy, **kwargs)
        self.cache_op = True

    def _weight_alloc(self, dtype, device):
        N, Dp = self.num_embeddings, self.embedding_dim_per_partition
        full = self.embedding_dim
        self.embedding_dim = Dp
        try:
            w = super()._weight_alloc(dtype, device)
        finally:
            self.embedding_dim = full
        return w

    def forward(self, indices, offsets=None, per_sample_weights=None, shape_hook=None, scatter_dim=0, gather_dim=-1):
        if self.cache_op:
            with torch.no_grad():
                shape = indices.shape
                indices = self.cache_weight_mgr.prepare_ids(indices).view(shape)
        output_shard = self._embed(indices, offsets, per_sample_weights)
        if shape_hook is not None:
            output_shard = shape_hook(output_shard)
        if self.world_size == 1:
            return output_shard
        W = self.world_size
        col_sizes = [get_partition(self.full_embedding_dim, r, W)[1] - get_partition(self.full_embedding_dim, r, W)[0]
                     for r in range(W)]
        row_sizes = split_sizes(output_shard.shape[scatter_dim], W)
        return dual_all_to_all(output_shard, self.process_group, scatter_dim, gather_dim,
                               fwd_gather_sizes=col_sizes, bwd_gather_sizes=row_sizes)

    @classmethod
    def from_pretrained(cls, embedding: torch.Tensor, freeze: bool = True, padding_idx: Optional[int] = None,
                        max_norm: Optional[float] = None, norm_type: float = 2., scale_grad_by_freq: bool = False,
                        sparse: bool = False, mode: str = 'mean', include_last_offset: bool = False,
                        cuda_row_num: int = 100_000, ids_freq_mapping=None, warmup_ratio: float = 0.7,
                        buffer_size: int = 0, **kwargs) -> 'ParallelCachedEmbeddingBag':
        """`embedding` is this rank's column shard [N, D_r] (as upstream); the full D is its width times... the sum
        over ranks, so pass full_dim=... when W does not divide D evenly."""
        rows, cols = embedding.shape
        rank, world = _world(kwargs.get("process_group"))
        full_dim = kwargs.pop("full_dim", cols * world)
        bag = cls(rows, full_dim, padding_idx=padding_idx, max_norm=max_norm, norm_type=norm_type,
                  scale_grad_by_freq=scale_grad_by_freq, sparse=sparse, _weight=embedding, mode=mode,
                  include_last_offset=include_last_offset, cache_ratio=min(cuda_row_num / rows, 1.0),
                  ids_freq_mapping=ids_freq_mapping, warmup_ratio=warmup_ratio, buffer_size=buffer_size,
                  cuda_row_num=cuda_row_num, **kwargs)
        bag.cache_weight_mgr.cuda_cached_weight.requires_grad_(not freeze)
        return bag
