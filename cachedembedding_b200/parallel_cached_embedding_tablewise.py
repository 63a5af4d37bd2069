"""ParallelCachedEmbeddingBagTablewise: table-wise sharded cached embedding bag.

Drop-in for ``colossalai.nn.parallel.layers.ParallelCachedEmbeddingBagTablewise`` as constructed at
/root/reference/recsys/models/dlrm.py:53-68: every rank owns whole tables (concatenated into one local table with
re-based ids), looks up the GLOBAL batch for its tables, and one all-to-all of pooled embeddings gives every rank its
slice of the batch with all features (rank-major feature order).

The bag kernel writes the local result directly in (B, F_loc, D) order, so upstream's
``torch.cat(out.split(B), 1)`` repack before the all-to-all disappears (forward and backward).
"""
from typing import List, Optional

import torch
import torch.distributed as dist

from .cached_embedding import CachedEmbeddingBag
from .collectives import dual_all_to_all_tablewise, split_sizes
from .embedding_config import TablewiseEmbeddingBagConfig
from .evict_strategy import EvictionStrategy
from .parallel_cached_embedding import _world


class ParallelCachedEmbeddingBagTablewise(CachedEmbeddingBag):
    """All tables assigned to this rank are handled by ONE cached bag (one host table, one slot cache)."""

    def __init__(self,
                 embedding_bag_config_list: List[TablewiseEmbeddingBagConfig],
                 embedding_dim: int,
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
                 warmup_ratio=0.7,
                 buffer_size=50_000,
                 pin_weight=False,
                 evict_strategy: EvictionStrategy = EvictionStrategy.LFU,
                 process_group=None,
                 **kwargs):
        self.process_group = process_group
        self.rank, self.world_size = _world(process_group)
        self.rank_of_tables = [config.assigned_rank for config in embedding_bag_config_list]
        self.global_table_num_embeddings_list = [config.num_embeddings for config in embedding_bag_config_list]
        self.global_tables_num = len(embedding_bag_config_list)
        self.global_tables_offsets = torch.cumsum(
            torch.tensor([0] + self.global_table_num_embeddings_list), 0)
        self.assigned_table_list: List[int] = [i for i, r in enumerate(self.rank_of_tables) if r == self.rank]
        assert self.assigned_table_list, f"rank {self.rank} owns no table"
        num_embeddings = sum(self.global_table_num_embeddings_list[i] for i in self.assigned_table_list)

        # ids_freq_mapping of the local tables, concatenated (None if any table has none)
        ids_freq_mapping = []
        for i in self.assigned_table_list:
            m = embedding_bag_config_list[i].ids_freq_mapping
            if m is None:
                ids_freq_mapping = None
                break
            ids_freq_mapping.append(torch.as_tensor(m))
        if ids_freq_mapping is not None:
            ids_freq_mapping = torch.cat(ids_freq_mapping)
        if _weight is None:
            ws = [embedding_bag_config_list[i].initial_weight for i in self.assigned_table_list]
            if all(w is not None for w in ws):
                _weight = torch.cat([w.detach().cpu().float() for w in ws], 0).contiguous()

        # global id -> local row: subtract the rows of the NON-local tables that precede each local table
        self.idx_offset_list = []
        local_prefix = 0
        for i in self.assigned_table_list:
            self.idx_offset_list.append(int(self.global_tables_offsets[i]) - local_prefix)
            local_prefix += self.global_table_num_embeddings_list[i]
        self.embedding_dim_per_rank = [0 for _ in range(self.world_size)]
        for r in self.rank_of_tables:
            self.embedding_dim_per_rank[r] += embedding_dim

        super().__init__(num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                         _weight, mode, include_last_offset, dtype, device, cache_ratio, ids_freq_mapping,
                         warmup_ratio, buffer_size, pin_weight, evict_strategy, **kwargs)
        self.cache_op = True

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor = None, per_sample_weights=None, shape_hook=None,
                already_split_along_rank=True):
        n_local = len(self.assigned_table_list)
        if not already_split_along_rank:
            n_off = offsets.shape[0] - (1 if self.include_last_offset else 0)
            batch_size = n_off // self.global_tables_num
            indices, offsets, per_sample_weights = self.split_along_rank(batch_size, indices, offsets,
                                                                         per_sample_weights)
        else:
            batch_size = offsets.shape[0] // n_local
        if self.cache_op:
            with torch.no_grad():
                indices = self.cache_weight_mgr.prepare_ids(indices)
        # (B, F_loc, D) written by the kernel == torch.cat(out.split(B), 1) of the bag-major result
        if self._use_fused_exchange(batch_size, per_sample_weights):
            output_full = self._forward_fused(indices, offsets, batch_size)
            if shape_hook is not None:
                output_full = shape_hook(output_full)
            return output_full
        local_out = self._embed(indices, offsets, per_sample_weights, layout="sample_major", layout_batch=batch_size)
        local_out = local_out.view(batch_size, n_local * self.embedding_dim)
        scatter_strides = split_sizes(batch_size, self.world_size)
        output_full = dual_all_to_all_tablewise(local_out, self.process_group, scatter_strides,
                                                self.embedding_dim_per_rank)
        if shape_hook is not None:
            output_full = shape_hook(output_full)
        return output_full

    # ---- fused exchange over NVLink peer memory (fused_exchange.py) ----------------------------------------------------
    def enable_fused_exchange(self, flag: bool = True):
        """Fold the pooled-embedding all-to-all (and its backward) into the gather / optimizer kernels.  Needs the fused
        optimizer (`set_fused_optimizer`), mode 'sum', no per-sample weights and more than one rank."""
        self._fused_exchange_on = bool(flag)
        if not flag:
            for exch in getattr(self, "_exchanges", {}).values():
                exch.close()
            self._exchanges = {}
            self._exchange = None

    def _use_fused_exchange(self, batch_size, per_sample_weights) -> bool:
        """The fused exchange hands out a view of a persistent peer buffer that the NEXT forward overwrites, and it is
        the backward's barrier that keeps a fast rank from doing so while a peer still reads: it is used for training
        steps only (grad enabled, module in training mode); evaluation takes the NCCL path, whose outputs are fresh
        tensors."""
        return (getattr(self, "_fused_exchange_on", False) and self.world_size > 1 and per_sample_weights is None
                and self.mode == "sum" and self._fused_optimizer is not None and batch_size >= self.world_size
                and self.training and torch.is_grad_enabled())

    def _exchange_for(self, batch_size):
        from .fused_exchange import FusedExchange
        if not hasattr(self, "_exchanges"):
            self._exchanges = {}
        exch = self._exchanges.get(batch_size)
        if exch is None:
            # one set of peer buffers per batch size, kept until enable_fused_exchange(False): outputs handed out
            # earlier alias them, so a change of batch size must not free them
            total_features = len(self.rank_of_tables)
            feature_offset = sum(1 for r in self.rank_of_tables if r < self.rank)
            exch = FusedExchange(batch_size, total_features, feature_offset, self.embedding_dim, self.process_group)
            self._exchanges[batch_size] = exch
        self._exchange = exch
        return exch

    def _forward_fused(self, slot_ids, offsets, batch_size):
        from .fused_exchange import _FusedTablewiseFunction
        exch = self._exchange_for(batch_size)
        offsets = offsets.to(slot_ids.device)
        if offsets.dtype not in (torch.int32, torch.int64):
            offsets = offsets.long()
        self.cache_weight_mgr.wait_rows()
        out = _FusedTablewiseFunction.apply(self.cache_weight_mgr.cuda_cached_weight, slot_ids.contiguous().view(-1),
                                            offsets.contiguous(), self, exch)
        return out

    def fused_step(self, slot_ids: torch.Tensor, offsets: torch.Tensor, consumer=None) -> torch.Tensor:
        """Forward + fused backward/optimizer of one batch as ONE CUDA-graph launch (fused exchange only).

        The operator step of the reference's isolation harness (benchmark/benchmark_cache.py:58-72: forward, then
        backward of a given gradient) when the gradient of the pooled embeddings is already where the backward reads it
        (`exchange.grad_tensor()`): gather + barrier + barrier + segment-reduce/optimizer are captured once per
        (slot-id buffer, offsets, backward plan) and replayed, which takes the ~0.2 ms of Python / autograd / ctypes per
        step down to one graph launch -- at 8 ranks a step is only ~0.3 ms of GPU work.  Needs cache_op off, slot ids in
        a buffer with a stable address (the look-ahead driver's static ring) and a backward plan made for them
        (`LookaheadPrefetcher.submit(..., offsets=...)`); anything else falls back to the eager forward + backward.
        `consumer(out)`, if given, is captured between the two barriers -- where the dense part of a model reads the
        pooled embeddings and leaves their gradient in `exchange.grad_tensor()`; it must only touch buffers with stable
        addresses.  Returns this rank's (B_rank, F * D) pooled embeddings, a view of the exchange's output buffer: like
        in the eager path its content is only guaranteed until the backward's barrier -- a faster peer's next forward
        may overwrite it afterwards -- so read it inside `consumer`."""
        import ctypes
        from . import _lib
        from .cache_mgr import _stream_ptr
        from .cached_embedding import _bag_args
        n_local = len(self.assigned_table_list)
        batch_size = offsets.shape[0] // n_local
        plan = getattr(self, "_bwd_plans", {}).get((slot_ids.data_ptr(), slot_ids.numel()))
        if self.cache_op or plan is None or not self._use_fused_exchange(batch_size, None) or offsets.device != slot_ids.device:
            out = self(slot_ids, offsets)
            out.backward(self._exchange.grad_tensor() if getattr(self, "_exchange", None) is not None else torch.zeros_like(out))
            return out
        if not hasattr(self, "_step_graphs"):
            self._step_graphs, self.graph_launches = {}, 0
        exch = self._exchange_for(batch_size)
        key = (slot_ids.data_ptr(), slot_ids.numel(), offsets.data_ptr(), plan.workspace.data_ptr(), plan.keys,
               id(consumer))
        entry = self._step_graphs.get(key)
        self.cache_weight_mgr.wait_rows()
        if plan.ready is not None:       # the plan was made on the look-ahead side stream and may still be running
            torch.cuda.current_stream().wait_event(plan.ready)
        if entry is None:
            lib = _lib.load()
            weight = self.cache_weight_mgr.cuda_cached_weight
            fused = self._fused_optimizer
            state = self.cache_weight_mgr.cuda_cached_state
            a = _bag_args(weight, slot_ids, offsets, None, self.include_last_offset, _lib.MODE_SUM, self.padding_idx,
                          _lib.LAYOUT_EXCHANGE, exch.B)
            nbytes = int(lib.cebag_backward_workspace_bytes(ctypes.byref(a)))
            before = _lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                a.exchange = ctypes.pointer(exch._xo)
                _lib.check(lib.cebag_bag_forward(ctypes.byref(a), None, _stream_ptr()))
                exch.barrier()          # every rank's rows have landed
                if consumer is not None:
                    consumer(exch.out_tensor())
                exch.barrier()          # every rank's gradient is in place
                a.exchange = ctypes.pointer(exch._xg)
                _lib.check(lib.cebag_bag_backward_fused(
                    ctypes.byref(a), None, weight.data_ptr(), state.data_ptr() if state is not None else None,
                    fused["kind"], fused["lr"], fused["eps"], plan.workspace.data_ptr(), nbytes, plan.apply(a),
                    _stream_ptr()))
            entry = self._step_graphs[key] = (graph, _lib.launch_count() - before, (slot_ids, offsets, plan, consumer))
        entry[0].replay()
        self.graph_launches += entry[1]
        return exch.out_tensor()

    def split_along_rank(self, batch_size, indices: torch.Tensor, offsets: torch.Tensor = None,
                         per_sample_weights=None):
        """Cut this rank's tables out of a global KJT (values = global ids, feature-major offsets)."""
        return split_kjt_along_rank(self.assigned_table_list, self.idx_offset_list, self.include_last_offset,
                                    batch_size, indices, offsets, per_sample_weights)

    def set_cache_op(self, cache_op: bool = True):
        self.cache_op = cache_op


def split_kjt_along_rank(assigned_table_list, idx_offset_list, include_last_offset, batch_size, indices, offsets,
                         per_sample_weights=None):
    """Host logic of upstream's split_along_rank (A.6): slice the id ranges of the local tables out of a global KJT
    (feature-major: bag g = table * batch_size + sample), re-base the ids by `idx_offset_list` and rebuild the local
    offsets.  Table boundaries are read back once (upstream does one .item() per table)."""
    li, lo, lw = [], [], []
    pre_end = 0
    bounds = offsets[torch.arange(0, offsets.shape[0], batch_size, device=offsets.device)].tolist()
    for k, t in enumerate(assigned_table_list):
        start = bounds[t]
        if (not include_last_offset) and batch_size * (t + 1) >= offsets.shape[0]:
            end = indices.shape[0]
        else:
            end = bounds[t + 1]
        li.append(indices.narrow(0, start, end - start) - idx_offset_list[k])
        if per_sample_weights is not None:
            lw.append(per_sample_weights.narrow(0, start, end - start))
        last = (k + 1 == len(assigned_table_list))
        take = batch_size + 1 if (last and include_last_offset) else batch_size
        lo.append(offsets.narrow(0, batch_size * t, take) + (pre_end - start))
        pre_end += end - start
    return (torch.cat(li), torch.cat(lo), torch.cat(lw) if per_sample_weights is not None else None)
