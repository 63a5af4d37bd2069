"""CPU restatement of the reference's cached embedding-bag path (test oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  PARITY UNPINNED at the
reference boundary (the reference holds no tests for this path).

What is restated (the implementation itself is the
third-party ColossalAI package named in /root/reference/README.md:37, absent
from /root/reference):

  * ``CachedParamMgr``           -> :class:`OracleCachedParamMgr`        (A.1, A.3, A.4, A.7)
  * ``CachedEmbeddingBag``       -> :class:`OracleCachedEmbeddingBag`    (A.2)
  * ``ParallelCachedEmbeddingBagTablewise`` (+ config)
                                 -> :class:`OracleTablewiseWorld`        (A.6)
  * ``ParallelCachedEmbeddingBag`` (column-wise)
                                 -> :class:`OracleColumnwiseWorld`       (A.5)

Reference call sites the behaviour is anchored on:
  recsys/dlrm_main.py:259 (prepare_ids on the concatenated look-ahead window),
  recsys/models/dlrm.py:58-81,99-110 (constructors, forward + shape hook),
  recsys/dlrm_main.py:455-461,279 (torch.optim.SGD on the cached rows),
  benchmark/benchmark_cache.py:39-72 (bare operator fwd + backward(grad)).

Every tensor here lives on the CPU and every op is a plain ATen CPU op
(index_select / unique / isin / sort / nonzero / index_copy_ / F.embedding_bag /
sparse SGD) -- the same op list the reference issues on CUDA -- which is also
why this file doubles as the "reference path on host cores" in bench.py.

One refinement of the reference: ``torch.topk`` does
not define how ties are broken, so LFU victims are taken in the canonical
order ``(freq ascending, slot ascending)``.  Any answer torch.topk may give is a
superset-equivalent of this rule; the upstream KATs are tie-free at the decision
points.  DATASET victims have unique keys (cpu_row_idx), so no rule is needed.
"""
from __future__ import annotations

import enum
import math
import sys
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class EvictionStrategy(enum.Enum):
    LFU = 1
    DATASET = 2


_MAXSIZE = sys.maxsize  # empty-slot LFU counter (A.1)


class OracleCachedParamMgr(nn.Module):
    """A.1/A.3/A.4: host table + slot cache + id maps, all on CPU tensors."""

    def __init__(self,
                 weight: torch.Tensor,
                 cuda_row_num: int = 0,
                 buffer_size: int = 0,
                 pin_weight: bool = True,
                 evict_strategy: EvictionStrategy = EvictionStrategy.DATASET,
                 async_copy: bool = False):
        super().__init__()
        assert weight.dim() == 2
        self.buffer_size = buffer_size
        self.num_embeddings, self.embedding_dim = weight.shape
        self.cuda_row_num = cuda_row_num
        self._cuda_available_row_num = cuda_row_num
        self.pin_weight = pin_weight
        self.elem_size_in_byte = weight.element_size()
        self._evict_strategy = evict_strategy
        self._async_copy = async_copy
        self.weight = weight  # host table, updated in place on evict / flush
        if cuda_row_num == 0:
            raise NotImplementedError("cuda_row_num == 0")
        self.cuda_cached_weight = nn.Parameter(
            torch.zeros(cuda_row_num, self.embedding_dim, dtype=weight.dtype))
        self.register_buffer("idx_map", torch.arange(self.num_embeddings, dtype=torch.long), persistent=False)
        self.register_buffer("cached_idx_map", torch.full((cuda_row_num,), -1, dtype=torch.long), persistent=False)
        self.register_buffer("inverted_cached_idx", torch.full((self.num_embeddings,), -1, dtype=torch.long),
                             persistent=False)
        if evict_strategy == EvictionStrategy.LFU:
            self.register_buffer("freq_cnter", torch.full((cuda_row_num,), _MAXSIZE, dtype=torch.long),
                                 persistent=False)
        self.evict_backlist = torch.tensor([], dtype=torch.long)
        # look-ahead extension (not in the reference): rows of the previous prepare_ids call stay protected too
        self.protect_windows = 1
        self._prev_rows = torch.tensor([], dtype=torch.long)
        self.num_hits_history: List[int] = []
        self.num_miss_history: List[int] = []
        self.num_write_back_history: List[int] = []
        self._cpu_to_cuda_numel = 0
        self._cuda_to_cpu_numel = 0
        self._cache_miss = 0
        self._total_cache = 0

    # -- small accessors ----------------------------------------------------------------------
    @property
    def cuda_available_row_num(self):
        return self._cuda_available_row_num

    def cpu_weight_data(self, row_idx: int) -> torch.Tensor:
        return self.weight.data.view(-1).narrow(0, int(row_idx) * self.embedding_dim,
                                                self.embedding_dim).view(1, self.embedding_dim)

    # -- A.1 reorder ----------------------------------------------------------------------------
    @torch.no_grad()
    def reorder(self, ids_freq_mapping=None, warmup_ratio: float = 0.7):
        if ids_freq_mapping is not None and not isinstance(ids_freq_mapping, torch.Tensor):
            ids_freq_mapping = torch.tensor(ids_freq_mapping)
        if ids_freq_mapping is not None and self._evict_strategy == EvictionStrategy.DATASET:
            # id -> rank by descending frequency; stable so that equal counts give one answer
            tmp_idx = torch.argsort(ids_freq_mapping.to(torch.long), descending=True, stable=True)
            sorted_idx = torch.argsort(tmp_idx, stable=True)
            self.idx_map.data.copy_(sorted_idx)
        preload_row_num = min(int(np.ceil(self.cuda_row_num * warmup_ratio)), self.num_embeddings)
        if preload_row_num > 0:
            if self._evict_strategy == EvictionStrategy.LFU and ids_freq_mapping is not None:
                f = ids_freq_mapping.to(torch.long)
                # top-k by frequency, canonical order (freq desc, row asc)
                order = torch.argsort(f, descending=True, stable=True)[:preload_row_num]
                preload_cpu_ids = order
                freq_value = f[order]
            else:
                preload_cpu_ids = torch.arange(preload_row_num)
                freq_value = None
            preload_cuda_row_idxs = torch.arange(preload_row_num)
            rows = self.weight.view(self.num_embeddings, -1).index_select(0, preload_cpu_ids)
            self.cuda_cached_weight.data.view(self.cuda_row_num, -1).index_copy_(0, preload_cuda_row_idxs, rows)
            self.cached_idx_map.index_copy_(0, preload_cuda_row_idxs, preload_cpu_ids)
            self.inverted_cached_idx.index_copy_(0, preload_cpu_ids, preload_cuda_row_idxs)
            self._cuda_available_row_num -= preload_row_num
            if self._evict_strategy == EvictionStrategy.LFU:
                if freq_value is None:
                    self.freq_cnter.index_fill_(0, preload_cuda_row_idxs, 0)
                else:
                    self.freq_cnter[preload_cuda_row_idxs] = freq_value

    # -- A.1 flush ------------------------------------------------------------------------------
    @torch.no_grad()
    def flush(self):
        slots = torch.nonzero(self.cached_idx_map > -1).squeeze(1)
        row_ids = self.cached_idx_map[slots]
        rows = self.cuda_cached_weight.data.view(self.cuda_row_num, -1).index_select(0, slots)
        self.weight.view(self.num_embeddings, -1).index_copy_(0, row_ids, rows)
        self.cached_idx_map.index_fill_(0, slots, -1)
        self.inverted_cached_idx.index_fill_(0, row_ids, -1)
        self._cuda_available_row_num += slots.numel()
        if self._evict_strategy == EvictionStrategy.LFU:
            self.freq_cnter.fill_(_MAXSIZE)
        self._prev_rows = torch.tensor([], dtype=torch.long)
        assert self._cuda_available_row_num == self.cuda_row_num
        assert torch.all(self.inverted_cached_idx == -1).item()
        assert torch.all(self.cached_idx_map == -1).item()

    # -- A.3 prepare_ids --------------------------------------------------------------------------
    def _id_to_cached_cuda_id(self, ids: torch.Tensor) -> torch.Tensor:
        ids = self.idx_map.index_select(0, ids.view(-1))
        return self.inverted_cached_idx.index_select(0, ids)

    @torch.no_grad()
    def prepare_ids(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(torch.long)
        cpu_row_idxs_original = self.idx_map.index_select(0, ids.view(-1))
        cpu_row_idxs, repeat_times = torch.unique(cpu_row_idxs_original, return_counts=True)
        assert len(cpu_row_idxs) <= self.cuda_row_num, \
            f"You move {len(cpu_row_idxs)} embedding rows from CPU to CUDA. " \
            f"It is larger than the capacity of the cache, which at most contains {self.cuda_row_num} rows, " \
            f"Please increase cuda_row_num or decrease the training batch size."
        self.evict_backlist = cpu_row_idxs if self.protect_windows == 1 else torch.cat([cpu_row_idxs, self._prev_rows])
        miss_mask = torch.isin(cpu_row_idxs, self.cached_idx_map, invert=True)
        comm_cpu_row_idxs = cpu_row_idxs[miss_mask]
        self._cache_miss += int(repeat_times[miss_mask].sum().item())
        self._total_cache += ids.numel()
        self.num_hits_history.append(len(cpu_row_idxs) - len(comm_cpu_row_idxs))
        self.num_miss_history.append(len(comm_cpu_row_idxs))
        self.num_write_back_history.append(0)
        self._prepare_rows_on_cuda(comm_cpu_row_idxs)
        self.evict_backlist = torch.tensor([], dtype=cpu_row_idxs.dtype)
        self._prev_rows = cpu_row_idxs
        gpu_row_idxs = self._id_to_cached_cuda_id(ids)
        if self._evict_strategy == EvictionStrategy.LFU:
            unique_gpu_row_idxs = self.inverted_cached_idx[cpu_row_idxs]
            self.freq_cnter.scatter_add_(0, unique_gpu_row_idxs, repeat_times)
        return gpu_row_idxs

    # -- A.7 bounded-staging copier (same bytes, chunked) -------------------------------------------
    def _chunked_index_copy(self, src_index, tgt_index, src, tgt):
        size = self.buffer_size
        for begin in range(0, len(src_index), size):
            piece = src.index_select(0, src_index[begin:begin + size])
            tgt.index_copy_(0, tgt_index[begin:begin + size], piece)

    # -- A.4 _prepare_rows_on_cuda --------------------------------------------------------------------
    @torch.no_grad()
    def _prepare_rows_on_cuda(self, cpu_row_idxs: torch.Tensor) -> None:
        cpu_row_idxs = cpu_row_idxs.to(torch.long)
        M = cpu_row_idxs.numel()
        evict_num = M - self._cuda_available_row_num
        cache_w = self.cuda_cached_weight.data.view(self.cuda_row_num, -1)
        host_w = self.weight.view(self.num_embeddings, -1)
        if evict_num > 0:
            mask_cpu_row_idx = torch.isin(self.cached_idx_map, self.evict_backlist)
            invalid_idxs = torch.nonzero(mask_cpu_row_idx).squeeze(1)
            evictable = int((self.cached_idx_map >= 0).sum()) - invalid_idxs.numel()
            assert evict_num <= evictable, \
                f"only {evictable} cached rows may be evicted but {evict_num} are needed: " \
                f"Please increase cuda_row_num or decrease the training batch size."

            if self._evict_strategy == EvictionStrategy.DATASET:
                backup_idxs = self.cached_idx_map[mask_cpu_row_idx].clone()
                self.cached_idx_map.index_fill_(0, invalid_idxs, -2)
                # keys are unique: no tie rule needed
                evict_gpu_row_idxs = torch.topk(self.cached_idx_map, evict_num, largest=True).indices
                self.cached_idx_map.index_copy_(0, invalid_idxs, backup_idxs)
            else:
                backup_freqs = self.freq_cnter[invalid_idxs].clone()
                self.freq_cnter.index_fill_(0, invalid_idxs, _MAXSIZE)
                # H1: canonical (freq asc, slot asc) == first E of a stable ascending sort
                evict_gpu_row_idxs = torch.sort(self.freq_cnter, stable=True).indices[:evict_num]
                self.freq_cnter.index_copy_(0, invalid_idxs, backup_freqs)
            evict_info = self.cached_idx_map[evict_gpu_row_idxs]
            assert torch.all(evict_info >= 0).item(), "evicted an empty or protected slot"
            if self.buffer_size > 0:
                self._chunked_index_copy(evict_gpu_row_idxs, evict_info, cache_w, host_w)
            else:
                host_w.index_copy_(0, evict_info, cache_w.index_select(0, evict_gpu_row_idxs))
            self.cached_idx_map.index_fill_(0, evict_gpu_row_idxs, -1)
            self.inverted_cached_idx.index_fill_(0, evict_info, -1)
            if self._evict_strategy == EvictionStrategy.LFU:
                self.freq_cnter.index_fill_(0, evict_gpu_row_idxs, _MAXSIZE)
            self._cuda_available_row_num += evict_num
            self._cuda_to_cpu_numel += evict_num * self.embedding_dim
            if self.num_write_back_history:
                self.num_write_back_history[-1] += int(evict_num)
        slots = torch.nonzero(self.cached_idx_map == -1).squeeze(1)[:M]
        if self.buffer_size > 0:
            self._chunked_index_copy(cpu_row_idxs, slots, host_w, cache_w)
        else:
            cache_w.index_copy_(0, slots, host_w.index_select(0, cpu_row_idxs))
        self.cached_idx_map[slots] = cpu_row_idxs
        self.inverted_cached_idx.index_copy_(0, cpu_row_idxs, slots)
        if self._evict_strategy == EvictionStrategy.LFU:
            self.freq_cnter.index_fill_(0, slots, 0)
        self._cuda_available_row_num -= M
        self._cpu_to_cuda_numel += M * self.embedding_dim

    # -- legacy single-row helpers used by upstream test_cachemgr (B.1) --------------------------------
    def _row_in_cuda(self, row_id: int) -> bool:
        return bool(self.inverted_cached_idx[row_id] != -1)

    def _find_free_cuda_row(self) -> int:
        if self._cuda_available_row_num == 0:
            return -1
        return int(torch.nonzero(self.cached_idx_map == -1).squeeze(1)[0].item())

    @torch.no_grad()
    def _evict(self) -> int:
        mask = torch.logical_or(torch.isin(self.cached_idx_map, self.evict_backlist), self.cached_idx_map == -1)
        masked = self.cached_idx_map.clone()
        masked[mask] = -1
        max_row, slot = torch.max(masked, dim=0)
        if max_row.item() == -1:
            raise RuntimeError("Can not evict a row")
        row, slot = int(max_row.item()), int(slot.item())
        self.cpu_weight_data(row).copy_(self.cuda_cached_weight.data[slot].view(1, -1))
        self.cached_idx_map[slot] = -1
        self.inverted_cached_idx[row] = -1
        if self._evict_strategy == EvictionStrategy.LFU:
            self.freq_cnter[slot] = _MAXSIZE
        self._cuda_available_row_num += 1
        self._cuda_to_cpu_numel += self.embedding_dim
        return slot

    @torch.no_grad()
    def _admit(self, row_id: int):
        slot = self._find_free_cuda_row()
        if slot == -1:
            slot = self._evict()
        self.cuda_cached_weight.data[slot].copy_(self.cpu_weight_data(row_id).view(-1))
        self.cached_idx_map[slot] = row_id
        self.inverted_cached_idx[row_id] = slot
        if self._evict_strategy == EvictionStrategy.LFU:
            self.freq_cnter[slot] = 0
        self._cuda_available_row_num -= 1
        self._cpu_to_cuda_numel += self.embedding_dim


class OracleCachedEmbeddingBag(nn.Module):
    """A.2: the nn.Module surface, every op on CPU."""

    def __init__(self,
                 num_embeddings: int,
                 embedding_dim: int,
                 padding_idx: Optional[int] = None,
                 max_norm: Optional[float] = None,
                 norm_type: float = 2.,
                 scale_grad_by_freq: bool = False,
                 sparse: bool = False,
                 _weight: Optional[torch.Tensor] = None,
                 mode: str = 'mean',
                 include_last_offset: bool = False,
                 dtype=None,
                 device=None,
                 cache_ratio: float = 0.01,
                 ids_freq_mapping=None,
                 warmup_ratio: float = 0.7,
                 buffer_size: int = 0,
                 pin_weight: bool = False,
                 evict_strategy: EvictionStrategy = EvictionStrategy.LFU,
                 cuda_row_num: Optional[int] = None):
        super().__init__()
        assert cache_ratio <= 1.0, f"cache ratio {cache_ratio} must less than 1.0"
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        if padding_idx is not None and padding_idx < 0:
            padding_idx = num_embeddings + padding_idx
        self.padding_idx = padding_idx
        self.max_norm, self.norm_type = max_norm, norm_type
        self.scale_grad_by_freq = scale_grad_by_freq
        self.sparse = sparse
        self.mode = mode
        self.include_last_offset = include_last_offset
        self.evict_strategy = evict_strategy
        self.cache_ratio = cache_ratio
        if _weight is None:
            _weight = torch.empty(num_embeddings, embedding_dim, dtype=dtype or torch.float32).uniform_(
                -1 / num_embeddings, 1 / num_embeddings)
            if padding_idx is not None:
                _weight[padding_idx].fill_(0)
        rows = int(num_embeddings * cache_ratio) if cuda_row_num is None else cuda_row_num
        self.cache_weight_mgr = OracleCachedParamMgr(_weight, rows, buffer_size, pin_weight,
                                                     evict_strategy=evict_strategy)
        self.cache_weight_mgr.reorder(ids_freq_mapping, warmup_ratio)
        self.cache_op = True

    def set_cache_op(self, cache_op: bool = True):
        self.cache_op = cache_op

    def forward(self, input, offsets=None, per_sample_weights=None, shape_hook=None):
        if self.cache_op:
            with torch.no_grad():
                input = self.cache_weight_mgr.prepare_ids(input)
        emb = F.embedding_bag(input, self.cache_weight_mgr.cuda_cached_weight, offsets, self.max_norm,
                              self.norm_type, self.scale_grad_by_freq, self.mode, self.sparse,
                              per_sample_weights, self.include_last_offset, self.padding_idx)
        if shape_hook is not None:
            emb = shape_hook(emb)
        return emb

    @property
    def weight(self):
        return self.cache_weight_mgr.weight

    def named_parameters(self, prefix: str = '', recurse: bool = True):
        yield 'weight', self.cache_weight_mgr.cuda_cached_weight

    def parameters(self, recurse: bool = True):
        yield self.cache_weight_mgr.cuda_cached_weight

    @property
    def num_hits_history(self):
        return self.cache_weight_mgr.num_hits_history

    @property
    def num_miss_history(self):
        return self.cache_weight_mgr.num_miss_history

    @property
    def num_write_back_history(self):
        return self.cache_weight_mgr.num_write_back_history


def get_partition(embedding_dim: int, rank: int, world_size: int):
    """Column split rule == torch.tensor_split (reference: recsys/utils/misc.py:138-154)."""
    if world_size == 1:
        return 0, embedding_dim, True
    assert embedding_dim >= world_size
    chunk, rem = divmod(embedding_dim, world_size)
    sizes = [chunk + 1 if i < rem else chunk for i in range(world_size)]
    start = sum(sizes[:rank])
    return start, start + sizes[rank], rem == 0


@dataclass
class OracleTablewiseConfig:
    """Mirror of TablewiseEmbeddingBagConfig (reference use: recsys/utils/misc.py:175-180)."""
    num_embeddings: int
    cuda_row_num: int
    assigned_rank: int = 0
    buffer_size: int = 50_000
    ids_freq_mapping: Optional[Sequence[int]] = None
    initial_weight: Optional[torch.Tensor] = None
    name: str = ""


class OracleTablewiseWorld:
    """A.6 simulated in ONE process: W per-rank bags + the table-wise all-to-all done with tensor ops.

    ``forward`` takes the *global* KJT (values = global ids over the concatenation of all tables, feature-major)
    and returns the list of per-rank outputs ``(B_r, sum_r F_r * D)`` with features in rank-major order.
    """

    def __init__(self, config_list: List[OracleTablewiseConfig], embedding_dim: int, world_size: int,
                 mode='mean', include_last_offset=False, sparse=True, cache_ratio=0.01, warmup_ratio=0.7,
                 buffer_size=0, evict_strategy=EvictionStrategy.LFU, per_sample_weights_supported=True):
        self.W = world_size
        self.D = embedding_dim
        self.mode = mode
        self.include_last_offset = include_last_offset
        self.rank_of_tables = [c.assigned_rank for c in config_list]
        self.table_rows = [c.num_embeddings for c in config_list]
        self.global_tables_num = len(config_list)
        self.global_offsets = np.concatenate([[0], np.cumsum(self.table_rows)]).astype(np.int64)
        self.assigned = [[i for i, r in enumerate(self.rank_of_tables) if r == rk] for rk in range(world_size)]
        self.dim_per_rank = [embedding_dim * len(a) for a in self.assigned]
        self.bags: List[Optional[OracleCachedEmbeddingBag]] = []
        self.idx_offset_list = []
        for rk in range(world_size):
            tabs = self.assigned[rk]
            if not tabs:
                self.bags.append(None)
                self.idx_offset_list.append([])
                continue
            n_local = sum(self.table_rows[t] for t in tabs)
            ws = [config_list[t].initial_weight for t in tabs]
            weight = torch.cat([w.clone() for w in ws], 0) if all(w is not None for w in ws) else None
            freqs = [config_list[t].ids_freq_mapping for t in tabs]
            freq = None
            if all(f is not None for f in freqs):
                freq = torch.cat([torch.as_tensor(f) for f in freqs])
            self.bags.append(
                OracleCachedEmbeddingBag(n_local, embedding_dim, sparse=sparse, _weight=weight, mode=mode,
                                         include_last_offset=include_last_offset, cache_ratio=cache_ratio,
                                         ids_freq_mapping=freq, warmup_ratio=warmup_ratio,
                                         buffer_size=buffer_size, evict_strategy=evict_strategy))
            # ids of local table k are re-based by the rows of the NON-local tables that precede it
            offs, local_prefix = [], 0
            for t in tabs:
                offs.append(int(self.global_offsets[t]) - local_prefix)
                local_prefix += self.table_rows[t]
            self.idx_offset_list.append(offs)

    def split_along_rank(self, rank, batch_size, indices, offsets, per_sample_weights=None):
        li, lo, lw = [], [], []
        pre_end = 0
        tabs = self.assigned[rank]
        for k, t in enumerate(tabs):
            start = int(offsets[batch_size * t])
            if (not self.include_last_offset) and batch_size * (t + 1) >= offsets.shape[0]:
                end = indices.shape[0]
            else:
                end = int(offsets[batch_size * (t + 1)])
            li.append(indices[start:end] - self.idx_offset_list[rank][k])
            if per_sample_weights is not None:
                lw.append(per_sample_weights[start:end])
            last = (k + 1 == len(tabs))
            take = batch_size + 1 if (last and self.include_last_offset) else batch_size
            seg = offsets[batch_size * t: batch_size * t + take] - int(offsets[batch_size * t]) + pre_end
            lo.append(seg)
            pre_end += end - start
        return (torch.cat(li), torch.cat(lo), torch.cat(lw) if per_sample_weights is not None else None)

    def forward(self, indices, offsets, per_sample_weights=None):
        n_off = offsets.shape[0] - (1 if self.include_last_offset else 0)
        B = n_off // self.global_tables_num
        strides = [B // self.W + int(i < B % self.W) for i in range(self.W)]
        self.local_outs = []
        for rk in range(self.W):
            if self.bags[rk] is None:
                self.local_outs.append(torch.zeros(B, 0))
                continue
            li, lo, lw = self.split_along_rank(rk, B, indices, offsets, per_sample_weights)
            out = self.bags[rk](li, lo, lw)                       # (F_loc*B, D)
            self.local_outs.append(torch.cat(out.split(B), 1))    # (B, F_loc*D)
        outs, begin = [], 0
        for j in range(self.W):
            outs.append(torch.cat([lo_[begin:begin + strides[j]] for lo_ in self.local_outs], 1))
            begin += strides[j]
        return outs

    def flush(self):
        for b in self.bags:
            if b is not None:
                b.cache_weight_mgr.flush()


class OracleColumnwiseWorld:
    """A.5 simulated in ONE process: rank r holds columns get_partition(D, r, W) of every row."""

    def __init__(self, weight: torch.Tensor, world_size: int, cuda_row_num: int, mode='mean',
                 include_last_offset=False, sparse=True, ids_freq_mapping=None, warmup_ratio=0.7,
                 buffer_size=0, evict_strategy=EvictionStrategy.DATASET):
        N, D = weight.shape
        self.W = world_size
        self.bags = []
        for r in range(world_size):
            s, e, _ = get_partition(D, r, world_size)
            self.bags.append(
                OracleCachedEmbeddingBag(N, e - s, sparse=sparse, _weight=weight[:, s:e].clone().contiguous(),
                                         mode=mode, include_last_offset=include_last_offset,
                                         ids_freq_mapping=ids_freq_mapping, warmup_ratio=warmup_ratio,
                                         buffer_size=buffer_size, evict_strategy=evict_strategy,
                                         cuda_row_num=cuda_row_num))

    def forward(self, indices, offsets, per_sample_weights=None, shape_hook=None, scatter_dim=0, gather_dim=-1):
        locals_ = [b(indices, offsets, per_sample_weights, shape_hook) for b in self.bags]
        self.local_outs = locals_
        outs = []
        for j in range(self.W):
            outs.append(torch.cat([torch.tensor_split(lo, self.W, dim=scatter_dim)[j] for lo in locals_],
                                  dim=gather_dim))
        return outs

    def flush(self):
        for b in self.bags:
            b.cache_weight_mgr.flush()

    def full_weight(self):
        return torch.cat([b.weight for b in self.bags], 1)


def rowwise_adagrad_reference(weight: np.ndarray, state: np.ndarray, indices: np.ndarray, offsets: np.ndarray,
                              grad_out: np.ndarray, lr: float, eps: float,
                              per_sample_weights: Optional[np.ndarray] = None, mode: str = 'sum'):
    """float64 restatement of one row-wise Adagrad step on the FULL table.

    g_r = sum over lookups i of row r of (w_i * grad_out[bag(i)]);  m_r += mean_j(g_rj^2);
    W_r -= lr * g_r / (sqrt(m_r) + eps).  PARITY UNPINNED: the reference trains the cached rows with plain SGD
    (recsys/dlrm_main.py:453-461); row-wise Adagrad is the north-star's extension, defined by this formula.
    ``offsets`` must include the last offset.  Returns (new_weight, new_state) as float64 arrays.
    """
    W = weight.astype(np.float64).copy()
    m = state.astype(np.float64).copy()
    G = len(offsets) - 1
    acc = {}
    for g in range(G):
        lo, hi = int(offsets[g]), int(offsets[g + 1])
        scale = 1.0 / (hi - lo) if (mode == 'mean' and hi > lo) else 1.0
        for i in range(lo, hi):
            w = scale * (float(per_sample_weights[i]) if per_sample_weights is not None else 1.0)
            r = int(indices[i])
            acc[r] = acc.get(r, 0.0) + w * grad_out[g].astype(np.float64)
    for r, g in acc.items():
        m[r] += float(np.mean(g * g))
        W[r] -= lr * g / (math.sqrt(m[r]) + eps)
    return W, m
