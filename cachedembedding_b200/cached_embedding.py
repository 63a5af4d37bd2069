"""CachedEmbeddingBag (a.k.a. FreqAwareEmbeddingBag): the nn.Module surface of the cached embedding bag.

Drop-in for ``colossalai.nn.parallel.layers.CachedEmbeddingBag`` as used by the reference
(/root/reference/benchmark/benchmark_cache.py:39-72, benchmark/benchmark_fbgemm_uvm.py:98-161; subclassed by the
parallel variants that /root/reference/recsys/models/dlrm.py:58-81 constructs).  Behaviour restated in SURVEY.md
Appendix A.2.  The forward is a hand-written sm_100a gather kernel over the slot cache; the backward is either

* ``fused``  : radix-sort by slot + segment reduce + optimizer update in place (SGD or row-wise Adagrad), nothing
               materialised -- enable with ``set_fused_optimizer``; ``torch.optim.SGD`` then sees ``grad is None`` and
               skips the parameter, so the reference training loop runs unchanged;
* ``sparse`` : the COO gradient ``sparse=True`` produces, for an external ``torch.optim.SGD`` (reference default,
               recsys/dlrm_main.py:455-461);
* ``dense``  : a dense [C, D] gradient (``sparse=False``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterator, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib
from .cache_mgr import CachedParamMgr, _stream_ptr
from .evict_strategy import EvictionStrategy

_MODES = {"sum": _lib.MODE_SUM, "mean": _lib.MODE_MEAN}
# tables above this many elements are initialised by the GPU writing straight into the pinned host table
_GPU_INIT_NUMEL = 1 << 28


def _bag_args(weight: torch.Tensor, slot_ids: torch.Tensor, offsets: torch.Tensor, psw: Optional[torch.Tensor],
              include_last_offset: bool, mode: int, padding_idx: Optional[int], layout: int,
              layout_batch: int) -> _lib.BagArgs:
    a = _lib.BagArgs()
    a.cache = weight.data_ptr()
    a.cache_rows, a.dim = weight.shape
    a.slot_ids = slot_ids.data_ptr()
    a.n = slot_ids.numel()
    a.offsets = offsets.data_ptr()
    a.offsets_are_64 = 1 if offsets.dtype == torch.int64 else 0
    a.include_last_offset = 1 if include_last_offset else 0
    a.num_bags = offsets.numel() - 1 if include_last_offset else offsets.numel()
    a.per_sample_weights = psw.data_ptr() if psw is not None else None
    a.mode = mode
    a.padding_idx = -1 if padding_idx is None else padding_idx
    a.layout = layout
    a.layout_batch = layout_batch
    return a


class BackwardPlan:
    """The gradient-independent half of a batch's fused backward, made ahead of time: either a private plan inside
    `workspace` (cebag_bag_backward_plan) or the batch's segment of a window plan (cebag_bag_backward_plan_window:
    `keys` / `vals` / `mask` point into the window's sort workspace, `workspace` is this batch's scratch)."""
    __slots__ = ("workspace", "nbytes", "offsets_ptr", "tag", "keep", "keys", "vals", "mask", "ready")

    def __init__(self, workspace, nbytes, offsets_ptr, tag, keep, keys=None, vals=None, mask=0, ready=None):
        self.workspace, self.nbytes, self.offsets_ptr, self.tag, self.keep = workspace, nbytes, offsets_ptr, tag, keep
        self.keys, self.vals, self.mask = keys, vals, mask
        self.ready = ready      # event recorded on the planning stream after the plan: its consumer waits for it

    def apply(self, args) -> int:
        """Point `args` at the plan; returns the workspace_has_plan value for cebag_bag_backward_fused."""
        if self.keys is None:
            return 1
        args.plan_keys, args.plan_vals, args.plan_key_mask = self.keys, self.vals, self.mask
        return 2


class _CachedBagFunction(torch.autograd.Function):
    """out = embedding_bag(cache[slot_ids], offsets); backward per the owning module's backward mode."""

    @staticmethod
    def forward(ctx, weight, slot_ids, offsets, psw, include_last_offset, mode, padding_idx, layout, layout_batch,
                owner):
        lib = _lib.load()
        a = _bag_args(weight, slot_ids, offsets, psw, include_last_offset, mode, padding_idx, layout, layout_batch)
        out = torch.empty(a.num_bags, weight.shape[1], dtype=weight.dtype, device=weight.device)
        _lib.check(lib.cebag_bag_forward(ctypes.byref(a), out.data_ptr(), _stream_ptr()))
        ctx.save_for_backward(weight, slot_ids, offsets, psw)
        ctx.cfg = (include_last_offset, mode, padding_idx, layout, layout_batch)
        ctx.owner = owner
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        weight, slot_ids, offsets, psw = ctx.saved_tensors
        include_last_offset, mode, padding_idx, layout, layout_batch = ctx.cfg
        owner = ctx.owner
        grad_out = grad_out.contiguous()
        a = _bag_args(weight, slot_ids, offsets, psw, include_last_offset, mode, padding_idx, layout, layout_batch)
        stream = _stream_ptr()
        grad_psw = None
        if psw is not None and ctx.needs_input_grad[3]:
            grad_psw = torch.empty_like(psw)
            _lib.check(lib.cebag_bag_backward_weights(ctypes.byref(a), grad_out.data_ptr(), grad_psw.data_ptr(), stream))
        grad_weight = None
        if ctx.needs_input_grad[0]:
            fused = owner._fused_optimizer if owner is not None else None
            if fused is not None:
                nbytes = int(lib.cebag_backward_workspace_bytes(ctypes.byref(a)))
                plan = owner._take_backward_plan(slot_ids, offsets, psw, mode, nbytes) \
                    if hasattr(owner, "_take_backward_plan") else None
                has_plan = 0
                if plan is not None:
                    ws = plan.workspace
                    if plan.ready is not None:       # made on another stream (look-ahead): may still be running
                        torch.cuda.current_stream().wait_event(plan.ready)
                    ws.record_stream(torch.cuda.current_stream())
                    has_plan = plan.apply(a)
                else:
                    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weight.device)
                state = owner.cache_weight_mgr.cuda_cached_state
                _lib.check(lib.cebag_bag_backward_fused(
                    ctypes.byref(a), grad_out.data_ptr(), weight.data_ptr(),
                    state.data_ptr() if state is not None else None, fused["kind"], fused["lr"], fused["eps"],
                    ws.data_ptr(), nbytes, has_plan, stream))
            elif owner is not None and not owner.sparse:
                nbytes = int(lib.cebag_backward_workspace_bytes(ctypes.byref(a)))
                ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weight.device)
                grad_weight = torch.empty_like(weight)
                _lib.check(lib.cebag_bag_backward_dense(ctypes.byref(a), grad_out.data_ptr(), grad_weight.data_ptr(),
                                                        ws.data_ptr(), nbytes, stream))
            else:
                values = torch.empty(slot_ids.numel(), weight.shape[1], dtype=weight.dtype, device=weight.device)
                _lib.check(lib.cebag_bag_backward_coo(ctypes.byref(a), grad_out.data_ptr(), values.data_ptr(), stream))
                grad_weight = torch.sparse_coo_tensor(slot_ids.view(1, -1), values, weight.shape)
        return grad_weight, None, None, grad_psw, None, None, None, None, None, None


def embedding_bag_cached(weight, slot_ids, offsets, per_sample_weights=None, include_last_offset=False, mode="sum",
                         padding_idx=None, layout="bag_major", layout_batch=0, owner=None):
    """Functional form: F.embedding_bag over a slot cache, run by the CUDA kernels of libcebag_b200."""
    if not weight.is_cuda:
        raise RuntimeError("embedding_bag_cached needs CUDA tensors: there is no CPU path")
    if mode not in _MODES:
        raise NotImplementedError(f"mode {mode!r}: only 'sum' and 'mean' are implemented")
    dev = weight.device
    slot_ids = slot_ids.to(device=dev, dtype=torch.long)
    if slot_ids.dim() == 2:
        if offsets is not None:
            raise ValueError("if input is 2D, then offsets has to be None")
        B, L = slot_ids.shape
        offsets = torch.arange(0, B * L + 1, L, dtype=torch.long, device=dev)
        include_last_offset = True
        if per_sample_weights is not None:
            per_sample_weights = per_sample_weights.reshape(-1)
    elif offsets is None:
        raise ValueError("offsets has to be a 1D Tensor for 1D input")
    slot_ids = slot_ids.contiguous().view(-1)
    offsets = offsets.to(dev)
    if offsets.dtype not in (torch.int32, torch.int64):
        offsets = offsets.long()
    offsets = offsets.contiguous()
    if per_sample_weights is not None:
        if mode != "sum":
            raise NotImplementedError("per_sample_weights is only supported for mode='sum'")
        per_sample_weights = per_sample_weights.to(device=dev, dtype=weight.dtype).contiguous()
    lay = _lib.LAYOUT_SAMPLE_MAJOR if layout == "sample_major" else _lib.LAYOUT_BAG_MAJOR
    return _CachedBagFunction.apply(weight, slot_ids, offsets, per_sample_weights, bool(include_last_offset),
                                    _MODES[mode], padding_idx, lay, int(layout_batch), owner)


class BaseEmbeddingBag(nn.Module):
    """Argument holder shared with torch.nn.EmbeddingBag (upstream base_embedding.py)."""

    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, max_norm=None, norm_type=2.,
                 scale_grad_by_freq=False, sparse=False, mode='mean', include_last_offset=False):
        super().__init__()
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        if padding_idx is not None:
            if padding_idx > 0:
                assert padding_idx < self.num_embeddings, 'Padding_idx must be within num_embeddings'
            elif padding_idx < 0:
                assert padding_idx >= -self.num_embeddings, 'Padding_idx must be within num_embeddings'
                padding_idx = self.num_embeddings + padding_idx
        self.padding_idx = padding_idx
        self.max_norm = max_norm
        self.norm_type = norm_type
        self.scale_grad_by_freq = scale_grad_by_freq
        self.sparse = sparse
        self.mode = mode
        self.include_last_offset = include_last_offset


class CachedEmbeddingBag(BaseEmbeddingBag):
    """Embedding bag whose table lives in pinned host DRAM with ``cache_ratio`` of its rows cached in HBM.

    Signature as upstream.  Extra keyword arguments (all optional): ``cuda_row_num`` overrides
    ``int(num_embeddings * cache_ratio)``; ``fused_optimizer`` / ``lr`` / ``eps`` enable the fused backward at
    construction (see ``set_fused_optimizer``); ``init_seed`` seeds the GPU-side initialiser of very large tables.
    """

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
                 cuda_row_num: Optional[int] = None,
                 fused_optimizer: Optional[str] = None,
                 lr: float = 0.0,
                 eps: float = 1e-8,
                 init_seed: int = 0):
        super().__init__(num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                         mode, include_last_offset)
        assert cache_ratio <= 1.0, f"cache ratio {cache_ratio} must less than 1.0"
        if max_norm is not None:
            raise NotImplementedError("max_norm renormalisation is not implemented")
        if scale_grad_by_freq:
            raise NotImplementedError("scale_grad_by_freq is not implemented")
        if dtype not in (None, torch.float32):
            raise NotImplementedError("the cached table is fp32")
        self.evict_strategy = evict_strategy
        self.cache_ratio = cache_ratio
        self._init_seed = init_seed
        if _weight is None:
            _weight = self._weight_alloc(dtype, device)
        if cuda_row_num is None:
            cuda_row_num = int(num_embeddings * cache_ratio)
        self._fused_optimizer = None
        env = os.environ.get("CEBAG_FUSED_OPTIMIZER")   # e.g. "sgd:lr=1.0" -- lets an unmodified training script opt in
        if fused_optimizer is None and env:
            kind, _, rest = env.partition(":")
            opts = dict(kv.split("=") for kv in rest.split(",") if "=" in kv)
            fused_optimizer, lr, eps = kind, float(opts.get("lr", lr)), float(opts.get("eps", eps))
        with_state = fused_optimizer in ("rowwise_adagrad", "adagrad")
        self._preprocess(_weight, cuda_row_num, ids_freq_mapping, warmup_ratio, buffer_size, pin_weight, with_state)
        self.cache_op = True
        if fused_optimizer is not None:
            self.set_fused_optimizer(fused_optimizer, lr=lr, eps=eps)

    # ---- construction ------------------------------------------------------------------------------------------------------
    def _weight_alloc(self, dtype, device) -> torch.Tensor:
        """U(-1/N, 1/N) like upstream.  Tables above 1 GiB are allocated pinned and filled by the GPU over PCIe."""
        N, D = self.num_embeddings, self.embedding_dim
        if N * D <= _GPU_INIT_NUMEL:
            w = torch.empty(N, D, dtype=torch.float32).uniform_(-1 / N, 1 / N)
        else:
            w = alloc_pinned_table(N, D)
            lib = _lib.load()
            dev = ctypes.c_void_p()
            _lib.check(lib.cebag_host_device_pointer(w.data_ptr(), ctypes.byref(dev)))
            _lib.check(lib.cebag_fill_uniform(dev.value, N * D, -1.0 / N, 1.0 / N, self._init_seed, _stream_ptr()))
            torch.cuda.current_stream().synchronize()
        if self.padding_idx is not None:
            w[self.padding_idx].fill_(0)
        return w

    def _preprocess(self, weight, cuda_row_num, ids_freq_mapping=None, warmup_ratio=0.7, buffer_size=50_000,
                    pin_weight=False, with_row_state=False):
        width = weight.shape[1]       # the row width the kernels will see (a column shard for the column-wise bag)
        limit = 512 if width % 4 == 0 else 128
        if width > limit:
            raise NotImplementedError(
                f"rows of {width} floats are not supported: the bag kernels keep a row in the registers of one warp, up "
                f"to 512 floats when the width is a multiple of 4 (128-bit accesses) and up to 128 floats otherwise")
        self.cache_weight_mgr = CachedParamMgr(weight, cuda_row_num, buffer_size, pin_weight,
                                               evict_strategy=self.evict_strategy, with_row_state=with_row_state)
        self.cache_weight_mgr.reorder(ids_freq_mapping, warmup_ratio)

    def set_fused_optimizer(self, kind: Optional[str], lr: float = 0.0, eps: float = 1e-8):
        """Apply the optimizer inside backward ('sgd' | 'rowwise_adagrad'); None returns to producing ``.grad``."""
        if kind is None:
            self._fused_optimizer = None
            return
        kinds = {"sgd": _lib.OPT_SGD, "rowwise_adagrad": _lib.OPT_ROWWISE_ADAGRAD, "adagrad": _lib.OPT_ROWWISE_ADAGRAD}
        if kind not in kinds:
            raise ValueError(f"unknown fused optimizer {kind!r}")
        if kinds[kind] == _lib.OPT_ROWWISE_ADAGRAD and self.cache_weight_mgr.cuda_cached_state is None:
            raise ValueError("row-wise Adagrad needs the per-row state: construct with fused_optimizer='rowwise_adagrad'")
        self._fused_optimizer = {"kind": kinds[kind], "lr": float(lr), "eps": float(eps)}

    # ---- backward plans (look-ahead) ----------------------------------------------------------------------------------------
    def plan_backward(self, slot_ids: torch.Tensor, offsets: torch.Tensor, layout="bag_major", layout_batch=0,
                      workspace_factory=None, tag=None) -> bool:
        """Run the gradient-independent half of the fused backward (lookup->bag map + radix sort by slot) for a batch NOW,
        on the current stream, and keep it until the backward of a forward over the very same `slot_ids` / `offsets`
        tensors picks it up.  A look-ahead driver calls this on its side stream right after prepare_ids
        (`workspace_factory(nbytes)` lets it supply a recycled buffer, `tag` names the group of plans that is dropped
        together when the buffers are recycled).  Returns False (and does nothing) when the fused backward is off or
        the bag is not in plain mode 'sum'."""
        if self._fused_optimizer is None or self.mode != "sum" or slot_ids.dim() != 1:
            return False
        lib = _lib.load()
        weight = self.cache_weight_mgr.cuda_cached_weight
        offsets = offsets.to(weight.device)
        if offsets.dtype not in (torch.int32, torch.int64):
            offsets = offsets.long()
        offsets = offsets.contiguous()
        lay = _lib.LAYOUT_SAMPLE_MAJOR if layout == "sample_major" else _lib.LAYOUT_BAG_MAJOR
        a = _bag_args(weight, slot_ids, offsets, None, self.include_last_offset, _lib.MODE_SUM, self.padding_idx, lay,
                      int(layout_batch))
        nbytes = int(lib.cebag_backward_workspace_bytes(ctypes.byref(a)))
        ws = workspace_factory(nbytes) if workspace_factory is not None else \
            torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weight.device)
        _lib.check(lib.cebag_bag_backward_plan(ctypes.byref(a), ws.data_ptr(), nbytes, _stream_ptr()))
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        if not hasattr(self, "_bwd_plans"):
            self._bwd_plans = {}
        # the plan is only valid for these very tensors: keep them alive so that their addresses cannot be recycled
        self._bwd_plans[(slot_ids.data_ptr(), slot_ids.numel())] = BackwardPlan(ws, nbytes, offsets.data_ptr(), tag,
                                                                               (slot_ids, offsets), ready=ready)
        return True

    def plan_backward_window(self, chunks, offsets_list, layout="bag_major", layout_batch=0, window_factory=None,
                             scratch_factory=None, tag=None) -> bool:
        """plan_backward for all batches of a look-ahead window at once: ONE radix sort over (batch, slot) instead of
        one per batch (cebag_bag_backward_plan_window).  `chunks[j]` are the slot ids of batch j (views of the window's
        slot-id tensor), `offsets_list[j]` its offsets; `window_factory(nbytes)` / `scratch_factory(j, nbytes)` supply
        recycled buffers."""
        if self._fused_optimizer is None or self.mode != "sum" or any(c.dim() != 1 for c in chunks):
            return False
        lib = _lib.load()
        weight = self.cache_weight_mgr.cuda_cached_weight
        P = len(chunks)
        lay = _lib.LAYOUT_SAMPLE_MAJOR if layout == "sample_major" else _lib.LAYOUT_BAG_MAJOR
        args = (_lib.BagArgs * P)()
        offs = []
        for j, (chunk, off) in enumerate(zip(chunks, offsets_list)):
            off = off.to(weight.device)
            if off.dtype not in (torch.int32, torch.int64):
                off = off.long()
            off = off.contiguous()
            offs.append(off)
            a = _bag_args(weight, chunk, off, None, self.include_last_offset, _lib.MODE_SUM, self.padding_idx, lay,
                          int(layout_batch))
            ctypes.memmove(ctypes.byref(args, j * ctypes.sizeof(_lib.BagArgs)), ctypes.byref(a), ctypes.sizeof(_lib.BagArgs))
        total = sum(int(c.numel()) for c in chunks)
        wbytes = int(lib.cebag_backward_window_plan_bytes(total))
        wws = window_factory(wbytes) if window_factory is not None else \
            torch.empty(max(wbytes, 16), dtype=torch.uint8, device=weight.device)
        keys = (ctypes.c_void_p * P)()
        vals = (ctypes.c_void_p * P)()
        mask = ctypes.c_uint32(0)
        _lib.check(lib.cebag_bag_backward_plan_window(args, P, wws.data_ptr(), wbytes, keys, vals, ctypes.byref(mask),
                                                      _stream_ptr()))
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        if not hasattr(self, "_bwd_plans"):
            self._bwd_plans = {}
        for j, (chunk, off) in enumerate(zip(chunks, offs)):
            nbytes = int(lib.cebag_backward_workspace_bytes(ctypes.byref(args[j])))
            scratch = scratch_factory(j, nbytes) if scratch_factory is not None else \
                torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weight.device)
            self._bwd_plans[(chunk.data_ptr(), chunk.numel())] = BackwardPlan(
                scratch, nbytes, off.data_ptr(), tag, (chunk, off, wws), keys[j], vals[j], int(mask.value), ready=ready)
        return True

    def drop_backward_plans(self, tag=None):
        """Forget plans that were never consumed (all of them, or the group `tag`): their workspace is about to be
        reused, or the batches they were made for will not run a backward."""
        plans = getattr(self, "_bwd_plans", None)
        if not plans:
            return
        for key in [k for k, v in plans.items() if tag is None or v.tag == tag]:
            del plans[key]

    def _take_backward_plan(self, slot_ids, offsets, psw, mode, nbytes):
        plans = getattr(self, "_bwd_plans", None)
        if not plans or psw is not None or mode != _lib.MODE_SUM:
            return None
        entry = plans.pop((slot_ids.data_ptr(), slot_ids.numel()), None)
        if entry is None:
            return None
        if entry.nbytes != nbytes or entry.offsets_ptr != offsets.data_ptr():
            return None
        return entry

    # ---- forward --------------------------------------------------------------------------------------------------------------
    def _embed(self, slot_ids, offsets, per_sample_weights, layout="bag_major", layout_batch=0):
        self.cache_weight_mgr.wait_rows()       # rows that a prepare_ids left moving on a copy stream
        return embedding_bag_cached(self.cache_weight_mgr.cuda_cached_weight, slot_ids, offsets, per_sample_weights,
                                    self.include_last_offset, self.mode, self.padding_idx, layout, layout_batch, self)

    def forward(self, input, offsets=None, per_sample_weights=None, shape_hook=None):
        if self.cache_op:
            with torch.no_grad():
                shape = input.shape
                input = self.cache_weight_mgr.prepare_ids(input).view(shape)
        embeddings = self._embed(input, offsets, per_sample_weights)
        if shape_hook is not None:
            embeddings = shape_hook(embeddings)
        return embeddings

    # ---- nn.Module surface the reference relies on ------------------------------------------------------------------------------
    @property
    def weight(self):
        return self.cache_weight_mgr.weight

    def named_parameters(self, prefix: str = '', recurse: bool = True) -> Iterator[Tuple[str, Parameter]]:
        yield 'weight', self.cache_weight_mgr.cuda_cached_weight

    def parameters(self, recurse: bool = True) -> Iterator[Parameter]:
        yield self.cache_weight_mgr.cuda_cached_weight

    def set_cache_op(self, cache_op: bool = True):
        self.cache_op = cache_op

    def set_cache_mgr_async_copy(self, flag):
        """Reference flag --use_cache_mgr_async_copy (recsys/dlrm_main.py:121,354): the PCIe row traffic of
        prepare_ids runs on the manager's copy stream; forward waits (on the device) for the missed rows only, the
        write-back of the victims proceeds under the following steps."""
        self.cache_weight_mgr._async_copy = bool(flag)

    def element_size(self):
        return self.weight.element_size()

    def print_comm_stats_(self):
        self.cache_weight_mgr.print_comm_stats()

    @property
    def num_hits_history(self):
        return self.cache_weight_mgr.num_hits_history

    @property
    def num_miss_history(self):
        return self.cache_weight_mgr.num_miss_history

    @property
    def num_write_back_history(self):
        return self.cache_weight_mgr.num_write_back_history

    @property
    def swap_in_bandwidth(self):
        t = max(self.cache_weight_mgr._elapsed_dict["cache_op"], 1e-12)
        return self.cache_weight_mgr._cpu_to_cuda_numel * self.cache_weight_mgr.elem_size_in_byte / 1e6 / t

    @property
    def swap_out_bandwidth(self):
        t = max(self.cache_weight_mgr._elapsed_dict["cache_op"], 1e-12)
        return self.cache_weight_mgr._cuda_to_cpu_numel * self.cache_weight_mgr.elem_size_in_byte / 1e6 / t

    @classmethod
    def from_pretrained(cls, embedding: torch.Tensor, freeze: bool = True, **kwargs):
        rows, cols = embedding.shape
        bag = cls(rows, cols, _weight=embedding, **kwargs)
        bag.cache_weight_mgr.cuda_cached_weight.requires_grad_(not freeze)
        return bag


# the name BASELINE.json uses for this module (the older upstream name of CachedEmbeddingBag)
FreqAwareEmbeddingBag = CachedEmbeddingBag


class _PinnedBlock:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.load().cebag_host_free(self.ptr)
        except Exception:
            pass


def alloc_pinned_table(num_rows: int, dim: int) -> torch.Tensor:
    """fp32[num_rows, dim] in page-locked, device-mapped host memory of exactly that size (torch's pinned allocator
    rounds large blocks up to a power of two: 128 GiB for the 91 GB Criteo-1TB table)."""
    lib = _lib.load()
    nbytes = num_rows * dim * 4
    ptr = ctypes.c_void_p()
    _lib.check(lib.cebag_host_alloc(ctypes.byref(ptr), nbytes))
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    buf._cebag_block = _PinnedBlock(ptr.value)   # torch keeps `buf` alive with the storage; freed with it
    return torch.frombuffer(buf, dtype=torch.float32).view(num_rows, dim)
