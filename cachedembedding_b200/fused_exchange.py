"""Fused pooled-embedding exchange over NVLink peer memory (table-wise sharding).

The reference's table-wise bag runs F.embedding_bag, repacks, and calls an NCCL all-to-all (forward) and its mirror
(backward) (SURVEY.md A.6 `dual_all_to_all_tablewise`).  At Criteo-1TB sizes that exchange -- pack copy, send/recv,
unpack copy, twice per step -- costs several times the embedding kernels themselves (measured 2.1 ms vs 0.34 ms per
step at W = 2).  Here the collective is folded into the kernels:

  forward : the gather kernel stores every pooled row straight into the (B_j, F, D) output buffer of the rank j that
            owns the row's sample, through a CUDA-IPC mapping of that buffer (NVLink stores);
  backward: the fused segment-reduce/optimizer kernel loads every gradient row straight from the owner's gradient
            buffer (NVLink loads), so no gradient is packed, sent or unpacked either.

What is left of the collective is ordering: a stream-ordered barrier (a 4-byte NCCL all-reduce) after the forward's
stores and before the backward's loads.  Buffers are plain cudaMalloc memory owned by the C library
(cebag_device_alloc) so that they can be exported with cudaIpcGetMemHandle.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from .cache_mgr import _stream_ptr
from .collectives import split_sizes


class _DevicePtr:
    """Minimal __cuda_array_interface__ carrier so torch can wrap memory it did not allocate."""

    def __init__(self, ptr: int, shape, keepalive):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}
        self._keepalive = keepalive


class PeerBuffer:
    """fp32 buffer of `numel` floats on every rank of `group`, each mapped into every other rank's address space."""

    def __init__(self, numel: int, group=None, zero: bool = False):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert self.world <= _lib.MAX_PEERS, f"at most {_lib.MAX_PEERS} ranks per node"
        self.lib = _lib.load()
        self.numel = int(numel)
        ptr = ctypes.c_void_p()
        _lib.check(self.lib.cebag_device_alloc(ctypes.byref(ptr), max(self.numel, 4) * 4))
        self.local_ptr = ptr.value
        if zero:
            self.tensor((max(self.numel, 4),)).zero_()
            torch.cuda.synchronize()
        handle = ctypes.create_string_buffer(64)
        _lib.check(self.lib.cebag_ipc_export(self.local_ptr, handle))
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs: List[int] = []
        self._imported: List[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            p = ctypes.c_void_p()
            _lib.check(self.lib.cebag_ipc_import(ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
            self.ptrs.append(p.value)
            self._imported.append(p.value)
        dist.barrier(group=group)

    def tensor(self, shape) -> torch.Tensor:
        """The local buffer as a torch tensor (no copy)."""
        n = 1
        for s in shape:
            n *= int(s)
        assert n <= max(self.numel, 4)
        return torch.as_tensor(_DevicePtr(self.local_ptr, shape, self), device=torch.device("cuda", torch.cuda.current_device()))

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._imported:
            self.lib.cebag_ipc_close(p)
        self._imported = []
        if self.local_ptr:
            self.lib.cebag_device_free(self.local_ptr)
            self.local_ptr = 0


class FusedExchange:
    """Output and gradient peer buffers of one table-wise bag + the stream-ordered barrier."""

    def __init__(self, global_batch: int, total_features: int, feature_offset: int, dim: int, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lib = _lib.load()
        self.B, self.F, self.D = int(global_batch), int(total_features), int(dim)
        self.feature_offset = int(feature_offset)
        self.strides = split_sizes(self.B, self.world)
        rows_max = max(self.strides)
        self.out_buf = PeerBuffer(rows_max * self.F * self.D, group)
        self.grad_buf = PeerBuffer(rows_max * self.F * self.D, group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._forward_pending = False      # a forward whose backward has not run yet
        self._out_view = self._grad_view = None
        self._bwd_bytes = {}
        # barrier over peer memory: one uint32 flag per peer on every rank (zeroed before anybody can write into it)
        self.flag_buf = PeerBuffer(_lib.MAX_PEERS, group, zero=True)
        self._xf = self._struct(self.flag_buf)
        self._barrier_seq = torch.zeros(1, dtype=torch.int32, device=dev)      # advanced on the device by every barrier
        self._barrier_failed = torch.zeros(1, dtype=torch.int32, device=dev)
        self.use_peer_barrier = os.environ.get("CEBAG_PEER_BARRIER", "1") != "0"
        self._xo = self._struct(self.out_buf)
        self._xg = self._struct(self.grad_buf)

    def _struct(self, buf: PeerBuffer) -> _lib.Exchange:
        x = _lib.Exchange()
        x.world, x.feature_offset, x.total_features = self.world, self.feature_offset, self.F
        for r in range(self.world):
            x.peer[r] = buf.ptrs[r]
        return x

    @property
    def local_rows(self) -> int:
        return self.strides[self.rank]

    def out_tensor(self) -> torch.Tensor:
        """This rank's (B_rank, F * D) slice of the pooled embeddings: a view of the peer-mapped output buffer (wrapping
        foreign memory in a tensor costs ~30 us of host time, so the view is made once)."""
        if self._out_view is None:
            self._out_view = self.out_buf.tensor((self.local_rows, self.F * self.D))
        return self._out_view

    def grad_tensor(self) -> torch.Tensor:
        if self._grad_view is None:
            self._grad_view = self.grad_buf.tensor((self.local_rows, self.F * self.D))
        return self._grad_view

    def barrier(self):
        """All ranks' work enqueued so far on their current streams is complete before anything enqueued after it
        starts.  One tiny kernel over peer memory (cebag_peer_barrier: a release store into every peer's flag array,
        then an acquire spin on the own one); CEBAG_PEER_BARRIER=0 falls back to a 4-byte NCCL all-reduce."""
        if not self.use_peer_barrier:
            dist.all_reduce(self._flag, group=self.group)
            return
        _lib.check(self.lib.cebag_peer_barrier(ctypes.byref(self._xf), self.rank, self._barrier_seq.data_ptr(),
                                               self._barrier_failed.data_ptr(), _stream_ptr()))

    def check_barriers(self):
        """Raises if a peer ever failed to arrive at a barrier (reads one int from the device)."""
        if int(self._barrier_failed.item()):
            raise RuntimeError("fused exchange: a peer did not arrive at a barrier within the timeout")

    def close(self):
        self.out_buf.close()
        self.grad_buf.close()
        self.flag_buf.close()


class _FusedTablewiseFunction(torch.autograd.Function):
    """forward: gather + store to the owners' output buffers; backward: load from the owners' gradient buffers +
    fused segment-reduce / optimizer.  The weight gets no gradient: the update happens in the kernel."""

    @staticmethod
    def forward(ctx, weight, slot_ids, offsets, bag, exch: FusedExchange):
        from .cached_embedding import _bag_args
        lib = _lib.load()
        a = _bag_args(weight, slot_ids, offsets, None, bag.include_last_offset, _lib.MODE_SUM, bag.padding_idx,
                      _lib.LAYOUT_EXCHANGE, exch.B)
        a.exchange = ctypes.pointer(exch._xo)
        if exch._forward_pending:
            exch.barrier()                  # two forwards in a row: peers may still be reading the previous output
        exch._forward_pending = True
        _lib.check(lib.cebag_bag_forward(ctypes.byref(a), None, _stream_ptr()))
        exch.barrier()                      # every rank's rows have landed in my buffer
        ctx.save_for_backward(weight, slot_ids, offsets)
        ctx.bag, ctx.exch, ctx.args = bag, exch, a
        return exch.out_tensor().view(exch.local_rows, exch.F * exch.D)

    @staticmethod
    def backward(ctx, grad_out):
        from .cached_embedding import _bag_args
        lib = _lib.load()
        weight, slot_ids, offsets = ctx.saved_tensors
        bag, exch = ctx.bag, ctx.exch
        exch._forward_pending = False
        fused = bag._fused_optimizer
        gbuf = exch.grad_tensor()
        if grad_out.data_ptr() != gbuf.data_ptr():
            gbuf.copy_(grad_out.reshape(gbuf.shape))        # producers that write into grad_tensor() skip this copy
        exch.barrier()                      # every rank's gradient is in place before anyone reads it
        a = ctx.args                        # the forward's arguments, now pointing at the gradient buffers
        a.exchange = ctypes.pointer(exch._xg)
        key = (int(a.n), int(a.dim))
        nbytes = exch._bwd_bytes.get(key)
        if nbytes is None:
            nbytes = exch._bwd_bytes[key] = int(lib.cebag_backward_workspace_bytes(ctypes.byref(a)))
        plan = bag._take_backward_plan(slot_ids, offsets, None, _lib.MODE_SUM, nbytes)
        ws = plan.workspace if plan is not None else torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weight.device)
        has_plan = plan.apply(a) if plan is not None else 0
        state = bag.cache_weight_mgr.cuda_cached_state
        _lib.check(lib.cebag_bag_backward_fused(
            ctypes.byref(a), None, weight.data_ptr(), state.data_ptr() if state is not None else None,
            fused["kind"], fused["lr"], fused["eps"], ws.data_ptr(), nbytes, has_plan, _stream_ptr()))
        return None, None, None, None, None
