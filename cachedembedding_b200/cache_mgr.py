"""CachedParamMgr: host-resident table + HBM slot cache + id maps, driven by libcebag_b200's CUDA kernels.

Mirrors ColossalAI's ``colossalai.nn.parallel.layers.cache_embedding.CachedParamMgr`` (the class behind
``embed.cache_weight_mgr`` at /root/reference/recsys/dlrm_main.py:259; behaviour restated in SURVEY.md Appendix
A.1/A.3/A.4): same constructor, same methods (``prepare_ids``, ``reorder``, ``flush``, ``print_comm_stats``, the legacy
single-row helpers), same attributes (``cuda_cached_weight``, ``weight``, ``idx_map``, ``cached_idx_map``,
``inverted_cached_idx``, ``freq_cnter``, hit/miss histories).  What differs is how the work is done:

* the maps are int32 on the device (``cached_idx_map`` & co. are int64 *views* materialised on access);
* ``prepare_ids`` is one C call that only ENQUEUES hand-written kernels (no sort-based unique/isin/topk, no wait for
  the GPU inside the call: misses, evictions and free slots are counted in device memory) and moves rows between the
  pinned host table and HBM with zero-copy 128-bit accesses from the GPU (no CPU gather/scatter);
* with ``async_copy`` (``set_cache_mgr_async_copy(True)``, reference flag ``--use_cache_mgr_async_copy``) the PCIe row
  traffic runs on a private copy stream: victims are parked in an HBM staging buffer, the missed rows are filled, and
  the write-back to the host table goes on under the following forward/backward steps;
* there is no CPU path: constructing a manager without a CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
import sys
import time
from collections import deque
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import CacheCapacityError  # noqa: F401  (re-export)
from .evict_strategy import EvictionStrategy


_RESULT_RING = 64     # result records of prepare_ids calls that may be in flight at once


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class _PinnedRegistration:
    """Keeps a host tensor page-locked (in place) for as long as the manager lives."""

    def __init__(self, tensor: torch.Tensor):
        self.ptr = tensor.data_ptr()
        self.registered = False
        lib = _lib.load()
        if not tensor.is_pinned():
            _lib.check(lib.cebag_host_register(self.ptr, tensor.numel() * tensor.element_size()))
            self.registered = True
        dev = ctypes.c_void_p()
        _lib.check(lib.cebag_host_device_pointer(self.ptr, ctypes.byref(dev)))
        self.device_ptr = dev.value

    def release(self):
        if self.registered:
            try:
                _lib.load().cebag_host_unregister(self.ptr)
            except Exception:
                pass
            self.registered = False

    def __del__(self):
        self.release()


class CachedParamMgr(nn.Module):
    """Manage an embedding table whose rows live in host DRAM with a fixed number of them cached in HBM.

    Args mirror upstream (A.1): ``weight`` is the host table fp32[N, D]; ``cuda_row_num`` the number of HBM slots.
    ``buffer_size`` is accepted for API compatibility: the reference used it to bound the host staging buffer of its
    chunked copier (A.7); rows here move between the pinned table and HBM without host staging, so there is nothing
    to bound.
    ``with_row_state`` adds one fp32 per row (row-wise Adagrad accumulator) that travels with the row.
    """

    def __init__(self,
                 weight: torch.Tensor,
                 cuda_row_num: int = 0,
                 buffer_size: int = 0,
                 pin_weight: bool = True,
                 evict_strategy: EvictionStrategy = EvictionStrategy.DATASET,
                 async_copy: bool = False,
                 with_row_state: bool = False):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("CachedParamMgr needs a CUDA device: the B200 build has no CPU path")
        assert weight.dim() == 2 and weight.dtype == torch.float32, "weight must be fp32 [N, D]"
        assert weight.device.type == "cpu", "the full table lives in host memory"
        if cuda_row_num == 0:
            raise NotImplementedError("cuda_row_num == 0")
        self._lib = _lib.load()
        self.buffer_size = buffer_size
        self.num_embeddings, self.embedding_dim = weight.shape
        assert self.num_embeddings < 2**31, "row ids are int32 on the device"
        self.cuda_row_num = int(cuda_row_num)
        self._cuda_available_row_num = self.cuda_row_num
        self.pin_weight = pin_weight
        self.elem_size_in_byte = weight.element_size()
        self._evict_strategy = evict_strategy
        self._async_copy = async_copy
        self.device = torch.device("cuda", torch.cuda.current_device())

        self.weight = weight.contiguous()
        # zero-copy row movement needs the table page-locked; register in place (no second copy of the table)
        self._pin = _PinnedRegistration(self.weight)

        N, C, D = self.num_embeddings, self.cuda_row_num, self.embedding_dim
        dev = self.device
        self.cuda_cached_weight = nn.Parameter(torch.zeros(C, D, dtype=torch.float32, device=dev))
        self.register_buffer("_row2slot", torch.full((N,), -1, dtype=torch.int32, device=dev), persistent=False)
        self.register_buffer("_slot2row", torch.full((C,), -1, dtype=torch.int32, device=dev), persistent=False)
        self.register_buffer("_slot_epoch", torch.zeros(C, dtype=torch.int32, device=dev), persistent=False)
        self.register_buffer("_miss_bitmap", torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev),
                             persistent=False)
        self._idx_map: Optional[torch.Tensor] = None     # int32[N]; None == identity
        if evict_strategy == EvictionStrategy.LFU:
            self.register_buffer("_freq", torch.full((C,), sys.maxsize, dtype=torch.int64, device=dev),
                                 persistent=False)
        else:
            self._freq = None
        self.with_row_state = with_row_state
        if with_row_state:
            self.row_state = torch.zeros(N, dtype=torch.float32)          # host, travels with the row
            self._state_pin = _PinnedRegistration(self.row_state)
            self.register_buffer("cuda_cached_state", torch.zeros(C, dtype=torch.float32, device=dev),
                                 persistent=False)
        else:
            self.row_state = None
            self._state_pin = None
            self.cuda_cached_state = None
        self.register_buffer("_hit_flags", torch.zeros((C + 15) // 16 * 16, dtype=torch.uint8, device=dev),
                             persistent=False)
        state = torch.zeros(_lib.STATE_WORDS, dtype=torch.int64)
        state[_lib.STATE_AVAIL] = C
        self.register_buffer("_dev_state", state.to(dev), persistent=False)
        self.protect_windows = 1      # 2 while a LookaheadPrefetcher overlaps prepare_ids with the previous window
        # stream plumbing of the row traffic (see _workspace)
        self._copy_stream = None      # a look-ahead driver's copy stream, for the calls it makes
        self._own_copy_stream = None  # created on first use when async_copy is on
        self._victims_ready = None    # event the victims' rows have to wait for (set by the look-ahead driver)
        self._rows_ready = None       # event after the last fill on a copy stream: the next forward waits for it
        self._last_writeback = None   # event after the last write-back on a copy stream: flush waits for it
        self._defer_results = False   # look-ahead driver: results are read when the window is waited for
        self.stage_rows = 0           # 0: max(65536, C // 4) rows, capped at C
        # DMA write-back (with a copy stream): parked victims leave through a copy engine into a pinned ring and host
        # threads scatter them into the table (csrc/writeback_pool.cpp).  Off by default (CEBAG_DMA_WRITEBACK=1 or this
        # attribute turn it on): at Criteo-1TB the row traffic is no longer the critical path, and the zero-copy kernel
        # keeps the host threads out of the loop
        self.dma_writeback = os.environ.get("CEBAG_DMA_WRITEBACK", "0") == "1"
        self._dma_stream = None
        self._dma_ring = None         # pinned [rows, D] (+ row list, + state): one ring, the DMA stream is serial
        self._last_evicted = 0        # E of the last call whose result has been read: the estimate for the next DMA
        self._marked = deque()        # calls whose victims still carry markers: (call index, ws buffer, dma event, stage entry)
        self._forwarding_from = None
        self._ws_ring = [None, None, None]       # [buffer, event after the call, write-back event]
        self._ws_next = 0
        self._stage_ring = [None, None, None]    # [rows, state, events to wait for before reuse]
        self._stage_next = 0
        self._results = torch.full((_RESULT_RING, 8), -1, dtype=torch.int64).pin_memory()
        self._result_next = 0
        self._pending = deque()       # (result index, event, n): calls whose result has not been read yet
        self._counters_pinned = torch.zeros(64, dtype=torch.int32).pin_memory()
        self._calls = 0

        self.evict_backlist = torch.tensor([], device=dev)
        self._num_hits_history: List[int] = []
        self._num_miss_history: List[int] = []
        self._num_write_back_history: List[int] = []
        self._cpu_to_cuda_numel = 0
        self._cuda_to_cpu_numel = 0
        self._cache_miss = 0
        self._total_cache = 0
        self._elapsed_dict = {"cache_op": 0.0}

    # ---- the C-ABI view of this manager ---------------------------------------------------------------------------
    def _table(self) -> _lib.Table:
        t = _lib.Table()
        t.num_rows = self.num_embeddings
        t.dim = self.embedding_dim
        t.cache_rows = self.cuda_row_num
        t.strategy = _lib.EVICT_LFU if self._evict_strategy == EvictionStrategy.LFU else _lib.EVICT_DATASET
        t.protect_windows = self.protect_windows
        t.host_table = self._pin.device_ptr
        t.host_state = self._state_pin.device_ptr if self._state_pin is not None else None
        t.cache = self.cuda_cached_weight.data_ptr()
        t.cache_state = self.cuda_cached_state.data_ptr() if self.cuda_cached_state is not None else None
        t.idx_map = self._idx_map.data_ptr() if self._idx_map is not None else None
        t.row2slot = self._row2slot.data_ptr()
        t.slot2row = self._slot2row.data_ptr()
        t.freq = self._freq.data_ptr() if self._freq is not None else None
        t.slot_epoch = self._slot_epoch.data_ptr()
        t.miss_bitmap = self._miss_bitmap.data_ptr()
        t.hit_flags = self._hit_flags.data_ptr()
        t.dev_state = self._dev_state.data_ptr()
        return t

    def _refresh_avail(self):
        """Free slots after an operation that changed them outside prepare_ids (waits for the stream)."""
        t = self._table()
        avail = ctypes.c_int64(0)
        _lib.check(self._lib.cebag_available_rows(ctypes.byref(t), ctypes.byref(avail), _stream_ptr()))
        self._cuda_available_row_num = int(avail.value)

    def _active_copy_stream(self):
        """Where the PCIe row traffic of the next prepare_ids goes: the look-ahead driver's copy stream, the
        manager's own one when async_copy is on, or None (= the calling stream, the reference's serial order)."""
        if self._copy_stream is not None:
            return self._copy_stream
        if self._async_copy:
            if self._own_copy_stream is None:
                self._own_copy_stream = torch.cuda.Stream(device=self.device, priority=-1)
            return self._own_copy_stream
        return None

    def _workspace(self, t: _lib.Table, n_ids: int, plumbing: bool = True):
        """Scratch + stream plumbing of one call.  Buffers come from small rings owned by the manager (nothing is
        allocated per call in steady state); a buffer is handed out again only after the calling stream has been made
        to wait for the kernels -- on either stream -- that still read it."""
        cur = torch.cuda.current_stream()
        nbytes = int(self._lib.cebag_prepare_workspace_bytes(ctypes.byref(t), n_ids))
        i = self._ws_next
        self._ws_next = (i + 1) % len(self._ws_ring)
        entry = self._ws_ring[i]
        if entry is not None:
            for ev in entry[1:]:
                if ev is not None:
                    cur.wait_event(ev)
        if entry is None or entry[0].numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        else:
            buf = entry[0]
        ws = _lib.Workspace()
        ws.device, ws.device_bytes, ws.pinned = buf.data_ptr(), buf.numel(), self._counters_pinned.data_ptr()
        entry = [buf, None, None]
        self._ws_ring[i] = entry
        if not plumbing:
            return ws, entry, None
        call = self._calls
        # markers left by earlier calls (DMA write-back): everything older than the previous call is retired now -- its
        # write-back is waited for on the device -- and the previous call's staging buffer is where re-admitted rows are
        # filled from
        while self._marked and self._marked[0][0] <= call - 2:
            _, old_buf, dma_done, _ = self._marked.popleft()
            ws.retire_device = old_buf.data_ptr()
            ws.retire_wait_event = dma_done.cuda_event
        prev_stage_entry = self._marked[-1][3] if self._marked else None
        self._forwarding_from = prev_stage_entry
        if prev_stage_entry is not None:
            ws.prev_stage = prev_stage_entry[0].data_ptr()
            ws.prev_stage_state = prev_stage_entry[1].data_ptr() if prev_stage_entry[1] is not None else None
        cstream = self._active_copy_stream()
        events = None
        if cstream is not None:
            # timing-enabled: a look-ahead driver that traces its pipeline reads the end of the fill from rows_done
            rows_done, wb_done = torch.cuda.Event(enable_timing=True), torch.cuda.Event()
            rows_done.record(cstream)          # instantiates the events; re-recorded by the library after the copies
            wb_done.record(cstream)
            ws.copy_stream = cstream.cuda_stream
            ws.copy_done_event = rows_done.cuda_event
            ws.writeback_done_event = wb_done.cuda_event
            entry[2] = wb_done
            events = (rows_done, wb_done)
            if prev_stage_entry is not None:
                prev_stage_entry[2].append(rows_done)      # this call's fill may read that buffer
            # staging buffer for the victims (a ring: the write-back of earlier calls may still read theirs)
            C, D = self.cuda_row_num, self.embedding_dim
            rows = self.stage_rows if self.stage_rows > 0 else max(65536, C // 4)
            rows = min(rows, C)
            j = self._stage_next
            self._stage_next = (j + 1) % len(self._stage_ring)
            st = self._stage_ring[j]
            if st is not None:
                for ev in st[2]:
                    cur.wait_event(ev)
            if st is None or st[0].shape[0] != rows:
                st = [torch.empty(rows, D, dtype=torch.float32, device=self.device),
                      torch.empty(rows, dtype=torch.float32, device=self.device) if self.with_row_state else None, []]
            st[2] = [wb_done]
            self._stage_ring[j] = st
            ws.stage = st[0].data_ptr()
            ws.stage_state = st[1].data_ptr() if st[1] is not None else None
            ws.stage_rows = rows
            dma_rows = min(rows, int(self._last_evicted * 1.25) + 256) if self.dma_writeback and self._last_evicted > 0 else 0
            if self.dma_writeback:
                # markers are kept for every parked victim whenever the DMA path is configured (dma_rows may be 0)
                if self._dma_stream is None:
                    self._dma_stream = torch.cuda.Stream(device=self.device)
                if self._dma_ring is None or self._dma_ring[0].shape[0] != rows:
                    self._dma_ring = (torch.empty(rows, D, dtype=torch.float32).pin_memory(),
                                      torch.empty(rows, dtype=torch.int32).pin_memory(),
                                      torch.empty(rows, dtype=torch.float32).pin_memory() if self.with_row_state else None)
                dma_done = torch.cuda.Event()
                dma_done.record(self._dma_stream)
                ws.dma_stream = self._dma_stream.cuda_stream
                ws.dma_done_event = dma_done.cuda_event
                if self._last_writeback is not None:
                    ws.dma_wait_event = self._last_writeback.cuda_event
                ws.dma_ring = self._dma_ring[0].data_ptr()
                ws.dma_ring_rows = self._dma_ring[1].data_ptr()
                ws.dma_ring_state = self._dma_ring[2].data_ptr() if self._dma_ring[2] is not None else None
                ws.dma_rows = max(dma_rows, 1)
                ws.host_table_hostptr = self.weight.data_ptr()
                ws.host_state_hostptr = self.row_state.data_ptr() if self.row_state is not None else None
                st[2].append(dma_done)
                self._marked.append((call, buf, dma_done, st))
        if self._victims_ready is not None:
            ws.victims_ready_event = self._victims_ready.cuda_event
        return ws, entry, events

    # ---- results of the asynchronous calls -----------------------------------------------------------------------------
    def _harvest(self, block: bool = True):
        """Read the result records of finished prepare_ids calls (all of them if `block`): histories and counters are
        updated in call order; a rejected call raises here (CacheCapacityError / IndexError)."""
        while self._pending:
            idx, ev, n = self._pending[0]
            if not block and not ev.query():
                break
            ev.synchronize()
            self._pending.popleft()
            rec = self._results[idx]
            t = self._table()
            stats = _lib.PrepareStats()
            rc = self._lib.cebag_prepare_result_status(ctypes.byref(t), rec.data_ptr(), ctypes.byref(stats))
            _lib.check(rc)
            self._cuda_available_row_num = int(rec[7])
            self._last_evicted = int(stats.evicted)
            self._cache_miss += stats.miss_lookups
            self._total_cache += n
            self._num_hits_history.append(int(stats.unique_hits))
            self._num_miss_history.append(int(stats.unique_misses))
            self._num_write_back_history.append(int(stats.evicted))
            self._cpu_to_cuda_numel += stats.unique_misses * self.embedding_dim
            self._cuda_to_cpu_numel += stats.evicted * self.embedding_dim

    @property
    def num_hits_history(self) -> List[int]:
        self._harvest()
        return self._num_hits_history

    @property
    def num_miss_history(self) -> List[int]:
        self._harvest()
        return self._num_miss_history

    @property
    def num_write_back_history(self) -> List[int]:
        self._harvest()
        return self._num_write_back_history

    def wait_rows(self):
        """Make the current stream wait for the last fill that ran on a copy stream (no-op otherwise)."""
        if self._rows_ready is not None:
            torch.cuda.current_stream().wait_event(self._rows_ready)
            self._rows_ready = None

    def _wait_writeback(self):
        if self._last_writeback is not None:
            torch.cuda.current_stream().wait_event(self._last_writeback)
            self._last_writeback = None

    # ---- reference-compatible views of the maps (int64, like upstream's buffers) ------------------------------------
    @property
    def idx_map(self) -> torch.Tensor:
        if self._idx_map is None:
            return torch.arange(self.num_embeddings, dtype=torch.long, device=self.device)
        return self._idx_map.long()

    @property
    def cached_idx_map(self) -> torch.Tensor:
        return self._slot2row.long()

    @property
    def inverted_cached_idx(self) -> torch.Tensor:
        return self._row2slot.long().clamp_(min=-1)

    @property
    def freq_cnter(self) -> torch.Tensor:
        if self._freq is None:
            raise AttributeError("freq_cnter exists only under EvictionStrategy.LFU")
        return self._freq

    @property
    def cuda_available_row_num(self) -> int:
        self._harvest()
        return self._cuda_available_row_num

    def cpu_weight_data(self, row_idx: int) -> torch.Tensor:
        return self.weight.data.view(-1).narrow(0, int(row_idx) * self.embedding_dim,
                                                self.embedding_dim).view(1, self.embedding_dim)

    @property
    def cuda_weight(self):
        return self.cuda_cached_weight

    # ---- A.1 reorder ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def reorder(self, ids_freq_mapping=None, warmup_ratio: float = 0.7):
        """Build the id -> row remap (DATASET) and preload the warm-up rows into slots 0..k-1."""
        dev = self.device
        freq = None
        if ids_freq_mapping is not None:
            freq = torch.as_tensor(ids_freq_mapping).to(device=dev, dtype=torch.long)
            assert freq.numel() == self.num_embeddings, "ids_freq_mapping needs one count per id"
        if freq is not None and self._evict_strategy == EvictionStrategy.DATASET:
            # id -> rank by descending frequency (stable, so equal counts have one answer)
            order = torch.argsort(freq, descending=True, stable=True)
            idx_map = torch.empty(self.num_embeddings, dtype=torch.int32, device=dev)
            idx_map[order] = torch.arange(self.num_embeddings, dtype=torch.int32, device=dev)
            self._idx_map = idx_map
        preload = min(int(np.ceil(self.cuda_row_num * warmup_ratio)), self.num_embeddings)
        if preload > 0:
            freq_init = None
            if self._evict_strategy == EvictionStrategy.LFU and freq is not None:
                order = torch.argsort(freq, descending=True, stable=True)[:preload]
                rows = order.to(torch.int32).contiguous()
                freq_init = freq[order].contiguous()
            else:
                rows = torch.arange(preload, dtype=torch.int32, device=dev)
            t = self._table()
            _lib.check(self._lib.cebag_preload(ctypes.byref(t), rows.data_ptr(),
                                               freq_init.data_ptr() if freq_init is not None else None,
                                               preload, _stream_ptr()))
            self._refresh_avail()                        # waits for the stream: rows / freq_init are locals

    # ---- A.1 flush ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def flush(self):
        """Write every resident row back to the host table and empty the cache."""
        self._harvest()
        self.wait_rows()
        self._wait_writeback()
        t = self._table()
        if self._dma_stream is not None:
            self._dma_stream.synchronize()      # every DMA write-back is in the table; flush clears the markers
        self._marked.clear()
        ws, entry, _ = self._workspace(t, 1, plumbing=False)
        written = ctypes.c_int64(0)
        _lib.check(self._lib.cebag_flush(ctypes.byref(t), ctypes.byref(ws), ctypes.byref(written), _stream_ptr()))
        if self._own_copy_stream is not None:
            self._own_copy_stream.synchronize()
        self._cuda_available_row_num = self.cuda_row_num
        self._cuda_to_cpu_numel += written.value * self.embedding_dim
        return written.value

    # ---- A.3 prepare_ids --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def prepare_ids(self, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Make every row that ``ids`` touches resident and return the slot of each id (int64, same order).

        One call covers a whole look-ahead window (reference: recsys/dlrm_main.py:259 passes the concatenation of
        ``prefetch_num`` batches).  Raises ``CacheCapacityError`` (an ``AssertionError``) with the reference's message
        when the window's unique rows exceed ``cuda_row_num``; the cache is left unchanged in that case.
        """
        start = time.perf_counter()
        ids = ids.to(device=self.device, dtype=torch.long).contiguous().view(-1)
        n = ids.numel()
        if out is None:
            out = torch.empty_like(ids)
        else:       # a caller-owned buffer (the look-ahead driver's static ring: stable addresses for CUDA graphs)
            assert out.is_cuda and out.dtype == torch.long and out.is_contiguous() and out.numel() == n
        self._calls += 1
        if self._calls >= 2**31 - 64:                 # the window stamps are int32
            self._reset_stamps()
        t = self._table()
        ws, entry, copy_events = self._workspace(t, n)
        idx = self._result_next
        self._result_next = (idx + 1) % _RESULT_RING
        if len(self._pending) >= _RESULT_RING - 1:
            self._harvest()
        rec = self._results[idx]
        rc = self._lib.cebag_prepare_ids_async(ctypes.byref(t), ids.data_ptr(), n, out.data_ptr(), ctypes.byref(ws),
                                               rec.data_ptr(), _stream_ptr())
        _lib.check(rc)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        entry[1] = done
        if copy_events is None and self._forwarding_from is not None:
            self._forwarding_from[2].append(done)      # this call's fill (on this stream) may have read that buffer
        self._forwarding_from = None
        self._pending.append((idx, done, n))
        if copy_events is not None:
            self._rows_ready, self._last_writeback = copy_events
        if not self._defer_results:
            self._harvest()       # the reference raises its capacity assert from this call
        self._elapsed_dict["cache_op"] += time.perf_counter() - start
        return out

    def _reset_stamps(self):
        torch.cuda.synchronize(self.device)
        self._harvest()
        self._slot_epoch.zero_()
        self._dev_state[_lib.STATE_EPOCH] = 0
        self._calls = 0

    def _id_to_cached_cuda_id(self, ids: torch.Tensor) -> torch.Tensor:
        ids = ids.to(self.device).view(-1)
        rows = ids if self._idx_map is None else self._idx_map[ids].long()
        return self._row2slot[rows].long().clamp_(min=-1)

    @torch.no_grad()
    def _prepare_rows_on_cuda(self, cpu_row_idxs: torch.Tensor) -> None:
        """Upstream's inner step (A.4), kept for tests: bring the given (already remapped, non-resident) rows in."""
        rows = cpu_row_idxs.to(self.device).long().view(-1)
        if rows.numel() == 0:
            return
        if self._idx_map is not None:
            inv = torch.empty_like(self._idx_map)
            inv[self._idx_map.long()] = torch.arange(self.num_embeddings, dtype=torch.int32, device=self.device)
            ids = inv[rows].long()
        else:
            ids = rows
        self._harvest()
        hits, misses, wb = len(self.num_hits_history), len(self.num_miss_history), len(self.num_write_back_history)
        freq_before = self._freq.clone() if self._freq is not None else None
        resident_before = self._row2slot[rows] >= 0
        slots = self.prepare_ids(ids)
        # this entry point is not a lookup: undo prepare_ids' bookkeeping (histories, LFU counts of the ids)
        del self._num_hits_history[hits:], self._num_miss_history[misses:], self._num_write_back_history[wb:]
        if self._freq is not None:
            freq_before[slots[~resident_before]] = 0      # A.4 step 6: admitted slots start at 0
            self._freq.copy_(freq_before)

    # ---- legacy single-row helpers (upstream test_cachemgr, B.1) ------------------------------------------------------------
    def _row_in_cuda(self, row_id: int) -> bool:
        return bool(self._row2slot[row_id].item() >= 0)

    def _find_free_cuda_row(self) -> int:
        if self.cuda_available_row_num == 0:
            return -1
        return int(torch.nonzero(self._slot2row == -1).squeeze(1)[0].item())

    @torch.no_grad()
    def _evict(self) -> int:
        masked = self._slot2row.clone()
        max_row, slot = torch.max(masked, dim=0)
        if max_row.item() == -1:
            raise RuntimeError("Can not evict a row")
        t = self._table()
        _lib.check(self._lib.cebag_evict_slot(ctypes.byref(t), int(slot.item()), _stream_ptr()))
        self._refresh_avail()
        self._cuda_to_cpu_numel += self.embedding_dim
        return int(slot.item())

    @torch.no_grad()
    def _admit(self, row_id: int):
        slot = self._find_free_cuda_row()
        if slot == -1:
            slot = self._evict()
        t = self._table()
        _lib.check(self._lib.cebag_admit_row(ctypes.byref(t), int(row_id), slot, _stream_ptr()))
        self._refresh_avail()
        self._cpu_to_cuda_numel += self.embedding_dim

    # ---- statistics ------------------------------------------------------------------------------------------------------------
    def print_comm_stats(self):
        self._harvest()
        elapsed = max(self._elapsed_dict["cache_op"], 1e-12)
        mb_out = self._cuda_to_cpu_numel * self.elem_size_in_byte / 1e6
        mb_in = self._cpu_to_cuda_numel * self.elem_size_in_byte / 1e6
        print(f"CUDA->CPU BWD {mb_out / elapsed:.1f} MB/s {mb_out / 1e3:.3f} GB moved in {elapsed:.3f} s of cache_op")
        print(f"CPU->CUDA BWD {mb_in / elapsed:.1f} MB/s {mb_in / 1e3:.3f} GB moved in {elapsed:.3f} s of cache_op")
        for k, v in self._elapsed_dict.items():
            print(f"{k}: {v:.4f} s")
        if self._total_cache:
            print(f"cache miss ratio {self._cache_miss / self._total_cache:.4f}")

    def extra_repr(self) -> str:
        return (f"rows={self.num_embeddings}, dim={self.embedding_dim}, slots={self.cuda_row_num}, "
                f"evict={self._evict_strategy.name}")
