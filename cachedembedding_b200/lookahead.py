"""Look-ahead driver: the cache operation of window k+1 overlaps the forward/backward of window k.

The reference gathers `prefetch_num` batches, runs ONE prepare_ids over their concatenated ids and only then steps
through them (/root/reference/recsys/dlrm_main.py:245-266); every cache operation is serialised with compute (its
timers end in torch.cuda.synchronize()).  Here three streams work at once:

  compute stream : forward / backward of window k (the caller's current stream)
  side stream    : prepare_ids of window k+1 -- id->slot probe, victim selection, map commit, LFU update -- and, if
                   asked, the gradient-independent half of each batch's fused backward (radix sort by slot)
  copy stream    : the PCIe row traffic of that prepare_ids (write-back of victims, fill of missed rows), which only the
                   forward of window k+1 has to wait for

Hazards and how they are closed:
  * rows the in-flight window k still reads/updates must not be evicted by prepare(k+1): the manager protects the
    slots stamped by the last TWO windows (`protect_windows = 2`); the capacity rule becomes
    |rows(k) U rows(k+1)| <= cuda_row_num, checked before anything is changed;
  * a victim may have been updated by window k-1's backward: the stream that moves rows (the copy stream, ordered
    after the side stream's map commit) waits for the event recorded after window k-1's compute was enqueued; the
    map-only head of prepare_ids does not touch rows and is not held back;
  * slot ids (side stream) and rows (copy stream) are consumed on the compute stream: `PrefetchHandle.wait()` makes it
    wait for both completion events;
  * backward-plan buffers are a ring of two windows owned by this object: the buffers of window k-1 are reused for
    window k+1 only after the same fence.
Pooled sums and updated rows are unaffected by which victims are chosen (the cache is transparent); the slot maps
follow the oracle run with the same two-window protection (tests/test_gpu_parity.py).
"""
from __future__ import annotations

from collections import deque
from typing import Optional

import torch


class PrefetchHandle:
    def __init__(self, slot_ids: torch.Tensor, done: torch.cuda.Event, rows_done: Optional[torch.cuda.Event]):
        self._slot_ids = slot_ids
        self._done = done
        self._rows_done = rows_done

    def wait(self) -> torch.Tensor:
        """Slot ids of the window; the current stream waits (on the device) for the cache operation to finish."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._done)
        if self._rows_done is not None:
            cur.wait_event(self._rows_done)
        self._slot_ids.record_stream(cur)
        return self._slot_ids


class LookaheadPrefetcher:
    """
    pf = LookaheadPrefetcher(bag)
    h = pf.submit(ids_of_window_0)
    for k in range(num_windows):
        slot_ids = h.wait()
        ... forward / backward of the window's batches with bag.set_cache_op(False) ...
        pf.window_enqueued()
        if k + 1 < num_windows:
            h = pf.submit(ids_of_window_k_plus_1)      # overlaps the work just enqueued
    pf.close()
    """

    def __init__(self, bag_or_mgr, priority: int = -1, copy_stream: bool = True):
        self.bag = bag_or_mgr if hasattr(bag_or_mgr, "cache_weight_mgr") else None
        self.mgr = getattr(bag_or_mgr, "cache_weight_mgr", bag_or_mgr)
        self.device = self.mgr.device
        self.stream = torch.cuda.Stream(device=self.device, priority=priority)
        self.copy_stream = torch.cuda.Stream(device=self.device, priority=priority) if copy_stream else None
        self._fences = deque(maxlen=2)     # events after the compute of the last two windows
        self._saved_protect = self.mgr.protect_windows
        self.mgr.protect_windows = max(2, self.mgr.protect_windows)
        self._plan_ring = [[], []]         # backward-plan workspaces of the even / odd windows
        self._window = 0

    def _plan_buffer(self, parity: int, j: int, nbytes: int) -> torch.Tensor:
        ring = self._plan_ring[parity]
        while len(ring) <= j:
            ring.append(None)
        if ring[j] is None or ring[j].numel() < nbytes:
            ring[j] = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        return ring[j]

    def submit(self, ids, ready: Optional[torch.cuda.Event] = None, offsets=None, layout="bag_major",
               layout_batch=0) -> PrefetchHandle:
        """Enqueue prepare_ids(ids) on the side stream.  `ids` is a tensor or a list of tensors (the batches of the
        window, concatenated here like recsys/dlrm_main.py:259 does); they may live in (pinned) host memory, in which
        case the H2D copies run on the side stream as well.  With `offsets` (one tensor for all batches, or one per
        batch) the fused backward of every batch is planned on the side stream too (CachedEmbeddingBag.plan_backward).

        Device ids must be complete when the side stream starts reading them.  Pass `ready` = an event recorded right
        after they were produced; without it nothing is waited for (recording an event here would be too late: the
        current stream already holds the whole window that this call is supposed to overlap)."""
        side = self.stream
        if ready is not None:
            side.wait_event(ready)
        # window k-1 must have finished before its rows can be written back (its updates have to be in them) and
        # before its plan buffers are recycled.  Only the ROW COPIES need that: with a copy stream the map-only head
        # of prepare_ids (probe, victim selection, commit) starts at once and just the copy stream is fenced.
        fence = self._fences[0] if len(self._fences) == 2 else None
        rows_done = None
        if self.copy_stream is not None:
            if fence is not None:
                self.copy_stream.wait_event(fence)
            rows_done = torch.cuda.Event()
            rows_done.record(self.copy_stream)     # instantiates the event; re-recorded after the row copies
            self.mgr._copy_stream, self.mgr._copy_done = self.copy_stream, rows_done
        elif fence is not None:
            side.wait_event(fence)
        parity = self._window & 1
        self._window += 1
        try:
            with torch.cuda.stream(side):
                parts = ids if isinstance(ids, (list, tuple)) else [ids]
                parts_dev = [t.to(self.device, non_blocking=True) for t in parts]
                ids_dev = parts_dev[0] if len(parts_dev) == 1 else torch.cat(parts_dev)
                slot_ids = self.mgr.prepare_ids(ids_dev)
                if offsets is not None and self.bag is not None:
                    # the gradient-independent half of every batch's fused backward also runs here, off the critical
                    # path; splitting by the batches' own sizes gives the views the training loop passes to forward
                    if fence is not None and self.copy_stream is not None:
                        side.wait_event(fence)     # plan buffers of window k-1 are free again
                    offs = offsets if isinstance(offsets, (list, tuple)) else [offsets] * len(parts)
                    for j, (chunk, off) in enumerate(zip(torch.split(slot_ids, [t.numel() for t in parts]), offs)):
                        self.bag.plan_backward(chunk, off, layout, layout_batch,
                                               workspace_factory=lambda n, p=parity, j=j: self._plan_buffer(p, j, n))
                done = torch.cuda.Event()
                done.record(side)
        finally:
            self.mgr._copy_stream, self.mgr._copy_done = None, None
        for t in parts:
            if t.is_cuda:
                t.record_stream(side)
        return PrefetchHandle(slot_ids, done, rows_done)

    def window_enqueued(self):
        """Call after the forward/backward of the current window has been enqueued on the compute stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._fences.append(ev)

    def drain(self):
        """Wait for everything submitted so far; the driver stays usable (streams, plan buffers and protection kept)."""
        self.stream.synchronize()
        if self.copy_stream is not None:
            self.copy_stream.synchronize()
        self._fences.clear()

    def close(self):
        """Back to the reference's one-window protection (waits for the side and copy streams)."""
        self.stream.synchronize()
        if self.copy_stream is not None:
            self.copy_stream.synchronize()
        self.mgr.protect_windows = self._saved_protect
        self._plan_ring = [[], []]
