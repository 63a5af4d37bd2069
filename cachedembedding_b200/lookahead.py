"""Look-ahead driver: the cache operation of window k+1 overlaps the forward/backward of window k.

The reference gathers `prefetch_num` batches, runs ONE prepare_ids over their concatenated ids and only then steps
through them (/root/reference/recsys/dlrm_main.py:245-266); every cache operation is serialised with compute (its
timers end in torch.cuda.synchronize()).  Here three streams work at once:

  compute stream : forward / backward of window k (the caller's current stream)
  side stream    : prepare_ids of window k+1 -- id->slot probe, victim selection, map commit, LFU update, parking of
                   the victims in an HBM staging buffer -- and, if asked, the gradient-independent half of each
                   batch's fused backward (radix sort by slot)
  copy stream    : the PCIe row traffic of that prepare_ids (fill of the missed rows -- a zero-copy gather kernel --
                   then the zero-copy write-back of whatever the DMA path does not take); only the fill is waited for
                   by the forward of window k+1
  DMA stream     : the parked victims go D2H as ONE contiguous copy-engine transfer into a pinned ring, in parallel
                   with the fill, and host threads scatter them into the table (cache_mgr.dma_writeback); a row that
                   is missed again while its write-back is in flight is filled from the staging buffer
  ids stream     : `stage()` -- the H2D copy of a window's ids from (pinned) host memory, one window further ahead than
                   its prepare_ids and (after_last_fill=True) behind the fill of the window submitted last: at
                   Criteo-1TB a window is 109 MB of int64 ids = 2 ms of PCIe, and the chain ids H2D -> map work -> fill
                   of one window does not fit under the previous window's compute when all three run back to back
                   (end to end 0.65 -> 0.56 ms per step); without the hold-back the early copy competes with the fill
                   for the PCIe read direction and is slower than no staging

prepare_ids never waits for the GPU (cebag_prepare_ids_async), so `submit` costs the host a few dozen kernel launches
and may be called anywhere inside window k -- the earlier the better: right after the window's first step has been
enqueued.

Hazards and how they are closed:
  * rows the in-flight window k still reads/updates must not be evicted by prepare(k+1): the manager protects the
    slots stamped by the last TWO windows (`protect_windows = 2`); the capacity rule becomes
    |rows(k) U rows(k+1)| <= cuda_row_num, checked on the device before anything is changed;
  * a victim of prepare(k+1) may have been updated by window k-1's backward: the side stream waits for the event
    recorded after window k-1's compute was enqueued before it parks the victims (cebag_workspace.victims_ready_event);
    the map-only part of prepare_ids does not touch rows and is not held back;
  * slot ids (side stream) and rows (copy stream) are consumed on the compute stream: `PrefetchHandle.wait()` makes it
    wait for both completion events -- and makes the HOST wait for the call's result record, so that a window that
    does not fit the cache raises there, before any of its batches is enqueued;
  * slot-id and backward-plan buffers are rings of three windows owned by this object (stable addresses, so that a
    step can be replayed as a CUDA graph): the buffers of window k-2 are reused for window k+1, after k-2's fence.
Pooled sums and updated rows are unaffected by which victims are chosen (the cache is transparent); the slot maps
follow the oracle run with the same two-window protection (tests/test_gpu_parity.py).
"""
from __future__ import annotations

from typing import Optional

import torch


_RING = 3     # windows whose slot ids / backward plans exist at once: w (being prepared), w-1 (computing), w-2 (draining)


class PrefetchHandle:
    def __init__(self, mgr, slot_ids: torch.Tensor, done: torch.cuda.Event, rows_done: Optional[torch.cuda.Event],
                 deferred_errors: bool = False):
        self._mgr = mgr
        self._slot_ids = slot_ids
        self._done = done
        self._rows_done = rows_done
        self._deferred = deferred_errors

    def wait(self) -> torch.Tensor:
        """Slot ids of the window.  The host waits for the result record of the call (a few bytes written by its last
        map kernel; raises CacheCapacityError / IndexError if the device rejected the window); the current stream
        waits, on the device, for the slot ids and for the missed rows.
        With `deferred_errors` the host does not wait: records that are already there are read, and a rejected window
        raises from a later wait() / drain() (its slot ids are -1, which forward and backward skip, and the table is
        untouched) -- the host keeps its lead over the GPU, which matters when a step is only a graph launch."""
        self._mgr._harvest(block=not self._deferred)
        cur = torch.cuda.current_stream()
        cur.wait_event(self._done)
        if self._rows_done is not None:
            cur.wait_event(self._rows_done)
        self._slot_ids.record_stream(cur)
        return self._slot_ids


class StagedIds:
    """Ids of one window on their way to the device (LookaheadPrefetcher.stage): `tensor` is the concatenation of the
    window's batches in a ring buffer owned by the driver, `sizes` the batches' lengths, `ready` the event recorded
    after the last H2D copy; `done` is set by submit() (the ring buffer is recycled after it)."""

    def __init__(self, tensor: torch.Tensor, sizes, ready: torch.cuda.Event):
        self.tensor, self.sizes, self.ready = tensor, list(sizes), ready
        self.done: Optional[torch.cuda.Event] = None


class LookaheadPrefetcher:
    """
    pf = LookaheadPrefetcher(bag)
    h = pf.submit(ids_of_window_0)
    for k in range(num_windows):
        slot_ids = h.wait()
        ... first forward / backward of the window, with bag.set_cache_op(False) ...
        if k + 1 < num_windows:
            h = pf.submit(ids_of_window_k_plus_1)      # overlaps this window's compute
        ... the other forward / backward steps of the window ...
        pf.window_enqueued()
    pf.close()
    (Submitting after `window_enqueued()` -- the round-1 order -- still works; it just starts the overlap later.)
    Ids that start in host memory can be sent ahead with `staged = pf.stage(ids_of_window_k_plus_2)` (own stream, no
    effect on the cache) and handed to `pf.submit(staged)` one window later.
    """

    def __init__(self, bag_or_mgr, priority: int = -1, copy_stream: bool = True, deferred_errors: bool = False):
        self.bag = bag_or_mgr if hasattr(bag_or_mgr, "cache_weight_mgr") else None
        self.mgr = getattr(bag_or_mgr, "cache_weight_mgr", bag_or_mgr)
        self.device = self.mgr.device
        self.stream = torch.cuda.Stream(device=self.device, priority=priority)
        self.copy_stream = torch.cuda.Stream(device=self.device, priority=priority) if copy_stream else None
        self.deferred_errors = deferred_errors
        self.window_plan = None            # None: one sort per window if it fits L2, else one per batch; True / False force
        # True: the handle's completion event is recorded BEFORE the backward plans, so the window's first forward does
        # not wait for the sorts of all its batches; every plan carries its own event, which its backward waits for
        self.early_done = True
        self.trace = None                  # a list: submit() appends timing events of its side / copy stream work
        self._saved_protect = self.mgr.protect_windows
        self._saved_defer = self.mgr._defer_results
        self.mgr.protect_windows = max(2, self.mgr.protect_windows)
        self.mgr._defer_results = True
        self._plan_ring = [[] for _ in range(_RING)]      # backward-plan workspaces of window w % 3
        self._slot_ring = [None] * _RING                  # slot ids of window w % 3: stable addresses (CUDA graphs)
        self._ids_stream = None                           # H2D copies of stage()
        self._ids_ring = [None] * _RING                   # the last StagedIds of every ring buffer
        self._staged = 0
        self._last_rows_done = None                       # fill event of the most recent submit (stage(after_last_fill))
        self._fences = {}                  # window index -> event recorded after its compute was enqueued
        self._submitted = 0                # windows submitted since the last drain
        self._enqueued = 0                 # windows whose compute has been enqueued since the last drain

    def _plan_buffer(self, slot: int, j: int, nbytes: int) -> torch.Tensor:
        ring = self._plan_ring[slot]
        while len(ring) <= j:
            ring.append(None)
        if ring[j] is None or ring[j].numel() < nbytes:
            ring[j] = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        return ring[j]

    def _victims_fence(self, w: int) -> torch.cuda.Event:
        """What the row traffic of window w has to wait for: the compute of window w-2 (victims are never taken from the
        last two windows).  Before the driver has seen two windows enqueued -- at start, after drain(), or when the
        caller does not report its windows -- an event recorded now on the current stream stands in: whatever ran on
        the cache before this driver took over is then finished as well."""
        ev = self._fences.get(w - 2)
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        return ev

    def stage(self, ids, after_last_fill: bool = False) -> StagedIds:
        """Start the H2D copies of a window's ids (a tensor or the list of the window's batches, normally in pinned host
        memory) on the driver's ids stream and return at once.  Nothing of the cache is touched: call it as early as
        the host has the batches -- one window before `submit(staged)` is what it takes to get the copy out of the
        window's critical chain.  The device buffer is one of three ring buffers; it is recycled once the submit that
        consumed its previous content has finished on the side stream.
        `after_last_fill`: the copies wait (on the device) for the fill of the most recently submitted window -- the
        missed rows and the ids share the PCIe read direction, and the fill is what the next forward waits for."""
        parts = ids if isinstance(ids, (list, tuple)) else [ids]
        sizes = [t.numel() for t in parts]
        total = sum(sizes)
        if self._ids_stream is None:
            self._ids_stream = torch.cuda.Stream(device=self.device)
        h2d = self._ids_stream
        k = self._staged % _RING
        self._staged += 1
        prev = self._ids_ring[k]
        with torch.cuda.stream(h2d):
            if after_last_fill and self._last_rows_done is not None:
                h2d.wait_event(self._last_rows_done)
            if prev is not None and prev.done is not None and prev.tensor.numel() == total:
                h2d.wait_event(prev.done)            # the readers of the old content: that window's prepare_ids
                buf = prev.tensor
            else:
                # first lap, another window size, or a staged window that was never submitted: a fresh buffer (the old
                # one goes back to the allocator, which holds it until the side stream has passed -- record_stream)
                buf = torch.empty(total, dtype=torch.long, device=self.device)
            off = 0
            for t, n in zip(parts, sizes):
                buf[off:off + n].copy_(t.reshape(-1), non_blocking=True)
                off += n
            ready = torch.cuda.Event()
            ready.record(h2d)
        staged = StagedIds(buf, sizes, ready)
        self._ids_ring[k] = staged
        return staged

    def submit(self, ids, ready: Optional[torch.cuda.Event] = None, offsets=None, layout="bag_major",
               layout_batch=0) -> PrefetchHandle:
        """Enqueue prepare_ids(ids) on the side stream.  `ids` is a tensor or a list of tensors (the batches of the
        window, concatenated here like recsys/dlrm_main.py:259 does); they may live in (pinned) host memory, in which
        case the H2D copies run on the side stream as well.  With `offsets` (one tensor for all batches, or one per
        batch) the fused backward of every batch is planned on the side stream too (CachedEmbeddingBag.plan_backward).

        Device ids must be complete when the side stream starts reading them.  Pass `ready` = an event recorded right
        after they were produced; without it nothing is waited for (recording an event here would be too late: the
        current stream already holds the window that this call is supposed to overlap)."""
        side = self.stream
        mgr = self.mgr
        staged = ids if isinstance(ids, StagedIds) else None
        if staged is not None:
            side.wait_event(staged.ready)
            staged.tensor.record_stream(side)
        if ready is not None:
            side.wait_event(ready)
        w = self._submitted
        self._submitted += 1
        fence = self._victims_fence(w)
        slot = w % _RING
        mgr._copy_stream, mgr._victims_ready = self.copy_stream, fence
        if self.copy_stream is None:
            side.wait_event(fence)
        tr = None
        if self.trace is not None:
            tr = {"window": w}
            self.trace.append(tr)

        def mark(name, stream):
            if tr is not None:
                tr[name] = torch.cuda.Event(enable_timing=True)
                tr[name].record(stream)
        try:
            with torch.cuda.stream(side):
                mark("side_start", side)
                if staged is not None:
                    parts = list(torch.split(staged.tensor, staged.sizes))
                    ids_dev = staged.tensor
                else:
                    parts = ids if isinstance(ids, (list, tuple)) else [ids]
                    parts_dev = [t.to(self.device, non_blocking=True) for t in parts]
                    ids_dev = parts_dev[0] if len(parts_dev) == 1 else torch.cat(parts_dev)
                # the slot ids of window w live in a ring buffer that is reused for window w+3: stable addresses (a
                # CUDA graph per (buffer, batch) can be replayed), and the readers of the old content -- the steps of
                # window w-3 -- finished a whole window ago, so waiting for their fence costs nothing
                ring = self._slot_ring[slot]
                if ring is None or ring.numel() != ids_dev.numel():
                    ring = self._slot_ring[slot] = torch.empty_like(ids_dev)
                elif self._fences.get(w - _RING) is not None:
                    side.wait_event(self._fences[w - _RING])
                slot_ids = mgr.prepare_ids(ids_dev, out=ring)
                rows_done = mgr._rows_ready
                mgr._rows_ready = None             # the handle carries it; forward() of the bag need not wait again
                self._last_rows_done = rows_done
                done = torch.cuda.Event()
                if self.early_done:
                    done.record(side)
                mark("prepared", side)
                if tr is not None and rows_done is not None:
                    tr["filled"] = rows_done
                if self.copy_stream is not None:
                    mark("copied", self.copy_stream)      # fill + write-back of this window's rows
                if offsets is not None and self.bag is not None:
                    # the gradient-independent half of every batch's fused backward also runs here, off the critical
                    # path (the side stream has passed the fence by now: the plan buffers of window w-2 are free);
                    # splitting by the batches' own sizes gives the views the training loop passes to forward
                    offs = offsets if isinstance(offsets, (list, tuple)) else [offsets] * len(parts)
                    self.bag.drop_backward_plans(slot)
                    chunks = list(torch.split(slot_ids, [t.numel() for t in parts]))
                    if self.window_plan is None:
                        # ONE radix sort for the whole window (key = batch index above the slot id) while its four
                        # arrays still fit the 126 MB L2; beyond that the window sort streams through HBM, which the
                        # fwd/bwd kernels it overlaps are bound by, and the L2-resident per-batch sorts win
                        # (Criteo-1TB on one GPU, 13.6 M lookups per window: 0.652 vs 0.625 ms per step)
                        use_window = len(parts) > 1 and slot_ids.numel() * 16 <= 64 << 20
                    else:
                        use_window = bool(self.window_plan)
                    if use_window:
                        self.bag.plan_backward_window(
                            chunks, offs, layout, layout_batch, tag=slot,
                            window_factory=lambda n, p=slot: self._plan_buffer(p, len(parts), n),
                            scratch_factory=lambda j, n, p=slot: self._plan_buffer(p, j, n))
                    else:
                        for j, (chunk, off) in enumerate(zip(chunks, offs)):
                            self.bag.plan_backward(chunk, off, layout, layout_batch, tag=slot,
                                                   workspace_factory=lambda n, p=slot, j=j: self._plan_buffer(p, j, n))
                if not self.early_done:
                    done.record(side)
                mark("planned", side)
        finally:
            mgr._copy_stream, mgr._victims_ready = None, None
        if staged is not None:
            staged.done = done
        else:
            for t in parts:
                if t.is_cuda:
                    t.record_stream(side)
        return PrefetchHandle(mgr, slot_ids, done, rows_done, self.deferred_errors)

    def window_enqueued(self):
        """Call after the forward/backward of the current window has been enqueued on the compute stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._fences[self._enqueued] = ev
        self._fences.pop(self._enqueued - 4, None)
        self._enqueued += 1

    def drain(self):
        """Wait for everything submitted so far -- side stream, copy stream and the compute the fences stand for; the
        driver stays usable (streams, plan buffers and protection kept)."""
        if self._ids_stream is not None:
            self._ids_stream.synchronize()
        self.stream.synchronize()
        if self.copy_stream is not None:
            self.copy_stream.synchronize()
        if self.mgr._dma_stream is not None:
            self.mgr._dma_stream.synchronize()
        for ev in self._fences.values():
            ev.synchronize()
        self.mgr._harvest()
        self._fences.clear()
        self._submitted = self._enqueued = 0

    def close(self):
        """Back to the reference's one-window protection (waits like drain())."""
        self.drain()
        self.mgr.protect_windows = self._saved_protect
        self.mgr._defer_results = self._saved_defer
        if self.bag is not None:
            self.bag.drop_backward_plans()
        self._plan_ring = [[] for _ in range(_RING)]
        self._slot_ring = [None] * _RING
        self._ids_ring = [None] * _RING
