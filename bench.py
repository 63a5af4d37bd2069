#!/usr/bin/env python
"""bench.py -- embedding lookups/s (fwd+bwd) of the cached embedding bag on synthetic Criteo-1TB-shape ids.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it is launched under
torch.distributed.run, one rank per GPU.  Rank 0 prints ONE JSON line.

A "step" is one batch through the hot path: forward (segment-sum gather over the slot cache) + backward fused with
the SGD update of the cached rows, plus -- once every `prefetch_num` steps -- the cache manager's prepare_ids over the
whole look-ahead window (id->slot lookup, LFU update, evict/admit, D2H write-back and H2D fill of rows).  This is the
loop of /root/reference/recsys/dlrm_main.py:235-279 restricted to the embedding operator, i.e. the isolation harness of
/root/reference/benchmark/benchmark_cache.py:58-72 plus the optimizer step.

Workload at N = 1: BASELINE.json configs[2] -- Criteo-1TB DLRM shape (26 tables, 177,944,275 rows, dim 128, the full
91.1 GB fp32 table pinned in host DRAM), cache_ratio 0.01, prefetch_num 8, batch 65536 -- whenever the host has the RAM
for it; otherwise the rows are scaled down and `config.row_scale` says by how much.
N > 1: the same tables sharded table-wise over the ranks (BASELINE.json configs[3]), global batch 65536 ("strong"
scaling: total work is fixed); the all-to-all of pooled embeddings and of their gradients is fused into the forward /
backward kernels over NVLink peer memory (--no-fused-exchange: NCCL all-to-all, the reference's way).
By default the cache operation of window k+1 runs on side streams under the compute of window k (look-ahead driver:
window k+1 is submitted right after the FIRST step of window k; --no-overlap for the reference's serial order).
N > 1 first runs an untimed parity leg (bench_verify.py: fused NVLink exchange vs NCCL exchange bit for bit, slot maps,
pooled sums and tables against the CPU oracle, table-wise and column-wise) and fails the run on a mismatch; the record
is the line's `parity_check`.  `--parallelism column` times the column-wise bag instead of the table-wise one.
End-to-end arm (`e2e`): the ids of every batch start in pinned host memory; on one GPU the ids of window k+2 are staged
H2D on their own stream behind the fill of window k+1 (`--stage-ids`, LookaheadPrefetcher.stage), one pooled row per
step is read back.  `--ab` / `--e2e-ab` / `--trace-steps` add interleaved in-process A/B samples of launch-shape
settings, of the ids-H2D orders, and a timeline of the pipeline to the line / stderr.
`--impl reference` times the CPU oracle port on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

# /root/reference/recsys/datasets/criteo.py:29-34 (table cardinalities; constants, not code)
CRITEO_1TB_ROWS = [45833188, 36746, 17245, 7413, 20243, 3, 7114, 1441, 62, 29275261, 1572176, 345138, 10, 2209, 11267,
                   128, 4, 974, 14, 48937457, 11316796, 40094537, 452104, 12606, 104, 35]
CRITEO_KAGGLE_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                      5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
# table -> rank maps of /root/reference/recsys/utils/misc.py:198-206 (criteo 1TB, world 2 and 4); world 8 has no map
# in the reference, ours is greedy by rows
REF_RANK_ARRANGE_1TB = {
    1: [0] * 26,
    2: [1, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0],
    4: [1, 3, 3, 3, 3, 0, 2, 2, 1, 2, 2, 2, 0, 1, 2, 1, 0, 1, 0, 0, 2, 3, 3, 3, 1, 0],
}
SKEW = 0.25        # long-tail exponent of the reference's generator (/root/reference/baselines/data/custom.py:23)
SEED = 1024        # default --seed of the reference (/root/reference/recsys/dlrm_main.py:139-144)

WORKLOADS = {
    "criteo1tb": dict(rows=CRITEO_1TB_ROWS, dim=128, batch=65536, prefetch=8, cache_ratio=0.01),
    "kaggle": dict(rows=CRITEO_KAGGLE_ROWS, dim=128, batch=4096, prefetch=8, cache_ratio=0.01),
    "plumbing": dict(rows=[100000] * 26, dim=16, batch=512, prefetch=1, cache_ratio=0.05),
    # BASELINE.json configs[4]: one 1e9-row table (512 GB at dim 128 -- rows are scaled to the host's RAM, see
    # config.row_scale); choose the slot fraction with --cache-ratio (0.001 / 0.01 / 0.1)
    "zipf1e9": dict(rows=[1_000_000_000], dim=128, batch=65536 * 26, prefetch=8, cache_ratio=0.01),
}


def rank_arrange(rows, world, placement="auto"):
    """table -> rank.  The reference's hard-coded maps where it has them (Criteo-1TB, world 2 and 4,
    /root/reference/recsys/utils/misc.py:198-206); otherwise (it has none for world 8) a snake over the tables sorted by
    rows: every rank gets ceil(F / world) or floor(F / world) tables -- lookups per step are equal per table, so this
    balances kernel time -- and one of the `world` largest tables each, which balances host memory and cache demand."""
    if placement == "auto" and world in REF_RANK_ARRANGE_1TB and len(rows) == 26:
        return REF_RANK_ARRANGE_1TB[world]
    order = sorted(range(len(rows)), key=lambda i: -rows[i])
    arrange = [0] * len(rows)
    for pos, t in enumerate(order):
        lap, k = divmod(pos, world)
        arrange[t] = k if lap % 2 == 0 else world - 1 - k
    return arrange


def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def sample_ids(rows_dev, batch, gen, device):
    """One batch of KJT-ordered global ids: values[f*B + b], one id per (feature, sample), per table the reference
    generator's long tail (cachedembedding_b200/synth_criteo.py; /root/reference/baselines/data/custom.py:76,89-91)."""
    from cachedembedding_b200.synth_criteo import sample_ids as _sample
    return _sample(rows_dev, batch, gen, device, SKEW)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_info():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, local, world


# ---------------------------------------------------------------------------------------------------------------- b200
def run_b200(args):
    import torch.distributed as dist
    import cachedembedding_b200 as ce
    from cachedembedding_b200 import _lib
    from cachedembedding_b200.collectives import split_sizes

    rank, local, world = dist_info()
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.stage_ids == "auto":
        args.stage_ids = "after-fill" if world == 1 else "off"
    wl = dict(WORKLOADS[args.workload])
    if args.cache_ratio > 0:
        wl["cache_ratio"] = args.cache_ratio
    if args.prefetch > 0:
        wl["prefetch"] = args.prefetch
    rows_all = list(wl["rows"])
    D, B, P = wl["dim"], wl["batch"], wl["prefetch"]
    column = args.parallelism == "column" and world > 1
    arrange = rank_arrange(rows_all, world, args.placement)

    # ---- untimed parity leg (N > 1): the parallel bags against the CPU oracle on small tables ------------------------
    parity = None
    if world > 1 and not args.no_verify:
        from bench_verify import verify_parallel
        t0 = time.time()
        parity = verify_parallel([300 + 37 * (t % 5) for t in range(len(rows_all))], arrange, D)
        parity["seconds"] = round(time.time() - t0, 1)
        if not parity["ok"] or args.verify_only:
            if rank == 0:
                print(json.dumps({"metric": "embedding lookups/sec (fwd+bwd)", "value": None, "n_gpus": world,
                                  "parity_check": parity}))
            dist.destroy_process_group()
            sys.exit(0 if parity["ok"] else 1)

    # every rank's tables must fit the host: scale rows down only if they do not
    need_gb = sum(rows_all) * D * 4 / 1e9
    avail_gb = host_mem_available_gb()
    row_scale = args.row_scale
    if row_scale == 0:
        row_scale = 1
        while need_gb / row_scale > 0.8 * avail_gb and row_scale < 1024:
            row_scale *= 2
    rows_all = [max(1, r // row_scale) for r in rows_all]
    # column-wise: every rank holds ALL rows (a D / W column slice of each) and looks up ALL ids of the global batch
    my_tables = list(range(len(rows_all))) if column else [t for t, r in enumerate(arrange) if r == rank]
    rows_loc = [rows_all[t] for t in my_tables]
    F, F_loc = len(rows_all), len(my_tables)
    N_loc = sum(rows_loc)
    # HBM slot budget = cache_ratio of the UNSCALED table (a window of full batches must still fit when rows are scaled
    # down).  With table-wise sharding the budget is split evenly over the ranks: the reference sizes each rank's
    # cache as cache_ratio x its local rows, which leaves a rank that holds only small tables with fewer slots than
    # one look-ahead window of a 65536 batch touches (its own capacity assert fires); demand per rank is set by the
    # number of tables, not by their rows.
    total_slots = max(int(sum(wl["rows"]) * wl["cache_ratio"]), 1)
    C_loc = min(N_loc, total_slots if column else total_slots // world)
    K, W = args.steps, args.warmup
    total_steps = W + K
    windows = (total_steps + P - 1) // P + 2      # + the windows the last timed one prefetches

    gen = torch.Generator(device=dev).manual_seed(SEED + (0 if column else rank))
    rows_dev = torch.tensor(rows_loc, dtype=torch.long, device=dev)

    # id frequencies counted over a sample of the synthetic "dataset" (the reference counts its training set:
    # /root/reference/recsys/datasets/feature_counter.py:21-29) -> LFU warm start
    t0 = time.time()
    counter = ce.IdFrequencyCounter(N_loc, dev)          # GPU histogram kernel (cebag_id_histogram)
    for _ in range(args.freq_batches):
        counter.update(sample_ids(rows_dev, B, gen, dev))
    freq = counter.result()
    del counter
    common = dict(sparse=True, mode="sum", include_last_offset=True, cache_ratio=wl["cache_ratio"],
                  warmup_ratio=args.warmup_ratio,
                  evict_strategy=ce.EvictionStrategy.LFU, cuda_row_num=C_loc, init_seed=SEED, fused_optimizer="sgd",
                  lr=1.0)
    if world == 1:
        model = ce.CachedEmbeddingBag(N_loc, D, ids_freq_mapping=freq, **common)
    elif column:
        # the reference's default module (recsys/models/dlrm.py:70-81)
        model = ce.ParallelCachedEmbeddingBag(N_loc, D, ids_freq_mapping=freq, **common)
    else:
        # the reference's table-wise module (recsys/models/dlrm.py:53-68): every table with its rank and frequencies
        table_freq = dict(zip(my_tables, torch.split(freq, rows_loc)))
        cfgs = [ce.TablewiseEmbeddingBagConfig(rows_all[t], 0, assigned_rank=arrange[t],
                                               ids_freq_mapping=table_freq.get(t, torch.zeros(1)))
                for t in range(len(rows_all))]
        model = ce.ParallelCachedEmbeddingBagTablewise(cfgs, D, **common)
        model.enable_fused_exchange(not args.no_fused_exchange)
    del freq
    mgr = model.cache_weight_mgr
    model.set_cache_op(False)
    setup_s = time.time() - t0

    n_b = F_loc * B                         # lookups per step on this rank (pooling factor 1)
    overlap = not args.no_overlap
    offsets = torch.arange(n_b + 1, dtype=torch.long, device=dev)
    # three arms (timed, per-kernel replay, end-to-end) each get their own fresh windows of ids
    arm_names = ("value", "profile", "e2e")
    arms = {a: [sample_ids(rows_dev, B, gen, dev) for _ in range(windows * P)] for a in arm_names}
    # the gradient of the pooled embeddings, fixed (benchmark_cache.py:64); for N > 1 it is what the dense part
    # returns for this rank's slice of the batch, all features
    strides = split_sizes(B, world)
    if world > 1:
        grad_full = torch.randn(strides[rank], F, D, device=dev) if column else torch.randn(strides[rank], F * D, device=dev)
    else:
        grad_full = torch.randn(n_b, D, device=dev)

    grad_holder = {"g": grad_full}
    # one look-ahead driver for the whole run: its streams and plan buffers are warmed once
    # N > 1: a step is one graph launch, so the host must not wait for every window's result record (it would lose its
    # lead over the GPU); the device-side verdict still protects the table and a rejected window raises one window later
    prefetcher = {"pf": ce.LookaheadPrefetcher(model, deferred_errors=world > 1) if overlap else None}
    col_hook = (lambda x: x.view(F, B, -1).transpose(0, 1)) if column else None     # recsys/models/dlrm.py:26-27

    graph_step = world > 1 and not column and not args.no_fused_exchange and not args.no_graph_step

    def embed_step(slots):
        # table-wise: (B / W, F * D) after the exchange; column-wise: (B / W, F, D); single GPU: (F * B, D)
        if graph_step:
            # forward + fused backward as one CUDA-graph launch (the gradient already sits in the exchange's buffer)
            return model.fused_step(slots, offsets)
        out = model(slots, offsets, shape_hook=col_hook) if column else model(slots, offsets)
        out.backward(grad_holder["g"])
        return out

    class Runner:
        """Steps of one arm, in order.  With the look-ahead driver, window w+1 is submitted to the side streams right
        after the FIRST step of window w has been enqueued (prepare_ids never waits for the GPU, so this costs the host
        a few dozen launches), i.e. every window's cache operation has the rest of the previous window to hide under,
        wherever the timed region starts; every timed window still submits exactly one prepare_ids."""

        def __init__(self, batches, host_inputs, overlap=overlap, stage_ids=None):
            self.batches, self.host, self.overlap = batches, host_inputs, overlap
            self.stage_ids = args.stage_ids if stage_ids is None else stage_ids        # "off" | "early" | "after-fill"
            self.submit_before = args.submit_before       # window w+1 goes out before (not after) the first step of w
            self.w, self.slots, self.handles, self.staged, self.h2d = -1, None, {}, {}, 0
            self.trace = []
            self.plan = dict(offsets=offsets) if not args.no_plan_side else {}

        def ids(self, w):
            return self.batches[w * P:(w + 1) * P]

        def stage(self, w):
            """Host ids only: their H2D copies start one window before the window's prepare_ids (own stream)."""
            if (self.host and self.stage_ids != "off" and w not in self.staged and w not in self.handles
                    and (w + 1) * P <= len(self.batches)):
                self.staged[w] = prefetcher["pf"].stage(self.ids(w), after_last_fill=self.stage_ids == "after-fill")
                self.h2d += P * n_b * 8

        def submit(self, w):
            if w not in self.handles and (w + 1) * P <= len(self.batches):
                self.stage(w)
                if w in self.staged:
                    src = self.staged.pop(w)
                else:
                    src = self.ids(w)
                    self.h2d += P * n_b * 8 if self.host else 0
                self.handles[w] = prefetcher["pf"].submit(src, **self.plan)

        def run(self, first, count):
            """Steps [first, first+count).  host inputs: ids start in pinned host memory and are copied H2D inside the
            region (every batch of a window before its prepare_ids, like recsys/dlrm_main.py:248-259); one pooled row
            is read back D2H per step."""
            self.h2d = d2h = 0
            pf = prefetcher["pf"] if self.overlap else None
            saved_protect = mgr.protect_windows
            if not self.overlap:
                mgr.protect_windows = 1          # reference order: only the current window is protected
            for s in range(first, first + count):
                w, j = divmod(s, P)
                if w != self.w:                  # entering a new window
                    if self.overlap:
                        self.submit(w)           # only the very first window of an arm is not in flight already
                        self.slots = torch.chunk(self.handles.pop(w).wait(), P)
                        if self.submit_before:
                            self.submit(w + 1)
                            self.stage(w + 2)
                    else:
                        win = torch.cat([b.to(dev, non_blocking=True) for b in self.ids(w)])
                        self.h2d += P * n_b * 8 if self.host else 0
                        self.slots = torch.chunk(mgr.prepare_ids(win), P)
                    self.w = w
                out = embed_step(self.slots[j])
                if args.trace_steps:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                    self.trace.append((s, ev, time.perf_counter()))
                if self.host:
                    result_host.copy_(out.view(-1)[:D], non_blocking=True)
                    d2h += D * 4
                if self.overlap and j == 0 and not self.submit_before:
                    self.submit(w + 1)
                    self.stage(w + 2)
                if self.overlap and j == P - 1:
                    pf.window_enqueued()
            mgr.protect_windows = saved_protect
            return self.h2d, d2h

        def finish(self):
            if self.overlap:
                prefetcher["pf"].drain()
                self.handles.clear()
                self.staged.clear()

    def timed(runner, first, count):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        h2d, d2h = runner.run(first, count)
        e1.record()
        if args.trace_steps and rank == 0 and runner.trace:
            torch.cuda.synchronize()
            s0, ev0, h0 = [t for t in runner.trace if t[0] >= first][0]
            print(f"first timed step {s0}: gpu +{e0.elapsed_time(ev0):.3f} ms after the start event, host "
                  f"+{(h0 - t_host) * 1e3:.3f} ms; whole region {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), h2d, d2h

    # ---- device-resident arm: warm-up, then exactly K timed steps -------------------------------------------------
    # warm-up is rounded to whole windows internally only for the *slot* bookkeeping: steps W..W+K-1 are timed.
    # two untimed windows on scratch ids before anything is measured: first-touch costs of the caching allocator
    # (cross-stream buffers of the look-ahead driver) and of the lazily created streams/events
    if world > 1 and not column and not args.no_fused_exchange:
        # the "dense part" leaves its gradient where the fused backward reads it (no staging copy per step)
        g = model._exchange_for(B).grad_tensor()
        g.copy_(grad_full)
        grad_holder["g"] = g
    # three windows: every (ring buffer, batch) pair has been seen once (CUDA graphs of the multi-GPU step captured)
    scratch = Runner([sample_ids(rows_dev, B, gen, dev) for _ in range(4 * P)], False)
    scratch.run(0, 3 * P)
    scratch.finish()
    del scratch
    # clocks are sampled from the warm-up to the end of the end-to-end arm: every arm runs the same steps, and the
    # K timed steps alone are shorter than nvidia-smi's sampling period
    sampler = ClockSampler(local) if rank == 0 else None
    value_runner = Runner(arms["value"], False)
    if args.trace_steps and prefetcher["pf"] is not None:
        prefetcher["pf"].trace = []
    value_runner.run(0, W)
    launches0 = _lib.launch_count() + getattr(model, "graph_launches", 0)
    hist0 = len(mgr.num_miss_history)
    ms_total, _, _ = timed(value_runner, W, K)
    value_runner.finish()
    gpu_launches = _lib.launch_count() + getattr(model, "graph_launches", 0) - launches0
    miss_u = sum(mgr.num_miss_history[hist0:])
    hit_u = sum(mgr.num_hits_history[hist0:])
    evicted = sum(mgr.num_write_back_history[hist0:])
    miss_ratio_lookups = mgr._cache_miss / max(mgr._total_cache, 1)
    if args.trace_steps and rank == 0:
        tr = [t for t in value_runner.trace if t[0] >= W]
        for (s0, e0, h0), (s1, e1, h1) in zip(tr[:-1], tr[1:]):
            print(f"step {s1}: gpu +{e0.elapsed_time(e1):.3f} ms, host +{(h1 - h0) * 1e3:.3f} ms", file=sys.stderr)
        if prefetcher["pf"] is not None and prefetcher["pf"].trace:
            # where the side / copy stream work of every submitted window sits between the steps (ms since the end of
            # the first timed step; step s ends at "step s" below)
            torch.cuda.synchronize()
            t0 = tr[0][1]
            print("step ends: " + " ".join(f"{s}:{t0.elapsed_time(e):.2f}" for s, e, _ in tr), file=sys.stderr)
            for rec in prefetcher["pf"].trace:
                try:
                    print("window %d: " % rec["window"] + " ".join(
                        f"{k}={t0.elapsed_time(rec[k]):.2f}" for k in ("side_start", "prepared", "planned", "filled", "copied")
                        if k in rec), file=sys.stderr)
                except Exception as exc:        # an event of a window outside the timed region
                    print("window %d: %s" % (rec["window"], exc), file=sys.stderr)
            prefetcher["pf"].trace = None
    # --ab: the same timed loop again under other settings (environment knobs the library reads per call, PRIORITY =
    # stream priority of the look-ahead driver), fresh ids each, same process and box: A/B records, not the headline
    ab = {}
    specs = [x for x in args.ab.split(";") if x] if overlap else []
    for rep in range(args.ab_reps if specs else 0):
        for spec in specs:      # interleaved: box drift hits every setting alike
            name, _, kv = spec.partition(":")
            env = dict(x.split("=") for x in kv.split(",") if x)
            prio = int(env.pop("PRIORITY", -1))
            early_done = env.pop("EARLY_DONE", None)
            submit_before = env.pop("SUBMIT_BEFORE", None)
            cprio = env.pop("COMPUTE_PRIORITY", None)
            dma = env.pop("DMA", None)                  # parked victims leave through the copy engine + host threads
            saved_dma = mgr.dma_writeback
            if dma is not None:
                mgr.dma_writeback = bool(int(dma))
            saved = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            main_pf = prefetcher["pf"]
            prefetcher["pf"] = ce.LookaheadPrefetcher(model, priority=prio, deferred_errors=world > 1)
            if early_done is not None:
                prefetcher["pf"].early_done = bool(int(early_done))
            # three untimed windows first: every ring buffer of the new driver has been allocated once
            r = Runner([sample_ids(rows_dev, B, gen, dev) for _ in range((windows + 3) * P)], False)
            if submit_before is not None:
                r.submit_before = bool(int(submit_before))
            torch.cuda.synchronize()
            # COMPUTE_PRIORITY: forward / backward on a stream of that priority instead of the default stream
            cstream = torch.cuda.Stream(priority=int(cprio)) if cprio is not None else torch.cuda.current_stream()
            with torch.cuda.stream(cstream):
                r.run(0, 3 * P + W)
                ab.setdefault(name, {"settings": kv, "ms_per_step": []})["ms_per_step"].append(
                    round(timed(r, 3 * P + W, K)[0] / K, 4))
                r.finish()
            torch.cuda.synchronize()
            prefetcher["pf"].close()
            prefetcher["pf"] = main_pf
            mgr.dma_writeback = saved_dma
            # the main driver's settings are back in force (close() restored what it found: two-window protection)
            mgr.protect_windows = max(2, mgr.protect_windows)
            mgr._defer_results = True
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    for rec in ab.values():
        xs = sorted(rec["ms_per_step"])
        rec["median"] = xs[len(xs) // 2]

    # ---- per-kernel timers on a replay of the same steps (CUDA events on the launching stream) ------------------
    # (look-ahead off for this replay: every kernel is alone on the GPU, so its event-bracketed time is its own)
    _lib.profile_enable(True)
    Runner(arms["profile"], False, overlap=False).run(W, K)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    _lib.profile_enable(False)

    # ---- end-to-end arm: ids come from pinned host memory, one pooled row goes back per step ---------------------
    host_batches = [b.cpu().pin_memory() for b in arms["e2e"]]
    result_host = torch.empty(D, dtype=torch.float32).pin_memory()
    e2e_ab = None
    if args.e2e_ab and overlap:      # both ids-H2D orders on the same box, interleaved, fresh ids (A/B record, not the headline)
        e2e_ab = {m: [] for m in args.e2e_ab_modes.split(",")}
        for rep in range(args.ab_reps):
            for mode in e2e_ab:
                hb = [sample_ids(rows_dev, B, gen, dev).cpu().pin_memory() for _ in range(windows * P)]
                r = Runner(hb, True, stage_ids=mode)
                r.run(0, W)
                e2e_ab[mode].append(round(timed(r, W, K)[0] / K, 4))
                r.finish()
                del r, hb
    e2e_runner = Runner(host_batches, True)
    e2e_runner.run(0, W)
    e2e_ms, h2d, d2h = timed(e2e_runner, W, K)
    e2e_runner.finish()
    clocks = sampler.stop() if sampler else None

    if prefetcher["pf"] is not None:
        prefetcher["pf"].close()
    lookups_per_step = B * F                                  # whole job, all ranks
    value = lookups_per_step * K / (ms_total / 1e3)
    e2e_value = lookups_per_step * K / (e2e_ms / 1e3)

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    uniq = [int(torch.unique(b).numel()) for b in arms["profile"][W:W + min(K, 8)]]
    u_avg = sum(uniq) / len(uniq)
    row_b = 4 * (D // world if column else D)
    # Bytes per launch.  "compulsory": what DRAM has to move at least -- ids, offsets, every UNIQUE
    # row once (a batch re-reads its hot rows from the 126 MB L2), every output / gradient row once.  "algorithmic":
    # SURVEY.md section 8d's per-lookup figure, which charges a full row read to every lookup.
    compulsory = {
        "bag_forward": n_b * 8 + (n_b + 1) * 8 + u_avg * row_b + n_b * row_b,
        "bag_backward_phase1": n_b * row_b + n_b * 8 + u_avg * 2 * row_b,
    }
    alg = {
        "bag_forward": n_b * (8 + row_b) + n_b * row_b + (n_b + 1) * 8,
        "bag_backward_phase1": n_b * row_b + u_avg * 2 * row_b + n_b * 8,
    }
    kernels = {}
    for name, (ms, cnt) in prof.items():
        kernels[name] = {"ms_per_step": ms / K, "launch_groups": cnt}
        if name in alg and cnt:
            kernels[name]["us_per_launch"] = ms / cnt * 1e3
            kernels[name]["compulsory_gbs"] = compulsory[name] / (ms / cnt / 1e3) / 1e9
            kernels[name]["algorithmic_gbs"] = alg[name] / (ms / cnt / 1e3) / 1e9
    kernels["swap_rows"] = {"ms_per_step": sum(kernels.get(k, {}).get("ms_per_step", 0.0)
                                               for k in ("fill_rows", "write_back", "park_victims")),
                            "note": "fill_rows + write_back + park_victims"}
    dom = max((k for k in kernels if k in alg), key=lambda k: kernels[k]["ms_per_step"])
    achieved = kernels[dom]["compulsory_gbs"]
    # DRAM traffic per launch of that kernel from an `ncu --set full` capture of this launch shape, if one has been
    # recorded with scripts/ncu_traffic.sh (profiles/ncu_traffic.json); never a constant in this file
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(traffic_path):
        rec = json.load(open(traffic_path)).get(f"{args.workload}:n{world}:{args.parallelism}", {})
        traffic = rec.get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": int(compulsory[dom]),
                "basis": ("compulsory DRAM bytes per launch (ids + offsets + unique rows + output / gradient rows) / the "
                          "kernel's CUDA-event time in this run, each kernel alone on the GPU"),
                "algorithmic_frac": round(kernels[dom]["algorithmic_gbs"] / peak, 4),
                "algorithmic_bytes_per_launch": int(alg[dom]),
                "note": ("algorithmic bytes charge a row read to every lookup (SURVEY.md 8d); a batch has only ~%d unique "
                         "rows and re-reads them from L2, so algorithmic_frac can exceed 1 -- frac is the honest one"
                         % round(u_avg))}
    if world > 1 and not column and not args.no_fused_exchange:
        # N > 1: the exchange is fused into these two kernels, and the rows that leave the GPU bound them -- every pooled
        # row (forward, stores) / gradient row (backward, loads) whose sample lives on another rank crosses NVLink once
        remote = n_b * row_b * (B - strides[rank]) / B
        link_peak = 770.0     # measured peer-copy bandwidth per direction on this pool (B200_PROFILING.md)
        roofline["nvlink"] = {"bytes_per_launch": int(remote), "peak": link_peak, "unit": "GB/s per direction",
                              "peak_source": "B200_PROFILING.md (measured peer copy)"}
        for name in ("bag_forward", "bag_backward_phase1"):
            if name in kernels and kernels[name].get("us_per_launch"):
                gbs = remote / (kernels[name]["us_per_launch"] / 1e6) / 1e9
                roofline["nvlink"][name] = {"achieved": round(gbs, 1), "frac": round(gbs / link_peak, 4)}
        roofline["note"] += ("; at N > 1 the dominant kernels are NVLink-bound (see nvlink), their HBM fraction is low by "
                             "construction")
    # whole step against the HBM roofline: compulsory bytes of forward + backward + the sort / probe traffic
    step_bytes = compulsory["bag_forward"] + compulsory["bag_backward_phase1"] + n_b * 40 + n_b * 20
    roofline["step_hbm_frac"] = round(step_bytes / (ms_total / K / 1e3) / 1e9 / peak, 4)

    if world > 1:     # every rank's kernel times: the step is the max over ranks
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: round(v["ms_per_step"], 4) for k, v in kernels.items()})
        for k in kernels:
            kernels[k]["ms_per_step_by_rank"] = [g.get(k) for g in gathered]

    par = "single GPU" if world == 1 else (
        f"column-wise x{world} (D / W columns per rank, every rank looks up all ids), NCCL all-to-all of pooled embeddings"
        if column else f"table-wise x{world}, pooled-embedding all-to-all " +
        ("via NCCL" if args.no_fused_exchange else "fused into the fwd/bwd kernels over NVLink peer memory"))
    line = {
        "metric": "embedding lookups/sec (fwd+bwd)", "value": value, "unit": "lookups/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {F} tables, {sum(rows_all):,} rows, dim {D}, batch {B}, "
                        f"prefetch_num {P}, cache_ratio {wl['cache_ratio']}, LFU + id-frequency warm start "
                        f"(warmup_ratio {args.warmup_ratio}), fused SGD lr=1",
            "lookahead": ("prepare_ids(window k+1) is enqueued on side streams right after the first step of window k "
                          "(no host wait inside prepare_ids); each timed window submits one prepare_ids, wherever the "
                          "timed region starts") if overlap else "serial (reference order)",
            "tables_per_rank": [F] * world if column else [sum(1 for a in arrange if a == q) for q in range(world)],
            "row_scale": row_scale, "host_table_gb": round(N_loc * row_b / 1e9, 2), "cache_rows_per_rank": C_loc,
            "ids": f"per-table power law s={SKEW} (reference generator), seed {SEED}",
            "parallelism": par,
            "l2": "inputs larger than L2: each step streams >= 2 x n_b x 512 B (1.7 GB at n_b = 1.7 M) vs 126 MB L2",
            "unique_rows_per_step": round(u_avg), "unique_hits": hit_u, "unique_misses": miss_u, "evicted_rows": evicted,
            "miss_ratio_lookups": round(miss_ratio_lookups, 5), "setup_s": round(setup_s, 1),
        },
        "roofline": roofline,
        "kernels": kernels,
        "e2e": {"value": e2e_value, "unit": "lookups/s", "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
                "ms_per_step": e2e_ms / K,
                "ids_h2d": ("side stream, right before the window's prepare_ids" if args.stage_ids == "off" or not overlap else
                            "own stream, one window ahead of the window's prepare_ids (LookaheadPrefetcher.stage, "
                            + args.stage_ids + "); every timed window still copies one window of ids"),
                "note": ("ids of every batch come from pinned host memory inside the timed region; the pooled embeddings "
                         "stay in HBM by design (their consumer is the dense part of the model), one pooled row per step "
                         "is read back as the result; the D2H traffic that matters -- evicted rows going back to the "
                         "host table -- is inside the region in both arms")},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
    }
    if ab:
        line["ab"] = ab
    if e2e_ab is not None:
        line["e2e"]["ab_ms_per_step"] = e2e_ab
    if parity is not None:
        line["parity_check"] = parity
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_oracle_arm(args.workload, steps=P, warmup=0, budget_s=40.0)["cpu_baseline"]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------- CPU oracle
def cpu_oracle_arm(workload, steps, warmup, budget_s=None):
    """The reference's own CPU implementation of the path = the oracle port (pure-PyTorch ATen ops on host cores; the
    reference's module itself is third-party, absent, and hard-wired to CUDA buffers).  Every step is one batch of the
    same shape on a ROW-SCALED table (the ATen op sequence and the ids-per-step are unchanged; the scale is stated)."""
    from oracle import EvictionStrategy as OStrategy
    from oracle import OracleCachedEmbeddingBag

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[workload]
    D, B, P = wl["dim"], wl["batch"], wl["prefetch"]
    scale = 1
    while sum(wl["rows"]) // scale * D * 4 > 6e9:      # keep the oracle's table under 6 GB
        scale *= 2
    rows = [max(1, r // scale) for r in wl["rows"]]
    N, F = sum(rows), len(rows)
    gen = torch.Generator().manual_seed(SEED)
    rows_t = torch.tensor(rows, dtype=torch.long)
    freq = torch.zeros(N, dtype=torch.long)
    for _ in range(2):
        freq += torch.bincount(sample_ids(rows_t, B, gen, "cpu"), minlength=N)
    model = OracleCachedEmbeddingBag(N, D, sparse=True, mode="sum", include_last_offset=True,
                                     cache_ratio=wl["cache_ratio"], ids_freq_mapping=freq, warmup_ratio=0.7,
                                     evict_strategy=OStrategy.LFU,
                                     # slots follow the UNSCALED table so that a window of full batches still fits
                                     cuda_row_num=min(N, max(int(sum(wl["rows"]) * wl["cache_ratio"]), 1)),
                                     _weight=torch.empty(N, D).uniform_(-1.0 / N, 1.0 / N))
    opt = torch.optim.SGD(model.parameters(), lr=1.0)
    model.set_cache_op(False)
    n_b = F * B
    offsets = torch.arange(n_b + 1)
    grad = torch.randn(n_b, D)
    done, step_s, prep_s, preps, slots_window = 0, 0.0, 0.0, 0, None
    total = warmup + steps
    for s in range(total):
        if s % P == 0:
            batch_ids = [sample_ids(rows_t, B, gen, "cpu") for _ in range(P)]
            t0 = time.perf_counter()
            slots_window = torch.chunk(model.cache_weight_mgr.prepare_ids(torch.cat(batch_ids)), P)
            if s >= warmup:
                prep_s += time.perf_counter() - t0
                preps += 1
        t0 = time.perf_counter()
        out = model(slots_window[s % P], offsets)
        out.backward(grad)
        opt.step()
        opt.zero_grad()
        if s >= warmup:
            step_s += time.perf_counter() - t0
            done += 1
            if budget_s is not None and step_s + prep_s > budget_s:
                break
    # prepare_ids covers a whole window of P steps: charge each measured step 1/P of the average prepare
    elapsed = step_s + (prep_s / max(preps, 1)) * (done / P)
    value = n_b * done / elapsed
    sample = (f"{done} steps of batch {B} x {F} tables (prefetch window {P}) on the {workload} shape with rows / {scale} "
              f"({N:,} rows, dim {D}); oracle port = pure-PyTorch CPU ops")
    return {"value": value, "ms_per_step": elapsed / done * 1e3, "steps": done,
            "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": cores, "kind": "port", "sample": sample}}


def run_reference(args):
    rank, local, world = dist_info()
    if rank != 0:
        return
    r = cpu_oracle_arm(args.workload, steps=args.steps, warmup=args.warmup, budget_s=args.reference_budget_s)
    wl = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "embedding lookups/sec (fwd+bwd)", "value": r["value"], "unit": "lookups/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {len(wl['rows'])} tables, dim {wl['dim']}, batch {wl['batch']}, "
                               f"prefetch_num {wl['prefetch']}, cache_ratio {wl['cache_ratio']}, LFU, SGD lr=1",
                   "note": "CPU arm: bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": r["cpu_baseline"],
        "e2e": {"value": r["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="criteo1tb", choices=sorted(WORKLOADS))
    ap.add_argument("--cache-ratio", type=float, default=0.0, help="override the workload's cache_ratio")
    ap.add_argument("--prefetch", type=int, default=0, help="override the workload's look-ahead window (prefetch_num)")
    ap.add_argument("--warmup-ratio", type=float, default=0.7, help="fraction of the slots preloaded at construction")
    ap.add_argument("--row-scale", type=int, default=0, help="divide table rows by this (0 = only if the host lacks RAM)")
    ap.add_argument("--freq-batches", type=int, default=8, help="batches counted for the id-frequency warm start")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="prepare_ids on the compute stream (reference order)")
    ap.add_argument("--placement", default="auto", choices=["auto", "snake"],
                    help="auto: the reference's table->rank map where it has one, else the snake; snake: always")
    ap.add_argument("--no-fused-exchange", action="store_true", help="N > 1: NCCL all-to-all instead of peer-memory kernels")
    ap.add_argument("--no-plan-side", action="store_true", help="keep the backward's radix sort on the compute stream")
    ap.add_argument("--stage-ids", default="auto", choices=["auto", "off", "early", "after-fill"],
                    help="end-to-end arm: copy a window's ids H2D one window earlier on their own stream "
                         "(LookaheadPrefetcher.stage) instead of on the side stream right before its prepare_ids. "
                         "'early' = as soon as the host has them: measured SLOWER at Criteo-1TB (0.78-0.82 vs 0.58-0.65 ms "
                         "per step) -- the copy shares the PCIe read direction with the previous window's fill, which "
                         "the forward waits for; 'after-fill' = the copy waits for that fill on the device: 0.56 vs "
                         "0.65 ms per step, the default on one GPU ('auto'; N > 1 keeps 'off': not measured there)")
    ap.add_argument("--parallelism", default="table", choices=["table", "column"],
                    help="N > 1: table-wise sharding (BASELINE.json configs[3]) or the reference's default column-wise bag")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the untimed parity leg")
    ap.add_argument("--no-graph-step", action="store_true",
                    help="N > 1: eager forward + backward through autograd instead of the CUDA-graph operator step")
    ap.add_argument("--ab", default="", help="'name:ENV=V,ENV2=V2;name2:PRIORITY=0' -- time the device-resident arm again "
                                             "under these settings in the same process")
    ap.add_argument("--ab-reps", type=int, default=3)
    ap.add_argument("--submit-before", action="store_true",
                    help="submit window w+1 to the look-ahead driver before the first step of window w, not after it")
    ap.add_argument("--e2e-ab-modes", default="off,after-fill,early")
    ap.add_argument("--e2e-ab", action="store_true", help="also time the end-to-end arm with the other ids-H2D order")
    ap.add_argument("--trace-steps", action="store_true", help="print GPU / host time between consecutive timed steps")
    ap.add_argument("--verify-only", action="store_true", help="N > 1: run the parity leg and stop")
    ap.add_argument("--reference-budget-s", type=float, default=150.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
