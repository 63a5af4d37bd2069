"""Timeline of the look-ahead pipeline (no nsys in this image): CUDA events on every stream, printed per window.
usage: python scripts/overlap_timeline.py [--plan-side] [--row-scale S] [--windows W]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import cachedembedding_b200 as ce

ap = argparse.ArgumentParser()
ap.add_argument("--plan-side", action="store_true")
ap.add_argument("--row-scale", type=int, default=4)
ap.add_argument("--windows", type=int, default=8)
ap.add_argument("--no-copy-stream", action="store_true")
ap.add_argument("--priority", type=int, default=-1)
ap.add_argument("--quiet", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda:0")
wl = bench.WORKLOADS["criteo1tb"]
rows = [max(1, r // args.row_scale) for r in wl["rows"]]
D, B, P = wl["dim"], wl["batch"], wl["prefetch"]
N = sum(rows)
C = min(N, int(sum(wl["rows"]) * wl["cache_ratio"]))
rows_dev = torch.tensor(rows, device=dev)
gen = torch.Generator(device=dev).manual_seed(1024)
freq = torch.zeros(N, dtype=torch.long, device=dev)
for _ in range(8):
    freq += torch.bincount(bench.sample_ids(rows_dev, B, gen, dev), minlength=N)
model = ce.CachedEmbeddingBag(N, D, sparse=True, mode="sum", include_last_offset=True, cache_ratio=wl["cache_ratio"],
                              ids_freq_mapping=freq, warmup_ratio=0.7, evict_strategy=ce.EvictionStrategy.LFU,
                              cuda_row_num=C, fused_optimizer="sgd", lr=1.0)
model.set_cache_op(False)
F = len(rows); n_b = F * B
offsets = torch.arange(n_b + 1, device=dev)
grad = torch.randn(n_b, D, device=dev)
W = args.windows + 2
wins = [[bench.sample_ids(rows_dev, B, gen, dev) for _ in range(P)] for _ in range(W)]
torch.cuda.synchronize()

def ev(stream=None):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream if stream is not None else torch.cuda.current_stream())
    return e

pf = ce.LookaheadPrefetcher(model, priority=args.priority, copy_stream=not args.no_copy_stream)
plan = dict(offsets=offsets) if args.plan_side else {}
base = ev()
cpu0 = time.perf_counter()
log = []
rec = {}
t_sub0 = time.perf_counter()
rec["side_start"] = ev(pf.stream)
h = pf.submit(wins[0], **plan)
rec["side_end"] = ev(pf.stream); rec["copy_end"] = ev(pf.copy_stream) if pf.copy_stream else None
rec["cpu_submit"] = (t_sub0 - cpu0, time.perf_counter() - cpu0)
recs = [rec]
for k in range(W):
    r = recs[k]
    slots = torch.chunk(h.wait(), P)
    r["comp_start"] = ev()
    t0 = time.perf_counter()
    for j in range(P):
        out = model(slots[j], offsets)
        out.backward(grad)
    r["comp_end"] = ev()
    r["cpu_enqueue"] = (t0 - cpu0, time.perf_counter() - cpu0)
    pf.window_enqueued()
    if k + 1 < W:
        nr = {}
        t_sub0 = time.perf_counter()
        nr["side_start"] = ev(pf.stream)
        h = pf.submit(wins[k + 1], **plan)
        nr["side_end"] = ev(pf.stream); nr["copy_end"] = ev(pf.copy_stream) if pf.copy_stream else None
        nr["cpu_submit"] = (t_sub0 - cpu0, time.perf_counter() - cpu0)
        recs.append(nr)
pf.close()
torch.cuda.synchronize()
print(f"rows/{args.row_scale} plan_side={args.plan_side} copy_stream={not args.no_copy_stream}  (all times ms since start)")
print(" win | cpu submit [start,end] | side [start,end] copy_end | cpu enqueue [start,end] | compute [start,end] | period")
prev_end = None
periods = []
for k, r in enumerate(recs):
    g = lambda e: base.elapsed_time(e) if e is not None else float('nan')
    ce_ = g(r.get("comp_end")); cs = g(r.get("comp_start"))
    period = (ce_ - prev_end) if prev_end is not None else float('nan')
    prev_end = ce_
    if 4 <= k <= len(recs) - 2: periods.append(period)
    if not args.quiet: print(f" {k:3d} | {r['cpu_submit'][0]*1e3:7.2f} {r['cpu_submit'][1]*1e3:7.2f} | {g(r['side_start']):7.2f} {g(r['side_end']):7.2f} {g(r['copy_end']):7.2f} | "
          f"{r['cpu_enqueue'][0]*1e3:7.2f} {r['cpu_enqueue'][1]*1e3:7.2f} | {cs:7.2f} {ce_:7.2f} | {period:6.2f}")
mgr = model.cache_weight_mgr
print("steady-state period per window: %.2f ms -> %.3f ms/step  env=%s" % (sum(periods)/len(periods), sum(periods)/len(periods)/P, {k:v for k,v in os.environ.items() if k.startswith("CEBAG_")}))
print("misses/window", mgr.num_miss_history[-4:], "evicted", mgr.num_write_back_history[-4:])
