"""Forward + fused backward of one Criteo-1TB-shape batch on a device-only slot cache (no host table), for ncu and
for knob sweeps.  Prints per-kernel event timings from the library's own timers.

usage: python scripts/profile_step.py [steps] [batch] ["name:ENV=V,..;name2:..."]   (settings the library reads per call)
       ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
           --log-file gpurun_out/launches.csv python scripts/profile_step.py 2      # launch list of the timed steps only
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
import cachedembedding_b200 as ce
from cachedembedding_b200 import _lib


class Owner:
    sparse = True

    def __init__(self, lr):
        self._fused_optimizer = {"kind": _lib.OPT_SGD, "lr": lr, "eps": 0.0}
        self.cache_weight_mgr = type("M", (), {"cuda_cached_state": None})()


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    dev = torch.device("cuda:0")
    rows = torch.tensor(bench.CRITEO_1TB_ROWS, device=dev)
    F, D = rows.numel(), 128
    C = int(sum(bench.CRITEO_1TB_ROWS) * 0.01)
    gen = torch.Generator(device=dev).manual_seed(1024)
    weight = (torch.rand(C, D, device=dev) - 0.5).requires_grad_(True)
    n = F * B
    offsets = torch.arange(n + 1, device=dev)
    grad = torch.randn(n, D, device=dev)
    owner = Owner(1.0)
    # slot ids with the duplication structure of the real ids: hash the global id into the cache
    batches = [(bench.sample_ids(rows, B, gen, dev) * 2654435761 % C) for _ in range(steps + 3)]
    uniq = torch.unique(batches[3]).numel()
    for i in range(3):
        out = ce.embedding_bag_cached(weight, batches[i], offsets, include_last_offset=True, mode="sum", owner=owner)
        out.backward(grad)
    torch.cuda.synchronize()
    sweep = [x for x in (sys.argv[3] if len(sys.argv) > 3 else "default:").split(";") if x]
    alg = {"bag_forward": n * (8 + 4 * D) + n * 4 * D + (n + 1) * 8,
           "bag_backward_phase1": n * 4 * D + uniq * 8 * D + n * 8}
    comp = {"bag_forward": n * 8 + (n + 1) * 8 + uniq * 4 * D + n * 4 * D,
            "bag_backward_phase1": n * 4 * D + uniq * 8 * D + n * 8}
    for spec in sweep:
        name, _, kv = spec.partition(":")
        env = dict(x.split("=") for x in kv.split(",") if x)
        os.environ.update(env)
        _lib.profile_enable(True)
        torch.cuda.cudart().cudaProfilerStart()       # `ncu --profile-from-start off` captures from here
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            out = ce.embedding_bag_cached(weight, batches[3 + i], offsets, include_last_offset=True, mode="sum", owner=owner)
            out.backward(grad)
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        for key in env:
            os.environ.pop(key, None)
        ms = e0.elapsed_time(e1) / steps
        knobs = {k: v for k, v in os.environ.items() if k.startswith("CEBAG_")}
        print(f"[{name}] {kv} n={n} unique={uniq} step={ms:.3f} ms  {n / ms / 1e6:.2f} G lookups/s  env={knobs}")
        for nm, (t, c) in prof.items():
            extra = (f"  {comp[nm] / (t / c / 1e3) / 1e9:7.0f} GB/s compulsory  {alg[nm] / (t / c / 1e3) / 1e9:7.0f} GB/s algorithmic"
                     if nm in alg else "")
            print(f"  {nm:22s} {t / steps * 1e3:8.1f} us/step{extra}")


if __name__ == "__main__":
    main()
