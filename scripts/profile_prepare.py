"""prepare_ids over Criteo-1TB-shape look-ahead windows with the map work at full size (177.9 M rows, 1 % slots,
13.6 M ids per window) and narrow rows (dim 4: a 2.8 GB host table instead of 91 GB, so the run starts in seconds and
the PCIe row traffic is negligible) -- for ncu captures and knob sweeps of the cache manager's integer kernels.
Prints the library's per-kernel event timers per window.

usage: python scripts/profile_prepare.py [windows] ["name:ENV=V,ENV2=V;name2:..."]   (settings the library reads per call)
       ncu --profile-from-start off --set full --clock-control none -k regex:'probe|lfu_count|select_hist|bitmap' \
           -o gpurun_out/prepare python scripts/profile_prepare.py 1
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench
import cachedembedding_b200 as ce
from cachedembedding_b200 import _lib


def main():
    windows = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    dev = torch.device("cuda:0")
    rows_list = bench.CRITEO_1TB_ROWS
    rows = torch.tensor(rows_list, device=dev)
    N, D, B, P = sum(rows_list), 4, 65536, 8
    C = int(N * 0.01)
    gen = torch.Generator(device=dev).manual_seed(bench.SEED)
    counter = ce.IdFrequencyCounter(N, dev)
    for _ in range(8):
        counter.update(bench.sample_ids(rows, B, gen, dev))
    freq = counter.result()
    model = ce.CachedEmbeddingBag(N, D, ids_freq_mapping=freq, sparse=True, mode="sum", include_last_offset=True,
                                  cache_ratio=0.01, warmup_ratio=0.7, evict_strategy=ce.EvictionStrategy.LFU,
                                  cuda_row_num=C, init_seed=bench.SEED)
    mgr = model.cache_weight_mgr
    warm = 3
    sweep = [x for x in (sys.argv[2] if len(sys.argv) > 2 else "default:").split(";") if x]
    n_win = warm + windows * len(sweep)
    wins = [torch.cat([bench.sample_ids(rows, B, gen, dev) for _ in range(P)]) for _ in range(n_win)]
    for w in wins[:warm]:
        mgr.prepare_ids(w)
    torch.cuda.synchronize()
    n = wins[0].numel()
    for k, spec in enumerate(sweep):
        name, _, kv = spec.partition(":")
        env = dict(x.split("=") for x in kv.split(",") if x)
        os.environ.update(env)
        _lib.profile_enable(True)
        torch.cuda.cudart().cudaProfilerStart()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for w in wins[warm + k * windows: warm + (k + 1) * windows]:
            mgr.prepare_ids(w)
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        for key in env:
            os.environ.pop(key, None)
        print(f"[{name}] {kv}  n={n} ids/window  C={C}  {e0.elapsed_time(e1) / windows * 1e3:.0f} us/window (host in the "
              f"loop)  misses/window={sum(mgr.num_miss_history[-windows:]) // windows}")
        print("   " + "  ".join(f"{nm} {t / windows * 1e3:.1f}" for nm, (t, c) in prof.items()))


if __name__ == "__main__":
    main()
