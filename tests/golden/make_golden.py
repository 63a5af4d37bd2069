"""Generates tests/golden/cache_scenarios.npz: seeded cache-manager scenarios and the answers of the CPU oracle.

The reference's implementation of this path (ColossalAI's cache_embedding package) cannot be imported here or anywhere
offline, and the reference repository has no fixtures of its own, so these vectors come from the oracle -- which is
itself pinned against torch.nn.EmbeddingBag + SGD and the upstream known-answer tests (tests/test_oracle.py).  They
freeze the oracle's behaviour (a regression in oracle/ fails tests/test_golden.py on CPU) and give the CUDA path a
fixed target that does not depend on the oracle being importable next to it.

    python tests/golden/make_golden.py        # rewrites cache_scenarios.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import EvictionStrategy, OracleCachedEmbeddingBag  # noqa: E402

SCENARIOS = [
    # name, N, D, cache_ratio, warmup, strategy, use_freq, protect_windows, calls, ids per call
    ("lfu_freq", 400, 8, 0.10, 0.7, "LFU", True, 1, 12, 26),
    ("lfu_nofreq", 300, 4, 0.15, 0.0, "LFU", False, 1, 12, 28),
    ("dataset_freq", 500, 8, 0.08, 0.7, "DATASET", True, 1, 12, 26),
    ("dataset_nofreq", 256, 16, 0.20, 1.0, "DATASET", False, 1, 10, 32),
    ("lfu_lookahead2", 600, 8, 0.20, 0.5, "LFU", True, 2, 12, 44),
]


def scenario_inputs(name, N, D, calls, n_ids):
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    weight = (torch.rand(N, D, generator=gen) - 0.5)
    freq = torch.randint(0, 40, (N,), generator=gen)
    ids = [(torch.rand(n_ids, generator=gen) ** 3 * N).long().clamp_(0, N - 1) for _ in range(calls)]
    return weight, freq, ids


def run_oracle(name, N, D, ratio, warm, strategy, use_freq, protect, calls, n_ids):
    weight, freq, ids = scenario_inputs(name, N, D, calls, n_ids)
    bag = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True,
                                   cache_ratio=ratio, ids_freq_mapping=freq if use_freq else None, warmup_ratio=warm,
                                   evict_strategy=getattr(EvictionStrategy, strategy))
    mgr = bag.cache_weight_mgr
    mgr.protect_windows = protect
    out = {"idx_map": mgr.idx_map.numpy().copy()}
    for k, x in enumerate(ids):
        out[f"slots_{k}"] = mgr.prepare_ids(x).numpy().copy()
        out[f"cached_idx_map_{k}"] = mgr.cached_idx_map.numpy().copy()
    out["inverted_cached_idx"] = mgr.inverted_cached_idx.numpy().copy()
    if strategy == "LFU":
        out["freq_cnter"] = mgr.freq_cnter.numpy().copy()
    out["hits"] = np.array(mgr.num_hits_history)
    out["misses"] = np.array(mgr.num_miss_history)
    out["write_backs"] = np.array(mgr.num_write_back_history)
    return out


def main():
    blob = {}
    for sc in SCENARIOS:
        for key, val in run_oracle(*sc).items():
            blob[f"{sc[0]}/{key}"] = val
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cache_scenarios.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes;", len(blob), "arrays")


if __name__ == "__main__":
    main()
