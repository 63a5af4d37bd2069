"""Golden fixtures (tests/golden/cache_scenarios.npz, made by tests/golden/make_golden.py): the oracle reproduces them
on CPU; the CUDA path reproduces them through the C ABI on the GPU -- bit-exact slot ids, maps, counters, histories."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "cache_scenarios.npz"))


@pytest.mark.parametrize("sc", mg.SCENARIOS, ids=[s[0] for s in mg.SCENARIOS])
def test_oracle_reproduces_golden(sc):
    got = mg.run_oracle(*sc)
    keys = [k for k in GOLD.files if k.startswith(sc[0] + "/")]
    assert len(keys) == len(got)
    for k in keys:
        assert np.array_equal(GOLD[k], got[k.split("/", 1)[1]]), k
    assert GOLD[sc[0] + "/write_backs"].sum() > 0, "the scenario must exercise eviction"


@pytest.mark.gpu
@pytest.mark.parametrize("sc", mg.SCENARIOS, ids=[s[0] for s in mg.SCENARIOS])
def test_cuda_reproduces_golden(sc):
    import cachedembedding_b200 as ce
    name, N, D, ratio, warm, strategy, use_freq, protect, calls, n_ids = sc
    weight, freq, ids = mg.scenario_inputs(name, N, D, calls, n_ids)
    bag = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True, cache_ratio=ratio,
                                ids_freq_mapping=freq if use_freq else None, warmup_ratio=warm,
                                evict_strategy=getattr(ce.EvictionStrategy, strategy))
    mgr = bag.cache_weight_mgr
    mgr.protect_windows = protect
    g = lambda key: GOLD[f"{name}/{key}"]
    assert np.array_equal(mgr.idx_map.cpu().numpy(), g("idx_map"))
    for k, x in enumerate(ids):
        slots = mgr.prepare_ids(x.cuda())
        assert np.array_equal(slots.cpu().numpy(), g(f"slots_{k}")), f"slot ids of call {k}"
        assert np.array_equal(mgr.cached_idx_map.cpu().numpy(), g(f"cached_idx_map_{k}")), f"slot->row map after call {k}"
    assert np.array_equal(mgr.inverted_cached_idx.cpu().numpy(), g("inverted_cached_idx"))
    if strategy == "LFU":
        assert np.array_equal(mgr.freq_cnter.cpu().numpy(), g("freq_cnter"))
    assert mgr.num_hits_history == g("hits").tolist()
    assert mgr.num_miss_history == g("misses").tolist()
    assert mgr.num_write_back_history == g("write_backs").tolist()
