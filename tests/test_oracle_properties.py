"""Property tests of the oracle's cache manager (hypothesis): invariants every valid implementation must keep, checked
over random tables, strategies and id sequences.  The same invariants are asserted on the CUDA path at full size in
tests/test_gpu_parity.py (where the oracle itself is too slow to be the comparison)."""
import torch
from hypothesis import given, settings, strategies as st

from oracle import EvictionStrategy, OracleCachedEmbeddingBag


@st.composite
def scenario(draw):
    N = draw(st.integers(8, 120))
    C = draw(st.integers(2, max(2, N // 2)))
    strategy = draw(st.sampled_from([EvictionStrategy.LFU, EvictionStrategy.DATASET]))
    use_freq = draw(st.booleans())
    warm = draw(st.sampled_from([0.0, 0.5, 1.0]))
    protect = draw(st.sampled_from([1, 2]))
    seed = draw(st.integers(0, 2**31 - 1))
    calls = draw(st.integers(1, 8))
    return N, C, strategy, use_freq, warm, protect, seed, calls


@settings(max_examples=60, deadline=None)
@given(scenario())
def test_cache_manager_invariants(sc):
    N, C, strategy, use_freq, warm, protect, seed, calls = sc
    gen = torch.Generator().manual_seed(seed)
    D = 3
    weight = torch.rand(N, D, generator=gen)
    original = weight.clone()
    freq = torch.randint(0, 20, (N,), generator=gen) if use_freq else None
    bag = OracleCachedEmbeddingBag(N, D, _weight=weight, mode="sum", include_last_offset=True, cuda_row_num=C,
                                   cache_ratio=1.0, ids_freq_mapping=freq, warmup_ratio=warm, evict_strategy=strategy)
    mgr = bag.cache_weight_mgr
    mgr.protect_windows = protect
    prev_rows = torch.tensor([], dtype=torch.long)
    for _ in range(calls):
        n = int(torch.randint(1, 3 * C, (1,), generator=gen))
        ids = torch.randint(0, N, (n,), generator=gen)
        rows = torch.unique(mgr.idx_map[ids])
        # keep the call inside the capacity contract: |rows| <= C, and with two protected windows
        # |rows(k-1) U rows(k)| <= C (rows shared with the previous window cost nothing)
        if protect == 2:
            shared = rows[torch.isin(rows, prev_rows)]
            fresh = rows[~torch.isin(rows, prev_rows)][: max(0, C - prev_rows.numel())]
            keep = torch.cat([shared, fresh])
        else:
            keep = rows[:C]
        if keep.numel() == 0:
            continue
        ids = ids[torch.isin(mgr.idx_map[ids], keep)]
        rows = torch.unique(mgr.idx_map[ids])
        resident_before = mgr.inverted_cached_idx[prev_rows] >= 0
        slots = mgr.prepare_ids(ids)
        # every id resolved to the slot that holds its row
        assert torch.equal(mgr.cached_idx_map[slots], mgr.idx_map[ids])
        occ = mgr.cached_idx_map >= 0
        assert int(occ.sum()) + mgr.cuda_available_row_num == C
        s = torch.nonzero(occ).squeeze(1)
        r = mgr.cached_idx_map[s]
        assert r.unique().numel() == r.numel()                      # a row is resident at most once
        assert torch.equal(mgr.inverted_cached_idx[r], s)
        assert int((mgr.inverted_cached_idx >= 0).sum()) == s.numel()
        # the cached copy of every resident row equals the host row (no training in this test)
        assert torch.equal(mgr.cuda_cached_weight.detach()[s], mgr.weight[r])
        if protect == 2:                                            # the previous window's rows survived this call
            assert bool((mgr.inverted_cached_idx[prev_rows][resident_before] >= 0).all())
        if strategy == EvictionStrategy.LFU:
            assert bool((mgr.freq_cnter[~occ] == torch.iinfo(torch.long).max).all())
            assert bool((mgr.freq_cnter[occ] >= 0).all())
        assert mgr.num_hits_history[-1] + mgr.num_miss_history[-1] == rows.numel()
        prev_rows = rows
    mgr.flush()
    assert torch.equal(mgr.weight, original)                        # write-backs and flush moved rows, never changed them
    assert mgr.cuda_available_row_num == C
