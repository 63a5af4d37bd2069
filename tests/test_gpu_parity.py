"""Parity of the CUDA path (through the C-ABI library) with the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): bit-exact cache-slot assignment and id->slot maps; pooled sums and updated rows
within 1e-5 relative fp32.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import EvictionStrategy as OStrategy
from oracle import OracleCachedEmbeddingBag, rowwise_adagrad_reference

# north_star: pooled sums and updated rows within 1e-5 relative fp32.  A sum of several rows that cancels has no
# meaningful element-wise relative error, so "relative" is taken against the magnitude of the summed operands:
# |got - want| <= 1e-5 * |want| + 1e-5 * scale, scale = max |operand| (the randn test tables have scale ~4).
RTOL, ATOL = 1e-5, 4e-5


def close(got, want, what=""):
    """1e-5 relative: element-wise against |want|, plus 1e-5 of the largest magnitude in `want` for sums that cancel."""
    want = torch.as_tensor(want)
    got = torch.as_tensor(got).to(want.dtype)
    torch.testing.assert_close(got, want, rtol=RTOL, atol=RTOL * max(float(want.abs().max()), 1e-30), msg=None)


def _mods():
    import cachedembedding_b200 as ce
    return ce


def make_bags(N, D, G, max_len, gen, include_last=True, empty_frac=0.2):
    lens = torch.randint(0, max_len + 1, (G,), generator=gen)
    lens[torch.rand(G, generator=gen) < empty_frac] = 0
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(lens, 0)])
    n = int(offsets[-1])
    ids = torch.randint(0, N, (n,), generator=gen)
    if not include_last:
        offsets = offsets[:-1]
    return ids, offsets


def assert_maps_equal(mgr, omgr):
    assert torch.equal(mgr.cached_idx_map.cpu(), omgr.cached_idx_map), "slot -> row map differs"
    assert torch.equal(mgr.inverted_cached_idx.cpu(), omgr.inverted_cached_idx), "row -> slot map differs"
    assert torch.equal(mgr.idx_map.cpu(), omgr.idx_map), "id -> row map differs"
    assert mgr.cuda_available_row_num == omgr.cuda_available_row_num
    if hasattr(omgr, "freq_cnter"):
        assert torch.equal(mgr.freq_cnter.cpu(), omgr.freq_cnter), "LFU counters differ"


# ---------------------------------------------------------------------------------------------------- forward
@pytest.mark.parametrize("D", [128, 16, 32, 64, 256, 512, 5, 36, 8, 1, 100])
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_forward_matches_oracle(D, mode):
    ce = _mods()
    gen = torch.Generator().manual_seed(D * 7 + len(mode))
    C, G = 300, 257
    weight = torch.randn(C, D, generator=gen)
    slots, offsets = make_bags(C, D, G, 9, gen)
    ref = torch.nn.functional.embedding_bag(slots, weight, offsets, mode=mode, include_last_offset=True)
    out = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.cuda(), include_last_offset=True, mode=mode)
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("D", [128, 16, 64, 256, 512, 8])
@pytest.mark.parametrize("knobs", [{"CEBAG_FWD_LD": "0"}, {"CEBAG_FWD_LD": "2"}, {"CEBAG_FWD_UNROLL": "8"},
                                   {"CEBAG_FWD_CTAS_PER_SM": "2", "CEBAG_FWD_LD": "0"}, {"CEBAG_FWD_THREADS": "128"}])
@pytest.mark.parametrize("kind", ["pooling1", "ragged"])
def test_forward_kernel_variants(D, knobs, kind, monkeypatch):
    """The fast-path variants of the forward (L1 policy of the row loads, rows in flight, grid size; the library reads
    these knobs per call) against torch: a pure gather (pooling factor 1, the DLRM call) bit for bit, ragged bags
    within 1e-5."""
    ce = _mods()
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    gen = torch.Generator().manual_seed(D + len(kind))
    C, G = 500, 256 * 9 + 77
    weight = torch.randn(C, D, generator=gen)
    if kind == "pooling1":
        slots, offsets = torch.randint(0, C, (G,), generator=gen), torch.arange(G + 1)
    else:
        slots, offsets = make_bags(C, D, G, 5, gen)
    ref = torch.nn.functional.embedding_bag(slots, weight, offsets, mode="sum", include_last_offset=True)
    out = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.cuda(), include_last_offset=True, mode="sum")
    if kind == "pooling1":
        assert torch.equal(out.cpu(), ref)
    else:
        torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("offset_dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("include_last", [True, False])
def test_forward_offsets_variants_weights_padding(offset_dtype, include_last):
    ce = _mods()
    gen = torch.Generator().manual_seed(5)
    C, D, G = 64, 128, 100
    weight = torch.randn(C, D, generator=gen)
    slots, offsets = make_bags(C, D, G, 6, gen, include_last=include_last)
    psw = torch.randn(slots.numel(), generator=gen)
    for pad in (None, 3):
        ref = torch.nn.functional.embedding_bag(slots, weight, offsets, mode="sum", per_sample_weights=psw,
                                                include_last_offset=include_last, padding_idx=pad)
        out = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.to(offset_dtype).cuda(), psw.cuda(),
                                      include_last_offset=include_last, mode="sum", padding_idx=pad)
        torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)
        refm = torch.nn.functional.embedding_bag(slots, weight, offsets, mode="mean",
                                                 include_last_offset=include_last, padding_idx=pad)
        outm = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.to(offset_dtype).cuda(),
                                       include_last_offset=include_last, mode="mean", padding_idx=pad)
        torch.testing.assert_close(outm.cpu(), refm, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("D", [128, 16, 64, 256, 36])
@pytest.mark.parametrize("kind", ["pooling1", "ragged", "mixed"])
def test_forward_tma_gather(D, kind, monkeypatch):
    """Calls large enough for the TMA row-gather kernel (bag_forward_tma.cu; D = 36 is not a multiple of 4 floats x 4 and
    falls back to the LDG kernel): pooling factor 1 is a pure bulk-copy gather and must be BIT-EXACT; ragged tiles are
    summed by the warp inside the same kernel; a padding index and a partial last tile are covered."""
    ce = _mods()
    import subprocess, sys, os
    if os.environ.get("CEBAG_FWD_TMA") != "1":
        # the kernel is opt-in (it is slower than the LDG kernel at 512 B rows): run this test in a child with it on
        env = dict(os.environ, CEBAG_FWD_TMA="1")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        out = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu",
                              f"{__file__}::test_forward_tma_gather[{kind}-{D}]"],
                             env=env, cwd=root, capture_output=True, text=True)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
        return
    gen = torch.Generator().manual_seed(D + len(kind))
    C, G = 3000, 9973                     # G not a multiple of the 32-row tile
    weight = torch.randn(C, D, generator=gen)
    if kind == "pooling1":
        slots = torch.randint(0, C, (G,), generator=gen)
        offsets = torch.arange(G + 1)
    elif kind == "ragged":
        slots, offsets = make_bags(C, D, G, 5, gen)
    else:                                 # mostly one id per bag, a few tiles with longer / empty bags
        lens = torch.ones(G, dtype=torch.long)
        lens[torch.randint(0, G, (60,), generator=gen)] = 3
        lens[torch.randint(0, G, (60,), generator=gen)] = 0
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(lens, 0)])
        slots = torch.randint(0, C, (int(offsets[-1]),), generator=gen)
    for pad in (None, 7):
        ref = torch.nn.functional.embedding_bag(slots, weight, offsets, mode="sum", include_last_offset=True,
                                                padding_idx=pad)
        out = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.cuda(), include_last_offset=True, mode="sum",
                                      padding_idx=pad)
        if kind == "pooling1":
            assert torch.equal(out.cpu(), ref), "a pure gather must be bit-exact"
        else:
            torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)


def test_forward_2d_input_and_empty():
    ce = _mods()
    gen = torch.Generator().manual_seed(9)
    weight = torch.randn(50, 16, generator=gen)
    idx2d = torch.randint(0, 50, (7, 3), generator=gen)
    ref = torch.nn.functional.embedding_bag(idx2d, weight, mode="sum")
    out = ce.embedding_bag_cached(weight.cuda(), idx2d.cuda(), None, mode="sum")
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=ATOL)
    # no lookups at all: G bags, all empty
    out = ce.embedding_bag_cached(weight.cuda(), torch.empty(0, dtype=torch.long).cuda(),
                                  torch.zeros(5, dtype=torch.long).cuda(), include_last_offset=True, mode="sum")
    assert out.shape == (4, 16) and float(out.abs().sum()) == 0.0


def test_forward_sample_major_layout():
    ce = _mods()
    gen = torch.Generator().manual_seed(10)
    C, D, F, B = 40, 128, 5, 12
    weight = torch.randn(C, D, generator=gen)
    slots, offsets = make_bags(C, D, F * B, 3, gen)
    ref = torch.nn.functional.embedding_bag(slots, weight, offsets, mode="sum", include_last_offset=True)
    ref = torch.cat(ref.split(B), 1)    # (B, F*D): what the table-wise bag sends to the all-to-all
    out = ce.embedding_bag_cached(weight.cuda(), slots.cuda(), offsets.cuda(), include_last_offset=True, mode="sum",
                                  layout="sample_major", layout_batch=B)
    torch.testing.assert_close(out.view(B, F * D).cpu(), ref, rtol=RTOL, atol=ATOL)


# ---------------------------------------------------------------------------------------------------- backward
class _Owner:
    """Minimal stand-in for the module the autograd function consults for its backward mode."""

    def __init__(self, sparse=True, fused=None, state=None):
        self.sparse = sparse
        self._fused_optimizer = fused
        self.cache_weight_mgr = type("M", (), {"cuda_cached_state": state})()


@pytest.mark.parametrize("D", [128, 16, 5, 256, 64])
@pytest.mark.parametrize("mode,use_psw", [("sum", False), ("sum", True), ("mean", False)])
def test_backward_sparse_dense_and_fused_sgd(D, mode, use_psw):
    ce = _mods()
    from cachedembedding_b200 import _lib
    gen = torch.Generator().manual_seed(D + 31 * use_psw)
    C, G = 97, 400          # few slots, many lookups: long runs of duplicates that cross chunk boundaries
    weight = torch.randn(C, D, generator=gen)
    slots, offsets = make_bags(C, D, G, 7, gen)
    slots[: slots.numel() // 3] = 11      # one very hot slot
    psw = torch.randn(slots.numel(), generator=gen) if use_psw else None
    grad = torch.randn(G, D, generator=gen)
    lr = 0.37

    wref = weight.clone().requires_grad_(True)
    pref = psw.clone().requires_grad_(True) if use_psw else None
    ref = torch.nn.functional.embedding_bag(slots, wref, offsets, mode=mode, per_sample_weights=pref,
                                            include_last_offset=True)
    ref.backward(grad)
    dense_ref = wref.grad
    updated_ref = weight - lr * dense_ref

    # sparse COO
    w = weight.cuda().requires_grad_(True)
    p = psw.cuda().requires_grad_(True) if use_psw else None
    out = ce.embedding_bag_cached(w, slots.cuda(), offsets.cuda(), p, include_last_offset=True, mode=mode,
                                  owner=_Owner(sparse=True))
    out.backward(grad.cuda())
    assert w.grad.is_sparse
    close(w.grad.to_dense().cpu(), dense_ref)
    if use_psw:
        close(p.grad.cpu(), pref.grad)
    # dense
    w = weight.cuda().requires_grad_(True)
    out = ce.embedding_bag_cached(w, slots.cuda(), offsets.cuda(), psw.cuda() if use_psw else None,
                                  include_last_offset=True, mode=mode, owner=_Owner(sparse=False))
    out.backward(grad.cuda())
    assert not w.grad.is_sparse
    close(w.grad.cpu(), dense_ref)
    # fused SGD (in place, no grad materialised)
    w = weight.cuda().requires_grad_(True)
    fused = {"kind": _lib.OPT_SGD, "lr": lr, "eps": 0.0}
    out = ce.embedding_bag_cached(w, slots.cuda(), offsets.cuda(), psw.cuda() if use_psw else None,
                                  include_last_offset=True, mode=mode, owner=_Owner(fused=fused))
    out.backward(grad.cuda())
    assert w.grad is None
    close(w.detach().cpu(), updated_ref)


@pytest.mark.parametrize("knobs", [{}, {"CEBAG_BWD_THREADS": "128"}, {"CEBAG_BWD_THREADS": "128", "CEBAG_BWD_UNROLL": "8"},
                                   {"CEBAG_SORT_CTAS": "0"}, {"CEBAG_SORT_CTAS": "3", "CEBAG_SORT_ITEMS": "16"}])
def test_backward_fused_is_deterministic_and_handles_padding(knobs, monkeypatch):
    """Also under the launch-shape knobs the library reads per call (CTA size of phase 1, positions in flight, CTAs of
    the one-sweep sort passes): the result does not depend on them, bit for bit."""
    ce = _mods()
    from cachedembedding_b200 import _lib
    gen = torch.Generator().manual_seed(77)
    C, D, G = 500, 128, 3000
    weight = torch.randn(C, D, generator=gen)
    slots, offsets = make_bags(C, D, G, 5, gen)
    grad = torch.randn(G, D, generator=gen)
    fused = {"kind": _lib.OPT_SGD, "lr": 0.5, "eps": 0.0}
    results = []
    for rep in range(3):
        if rep == 1:                     # the first run uses the defaults
            for k, v in knobs.items():
                monkeypatch.setenv(k, v)
        w = weight.cuda().requires_grad_(True)
        out = ce.embedding_bag_cached(w, slots.cuda(), offsets.cuda(), include_last_offset=True, mode="sum",
                                      padding_idx=7, owner=_Owner(fused=fused))
        out.backward(grad.cuda())
        results.append(w.detach().cpu())
    assert torch.equal(results[0], results[1]) and torch.equal(results[0], results[2]), "fused backward not bitwise repeatable"
    torch.testing.assert_close(results[0][7], weight[7], rtol=0, atol=0)   # padding slot untouched
    wref = weight.clone().requires_grad_(True)
    torch.nn.functional.embedding_bag(slots, wref, offsets, mode="sum", include_last_offset=True,
                                      padding_idx=7).backward(grad)
    close(results[0], weight - 0.5 * wref.grad)


@pytest.mark.parametrize("D", [128, 16])
def test_backward_fused_rowwise_adagrad(D):
    ce = _mods()
    from cachedembedding_b200 import _lib
    gen = torch.Generator().manual_seed(3 + D)
    C, G = 60, 300
    weight = torch.randn(C, D, generator=gen)
    state = torch.rand(C, generator=gen)
    slots, offsets = make_bags(C, D, G, 4, gen)
    grad = torch.randn(G, D, generator=gen)
    lr, eps = 0.1, 1e-8
    Wref, mref = rowwise_adagrad_reference(weight.numpy(), state.numpy(), slots.numpy(), offsets.numpy(), grad.numpy(),
                                           lr, eps)
    w = weight.cuda().requires_grad_(True)
    st = state.cuda()
    fused = {"kind": _lib.OPT_ROWWISE_ADAGRAD, "lr": lr, "eps": eps}
    out = ce.embedding_bag_cached(w, slots.cuda(), offsets.cuda(), include_last_offset=True, mode="sum",
                                  owner=_Owner(fused=fused, state=st))
    out.backward(grad.cuda())
    close(st.cpu().double(), torch.from_numpy(mref))
    close(w.detach().cpu().double(), torch.from_numpy(Wref))


# ---------------------------------------------------------------------------------------------------- cache manager
@pytest.mark.parametrize("strategy", ["LFU", "DATASET"])
@pytest.mark.parametrize("use_freq", [True, False])
@pytest.mark.parametrize("N,D,ratio,warm", [(1000, 16, 0.1, 0.7), (5000, 128, 0.05, 0.0), (333, 5, 0.3, 1.0)])
def test_prepare_ids_maps_bit_exact(strategy, use_freq, N, D, ratio, warm):
    """Slot assignment and every map identical to the oracle after each call, under heavy eviction traffic."""
    ce = _mods()
    gen = torch.Generator().manual_seed(N + D)
    weight = torch.randn(N, D, generator=gen)
    freq = torch.randint(0, 50, (N,), generator=gen) if use_freq else None
    C = int(N * ratio)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True,
                                  cache_ratio=ratio, ids_freq_mapping=freq, warmup_ratio=warm,
                                  evict_strategy=getattr(ce.EvictionStrategy, strategy))
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True,
                                      cache_ratio=ratio, ids_freq_mapping=freq, warmup_ratio=warm,
                                      evict_strategy=getattr(OStrategy, strategy))
    mgr, omgr = model.cache_weight_mgr, omodel.cache_weight_mgr
    assert_maps_equal(mgr, omgr)
    torch.testing.assert_close(mgr.cuda_cached_weight.detach().cpu(), omgr.cuda_cached_weight.detach(), rtol=0, atol=0)
    for step in range(25):
        n = int(torch.randint(1, 3 * C, (1,), generator=gen))
        # zipf-ish ids so that hits, misses, duplicates and ties all occur
        ids = (torch.rand(n, generator=gen) ** 3 * N).long().clamp_(0, N - 1)
        if torch.unique(omgr.idx_map[ids]).numel() > C:
            ids = ids[: C // 2]
        slots = mgr.prepare_ids(ids.cuda())
        oslots = omgr.prepare_ids(ids)
        assert torch.equal(slots.cpu(), oslots), f"slot ids differ at step {step}"
        assert_maps_equal(mgr, omgr)
        assert mgr.num_hits_history == omgr.num_hits_history
        assert mgr.num_miss_history == omgr.num_miss_history
        assert mgr.num_write_back_history == omgr.num_write_back_history
        torch.testing.assert_close(mgr.cuda_cached_weight.detach().cpu(), omgr.cuda_cached_weight.detach(),
                                   rtol=0, atol=0)
    assert sum(mgr.num_write_back_history) > 0, "the scenario must exercise eviction"
    assert mgr._cache_miss == omgr._cache_miss and mgr._total_cache == omgr._total_cache
    mgr.flush()
    omgr.flush()
    assert_maps_equal(mgr, omgr)
    torch.testing.assert_close(mgr.weight, omgr.weight, rtol=0, atol=0)


def test_capacity_overflow_leaves_cache_untouched():
    ce = _mods()
    model = ce.CachedEmbeddingBag(100, 4, cache_ratio=0.05, evict_strategy=ce.EvictionStrategy.LFU, warmup_ratio=0.4)
    mgr = model.cache_weight_mgr
    before = (mgr.cached_idx_map.clone(), mgr.inverted_cached_idx.clone(), mgr.freq_cnter.clone())
    with pytest.raises(AssertionError, match="increase cuda_row_num or decrease the training batch size"):
        mgr.prepare_ids(torch.arange(50, 56).cuda())
    assert torch.equal(before[0], mgr.cached_idx_map) and torch.equal(before[1], mgr.inverted_cached_idx)
    assert torch.equal(before[2], mgr.freq_cnter)
    # and the manager still works afterwards
    s = mgr.prepare_ids(torch.tensor([50, 51, 50]).cuda())
    assert s[0] == s[2] and s[0] != s[1]
    with pytest.raises(IndexError):
        mgr.prepare_ids(torch.tensor([1, 100]).cuda())
    s = mgr.prepare_ids(torch.tensor([50, 3]).cuda())
    assert int(mgr.cached_idx_map[s[1]]) == 3


def test_cachemgr_kat():  # upstream B.1
    ce = _mods()
    model = torch.nn.EmbeddingBag(10000, 128)
    mgr = ce.CachedParamMgr(model.weight.detach().clone(), 5)
    assert mgr.cuda_row_num == 5
    mgr._admit(1)
    assert not mgr._row_in_cuda(2)
    assert mgr._row_in_cuda(1)
    mgr._admit(8)
    assert mgr.cuda_available_row_num == 3
    mgr._evict()
    assert mgr.cuda_available_row_num == 4
    mgr._prepare_rows_on_cuda(torch.tensor([9, 6, 5], dtype=torch.long, device=0))
    mgr._prepare_rows_on_cuda(torch.tensor([3, 4, 5], dtype=torch.long, device=0))
    assert mgr.cuda_available_row_num == 0
    assert int((mgr.cached_idx_map == 5).sum()) == 1
    torch.testing.assert_close(mgr.cuda_cached_weight[mgr.inverted_cached_idx[5]].cpu(), model.weight[5].detach())
    mgr.flush()
    assert mgr.cuda_available_row_num == 5
    assert torch.all(mgr.cached_idx_map == -1) and torch.all(mgr.inverted_cached_idx == -1)
    torch.testing.assert_close(mgr.weight, model.weight.detach(), rtol=0, atol=0)


def test_reorder_with_freq_kat():  # upstream B.2
    ce = _mods()
    num_embed, num_chunks = 100, 5
    g = torch.Generator().manual_seed(3)
    freq = torch.randint(10000, size=(num_embed,), generator=g)
    sorted_idx = torch.argsort(freq, descending=True, stable=True).tolist()
    mgr = ce.CachedParamMgr(torch.rand(num_embed, 2), num_chunks)
    mgr.reorder(freq)
    got = mgr.idx_map.cpu().tolist()
    assert got == [sorted_idx.index(i) for i in range(num_embed)]


@pytest.mark.parametrize("init_freq", [True, False])
def test_lfu_strategy_kat(init_freq):  # upstream B.4
    ce = _mods()
    Bag = ce.CachedEmbeddingBag(5, 5, cache_ratio=3 / 5, buffer_size=0, pin_weight=True,
                                ids_freq_mapping=[4, 2, 1, 3, 1] if init_freq else None, warmup_ratio=1.0,
                                evict_strategy=ce.EvictionStrategy.LFU)
    offsets = torch.tensor([0], device="cuda:0")
    seq = [[2], [1, 2], [0, 2]] + [[0, 1, 2]] * 4 + [[0, 2]] * 4 + [[0]] * 4 + \
          [[0, 1, 2], [0, 1, 2], [3], [2], [4], [2], [0]]
    for ids in seq:
        Bag.forward(torch.tensor(ids, device="cuda:0"), offsets)
    assert Bag.cache_weight_mgr.num_hits_history[-6:] == [3, 0, 1, 0, 1, 1]


# ---------------------------------------------------------------------------------------------------- end to end
@pytest.mark.parametrize("use_LFU", [True, False])
@pytest.mark.parametrize("backward", ["sparse", "fused"])
def test_freq_aware_embed_matches_full_table(use_LFU, backward):  # upstream B.3, against torch.nn.EmbeddingBag
    ce = _mods()
    NUM_EMBED, EMBED_DIM, BATCH = 10, 8, 8
    gen = torch.Generator().manual_seed(4)
    strategy = ce.EvictionStrategy.LFU if use_LFU else ce.EvictionStrategy.DATASET
    model = ce.CachedEmbeddingBag(NUM_EMBED, EMBED_DIM, mode='mean', include_last_offset=True, sparse=True,
                                  cache_ratio=min(BATCH * 2 / NUM_EMBED, 1.0), ids_freq_mapping=None,
                                  evict_strategy=strategy)
    assert model.weight.shape[0] == NUM_EMBED
    ref_model = torch.nn.EmbeddingBag.from_pretrained(model.weight.detach().clone().cuda(), mode='mean',
                                                      include_last_offset=True, freeze=False)
    lr = 1e-3
    if backward == "fused":
        model.set_fused_optimizer("sgd", lr=lr)
    optimizer = torch.optim.SGD(model.parameters(), lr=lr)
    ref_optimizer = torch.optim.SGD(ref_model.parameters(), lr=lr)
    for i in range(5):
        n = BATCH * 2
        indices = torch.randint(0, NUM_EMBED, (n,), generator=gen).cuda()
        cuts = torch.sort(torch.randint(1, n, (BATCH - 1,), generator=gen)).values
        offsets = torch.cat([torch.tensor([0]), cuts, torch.tensor([n])]).cuda()
        res = model(indices, offsets)
        ref_res = ref_model(indices, offsets)
        torch.testing.assert_close(res, ref_res, rtol=RTOL, atol=ATOL)
        grad = torch.rand(res.shape, generator=gen).cuda()
        res.backward(grad)
        ref_res.backward(grad)
        optimizer.step(); optimizer.zero_grad()
        ref_optimizer.step(); ref_optimizer.zero_grad()
    model.cache_weight_mgr.flush()
    torch.testing.assert_close(model.weight.detach().cuda(), ref_model.weight.detach(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("strategy", ["LFU", "DATASET"])
def test_training_under_eviction_matches_oracle(strategy):
    """Several look-ahead windows with fwd + bwd + SGD per batch, cache much smaller than the table: pooled sums,
    slot ids and the final host table all match the oracle (which itself matches full-table SGD)."""
    ce = _mods()
    gen = torch.Generator().manual_seed(21)
    N, D, F, B, P = 4000, 128, 4, 32, 3
    weight = torch.randn(N, D, generator=gen) * 0.01
    freq = torch.randint(0, 100, (N,), generator=gen)
    kw = dict(mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.1, ids_freq_mapping=freq,
              warmup_ratio=0.7)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(ce.EvictionStrategy, strategy), **kw)
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(OStrategy, strategy), **kw)
    model.set_fused_optimizer("sgd", lr=0.5)
    oopt = torch.optim.SGD(omodel.parameters(), lr=0.5)
    offsets = torch.arange(F * B + 1)
    for window in range(6):
        batches = [(torch.rand(F * B, generator=gen) ** 2 * N).long().clamp_(0, N - 1) for _ in range(P)]
        slots = model.cache_weight_mgr.prepare_ids(torch.cat(batches).cuda())
        oslots = omodel.cache_weight_mgr.prepare_ids(torch.cat(batches))
        assert torch.equal(slots.cpu(), oslots)
        model.set_cache_op(False); omodel.set_cache_op(False)
        for s, os_ in zip(torch.chunk(slots, P), torch.chunk(oslots, P)):
            out = model(s, offsets.cuda())
            oout = omodel(os_, offsets)
            torch.testing.assert_close(out.cpu(), oout, rtol=RTOL, atol=ATOL)
            grad = torch.randn(out.shape, generator=gen)
            out.backward(grad.cuda())
            oout.backward(grad)
            oopt.step(); oopt.zero_grad()
    assert sum(model.num_write_back_history) > 0
    assert_maps_equal(model.cache_weight_mgr, omodel.cache_weight_mgr)
    model.cache_weight_mgr.flush(); omodel.cache_weight_mgr.flush()
    torch.testing.assert_close(model.weight, omodel.weight, rtol=RTOL, atol=ATOL)


def test_rowwise_adagrad_state_travels_with_rows():
    """Row-wise Adagrad through the module with evictions: matches the float64 full-table restatement."""
    ce = _mods()
    gen = torch.Generator().manual_seed(8)
    N, D, G = 300, 16, 40
    weight = torch.randn(N, D, generator=gen)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True, sparse=True,
                                  cache_ratio=0.2, warmup_ratio=0.5, evict_strategy=ce.EvictionStrategy.LFU,
                                  fused_optimizer="rowwise_adagrad", lr=0.05, eps=1e-8)
    W, m = weight.numpy().astype(np.float64), np.zeros(N)
    offsets = torch.arange(G + 1)
    for _ in range(12):
        ids = torch.randint(0, N, (G,), generator=gen)
        grad = torch.randn(G, D, generator=gen)
        out = model(ids.cuda(), offsets.cuda())
        out.backward(grad.cuda())
        W, m = rowwise_adagrad_reference(W, m, ids.numpy(), offsets.numpy(), grad.numpy(), 0.05, 1e-8)
    assert sum(model.num_write_back_history) > 0
    model.cache_weight_mgr.flush()
    close(model.weight.double(), torch.from_numpy(W))
    close(model.cache_weight_mgr.row_state.double(), torch.from_numpy(m))


def test_module_surface():
    ce = _mods()
    model = ce.CachedEmbeddingBag(50, 8, sparse=True, include_last_offset=True,
                                  evict_strategy=ce.EvictionStrategy.DATASET, cache_ratio=0.5).to(torch.device("cuda:0"))
    params = list(model.parameters())
    assert len(params) == 1 and params[0] is model.cache_weight_mgr.cuda_cached_weight
    assert [n for n, _ in model.named_parameters()] == ["weight"]
    assert model.weight.device.type == "cpu" and model.weight.shape == (50, 8)
    assert model.element_size() == 4
    assert sum(b.numel() for b in model.buffers()) > 0
    model.set_cache_mgr_async_copy(True)
    model.zero_grad()
    model.print_comm_stats_()


# ---------------------------------------------------------------------------------------------------- look-ahead overlap
@pytest.mark.parametrize("strategy", ["LFU", "DATASET"])
@pytest.mark.parametrize("early,stage_rows,dma", [(True, 0, True), (False, 0, True), (True, 7, True), (True, 0, False),
                                                   (True, 0, "slow"), ("staged", 0, True)])
def test_lookahead_prefetcher_matches_oracle_with_two_window_protection(strategy, early, stage_rows, dma, monkeypatch):
    """prepare_ids of window k+1 on a side stream while window k trains: slot ids and maps bit-exact against the
    oracle run with the same two-window protection; pooled sums / final table within 1e-5 of it.
    early: window k+1 is submitted right after the FIRST step of window k (the order bench.py uses), otherwise after its
    last step.  stage_rows = 7: most victims do not fit the staging buffer and are written back straight from their
    slots before the fill.  dma: parked victims leave through the copy engine + host threads ("slow": every host
    scatter is delayed by 20 ms, so rows that are missed again are served from the staging buffer or not at all).
    early = "staged": the ids of window k+2 leave pinned host memory on the driver's ids stream (pf.stage) while window
    k trains -- behind the fill of window k+1, the order bench.py's end-to-end arm uses -- and window k+1 is submitted
    from its staged ids (three ring buffers, each reused twice here)."""
    ce = _mods()
    if dma == "slow":
        monkeypatch.setenv("CEBAG_WB_DELAY_US", "20000")
    gen = torch.Generator().manual_seed(33)
    N, D, F, B, P = 6000, 128, 4, 64, 2
    weight = torch.randn(N, D, generator=gen) * 0.01
    freq = torch.randint(0, 100, (N,), generator=gen)
    kw = dict(mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.25, ids_freq_mapping=freq,
              warmup_ratio=0.7)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(ce.EvictionStrategy, strategy), **kw)
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(OStrategy, strategy), **kw)
    omodel.cache_weight_mgr.protect_windows = 2
    model.set_fused_optimizer("sgd", lr=0.5)
    model.set_cache_op(False); omodel.set_cache_op(False)
    oopt = torch.optim.SGD(omodel.parameters(), lr=0.5)
    offsets = torch.arange(F * B + 1)
    windows = [[(torch.rand(F * B, generator=gen) ** 2 * N).long().clamp_(0, N - 1) for _ in range(P)] for _ in range(8)]
    grads = [[torch.randn(F * B, D, generator=gen) for _ in range(P)] for _ in range(8)]
    model.cache_weight_mgr.stage_rows = stage_rows
    model.cache_weight_mgr.dma_writeback = bool(dma)
    pf = ce.LookaheadPrefetcher(model)
    assert model.cache_weight_mgr.protect_windows == 2
    staged = {}
    if early == "staged":
        pinned = [[w.pin_memory() for w in win] for win in windows]
        staged = {0: pf.stage(pinned[0]), 1: pf.stage(pinned[1])}
        h = pf.submit(staged.pop(0), offsets=offsets.cuda())
    else:
        h = pf.submit([w.pin_memory() for w in windows[0]], offsets=offsets.cuda())   # host ids: H2D rides the side stream
    for k in range(len(windows)):
        slots = h.wait()
        oslots = omodel.cache_weight_mgr.prepare_ids(torch.cat(windows[k]))
        outs = []
        for j, (s, g) in enumerate(zip(torch.chunk(slots, P), grads[k])):
            out = model(s, offsets.cuda())
            out.backward(g.cuda())
            outs.append(out)
            if early == "staged" and j == 0:
                if k + 1 < len(windows):
                    h = pf.submit(staged.pop(k + 1), offsets=offsets.cuda())
                if k + 2 < len(windows):      # held back on the device until the fill of window k+1 has finished
                    staged[k + 2] = pf.stage(pinned[k + 2], after_last_fill=True)
            elif early and j == 0 and k + 1 < len(windows):
                h = pf.submit([w.cuda() for w in windows[k + 1]], offsets=offsets.cuda())
        pf.window_enqueued()
        if not early and k + 1 < len(windows):
            h = pf.submit([w.cuda() for w in windows[k + 1]], offsets=offsets.cuda())
        assert torch.equal(slots.cpu(), oslots), f"slot ids differ in window {k}"
        for s, g, out in zip(torch.chunk(oslots, P), grads[k], outs):
            oout = omodel(s, offsets)
            close(out.cpu(), oout.detach())
            oout.backward(g)
            oopt.step(); oopt.zero_grad()
    pf.close()
    assert model.cache_weight_mgr.protect_windows == 1
    assert sum(model.num_write_back_history) > 0
    # the oracle is one prepare behind in wall-clock order only; the maps agree once both have seen every window
    assert_maps_equal(model.cache_weight_mgr, omodel.cache_weight_mgr)
    model.cache_weight_mgr.flush(); omodel.cache_weight_mgr.flush()
    close(model.weight, omodel.weight)


def test_planned_backward_is_bitwise_identical_to_inline():
    """cebag_bag_backward_plan on a side stream + backward with workspace_has_plan gives the same bits as the inline path."""
    ce = _mods()
    gen = torch.Generator().manual_seed(12)
    N, D, G = 3000, 128, 5000
    weight = torch.randn(N, D, generator=gen)
    res = []
    for planned in (False, True):
        model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True, sparse=True,
                                      cache_ratio=1.0, warmup_ratio=1.0, evict_strategy=ce.EvictionStrategy.DATASET,
                                      fused_optimizer="sgd", lr=0.3)
        model.set_cache_op(False)
        g2 = torch.Generator().manual_seed(5)
        ids = torch.randint(0, N, (G,), generator=g2).cuda()
        offsets = torch.arange(G + 1).cuda()
        grad = torch.randn(G, D, generator=g2).cuda()
        slots = model.cache_weight_mgr.prepare_ids(ids)
        if planned:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                assert model.plan_backward(slots, offsets)
            torch.cuda.current_stream().wait_stream(side)
            assert len(model._bwd_plans) == 1
        out = model(slots, offsets)
        out.backward(grad)
        if planned:
            assert len(model._bwd_plans) == 0, "the plan was not consumed"
        res.append(model.cache_weight_mgr.cuda_cached_weight.detach().cpu())
    assert torch.equal(res[0], res[1])


def test_window_plan_is_bitwise_identical_to_inline():
    """cebag_bag_backward_plan_window (ONE radix sort over (batch, slot) for several ragged batches, one of them empty)
    followed by backward with workspace_has_plan = 2 gives the same bits as the inline per-batch sort."""
    ce = _mods()
    gen = torch.Generator().manual_seed(19)
    N, D = 2500, 128
    weight = torch.randn(N, D, generator=gen)
    batches = []
    for G in (700, 0, 1300, 257, 900):
        ids, offsets = make_bags(N, D, G, 4, gen) if G else (torch.zeros(0, dtype=torch.long), torch.zeros(1, dtype=torch.long))
        batches.append((ids, offsets, torch.randn(G, D, generator=gen)))
    res = []
    for planned in (False, True):
        model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), mode="sum", include_last_offset=True, sparse=True,
                                      cache_ratio=1.0, warmup_ratio=1.0, evict_strategy=ce.EvictionStrategy.DATASET,
                                      fused_optimizer="sgd", lr=0.3, padding_idx=3)
        model.set_cache_op(False)
        slots = model.cache_weight_mgr.prepare_ids(torch.cat([b[0] for b in batches]).cuda())
        chunks = list(torch.split(slots, [b[0].numel() for b in batches]))
        offs = [b[1].cuda() for b in batches]
        if planned:
            assert model.plan_backward_window(chunks, offs)
            assert len(model._bwd_plans) == len([c for c in chunks if c.numel()]) or len(model._bwd_plans) >= 3
        for chunk, off, (_, _, grad) in zip(chunks, offs, batches):
            if chunk.numel() == 0:
                continue
            out = model(chunk, off)
            out.backward(grad.cuda())
        res.append(model.cache_weight_mgr.cuda_cached_weight.detach().cpu())
    assert torch.equal(res[0], res[1])
    assert not torch.equal(res[0], weight)


def test_two_window_protection_capacity_error():
    ce = _mods()
    model = ce.CachedEmbeddingBag(100, 4, cache_ratio=0.1, warmup_ratio=0.0, evict_strategy=ce.EvictionStrategy.LFU)
    mgr = model.cache_weight_mgr
    mgr.protect_windows = 2
    mgr.prepare_ids(torch.arange(0, 6).cuda())
    mgr.prepare_ids(torch.arange(10, 14).cuda())             # fills the cache: 6 + 4 rows
    with pytest.raises(AssertionError, match="increase cuda_row_num or decrease the training batch size"):
        mgr.prepare_ids(torch.arange(20, 27).cuda())         # needs 7 victims, only 6 are unprotected
    assert int((mgr.cached_idx_map >= 0).sum()) == 10
    s = mgr.prepare_ids(torch.arange(20, 25).cuda())         # 5 victims fit
    assert torch.equal(mgr.cached_idx_map[s].cpu(), torch.arange(20, 25))


def test_rejected_call_leaves_no_trace_with_two_window_protection():
    """A window the device rejects (capacity, id out of range) changes nothing: maps, counters, stamps and the
    protected windows are what they were, so the retry sees the same evictable set as the oracle."""
    ce = _mods()
    gen = torch.Generator().manual_seed(4)
    N, D = 400, 8
    weight = torch.randn(N, D, generator=gen)
    kw = dict(mode="sum", include_last_offset=True, cache_ratio=0.1, warmup_ratio=0.0)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=ce.EvictionStrategy.LFU, **kw)
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=OStrategy.LFU, **kw)
    mgr, omgr = model.cache_weight_mgr, omodel.cache_weight_mgr
    mgr.protect_windows = omgr.protect_windows = 2
    calls = [torch.arange(0, 20), torch.arange(15, 35), torch.arange(100, 130), torch.arange(30, 45),
             torch.tensor([3, N + 5]), torch.arange(200, 215), torch.arange(0, 10)]
    for ids in calls:
        before = (mgr.cached_idx_map.clone(), mgr.inverted_cached_idx.clone(), mgr.freq_cnter.clone(),
                  mgr._slot_epoch.clone(), mgr._dev_state.clone(), list(mgr.num_hits_history))
        if ids.numel() == 30:                      # 30 new rows while 35 slots are occupied and 20 of them protected
            with pytest.raises(AssertionError, match="increase cuda_row_num or decrease the training batch size"):
                mgr.prepare_ids(ids.cuda())
        elif int(ids.max()) >= N:
            with pytest.raises(IndexError):
                mgr.prepare_ids(ids.cuda())
        else:
            assert torch.equal(mgr.prepare_ids(ids.cuda()).cpu(), omgr.prepare_ids(ids))
            assert_maps_equal(mgr, omgr)
            continue
        after = (mgr.cached_idx_map, mgr.inverted_cached_idx, mgr.freq_cnter, mgr._slot_epoch, mgr._dev_state)
        for b, a in zip(before, after):
            b[_lib_state_mask(b)] = 0
            a = a.clone()
            a[_lib_state_mask(a)] = 0
            assert torch.equal(b, a)
        assert before[5] == mgr.num_hits_history
        assert int(mgr._miss_bitmap.abs().sum()) == 0 and int(mgr._hit_flags.sum()) == 0
    assert sum(mgr.num_write_back_history) > 0


def _lib_state_mask(tensor):
    """dev_state counts the calls the device has completed, rejected ones included: ignore that word."""
    from cachedembedding_b200 import _lib
    mask = torch.zeros_like(tensor, dtype=torch.bool)
    if tensor.numel() == _lib.STATE_WORDS and tensor.dtype == torch.int64:
        mask[_lib.STATE_CALLS] = True
    return mask


@pytest.mark.parametrize("strategy", ["LFU", "DATASET"])
@pytest.mark.parametrize("stage_rows,adagrad,dma", [(0, False, True), (5, False, True), (0, True, True), (0, False, False),
                                                    (0, True, "slow")])
def test_reference_loop_with_async_copy_flag(strategy, stage_rows, adagrad, dma, monkeypatch):
    """The reference's loop, unchanged (/root/reference/recsys/dlrm_main.py:245-279 with the flag of :121,354):
    set_cache_mgr_async_copy(True); one prepare_ids over the concatenated window on the CURRENT stream; then P forward /
    backward steps with cache_op off.  The row traffic runs on the manager's copy stream (victims parked in HBM, fill,
    write-back under the following steps); slot ids, maps and counters stay bit-exact against the oracle, pooled sums
    and the flushed table within 1e-5."""
    ce = _mods()
    gen = torch.Generator().manual_seed(21)
    N, D, F, B, P = 5000, 128, 4, 48, 3
    weight = torch.randn(N, D, generator=gen) * 0.01
    freq = torch.randint(0, 100, (N,), generator=gen)
    kw = dict(mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.2, ids_freq_mapping=freq, warmup_ratio=0.7)
    model = ce.CachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(ce.EvictionStrategy, strategy),
                                  fused_optimizer="rowwise_adagrad" if adagrad else "sgd", lr=0.5, eps=1e-6, **kw)
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=getattr(OStrategy, strategy), **kw)
    model.set_cache_mgr_async_copy(True)
    mgr, omgr = model.cache_weight_mgr, omodel.cache_weight_mgr
    mgr.stage_rows = stage_rows
    mgr.dma_writeback = bool(dma)
    if dma == "slow":
        monkeypatch.setenv("CEBAG_WB_DELAY_US", "20000")
    oopt = torch.optim.SGD(omodel.parameters(), lr=0.5)
    offsets = torch.arange(F * B + 1)
    state = np.zeros(N)
    ref_w = weight.double().numpy().copy()
    for k in range(7):
        window = [(torch.rand(F * B, generator=gen) ** 2 * N).long().clamp_(0, N - 1) for _ in range(P)]
        slots = mgr.prepare_ids(torch.cat([w.cuda() for w in window]))           # dlrm_main.py:259
        oslots = omgr.prepare_ids(torch.cat(window))
        assert torch.equal(slots.cpu(), oslots), f"slot ids differ in window {k}"
        assert_maps_equal(mgr, omgr)
        model.set_cache_op(False); omodel.set_cache_op(False)
        for j, (s, os_) in enumerate(zip(torch.chunk(slots, P), torch.chunk(oslots, P))):
            g = torch.randn(F * B, D, generator=gen)
            out = model(s, offsets.cuda())
            out.backward(g.cuda())
            if adagrad:
                want = torch.from_numpy(ref_w[omgr.idx_map[window[j]].numpy()]).float()
                close(out.cpu(), want)
                ref_w, state = rowwise_adagrad_reference(ref_w, state, omgr.idx_map[window[j]].numpy(), offsets.numpy(),
                                                         g.double().numpy(), 0.5, 1e-6)
            else:
                oout = omodel(os_, offsets)
                close(out.cpu(), oout.detach())
                oout.backward(g)
                oopt.step(); oopt.zero_grad()
    assert sum(mgr.num_write_back_history) > 0 and mgr._own_copy_stream is not None
    mgr.flush(); omgr.flush()
    if adagrad:
        close(mgr.weight, torch.from_numpy(ref_w).float())
        close(mgr.row_state, torch.from_numpy(state).float())
    else:
        close(mgr.weight, omgr.weight)
        assert_maps_equal(mgr, omgr)


def test_integration_stub_runs_verbatim():
    """The binding stub of INTEGRATION.md, executed as printed: a cached table driven through the C ABI alone
    (prepare_ids -> forward -> fused backward -> flush), against torch.nn.functional.embedding_bag + SGD."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"<!-- stub:begin -->\s*```python\n(.*?)```\s*<!-- stub:end -->", text, flags=re.S).group(1)
    ns = {}
    cwd = os.getcwd()
    os.chdir(root)                      # the stub loads the library by its repo-relative path
    try:
        exec(compile(block, "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    gen = torch.Generator().manual_seed(17)
    N, D, C, G = 3000, 128, 400, 300
    host = (torch.randn(N, D, generator=gen) * 0.1).pin_memory()
    ref = host.clone()
    t, arrays = ns["make_table"](host, C)
    pinned = torch.zeros(64, dtype=torch.int64).pin_memory()
    offsets = torch.arange(G + 1)
    evicted = 0
    for _ in range(5):
        ids = torch.randint(0, N, (G,), generator=gen)
        grad = torch.randn(G, D, generator=gen)
        out, stats = ns["train_step"](t, arrays, pinned, ids.cuda(), offsets.cuda(), grad.cuda(), 0.25)
        want = torch.nn.functional.embedding_bag(ids, ref, offsets, mode="sum", include_last_offset=True)
        close(out.cpu(), want)
        ref.index_add_(0, ids, grad, alpha=-0.25)
        assert stats.unique_hits + stats.unique_misses == int(torch.unique(ids).numel())
        evicted += stats.evicted
    assert evicted > 0
    assert ns["flush"](t, pinned) == C
    close(host, ref)


# ---------------------------------------------------------------------------------------------------- BASELINE configs
def test_baseline_config0_plumbing_matches_oracle():
    """BASELINE.json configs[0]: synthetic Kaggle-shape DLRM -- 26 tables x 1e5 rows, dim 16, batch 512 -- a few
    training steps of the FreqAwareEmbeddingBag against the CPU oracle (exact maps, 1e-5 values)."""
    ce = _mods()
    import bench
    F, rows, D, B = 26, 100_000, 16, 512
    N = F * rows
    gen = torch.Generator().manual_seed(1024)
    weight = torch.empty(N, D).uniform_(-1.0 / N, 1.0 / N, generator=gen)
    rows_t = torch.full((F,), rows, dtype=torch.long)
    kw = dict(mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.01, warmup_ratio=0.7)
    model = ce.FreqAwareEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=ce.EvictionStrategy.LFU, **kw)
    omodel = OracleCachedEmbeddingBag(N, D, _weight=weight.clone(), evict_strategy=OStrategy.LFU, **kw)
    model.set_fused_optimizer("sgd", lr=1.0)
    oopt = torch.optim.SGD(omodel.parameters(), lr=1.0)
    offsets = torch.arange(F * B + 1)
    for step in range(6):
        ids = bench.sample_ids(rows_t, B, gen, "cpu")
        grad = torch.randn(F * B, D, generator=gen) * 1e-3
        out = model(ids.cuda(), offsets.cuda())
        oout = omodel(ids, offsets)
        close(out.cpu(), oout.detach())
        out.backward(grad.cuda())
        oout.backward(grad)
        oopt.step(); oopt.zero_grad()
        assert_maps_equal(model.cache_weight_mgr, omodel.cache_weight_mgr)
    model.cache_weight_mgr.flush(); omodel.cache_weight_mgr.flush()
    close(model.weight, omodel.weight)


def test_baseline_config1_kaggle_full_size_properties():
    """BASELINE.json configs[1]: Criteo-Kaggle shape at FULL size -- 33,762,577 rows x 128 (17.3 GB pinned host table,
    GPU-initialised), cache_ratio 0.01, batch 4096 -- through size-independent properties."""
    ce = _mods()
    import bench
    rows = bench.CRITEO_KAGGLE_ROWS
    N, D, B, F = sum(rows), 128, 4096, len(rows)
    model = ce.CachedEmbeddingBag(N, D, mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.01,
                                  warmup_ratio=0.7, evict_strategy=ce.EvictionStrategy.LFU, init_seed=7)
    mgr = model.cache_weight_mgr
    assert mgr.cuda_row_num == 337_625 and tuple(model.weight.shape) == (N, D)
    # GPU-side uniform init: right range, not constant
    sample = model.weight[:: N // 1000]
    assert float(sample.abs().max()) <= 1.0 / N and float(sample.std()) > 0
    flat = model.weight.view(-1)
    span = 4096 * 8192                    # rows 0 .. 262143: the hottest rows of the first tables live here
    checksum = flat[:span].view(4096, 8192).sum(1, dtype=torch.float64).clone()
    tail_checksum = flat[-span:].view(4096, 8192).sum(1, dtype=torch.float64).clone()
    rows_t = torch.tensor(rows, device="cuda")
    gen = torch.Generator(device="cuda").manual_seed(3)
    offsets = torch.arange(F * B + 1, device="cuda")
    model.set_fused_optimizer("sgd", lr=1.0)
    touched = []
    for it in range(6):
        window = [bench.sample_ids(rows_t, B, gen, "cuda") for _ in range(4)]
        slots = mgr.prepare_ids(torch.cat(window))
        assert torch.equal(mgr._slot2row[slots].long(), torch.cat(window))        # id -> slot -> row round trip
        occ = mgr._slot2row[mgr._slot2row >= 0]
        assert occ.unique().numel() == occ.numel()                               # every row resident once
        model.set_cache_op(False)
        for ids, s in zip(window, torch.chunk(slots, 4)):
            out = model(s, offsets)
            assert torch.isfinite(out).all()
            out.backward(torch.zeros_like(out))                                  # zero gradient: tables must not move
        touched.append(torch.cat(window))
    mgr.flush()
    assert torch.equal(flat[:span].view(4096, 8192).sum(1, dtype=torch.float64), checksum)
    assert torch.equal(flat[-span:].view(4096, 8192).sum(1, dtype=torch.float64), tail_checksum)
    assert mgr.cuda_available_row_num == mgr.cuda_row_num


# ---------------------------------------------------------------------------------------------------- full-size properties
def test_large_table_round_trip_properties():
    """BASELINE-scale shapes through size-independent properties (the oracle would take minutes here):
    (1) cached_idx_map[slot_ids] == idx_map[ids]; (2) every id of the window is resident exactly once;
    (3) prepare -> flush without updates leaves the host table bit-identical (checksum of checksums);
    (4) fused SGD with lr=1 and grad g moves the flushed rows by exactly -count*g for a constant grad."""
    ce = _mods()
    N, D, B, F = 4_000_000, 128, 65536, 4
    model = ce.CachedEmbeddingBag(N, D, mode="sum", include_last_offset=True, sparse=True, cache_ratio=0.05,
                                  warmup_ratio=0.7, evict_strategy=ce.EvictionStrategy.DATASET)
    mgr = model.cache_weight_mgr
    before = model.weight.view(-1, 1024).sum(1).clone()
    gen = torch.Generator(device="cuda").manual_seed(1)
    for it in range(4):
        ids = (torch.rand(B * F, generator=gen, device="cuda") ** 4 * N).long().clamp_(0, N - 1)
        slots = mgr.prepare_ids(ids)
        assert torch.equal(mgr._slot2row[slots].long(), ids)
        occ = mgr._slot2row[mgr._slot2row >= 0]
        assert occ.unique().numel() == occ.numel()
        assert mgr.cuda_available_row_num + occ.numel() == mgr.cuda_row_num
    assert sum(mgr.num_write_back_history) > 0
    mgr.flush()
    assert torch.equal(model.weight.view(-1, 1024).sum(1), before)
    # constant-grad linearity
    model.set_fused_optimizer("sgd", lr=1.0)
    ids = (torch.rand(B * F, generator=gen, device="cuda") ** 4 * N).long().clamp_(0, N - 1)
    w0 = model.weight[ids.cpu()[:1000]].clone()
    counts = torch.bincount(ids, minlength=N)[ids[:1000]].cpu().float()
    out = model(ids, torch.arange(B * F + 1, device="cuda"))
    out.backward(torch.full_like(out, 0.25))
    mgr.flush()
    torch.testing.assert_close(model.weight[ids.cpu()[:1000]], w0 - 0.25 * counts[:, None], rtol=1e-5, atol=1e-6)
