"""Untimed parity leg of bench.py for N > 1 ranks (and of tests/test_gpu_multi.py): the parallel bags against the CPU oracle.

The driver's `pytest -m gpu` box has one GPU, so the multi-GPU paths are checked where the driver does run them: every
`bench.py --gpus N` (N > 1) first runs `verify_parallel` on small row-scaled tables and fails the run on a mismatch.

Per rank, with evictions in every step:
  table-wise (reference ctor recsys/models/dlrm.py:53-68)
    * model A: NCCL all-to-all (collectives.dual_all_to_all_tablewise), model B: exchange fused into the fwd/bwd kernels
      over NVLink peer memory -- pooled outputs of A and B must be BIT-EQUAL, and within 1e-5 of the oracle world's;
    * slot maps (cached_idx_map / inverted_cached_idx / LFU counters) of A and B bit-exact against the oracle bag of
      this rank after every step; host tables after flush within 1e-5 of the oracle's (A and B bit-equal);
  column-wise (A.5; reference ctor recsys/models/dlrm.py:70-81)
    * ParallelCachedEmbeddingBag against OracleColumnwiseWorld: outputs, maps, flushed column shard.
The oracle is the checker here, never the thing measured (it runs on the CPU, from the same seeds on every rank).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RTOL = 1e-5


def _close(got: torch.Tensor, want: torch.Tensor) -> float:
    """max |got - want| relative to the 1e-5 bar (<= 1 passes): |d| <= 1e-5 |want| + 1e-5 max|want|."""
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    bar = RTOL * want.abs() + RTOL * max(float(want.abs().max()), 1e-30)
    return float(((got - want).abs() / bar).max()) if want.numel() else 0.0


def _maps_equal(mgr, omgr) -> bool:
    ok = torch.equal(mgr.cached_idx_map.cpu(), omgr.cached_idx_map)
    ok = ok and torch.equal(mgr.inverted_cached_idx.cpu(), omgr.inverted_cached_idx)
    ok = ok and mgr.cuda_available_row_num == omgr.cuda_available_row_num
    if hasattr(omgr, "freq_cnter"):
        ok = ok and torch.equal(mgr.freq_cnter.cpu(), omgr.freq_cnter)
    return bool(ok)


def verify_parallel(rows, arrange, dim: int, steps: int = 6, seed: int = 77, lr: float = 0.5, group=None) -> dict:
    """`rows[t]` rows of table t (already small), `arrange[t]` its rank.  Returns the parity record of THIS rank merged
    over all ranks (ok = every check on every rank)."""
    import cachedembedding_b200 as ce
    from oracle import EvictionStrategy as OS
    from oracle import OracleColumnwiseWorld, OracleTablewiseConfig, OracleTablewiseWorld

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    F = len(rows)
    B = 8 * world + 3                                   # not divisible by the world: remainder to the low ranks
    gen = torch.Generator().manual_seed(seed)           # same stream of numbers on every rank
    weights = [torch.randn(n, dim, generator=gen) * 0.05 for n in rows]
    goff = torch.cumsum(torch.tensor([0] + list(rows)), 0)
    strides = [B // world + int(i < B % world) for i in range(world)]
    begin = sum(strides[:rank])
    # small caches: every rank must hold one batch of its tables, and evicts in every step after the first
    need = []
    for q in range(world):
        tabs = [t for t, r in enumerate(arrange) if r == q]
        need.append((len(tabs) * B + 4) / sum(rows[t] for t in tabs))
    ratio = min(1.0, max(max(need), 0.05))
    rec = {"world": world, "steps": steps, "batch": B, "tables": F, "dim": dim, "cache_ratio": round(ratio, 4)}
    ok = True
    worst = 0.0

    # ---- table-wise: NCCL exchange (A) and fused exchange (B) vs the oracle world ----------------------------------
    def tw_model(fused_exchange: bool):
        cfgs = [ce.TablewiseEmbeddingBagConfig(rows[t], 0, assigned_rank=arrange[t], initial_weight=weights[t].clone())
                for t in range(F)]
        m = ce.ParallelCachedEmbeddingBagTablewise(cfgs, dim, sparse=True, mode="sum", include_last_offset=True,
                                                   cache_ratio=ratio, warmup_ratio=0.7, buffer_size=0,
                                                   evict_strategy=ce.EvictionStrategy.LFU, fused_optimizer="sgd", lr=lr,
                                                   process_group=group)
        m.enable_fused_exchange(fused_exchange)
        return m

    model_a, model_b = tw_model(False), tw_model(True)
    ocfgs = [OracleTablewiseConfig(rows[t], 0, assigned_rank=arrange[t], initial_weight=weights[t].clone())
             for t in range(F)]
    oworld = OracleTablewiseWorld(ocfgs, dim, world, mode="sum", include_last_offset=True, sparse=True,
                                  cache_ratio=ratio, warmup_ratio=0.7, evict_strategy=OS.LFU)
    oopts = [torch.optim.SGD(b.parameters(), lr=lr) for b in oworld.bags]
    offsets = torch.arange(F * B + 1)
    bit_equal = maps_ok = True
    for _ in range(steps):
        ids = torch.cat([torch.randint(0, rows[t], (B,), generator=gen) + goff[t] for t in range(F)])
        grad = torch.randn(B, F * dim, generator=gen)
        out_a = model_a(ids.to(dev), offsets.to(dev), already_split_along_rank=False)
        out_b = model_b(ids.to(dev), offsets.to(dev), already_split_along_rank=False)
        oouts = oworld.forward(ids, offsets)
        bit_equal = bit_equal and torch.equal(out_a, out_b)
        worst = max(worst, _close(out_a, oouts[rank]), _close(out_b, oouts[rank]))
        g = grad[begin:begin + strides[rank]].to(dev)
        out_a.backward(g)
        out_b.backward(g.clone())
        torch.cat(oouts, 0).backward(grad)
        for o in oopts:
            o.step()
            o.zero_grad()
        omgr = oworld.bags[rank].cache_weight_mgr
        maps_ok = maps_ok and _maps_equal(model_a.cache_weight_mgr, omgr) and _maps_equal(model_b.cache_weight_mgr, omgr)
    evicted = sum(model_b.cache_weight_mgr.num_write_back_history)
    model_a.cache_weight_mgr.flush()
    model_b.cache_weight_mgr.flush()
    oworld.flush()
    want_w = oworld.bags[rank].weight
    worst_w = max(_close(model_a.weight, want_w), _close(model_b.weight, want_w))
    tables_equal = torch.equal(model_a.weight, model_b.weight)
    rec["tablewise"] = {"fused_vs_nccl_outputs_bit_equal": bool(bit_equal), "fused_vs_nccl_tables_bit_equal": bool(tables_equal),
                        "slot_maps_bit_exact_vs_oracle": bool(maps_ok), "evicted_rows_this_rank": int(evicted)}
    ok = ok and bit_equal and maps_ok and tables_equal and worst <= 1.0 and worst_w <= 1.0 and evicted > 0
    model_b.enable_fused_exchange(False)            # frees the peer buffers
    del model_a, model_b

    # ---- column-wise vs the oracle world -------------------------------------------------------------------------------
    full = torch.cat(weights, 0)
    N = full.shape[0]
    s, e, _ = ce.get_partition(dim, rank, world)
    crow = max(F * B + 4, N // 8)
    model_c = ce.ParallelCachedEmbeddingBag.from_pretrained(full[:, s:e].clone().contiguous(), freeze=False, mode="sum",
                                                            include_last_offset=True, cuda_row_num=crow, sparse=True,
                                                            full_dim=dim, evict_strategy=ce.EvictionStrategy.LFU,
                                                            fused_optimizer="sgd", lr=lr, process_group=group)
    cworld = OracleColumnwiseWorld(full, world, crow, mode="sum", include_last_offset=True, sparse=True,
                                   evict_strategy=OS.LFU)
    copts = [torch.optim.SGD(b.parameters(), lr=lr) for b in cworld.bags]
    hook = lambda x: x.view(F, B, -1).transpose(0, 1)          # recsys/models/dlrm.py:26-27
    cmaps_ok = True
    worst_c = 0.0
    for _ in range(steps):
        ids = torch.cat([torch.randint(0, rows[t], (B,), generator=gen) + goff[t] for t in range(F)])
        grad = torch.randn(B, F, dim, generator=gen)
        out_c = model_c(ids.to(dev), offsets.to(dev), shape_hook=hook)          # (B_rank, F, D)
        couts = cworld.forward(ids, offsets, shape_hook=hook)
        worst_c = max(worst_c, _close(out_c, couts[rank]))
        out_c.backward(grad[begin:begin + strides[rank]].to(dev))
        torch.cat(couts, 0).backward(grad)
        for o in copts:
            o.step()
            o.zero_grad()
        cmaps_ok = cmaps_ok and _maps_equal(model_c.cache_weight_mgr, cworld.bags[rank].cache_weight_mgr)
    evicted_c = sum(model_c.cache_weight_mgr.num_write_back_history)
    model_c.cache_weight_mgr.flush()
    cworld.flush()
    worst_cw = _close(model_c.weight, cworld.bags[rank].weight)
    rec["columnwise"] = {"slot_maps_bit_exact_vs_oracle": bool(cmaps_ok), "evicted_rows_this_rank": int(evicted_c),
                         "columns_this_rank": [s, e]}
    ok = ok and cmaps_ok and worst_c <= 1.0 and worst_cw <= 1.0 and evicted_c > 0
    del model_c

    # ---- merge over ranks ----------------------------------------------------------------------------------------------
    t = torch.tensor([1.0 if ok else 0.0, -max(worst, worst_c), -max(worst_w, worst_cw)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    rec["ok"] = bool(t[0].item() == 1.0)
    rec["max_output_err_over_1e-5_bar"] = round(-float(t[1]), 4)
    rec["max_table_err_over_1e-5_bar"] = round(-float(t[2]), 4)
    rec["checker"] = "CPU oracle worlds (oracle/cache_oracle.py), same seeds on every rank"
    return rec
