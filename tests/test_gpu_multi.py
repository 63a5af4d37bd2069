"""Two-rank tests of the parallel bags on real GPUs (NCCL): table-wise (upstream B.5 restated) and column-wise (B.6)
against the single-process oracle worlds.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _tablewise_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cachedembedding_b200 as ce
    from oracle import EvictionStrategy as OS, OracleTablewiseConfig, OracleTablewiseWorld
    torch.manual_seed(0)
    weights = torch.rand(18, 5)
    rows, ranks = [6, 5, 7], [0, 0, 1]
    starts = [0, 6, 11]
    cfgs = [ce.TablewiseEmbeddingBagConfig(n, 4, assigned_rank=r, initial_weight=weights[s:s + n].clone())
            for n, r, s in zip(rows, ranks, starts)]
    ocfgs = [OracleTablewiseConfig(n, 4, assigned_rank=r, initial_weight=weights[s:s + n].clone())
             for n, r, s in zip(rows, ranks, starts)]
    model = ce.ParallelCachedEmbeddingBagTablewise(cfgs, embedding_dim=5, include_last_offset=True, cache_ratio=0.8,
                                                   buffer_size=0, evict_strategy=ce.EvictionStrategy.LFU, sparse=True)
    oworld = OracleTablewiseWorld(ocfgs, 5, world, mode='mean', include_last_offset=True, cache_ratio=0.8,
                                  evict_strategy=OS.LFU)
    values = torch.tensor([1, 2, 3, 1, 5, 6, 7, 9, 6, 8, 13, 15, 11])
    offsets = torch.tensor([0, 3, 3, 5, 7, 8, 10, 10, 12, 13])
    res = model(values.cuda(), offsets.cuda(), already_split_along_rank=False)
    want = oworld.forward(values, offsets)[rank]
    ok = torch.allclose(res.cpu(), want, rtol=1e-5, atol=1e-6)
    optimizer = torch.optim.SGD(model.parameters(), lr=1e-2)
    rand_grad = torch.rand(3, 15)
    fake_grad = rand_grad[0:2] if rank == 0 else rand_grad[2:]
    res.backward(fake_grad.cuda())
    optimizer.step()
    optimizer.zero_grad()
    model.cache_weight_mgr.flush()
    ref = torch.nn.EmbeddingBag.from_pretrained(weights.clone(), include_last_offset=True, freeze=False)
    ref_opt = torch.optim.SGD(ref.parameters(), lr=1e-2)
    ref(values, offsets).backward(torch.cat(rand_grad.split(5, 1), 0))
    ref_opt.step()
    want_w = ref.weight.detach()[:11] if rank == 0 else ref.weight.detach()[11:]
    ok = ok and torch.allclose(model.cache_weight_mgr.weight, want_w, rtol=1e-5, atol=1e-6)
    results[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _columnwise_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cachedembedding_b200 as ce
    torch.manual_seed(1)
    gen = torch.Generator().manual_seed(5)
    N, D, B = 100, 8, 8
    weight = torch.rand(N, D)
    s, e, _ = ce.get_partition(D, rank, world)
    model = ce.ParallelCachedEmbeddingBag.from_pretrained(weight[:, s:e].clone().contiguous(), freeze=False, mode='mean',
                                                          include_last_offset=True, cuda_row_num=64, sparse=True,
                                                          full_dim=D)
    ref = torch.nn.EmbeddingBag.from_pretrained(weight.clone(), mode='mean', include_last_offset=True, freeze=False)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    ropt = torch.optim.SGD(ref.parameters(), lr=1e-3)
    ok = True
    for _ in range(5):
        n = 2 * B
        indices = torch.randint(0, N, (n,), generator=gen)
        cuts = torch.sort(torch.randint(1, n, (B - 1,), generator=gen)).values
        offsets = torch.cat([torch.tensor([0]), cuts, torch.tensor([n])])
        out = model(indices.cuda(), offsets.cuda())          # (B / W, D): my slice of the bags, all columns
        rres = ref(indices, offsets)
        mine = torch.tensor_split(rres, world, 0)[rank]
        ok = ok and torch.allclose(out.cpu(), mine.detach(), rtol=1e-5, atol=1e-6)
        grad = torch.rand(rres.shape, generator=gen)
        out.backward(torch.tensor_split(grad, world, 0)[rank].cuda())
        rres.backward(grad)
        opt.step(); opt.zero_grad(); ropt.step(); ropt.zero_grad()
    model.cache_weight_mgr.flush()
    ok = ok and torch.allclose(model.cache_weight_mgr.weight, ref.weight.detach()[:, s:e], rtol=1e-5, atol=1e-6)
    results[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def _fused_exchange_worker(rank, world, port, results):
    """Three ways through the same steps: (a) NCCL all-to-all, serial cache op (the reference's order); (b) exchange
    fused into the kernels over peer memory; (c) fused exchange + look-ahead driver (windows of two batches, cache op
    and backward plans on side streams).  Pooled outputs equal bit for bit for (b) and within 1e-5 for (c), tables
    equal within 1e-5 after flush, with evictions."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cachedembedding_b200 as ce
    gen = torch.Generator().manual_seed(3)
    rows = [500, 40, 3000, 7, 1200]
    ranks = [0, 1, 1, 0, 1]
    D, B, P, STEPS = 128, 37, 2, 8                   # B not divisible by the world
    weights = [torch.randn(n, D, generator=gen) * 0.1 for n in rows]
    mine = [t for t, r in enumerate(ranks) if r == rank]
    strides = [B // world + int(i < B % world) for i in range(world)]
    # inputs of every step, drawn identically on all ranks (each keeps its own tables / its own batch rows)
    steps = []
    for step in range(STEPS):
        per_table = []
        for t in range(len(rows)):
            lens = torch.randint(0, 3, (B,), generator=gen)
            ids = (torch.rand(int(lens.sum()), generator=gen) ** 2 * rows[t]).long().clamp_(0, rows[t] - 1)
            per_table.append((lens, ids))
        grad = torch.randn(B, len(rows) * D, generator=gen)
        local_off, parts, lens_all = 0, [], []
        for t in mine:                               # local KJT: ids re-based to the local concatenated table (A.6)
            lens, ids = per_table[t]
            parts.append(ids + local_off); lens_all.append(lens); local_off += rows[t]
        values = torch.cat(parts).cuda()
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.cat(lens_all), 0)]).cuda()
        steps.append((values, offsets, grad.split(strides, 0)[rank].cuda()))

    def make_bag(fused):
        cfgs = [ce.TablewiseEmbeddingBagConfig(n, 0, assigned_rank=r, initial_weight=w.clone())
                for n, r, w in zip(rows, ranks, weights)]
        bag = ce.ParallelCachedEmbeddingBagTablewise(cfgs, embedding_dim=D, include_last_offset=True, mode="sum",
                                                     cache_ratio=0.3, warmup_ratio=1.0, sparse=True,   # starts full
                                                     evict_strategy=ce.EvictionStrategy.LFU,
                                                     fused_optimizer="sgd", lr=0.25)
        bag.enable_fused_exchange(fused)
        return bag

    outs = {}
    bags = {}
    for variant in ("nccl", "fused"):
        bag = bags[variant] = make_bag(variant == "fused")
        outs[variant] = []
        for values, offsets, my_grad in steps:
            out = bag(values, offsets)
            outs[variant].append(out.detach().clone())
            out.backward(my_grad)
    # (c) look-ahead: windows of P batches
    bag = bags["lookahead"] = make_bag(True)
    bag.set_cache_op(False)
    outs["lookahead"] = []
    pf = ce.LookaheadPrefetcher(bag)
    windows = [steps[w * P:(w + 1) * P] for w in range(STEPS // P)]
    h = pf.submit([v for v, _, _ in windows[0]], offsets=[o for _, o, _ in windows[0]])
    for w, win in enumerate(windows):
        slots = h.wait()
        sizes = [v.numel() for v, _, _ in win]
        for s, (values, offsets, my_grad) in zip(torch.split(slots, sizes), win):
            out = bag(s, offsets)
            outs["lookahead"].append(out.detach().clone())
            out.backward(my_grad)
        pf.window_enqueued()
        if w + 1 < len(windows):
            h = pf.submit([v for v, _, _ in windows[w + 1]], offsets=[o for _, o, _ in windows[w + 1]])
    pf.close()

    # (d) the CUDA-graph operator step (fused_step) against the eager autograd path, both under the look-ahead driver
    # with early submission; pooling factor 1 so that every window has the same size (the graphs are REPLAYED)
    g2 = torch.Generator().manual_seed(11)
    F_loc = len(mine)
    local_rows = [rows[t] for t in mine]
    base = torch.tensor([sum(local_rows[:k]) for k in range(F_loc)]).view(F_loc, 1)
    steps1 = []
    for _ in range(12):
        ids = torch.stack([(torch.rand(B, generator=g2) ** 2 * n).long().clamp_(0, n - 1) for n in local_rows]) + base
        steps1.append((ids.view(-1).cuda(), torch.randn(B, len(rows) * D, generator=g2).split(strides, 0)[rank].cuda()))
    off1 = torch.arange(F_loc * B + 1).cuda()
    outs1 = {}
    for variant in ("eager", "graph"):
        bag = bags["d_" + variant] = make_bag(True)
        bag.set_cache_op(False)
        outs1[variant] = []
        pf = ce.LookaheadPrefetcher(bag)
        snapshot = torch.empty(strides[rank], len(rows) * D, device="cuda")
        grad_src = torch.empty_like(snapshot)
        gbuf = bag._exchange_for(B).grad_tensor()

        def dense_part(o):       # captured between the two barriers: reads the output, leaves the gradient for the peers
            snapshot.copy_(o)
            gbuf.copy_(grad_src)
        wins = [steps1[w * P:(w + 1) * P] for w in range(len(steps1) // P)]
        h = pf.submit([v for v, _ in wins[0]], offsets=off1)
        for w, win in enumerate(wins):
            slots = h.wait()
            for j, (s, (_, my_grad)) in enumerate(zip(torch.chunk(slots, P), win)):
                if variant == "graph":
                    grad_src.copy_(my_grad)
                    bag.fused_step(s, off1, consumer=dense_part)
                    outs1[variant].append(snapshot.clone())
                else:
                    out = bag(s, off1)
                    outs1[variant].append(out.detach().clone())     # before the backward's barrier lets the peers go on
                    out.backward(my_grad)
                if j == 0 and w + 1 < len(wins):
                    h = pf.submit([v for v, _ in wins[w + 1]], offsets=off1)
            pf.window_enqueued()
        pf.close()
    ok = True
    if getattr(bags["d_graph"], "graph_launches", 0) == 0 or len(bags["d_graph"]._step_graphs) >= len(steps1):
        ok = False
        print(f"[rank {rank}] fused_step did not replay graphs: {len(bags['d_graph']._step_graphs)} graphs, "
              f"{getattr(bags['d_graph'], 'graph_launches', 0)} launches", flush=True)
    for k in range(len(steps1)):
        if not torch.equal(outs1["eager"][k], outs1["graph"][k]):
            ok = False
            print(f"[rank {rank}] graph step {k}: pooled outputs differ from the eager path", flush=True)
    for variant in ("fused", "lookahead"):
        for k in range(STEPS):
            # same slot assignment -> same summation grouping -> same bits.  The look-ahead driver protects two windows,
            # picks other victims, so rows sit in other slots and the backward groups duplicate-slot partial sums
            # differently: equal within fp32 rounding (1e-5 relative), not bit for bit.
            same = torch.equal(outs["nccl"][k], outs[variant][k]) if variant == "fused" else \
                torch.allclose(outs["nccl"][k], outs[variant][k], rtol=1e-5, atol=2e-6)
            if not same:
                ok = False
                print(f"[rank {rank}] {variant} step {k}: pooled outputs differ, max abs diff "
                      f"{(outs['nccl'][k] - outs[variant][k]).abs().max().item():.3e}", flush=True)
    for b in bags.values():
        b.cache_weight_mgr.flush()
    if not torch.equal(bags["d_eager"].weight, bags["d_graph"].weight):
        ok = False
        print(f"[rank {rank}] graph step: tables differ from the eager path", flush=True)
    for variant in ("fused", "lookahead"):
        if not torch.allclose(bags["nccl"].weight, bags[variant].weight, rtol=1e-5, atol=1e-6):
            ok = False
            print(f"[rank {rank}] {variant}: tables differ, max abs diff "
                  f"{(bags['nccl'].weight - bags[variant].weight).abs().max().item():.3e}", flush=True)
    if not all(sum(b.num_write_back_history) > 0 for k, b in bags.items() if not k.startswith("d_")):
        ok = False
        print(f"[rank {rank}] scenario did not evict: {[sum(b.num_write_back_history) for b in bags.values()]}", flush=True)
    if torch.equal(bags["fused"].weight, torch.cat([weights[t] for t in mine])):
        ok = False
        print(f"[rank {rank}] table did not change", flush=True)
    results[rank] = bool(ok)
    dist.barrier()
    for b in bags.values():
        b.enable_fused_exchange(False)
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_tablewise_worker, _columnwise_worker, _fused_exchange_worker])
def test_parallel_bags_two_ranks(worker):
    _need_two_gpus()
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r, False) for r in range(world)), dict(results)


def _verify_leg_worker(rank, world, port, results):
    """bench.py's untimed parity leg (bench_verify.verify_parallel) on the Criteo-1TB table placement for `world`."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bench
    from bench_verify import verify_parallel
    arrange = bench.rank_arrange(bench.CRITEO_1TB_ROWS, world)
    rec = verify_parallel([300 + 37 * (t % 5) for t in range(26)], arrange, 128)
    results[rank] = rec
    dist.barrier()
    dist.destroy_process_group()


def test_bench_parity_leg_two_ranks():
    _need_two_gpus()
    world, port = 2, _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_verify_leg_worker, args=(world, port, results), nprocs=world, join=True)
    for r in range(world):
        assert results[r]["ok"], dict(results[r])
        assert results[r]["tablewise"]["evicted_rows_this_rank"] > 0

