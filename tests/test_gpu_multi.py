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
    """Fused peer-memory exchange == NCCL all-to-all path, bit for bit on the pooled output, and within 1e-5 on the
    tables after several steps with evictions."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cachedembedding_b200 as ce
    gen = torch.Generator().manual_seed(3)
    rows = [500, 40, 3000, 7, 1200]
    ranks = [0, 1, 1, 0, 1]
    D, B = 128, 37                                   # B not divisible by the world
    weights = [torch.randn(n, D, generator=gen) * 0.1 for n in rows]
    mine = [t for t, r in enumerate(ranks) if r == rank]
    bags = []
    for fused in (False, True):
        cfgs = [ce.TablewiseEmbeddingBagConfig(n, 0, assigned_rank=r, initial_weight=w.clone())
                for n, r, w in zip(rows, ranks, weights)]
        bag = ce.ParallelCachedEmbeddingBagTablewise(cfgs, embedding_dim=D, include_last_offset=True, mode="sum",
                                                     cache_ratio=0.2, warmup_ratio=1.0, sparse=True,   # cache starts full: every miss evicts
                                                     evict_strategy=ce.EvictionStrategy.LFU,
                                                     fused_optimizer="sgd", lr=0.25)
        bag.enable_fused_exchange(fused)
        bags.append(bag)
    strides = [B // world + int(i < B % world) for i in range(world)]
    ok = True
    for step in range(4):
        # the local KJT: my tables only, ids already re-based to the local concatenated table (A.6)
        local_off, parts, lens_all = 0, [], []
        for t in mine:
            lens = torch.randint(0, 3, (B,), generator=gen)
            ids = (torch.rand(int(lens.sum()), generator=gen) ** 2 * rows[t]).long().clamp_(0, rows[t] - 1) + local_off
            parts.append(ids); lens_all.append(lens); local_off += rows[t]
        # every rank draws from the same generator stream: advance it for the other rank's tables too
        for t in range(len(rows)):
            if t not in mine:
                lens = torch.randint(0, 3, (B,), generator=gen)
                torch.rand(int(lens.sum()), generator=gen)
        values = torch.cat(parts).cuda()
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(torch.cat(lens_all), 0)]).cuda()
        grad = torch.randn(B, len(rows) * D, generator=gen)
        my_grad = grad.split(strides, 0)[rank].cuda()
        outs = []
        for bag in bags:
            out = bag(values, offsets)
            outs.append(out.detach().clone())
            out.backward(my_grad)
        if not torch.equal(outs[0], outs[1]):
            ok = False
            print(f"[rank {rank}] step {step}: pooled outputs differ, max abs diff "
                  f"{(outs[0] - outs[1]).abs().max().item():.3e}", flush=True)
    for bag in bags:
        bag.cache_weight_mgr.flush()
    if not torch.allclose(bags[0].weight, bags[1].weight, rtol=1e-5, atol=1e-6):
        ok = False
        print(f"[rank {rank}] tables differ, max abs diff {(bags[0].weight - bags[1].weight).abs().max().item():.3e}",
              flush=True)
    if not sum(bags[1].num_write_back_history) > 0:
        ok = False
        print(f"[rank {rank}] scenario did not evict", flush=True)
    if torch.equal(bags[1].weight, torch.cat([weights[t] for t in mine])):
        ok = False
        print(f"[rank {rank}] table did not change", flush=True)
    results[rank] = bool(ok)
    dist.barrier()
    bags[1].enable_fused_exchange(False)
    dist.destroy_process_group()


@pytest.mark.parametrize("worker", [_tablewise_worker, _columnwise_worker, _fused_exchange_worker])
def test_parallel_bags_two_ranks(worker):
    _need_two_gpus()
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r, False) for r in range(world)), dict(results)
