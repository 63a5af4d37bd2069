"""Probe (torchrun, N ranks): cost of the pooled-embedding exchange pieces at the bench's sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from cachedembedding_b200.collectives import dual_all_to_all_tablewise, split_sizes
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, F, D = 65536, 26, 128
F_loc = F // world + (1 if rank < F % world else 0)
dim_per_rank = [D * (F // world + (1 if r < F % world else 0)) for r in range(world)]
strides = split_sizes(B, world)
x = torch.randn(B, F_loc * D, device=dev, requires_grad=True)
g = torch.randn(strides[rank], F * D, device=dev)

def timeit(name, fn, iters=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:55s} {e0.elapsed_time(e1)/iters:8.3f} ms")

send = [c.contiguous() for c in x.detach().split(strides, 0)]
recv = [torch.empty(strides[rank], dim_per_rank[r], device=dev) for r in range(world)]
timeit("dist.all_to_all (list) fwd payload", lambda: dist.all_to_all(recv, send))
flat_in = x.detach().reshape(-1)
flat_out = torch.empty(sum(r.numel() for r in recv), device=dev)
in_splits = [s * F_loc * D for s in strides]
out_splits = [strides[rank] * d for d in dim_per_rank]
timeit("dist.all_to_all_single fwd payload", lambda: dist.all_to_all_single(flat_out, flat_in, out_splits, in_splits))
timeit("torch.cat(recv, 1) unpack", lambda: torch.cat(recv, 1))
timeit("grad.split(dim1)+contiguous pack", lambda: [c.contiguous() for c in g.split(dim_per_rank, 1)])
def fb():
    out = dual_all_to_all_tablewise(x, None, strides, dim_per_rank)
    out.backward(g)
timeit("dual_all_to_all_tablewise fwd+bwd (autograd)", fb)
timeit("local copy same bytes (x.clone)", lambda: x.detach().clone())
dist.destroy_process_group()
