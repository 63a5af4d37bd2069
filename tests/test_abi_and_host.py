"""CPU-side checks: the C-ABI library loads and exports every symbol include/cebag.h declares (no compute calls),
host logic of the parallel bags, and the world_size-2 exchange over gloo."""
import os
import re
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cebag.h")).read()
    return sorted(set(re.findall(r"CEBAG_API\s+[\w\s\*]+?\b(cebag_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cachedembedding_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.EXPORTS) == declared, "binding and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cebag_abi_version() == _lib.ABI_VERSION
    # the library's own symbol table holds nothing else with the cebag_ prefix
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T cebag_" in l)
    assert exported == declared


def test_struct_layouts_and_constants_match_header(tmp_path):
    """A C program compiled against include/cebag.h prints sizeof / offsetof of every struct and the value of every
    constant; the ctypes mirror in _lib.py (what every call in this repo goes through) must agree byte for byte."""
    import ctypes
    import subprocess
    from cachedembedding_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cebag.h")).read()
    structs = {"cebag_table": _lib.Table, "cebag_workspace": _lib.Workspace, "cebag_prepare_stats": _lib.PrepareStats,
               "cebag_prepare_result": _lib.PrepareResult, "cebag_exchange": _lib.Exchange,
               "cebag_bag_args": _lib.BagArgs}
    declared = re.findall(r"typedef struct (\w+) \{(.*?)\} \1;", header, flags=re.S)
    assert sorted(n for n, _ in declared) == sorted(structs), "a struct of the header has no ctypes mirror"
    for name, body in declared:   # same field names, same order
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = re.search(r"(\w+)\s*(\[\w+\])?$", names[0].strip()).group(1)
            fields.append(first)
            fields += [re.search(r"(\w+)", n).group(1) for n in names[1:]]
        assert fields == [f for f, _ in structs[name]._fields_], name
    consts = {"CEBAG_ABI_VERSION": _lib.ABI_VERSION, "CEBAG_OK": _lib.OK, "CEBAG_ERR_INVALID": _lib.ERR_INVALID,
              "CEBAG_ERR_CUDA": _lib.ERR_CUDA, "CEBAG_ERR_CAPACITY": _lib.ERR_CAPACITY, "CEBAG_ERR_INDEX": _lib.ERR_INDEX,
              "CEBAG_PREPARE_PENDING": _lib.PREPARE_PENDING, "CEBAG_EVICT_LFU": _lib.EVICT_LFU,
              "CEBAG_EVICT_DATASET": _lib.EVICT_DATASET, "CEBAG_MODE_SUM": _lib.MODE_SUM, "CEBAG_MODE_MEAN": _lib.MODE_MEAN,
              "CEBAG_OPT_SGD": _lib.OPT_SGD, "CEBAG_OPT_ROWWISE_ADAGRAD": _lib.OPT_ROWWISE_ADAGRAD,
              "CEBAG_MAX_PEERS": _lib.MAX_PEERS, "CEBAG_LAYOUT_BAG_MAJOR": _lib.LAYOUT_BAG_MAJOR,
              "CEBAG_LAYOUT_SAMPLE_MAJOR": _lib.LAYOUT_SAMPLE_MAJOR, "CEBAG_LAYOUT_EXCHANGE": _lib.LAYOUT_EXCHANGE,
              "CEBAG_STATE_AVAIL": _lib.STATE_AVAIL, "CEBAG_STATE_EPOCH": _lib.STATE_EPOCH,
              "CEBAG_STATE_CALLS": _lib.STATE_CALLS, "CEBAG_STATE_MAXFREQ": _lib.STATE_MAXFREQ,
              "CEBAG_STATE_WORDS": _lib.STATE_WORDS}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "cebag.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    for cname in consts:
        lines.append(f'  printf("{cname} %lld\\n", (long long)({cname}));')
    lines.append('  printf("CEBAG_FREQ_EMPTY %lld\\n", (long long)CEBAG_FREQ_EMPTY);')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)   # the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
    for cname, value in consts.items():
        assert int(got[cname]) == value, cname
    assert int(got["CEBAG_FREQ_EMPTY"]) == _lib.FREQ_EMPTY


def test_integration_stub_matches_binding():
    """The ctypes stub printed in INTEGRATION.md declares the same struct layouts and argument lists as _lib.py (which
    the layout test above ties to the header); the GPU suite executes the same block end to end."""
    import ctypes
    from cachedembedding_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"<!-- stub:begin -->\s*```python\n(.*?)```\s*<!-- stub:end -->", text, flags=re.S).group(1)
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        exec(compile(block, "INTEGRATION.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    pairs = {"cebag_table": _lib.Table, "cebag_workspace": _lib.Workspace, "cebag_prepare_stats": _lib.PrepareStats,
             "cebag_bag_args": _lib.BagArgs}
    for name, cls in pairs.items():
        stub = ns[name]
        assert ctypes.sizeof(stub) == ctypes.sizeof(cls), name
        assert [(f, getattr(stub, f).offset) for f, _ in stub._fields_] == \
               [(f, getattr(cls, f).offset) for f, _ in cls._fields_], name
    lib = _lib.load()
    for fn in ("cebag_prepare_ids", "cebag_flush", "cebag_bag_forward", "cebag_bag_backward_fused",
               "cebag_prepare_workspace_bytes", "cebag_backward_workspace_bytes"):
        a, b = getattr(ns["lib"], fn).argtypes, getattr(lib, fn).argtypes
        assert len(a) == len(b), fn
        assert [ctypes.sizeof(x) for x in a] == [ctypes.sizeof(x) for x in b], fn


def test_no_cpu_fallback():
    import cachedembedding_b200 as ce
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        ce.CachedParamMgr(torch.zeros(8, 4), 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ce.embedding_bag_cached(torch.zeros(4, 4), torch.zeros(2, dtype=torch.long), torch.zeros(1, dtype=torch.long))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cachedembedding_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


@pytest.mark.parametrize("D,W", [(128, 8), (10, 4), (7, 3), (16, 1)])
def test_get_partition_is_tensor_split(D, W):
    from cachedembedding_b200 import get_partition
    from oracle import get_partition as oget
    chunks = torch.tensor_split(torch.arange(D), W)
    for r in range(W):
        s, e, div = get_partition(D, r, W)
        assert (s, e) == (int(chunks[r][0]), int(chunks[r][-1]) + 1)
        assert (s, e, div) == oget(D, r, W)


def test_split_kjt_along_rank_matches_oracle():
    from cachedembedding_b200.parallel_cached_embedding_tablewise import split_kjt_along_rank
    from oracle import OracleTablewiseConfig, OracleTablewiseWorld
    gen = torch.Generator().manual_seed(0)
    rows = [6, 5, 7, 3]
    ranks = [0, 1, 0, 1]
    cfgs = [OracleTablewiseConfig(n, 0, assigned_rank=r, initial_weight=torch.rand(n, 4)) for n, r in zip(rows, ranks)]
    for include_last in (True, False):
        world = OracleTablewiseWorld(cfgs, 4, 2, include_last_offset=include_last, cache_ratio=1.0)
        B = 5
        lens = torch.randint(0, 4, (len(rows) * B,), generator=gen)
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(lens, 0)])
        goff = torch.cumsum(torch.tensor([0] + rows), 0)
        ids = torch.cat([torch.randint(0, rows[t], (int(lens[t * B:(t + 1) * B].sum()),), generator=gen) + goff[t]
                         for t in range(len(rows))])
        psw = torch.rand(ids.numel(), generator=gen)
        offs = offsets if include_last else offsets[:-1]
        for rk in range(2):
            want = world.split_along_rank(rk, B, ids, offs, psw)
            got = split_kjt_along_rank(world.assigned[rk], world.idx_offset_list[rk], include_last, B, ids, offs, psw)
            for a, b in zip(got, want):
                assert torch.equal(a, b)


def test_bench_placement_and_sampler():
    import bench
    for W in (1, 2, 4, 8):
        arr = bench.rank_arrange(bench.CRITEO_1TB_ROWS, W)
        assert len(arr) == 26 and set(arr) == set(range(W))
    rows = torch.tensor([1000, 3, 50])
    ids = bench.sample_ids(rows, 4096, torch.Generator().manual_seed(1), "cpu").view(3, 4096)
    assert int(ids[0].min()) >= 0 and int(ids[0].max()) < 1000
    assert int(ids[1].min()) >= 1000 and int(ids[1].max()) < 1003
    assert int(ids[2].min()) >= 1003 and int(ids[2].max()) < 1053
    # long tail: row 0 of a table is by far the most frequent
    assert (ids[0] == 0).float().mean() > 0.15


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cachedembedding_b200.collectives import dual_all_to_all, dual_all_to_all_tablewise, split_sizes
    torch.manual_seed(0)
    B, D = 5, 3                       # B not divisible by the world: remainder goes to the low ranks
    dim_per_rank = [2 * D, 1 * D]     # rank 0 owns two tables, rank 1 one
    locals_ = [torch.arange(B * d, dtype=torch.float32).view(B, d) + 100 * r for r, d in enumerate(dim_per_rank)]
    x = locals_[rank].clone().requires_grad_(True)
    strides = split_sizes(B, world)
    out = dual_all_to_all_tablewise(x, None, strides, dim_per_rank)
    begin = sum(strides[:rank])
    want = torch.cat([l[begin:begin + strides[rank]] for l in locals_], 1)
    ok = torch.equal(out.detach(), want)
    # backward: grad of rank j's output rows flows back to every rank's local columns
    grads = [torch.arange(strides[j] * sum(dim_per_rank), dtype=torch.float32).view(strides[j], -1) * (j + 1)
             for j in range(world)]
    out.backward(grads[rank])
    col0 = sum(dim_per_rank[:rank])
    want_g = torch.cat([g[:, col0:col0 + dim_per_rank[rank]] for g in grads], 0)
    ok = ok and torch.equal(x.grad, want_g)
    # column-wise exchange: scatter rows, gather columns of unequal width
    widths = [2, 1]
    shard = (torch.arange(6 * widths[rank], dtype=torch.float32).view(6, widths[rank]) + 10 * rank).requires_grad_(True)
    full = dual_all_to_all(shard, None, 0, -1, fwd_gather_sizes=widths, bwd_gather_sizes=split_sizes(6, world))
    shards = [torch.arange(6 * w, dtype=torch.float32).view(6, w) + 10 * r for r, w in enumerate(widths)]
    want_full = torch.cat([torch.tensor_split(s, world, 0)[rank] for s in shards], -1)
    ok = ok and torch.equal(full.detach(), want_full)
    full.sum().backward()
    ok = ok and torch.equal(shard.grad, torch.ones_like(shard))
    results[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_all_to_all_exchange_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_exchange_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results[r] for r in range(world)), dict(results)


def _kjt_worker(rank, world, port, results):
    """FusedKJTAllToAll against the semantics of the reference's KJTAllToAll (recsys/datasets/utils.py:21-54): every rank
    gets, per key, rank 0's values then rank 1's ..., and lengths laid out [key][rank][sample]."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "shims"))
    from torchrec.sparse.jagged_tensor import KeyedJaggedTensor
    from cachedembedding_b200 import FusedKJTAllToAll
    ok = True
    for ragged in (True, False):
        gen = torch.Generator().manual_seed(100 + rank + (7 if ragged else 0))
        F, B = 3, 4
        lengths = torch.randint(0, 4, (F * B,), generator=gen, dtype=torch.int32) if ragged \
            else torch.ones(F * B, dtype=torch.int32)
        values = torch.randint(0, 1000, (int(lengths.sum()),), generator=gen) + 10000 * rank
        kjt = KeyedJaggedTensor(keys=[f"k{f}" for f in range(F)], values=values, lengths=lengths, stride=B)
        coll = FusedKJTAllToAll(None, capacity=F * B * 3 if ragged else F * B, fixed_lengths=not ragged,
                                kjt_factory=KeyedJaggedTensor.from_lengths_sync)
        out = coll.all_to_all(kjt)
        # expected, from everybody's inputs
        everyone = [None] * world
        dist.all_gather_object(everyone, (values.tolist(), lengths.view(F, B).tolist()))
        want_values, want_lengths = [], []
        for f in range(F):
            for vals, lens in everyone:
                start = sum(sum(lens[k]) for k in range(f))
                want_values += vals[start:start + sum(lens[f])]
                want_lengths += lens[f]
        ok = ok and out.values().tolist() == want_values and out.lengths().tolist() == want_lengths
        ok = ok and out.stride() == B * world and out.keys() == kjt.keys()
        ok = ok and out.values().dtype == values.dtype and out.lengths().dtype == lengths.dtype
    results[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_fused_kjt_all_to_all_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_kjt_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results[r] for r in range(world)), dict(results)


def test_reference_import_lines_resolve_after_install():
    """The reference's own import statements (recsys/models/dlrm.py:15-16, recsys/utils/misc.py:8,
    benchmark/benchmark_cache.py:16, benchmark/benchmark_fbgemm_uvm.py:3) work against this package."""
    import subprocess
    import sys
    code = (
        "from cachedembedding_b200 import colossalai_compat; colossalai_compat.install()\n"
        "from colossalai.nn.parallel.layers import ParallelCachedEmbeddingBag, EvictionStrategy, "
        "TablewiseEmbeddingBagConfig, ParallelCachedEmbeddingBagTablewise\n"
        "from colossalai.nn.parallel.layers import CachedEmbeddingBag, EvictionStrategy\n"
        "from colossalai.nn.parallel.layers.cache_embedding import CachedEmbeddingBag as C2\n"
        "import cachedembedding_b200 as ce\n"
        "assert CachedEmbeddingBag is ce.CachedEmbeddingBag is C2 and EvictionStrategy.LFU.value == 1\n"
        "cfg = TablewiseEmbeddingBagConfig(num_embeddings=10, cuda_row_num=4, assigned_rank=0, ids_freq_mapping=None)\n"
        "assert cfg.buffer_size == 50_000\n"
        "print('ok')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


@pytest.mark.parametrize("D", [516, 130, 1024])
def test_unsupported_row_width_is_rejected_at_construction(D):
    """Rows live in the registers of one warp: up to 512 floats (multiple of 4) / 128 floats otherwise.  Wider tables
    fail when the module is built -- before any device memory is touched -- not at the first forward."""
    import cachedembedding_b200 as ce
    with pytest.raises(NotImplementedError, match="floats are not supported"):
        ce.CachedEmbeddingBag(10, D, _weight=torch.zeros(10, D))


def test_limit_buff_index_copyer_chunks():
    from cachedembedding_b200 import LimitBuffIndexCopyer
    gen = torch.Generator().manual_seed(0)
    src = torch.rand(50, 6, generator=gen)
    tgt = torch.zeros(20, 6)
    src_idx = torch.randperm(50, generator=gen)[:13]
    tgt_idx = torch.randperm(20, generator=gen)[:13]
    LimitBuffIndexCopyer(4).index_copy(0, src_idx, tgt_idx, src, tgt)
    want = torch.zeros(20, 6)
    want.index_copy_(0, tgt_idx, src.index_select(0, src_idx))
    assert torch.equal(tgt, want)
