"""Make the reference's import lines resolve to this package without touching its files.

The reference imports the path as ``colossalai.nn.parallel.layers[.cache_embedding]``
(/root/reference/recsys/models/dlrm.py:15-16, recsys/utils/misc.py:8, benchmark/benchmark_cache.py:16,
benchmark/benchmark_fbgemm_uvm.py:3, baselines/data/synth.py:6).  ``install()`` registers alias modules under those
names; if a real ``colossalai`` is importable only the ``layers`` sub-modules are overridden, otherwise light parent
packages are created so that the ``from ... import ...`` statements work.  The rest of ColossalAI that the training
script uses (launch, logger, global context) is out of scope and is not provided here.
"""
import importlib
import sys
import types

_LAYER_NAMES = ("CachedEmbeddingBag", "FreqAwareEmbeddingBag", "ParallelCachedEmbeddingBag",
                "ParallelCachedEmbeddingBagTablewise", "CachedParamMgr", "EvictionStrategy",
                "TablewiseEmbeddingBagConfig", "LimitBuffIndexCopyer")


def install(force: bool = True):
    """Alias ``colossalai.nn.parallel.layers`` and ``...layers.cache_embedding`` to cachedembedding_b200."""
    import cachedembedding_b200 as ce
    layers = types.ModuleType("colossalai.nn.parallel.layers")
    cache_embedding = types.ModuleType("colossalai.nn.parallel.layers.cache_embedding")
    for name in _LAYER_NAMES:
        setattr(layers, name, getattr(ce, name))
        setattr(cache_embedding, name, getattr(ce, name))
    layers.cache_embedding = cache_embedding
    parents = ["colossalai", "colossalai.nn", "colossalai.nn.parallel"]
    for i, pname in enumerate(parents):
        if pname not in sys.modules:
            try:
                importlib.import_module(pname)
            except Exception:
                mod = types.ModuleType(pname)
                mod.__path__ = []          # a package, so that sub-module imports are looked up in sys.modules
                sys.modules[pname] = mod
                if i > 0:
                    setattr(sys.modules[parents[i - 1]], pname.rsplit(".", 1)[1], mod)
    if force or "colossalai.nn.parallel.layers" not in sys.modules:
        sys.modules["colossalai.nn.parallel.layers"] = layers
        sys.modules["colossalai.nn.parallel.layers.cache_embedding"] = cache_embedding
        sys.modules["colossalai.nn.parallel"].layers = layers
    return layers
