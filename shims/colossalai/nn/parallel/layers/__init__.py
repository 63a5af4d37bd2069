"""colossalai.nn.parallel.layers: the cached embedding bag -- here the B200-native one."""
from cachedembedding_b200 import (CachedEmbeddingBag, CachedParamMgr, EvictionStrategy, FreqAwareEmbeddingBag,  # noqa: F401
                                  LimitBuffIndexCopyer, ParallelCachedEmbeddingBag,
                                  ParallelCachedEmbeddingBagTablewise, TablewiseEmbeddingBagConfig)
from . import cache_embedding  # noqa: F401
