from cachedembedding_b200 import (CachedEmbeddingBag, CachedParamMgr, EvictionStrategy, FreqAwareEmbeddingBag,  # noqa: F401
                                  LimitBuffIndexCopyer, ParallelCachedEmbeddingBag,
                                  ParallelCachedEmbeddingBagTablewise, TablewiseEmbeddingBagConfig)
