"""cachedembedding_b200: a B200-native cached embedding bag (host-resident table, HBM slot cache).

Public names mirror ``colossalai.nn.parallel.layers`` as imported by the reference
(/root/reference/recsys/models/dlrm.py:15-16, recsys/utils/misc.py:8, benchmark/benchmark_cache.py:16).
"""
from .evict_strategy import EvictionStrategy
from .embedding_config import TablewiseEmbeddingBagConfig
from .copyer import LimitBuffIndexCopyer
from .cache_mgr import CachedParamMgr, CacheCapacityError
from .cached_embedding import (CachedEmbeddingBag, FreqAwareEmbeddingBag, BaseEmbeddingBag, embedding_bag_cached,
                               alloc_pinned_table)
from .parallel_cached_embedding import ParallelCachedEmbeddingBag
from .parallel_cached_embedding_tablewise import ParallelCachedEmbeddingBagTablewise
from .lookahead import LookaheadPrefetcher, PrefetchHandle
from .fused_exchange import FusedExchange, PeerBuffer
from .collectives import dual_all_to_all, dual_all_to_all_tablewise, get_partition
from .kjt_exchange import FusedKJTAllToAll
from .synth_criteo import IdFrequencyCounter, SyntheticCriteo, write_kaggle_format

__all__ = [
    'EvictionStrategy', 'TablewiseEmbeddingBagConfig', 'LimitBuffIndexCopyer', 'CachedParamMgr', 'CacheCapacityError',
    'CachedEmbeddingBag', 'FreqAwareEmbeddingBag', 'BaseEmbeddingBag', 'embedding_bag_cached', 'alloc_pinned_table',
    'ParallelCachedEmbeddingBag', 'ParallelCachedEmbeddingBagTablewise', 'dual_all_to_all',
    'dual_all_to_all_tablewise', 'get_partition', 'LookaheadPrefetcher', 'PrefetchHandle', 'FusedKJTAllToAll',
    'IdFrequencyCounter', 'SyntheticCriteo', 'write_kaggle_format',
]
