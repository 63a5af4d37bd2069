"""EvictionStrategy enum (reference use: recsys/models/dlrm.py:66,80; benchmark/benchmark_cache.py:40)."""
from enum import Enum


class EvictionStrategy(Enum):
    LFU = 1
    # dataset aware eviction strategy: evict the row whose id is the least frequent in the dataset
    DATASET = 2
