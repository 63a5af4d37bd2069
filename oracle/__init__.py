"""CPU oracle for the cached embedding-bag hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cachedembedding_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the CPU arm, never as the thing shipped.

PARITY UNPINNED at the reference boundary: the reference repository
(hpcaitech/CachedEmbedding @ a2af3d7e) holds no tests, golden vectors or
fixtures for this path, and the implementation of the path lives in a
third-party dependency that is absent from /root/reference
(ColossalAI @ e8d8eda5e7a0619bd779e35065397679e1536dcd, prose pin README.md:37;
package ``colossalai/nn/parallel/layers/cache_embedding``).  The oracle restates
that package's published algorithm and is pinned instead
against (1) ``torch.nn.EmbeddingBag`` + ``torch.optim.SGD`` on the full table,
which is upstream's own definition of value parity (Appendix B.3), and (2) the
upstream known-answer tests restated in tests/ (Appendix B.1, B.2, B.4, B.5).
"""
from .cache_oracle import (  # noqa: F401
    EvictionStrategy,
    OracleCachedParamMgr,
    OracleCachedEmbeddingBag,
    OracleTablewiseConfig,
    OracleTablewiseWorld,
    OracleColumnwiseWorld,
    rowwise_adagrad_reference,
    get_partition,
)
