"""colossalai.core.global_context: the one-group context recsys/models/dlrm.py:87,196-197 asks for."""
import torch.distributed as dist


class _GlobalContext:
    def __init__(self):
        self._rank, self._local_rank, self._world = 0, 0, 1

    def _set(self, rank, local_rank, world):
        self._rank, self._local_rank, self._world = rank, local_rank, world

    def get_group(self, parallel_mode=None):
        return dist.group.WORLD if dist.is_initialized() else None

    def get_global_rank(self):
        return self._rank

    def get_local_rank(self, parallel_mode=None):
        return self._rank

    def get_world_size(self, parallel_mode=None):
        return self._world


global_context = _GlobalContext()
