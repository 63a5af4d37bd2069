"""KeyedJaggedTensor: values of all keys concatenated key-major (feature-major), one length per (key, sample).

Only what the reference touches (recsys/datasets/criteo.py:200-213, recsys/dlrm_main.py:253, recsys/models/dlrm.py:104-110,
recsys/datasets/utils.py:24-54): construction from values + lengths / offsets (+ the precomputed per-key tables),
values / lengths / offsets / stride / keys, to / pin_memory / record_stream, from_lengths_sync."""
from typing import Dict, List, Optional

import torch


class KeyedJaggedTensor:

    def __init__(self, keys: List[str], values: torch.Tensor, weights: Optional[torch.Tensor] = None,
                 lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                 stride: Optional[int] = None, length_per_key: Optional[List[int]] = None,
                 offset_per_key: Optional[List[int]] = None, index_per_key: Optional[Dict[str, int]] = None, **_):
        assert lengths is not None or offsets is not None, "lengths or offsets are needed"
        self._keys, self._values, self._weights = list(keys), values, weights
        self._lengths, self._offsets = lengths, offsets
        n = (lengths.numel() if lengths is not None else offsets.numel() - 1)
        self._stride = stride if stride is not None else (n // max(len(self._keys), 1))
        self._length_per_key, self._offset_per_key, self._index_per_key = length_per_key, offset_per_key, index_per_key

    @staticmethod
    def from_lengths_sync(keys, values, lengths, weights=None):
        return KeyedJaggedTensor(keys=keys, values=values, weights=weights, lengths=lengths)

    @staticmethod
    def from_offsets_sync(keys, values, offsets, weights=None):
        return KeyedJaggedTensor(keys=keys, values=values, weights=weights, offsets=offsets)

    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def weights_or_none(self):
        return self._weights

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            self._lengths = self._offsets[1:] - self._offsets[:-1]
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            zero = torch.zeros(1, dtype=self._lengths.dtype, device=self._lengths.device)
            self._offsets = torch.cat([zero, torch.cumsum(self._lengths, 0).to(self._lengths.dtype)])
        return self._offsets

    def stride(self) -> int:
        return self._stride

    def length_per_key(self) -> List[int]:
        if self._length_per_key is None:
            self._length_per_key = self.lengths().view(len(self._keys), -1).sum(1).tolist()
        return self._length_per_key

    def _map(self, fn):
        opt = lambda t: fn(t) if t is not None else None
        return KeyedJaggedTensor(self._keys, fn(self._values), opt(self._weights), opt(self._lengths), opt(self._offsets),
                                 self._stride, self._length_per_key, self._offset_per_key, self._index_per_key)

    def to(self, device, non_blocking: bool = False):
        return self._map(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._map(lambda t: t.pin_memory())

    def record_stream(self, stream) -> None:
        for t in (self._values, self._weights, self._lengths, self._offsets):
            if t is not None and t.is_cuda:
                t.record_stream(stream)

    def __repr__(self):
        return f"KeyedJaggedTensor(keys={len(self._keys)}, values={tuple(self._values.shape)}, stride={self._stride})"


class KeyedTensor:
    """Dense per-key output of an EmbeddingBagCollection (only named in type annotations of the dense arches)."""

    def __init__(self, keys, length_per_key, values, key_dim: int = 1):
        self._keys, self._length_per_key, self._values, self._key_dim = keys, length_per_key, values, key_dim

    def keys(self):
        return self._keys

    def values(self):
        return self._values

    def to_dict(self):
        return dict(zip(self._keys, torch.split(self._values, self._length_per_key, dim=self._key_dim)))
