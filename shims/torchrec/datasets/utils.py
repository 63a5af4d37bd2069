"""torchrec.datasets.utils: the Batch record the loaders yield (dense features, sparse KJT, labels)."""
from dataclasses import dataclass

import torch

from ..sparse.jagged_tensor import KeyedJaggedTensor

PATH_MANAGER_KEY = "torchrec"


@dataclass
class Batch:
    dense_features: torch.Tensor
    sparse_features: KeyedJaggedTensor
    labels: torch.Tensor

    def to(self, device, non_blocking: bool = False) -> "Batch":
        return Batch(self.dense_features.to(device, non_blocking=non_blocking),
                     self.sparse_features.to(device, non_blocking=non_blocking),
                     self.labels.to(device, non_blocking=non_blocking))

    def record_stream(self, stream) -> None:
        if self.dense_features.is_cuda:
            self.dense_features.record_stream(stream)
            self.labels.record_stream(stream)
        self.sparse_features.record_stream(stream)

    def pin_memory(self) -> "Batch":
        return Batch(self.dense_features.pin_memory(), self.sparse_features.pin_memory(), self.labels.pin_memory())


class LoadFiles:       # csv datapipes of the Avazu tsv path: not provided (the npy path is)
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchrec LoadFiles is not available in the shim")


class ReadLinesFromCSV:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchrec ReadLinesFromCSV is not available in the shim")
