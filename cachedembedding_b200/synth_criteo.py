"""Synthetic Criteo-shape data for the cached embedding bag: batches generated on the GPU, the id-frequency counter on the
GPU, and a writer for the on-disk formats the reference's loaders read.

What it replaces / feeds in the reference:
  * recsys/datasets/criteo.py:38-249 (InMemoryBinaryCriteoIterDataPipe) reads ``day_{d}_dense.npy`` (float32 [rows, 13]),
    ``day_{d}_sparse.npy`` (int64 [rows, 26], raw per-table ids -- the loader applies ``% hashes`` and the table offsets)
    and ``day_{d}_labels.npy`` (int32 [rows, 1]); `write_kaggle_format` writes exactly those files, so the reference's
    own data pipeline runs on synthetic data;
  * recsys/datasets/feature_counter.py:21-29 + criteo.py:461-486 (``np.bincount`` over the training files, cached as
    ``id_freq_map.pt``): `IdFrequencyCounter` is the same count as a CUDA histogram kernel (`cebag_id_histogram`);
    `write_kaggle_format` also writes ``id_freq_map.pt`` so that the reference's ``get_id_freq_map`` finds it;
  * the id distribution is the reference's own long-tail generator, ``idx = floor(u ** (-1 / s)) - 1`` with
    ``u ~ U[(1 / N_f) ** s, 1]`` in float64 (baselines/data/custom.py:23,76,89-91), per table.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

# recsys/datasets/criteo.py:29-34 (table cardinalities; constants)
CRITEO_1TB_ROWS = [45833188, 36746, 17245, 7413, 20243, 3, 7114, 1441, 62, 29275261, 1572176, 345138, 10, 2209, 11267,
                   128, 4, 974, 14, 48937457, 11316796, 40094537, 452104, 12606, 104, 35]
CRITEO_KAGGLE_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                      5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
INT_FEATURE_COUNT = 13
DEFAULT_SKEW = 0.25      # baselines/data/custom.py:23


def sample_table_ids(rows: torch.Tensor, batch: int, gen: Optional[torch.Generator], device, skew: float = DEFAULT_SKEW
                     ) -> torch.Tensor:
    """int64 [F, batch]: one id per (table, sample), row 0 of every table the most frequent (long tail)."""
    F = rows.numel()
    n_f = rows.to(torch.float64).view(F, 1)
    lo = (1.0 / n_f) ** skew
    u = torch.rand(F, batch, dtype=torch.float64, device=device, generator=gen) * (1.0 - lo) + lo
    idx = torch.floor(u ** (-1.0 / skew)).long() - 1
    return torch.minimum(idx.clamp_(min=0), rows.view(F, 1) - 1)


def sample_ids(rows: torch.Tensor, batch: int, gen: Optional[torch.Generator], device, skew: float = DEFAULT_SKEW
               ) -> torch.Tensor:
    """One batch of KJT-ordered GLOBAL ids: values[f * B + b] = id of sample b in table f + the table's row offset
    (recsys/datasets/criteo.py:118-119,165-173)."""
    offsets = torch.cumsum(rows, 0) - rows
    return (sample_table_ids(rows, batch, gen, device, skew) + offsets.view(-1, 1)).view(-1)


class IdFrequencyCounter:
    """freq[id] = how often the id occurred in the batches seen so far (int64[N] on the GPU)."""

    def __init__(self, num_rows: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("IdFrequencyCounter needs a CUDA device: there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_rows = int(num_rows)
        self.freq = torch.zeros(self.num_rows, dtype=torch.int64, device=self.device)
        self._bad = torch.zeros(1, dtype=torch.int32, device=self.device)

    def update(self, ids: torch.Tensor) -> None:
        ids = ids.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        _lib.check(_lib.load().cebag_id_histogram(ids.data_ptr(), ids.numel(), self.freq.data_ptr(), self.num_rows,
                                                  self._bad.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def result(self) -> torch.Tensor:
        if int(self._bad.item()):
            raise IndexError(f"an id outside [0, {self.num_rows}) was counted")
        return self.freq


class SyntheticCriteo:
    """Criteo-shape batches made on the GPU: `dense` float32 [B, 13], KJT pieces (`values` int64 [F_loc * B] global or
    rank-local ids, `offsets` int32 [F_loc * B + 1], stride B), `labels` int32 [B].

    assigned_tables: table-wise mode (recsys/datasets/criteo.py:91-96,230): only those tables are emitted and their ids
    are re-based to the concatenation of the assigned tables."""

    def __init__(self, num_embeddings_per_feature: Sequence[int], batch_size: int, skew: float = DEFAULT_SKEW,
                 seed: int = 1024, device=None, assigned_tables: Optional[Sequence[int]] = None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.all_rows = list(num_embeddings_per_feature)
        self.tables = list(range(len(self.all_rows))) if assigned_tables is None else list(assigned_tables)
        self.rows = torch.tensor([self.all_rows[t] for t in self.tables], dtype=torch.long, device=self.device)
        self.batch_size, self.skew = int(batch_size), skew
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        n = len(self.tables) * self.batch_size
        self.offsets = torch.arange(n + 1, dtype=torch.int32, device=self.device)
        self.num_rows = int(self.rows.sum())

    def batch(self) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, int], torch.Tensor]:
        values = sample_ids(self.rows, self.batch_size, self.gen, self.device, self.skew)
        dense = torch.rand(self.batch_size, INT_FEATURE_COUNT, device=self.device, generator=self.gen)
        labels = (torch.rand(self.batch_size, device=self.device, generator=self.gen) < 0.25).int()
        return dense, (values, self.offsets, self.batch_size), labels

    def batches(self, count: int) -> Iterator:
        for _ in range(count):
            yield self.batch()

    def id_freq_map(self, num_batches: int) -> torch.Tensor:
        """Frequencies over `num_batches` fresh batches of this generator (the reference counts its training set)."""
        counter = IdFrequencyCounter(self.num_rows, self.device)
        for _ in range(num_batches):
            counter.update(sample_ids(self.rows, self.batch_size, self.gen, self.device, self.skew))
        return counter.result()


def write_kaggle_format(path: str, rows_per_day: int, num_embeddings_per_feature: Sequence[int] = CRITEO_KAGGLE_ROWS,
                        days: int = 7, skew: float = DEFAULT_SKEW, seed: int = 1024, device=None) -> List[str]:
    """Write `days` days of synthetic samples in the npy format of scripts/preprocess/npy_preproc_criteo.py /
    split_criteo_kaggle.py:26-30 (what recsys/datasets/criteo.py:377-412 lists and loads), plus ``id_freq_map.pt``
    counted on the GPU over the training days (all but the last, like the reference's split).  Returns the file names."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    os.makedirs(path, exist_ok=True)
    rows = torch.tensor(list(num_embeddings_per_feature), dtype=torch.long, device=device)
    gen = torch.Generator(device=device).manual_seed(seed)
    counter = IdFrequencyCounter(int(rows.sum()), device)
    table_offsets = (torch.cumsum(rows, 0) - rows).view(1, -1)
    written = []
    for d in range(days):
        ids = sample_table_ids(rows, rows_per_day, gen, device, skew).t().contiguous()       # [rows, 26] per-table ids
        dense = torch.rand(rows_per_day, INT_FEATURE_COUNT, device=device, generator=gen)
        labels = (torch.rand(rows_per_day, 1, device=device, generator=gen) < 0.25).int()
        if d < days - 1:
            counter.update(ids + table_offsets)
        for kind, arr in (("dense", dense.float()), ("sparse", ids), ("labels", labels)):
            name = os.path.join(path, f"day_{d}_{kind}.npy")
            np.save(name, arr.cpu().numpy())
            written.append(name)
    name = os.path.join(path, "id_freq_map.pt")
    torch.save(counter.result().cpu(), name)
    written.append(name)
    return written
