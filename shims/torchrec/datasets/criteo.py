"""torchrec.datasets.criteo: the Criteo constants and the npy helpers the in-memory loader calls
(recsys/datasets/criteo.py:13-14,151-178).  On-disk format: day_{d}_dense.npy float32 [rows, 13], day_{d}_sparse.npy
int64/int32 [rows, 26], day_{d}_labels.npy int32 [rows, 1] (scripts/preprocess/npy_preproc_criteo.py)."""
from typing import Dict, List, Tuple

import numpy as np

INT_FEATURE_COUNT = 13
CAT_FEATURE_COUNT = 26
DAYS = 24
DEFAULT_LABEL_NAME = "label"
DEFAULT_INT_NAMES: List[str] = [f"int_{idx}" for idx in range(INT_FEATURE_COUNT)]
DEFAULT_CAT_NAMES: List[str] = [f"cat_{idx}" for idx in range(CAT_FEATURE_COUNT)]
DEFAULT_COLUMN_NAMES: List[str] = [DEFAULT_LABEL_NAME, *DEFAULT_INT_NAMES, *DEFAULT_CAT_NAMES]


class BinaryCriteoUtils:

    @staticmethod
    def get_shape_from_npy(path: str, path_manager_key: str = "torchrec") -> Tuple[int, ...]:
        """Shape of the array stored in `path`, read from the npy header only."""
        return tuple(np.load(path, mmap_mode="r").shape)

    @staticmethod
    def get_file_idx_to_row_range(lengths: List[int], rank: int, world_size: int) -> Dict[int, Tuple[int, int]]:
        """The rows of the concatenation of all files are dealt out to the ranks in contiguous, near-equal shares (the
        first `total % world_size` ranks get one more).  Returns, for this rank, file index -> (first row, last row),
        both inclusive and relative to that file."""
        total = sum(lengths)
        share, rem = divmod(total, world_size)
        begin = rank * share + min(rank, rem)
        end = begin + share + (1 if rank < rem else 0)          # exclusive, in global rows
        out: Dict[int, Tuple[int, int]] = {}
        file_begin = 0
        for idx, n in enumerate(lengths):
            lo, hi = max(begin, file_begin), min(end, file_begin + n)
            if lo < hi:
                out[idx] = (lo - file_begin, hi - 1 - file_begin)
            file_begin += n
        return out

    @staticmethod
    def load_npy_range(fname: str, start_row: int, num_rows: int, path_manager_key: str = "torchrec",
                       mmap_mode: bool = False) -> np.ndarray:
        """Rows [start_row, start_row + num_rows) of a 2-D npy file."""
        data = np.load(fname, mmap_mode="r")
        if start_row + num_rows > data.shape[0]:
            raise ValueError(f"rows {start_row}..{start_row + num_rows} are outside {fname} ({data.shape[0]} rows)")
        view = data[start_row:start_row + num_rows]
        return view if mmap_mode else np.ascontiguousarray(view)
