"""Stand-in for the parts of torchrec the reference's recsys/ tree imports: the jagged sparse-input
container, the Batch record, the Criteo constants + npy helpers, and the MLP block of the dense arches."""
from .sparse.jagged_tensor import KeyedJaggedTensor, KeyedTensor  # noqa: F401
from . import datasets, modules, sparse  # noqa: F401
