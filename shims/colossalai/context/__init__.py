from . import parallel_mode  # noqa: F401
from .parallel_mode import ParallelMode  # noqa: F401
