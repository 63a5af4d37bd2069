from . import jagged_tensor  # noqa: F401
