from . import parallel  # noqa: F401
