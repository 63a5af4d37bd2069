from . import criteo, utils  # noqa: F401
