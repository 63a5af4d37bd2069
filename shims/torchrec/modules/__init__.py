from . import embedding_modules, mlp  # noqa: F401
