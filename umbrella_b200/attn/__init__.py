from .cache import KV_Cache  # noqa: F401
