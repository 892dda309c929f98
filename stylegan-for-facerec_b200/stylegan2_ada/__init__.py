from .generator import Generator  # noqa: F401
