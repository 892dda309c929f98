from . import op, model  # noqa: F401
