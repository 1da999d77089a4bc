from . import basic, op  # noqa: F401
