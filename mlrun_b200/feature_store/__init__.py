from . import steps  # noqa: F401
