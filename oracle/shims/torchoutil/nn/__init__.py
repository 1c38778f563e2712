from . import functional, modules  # noqa: F401
