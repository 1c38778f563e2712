from . import collections  # noqa: F401
