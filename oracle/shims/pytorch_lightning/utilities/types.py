from typing import Any

_METRIC_COLLECTION = Any
