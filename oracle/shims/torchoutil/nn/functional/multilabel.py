from typing import Mapping, Union

from torch import Tensor


def probs_to_names(probs: Tensor, threshold: Union[float, Tensor], idx_to_name: Mapping[int, str]) -> list:
    """(B, C) probabilities -> per-row list of class names with prob >= threshold (reference call: model.py:204)."""
    multihot = probs >= threshold
    return [[idx_to_name[int(i)] for i in row.nonzero().flatten().tolist()] for row in multihot]
