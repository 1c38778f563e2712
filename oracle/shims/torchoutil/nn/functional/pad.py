import torch.nn.functional as F
from torch import Tensor


def pad_dim(x: Tensor, target_length: int, *, dim: int = -1, pad_value: float = 0.0, align: str = "left",
            mode: str = "constant") -> Tensor:
    """Right-pad ``x`` with ``pad_value`` along ``dim`` up to ``target_length`` (reference call: pad.py:15)."""
    missing = max(target_length - x.shape[dim], 0)
    if missing == 0:
        return x
    dim = dim % x.ndim
    pads = [0, 0] * x.ndim
    pads[2 * (x.ndim - 1 - dim) + 1] = missing
    return F.pad(x, pads, mode=mode, value=pad_value)
