"""torchoutil.nn.functional: helpers restated from their documented behaviour."""
from typing import Iterable, Optional, Union

import torch
from torch import Tensor, nn

from .get import get_device  # noqa: F401
from .mask import masked_mean  # noqa: F401
from .multilabel import probs_to_names  # noqa: F401
from .pad import pad_dim  # noqa: F401


def generate_square_subsequent_mask(size: int, device=None, dtype=torch.float32) -> Tensor:
    """(size, size) float mask: 0 on/below the diagonal, -inf above (nn.Transformer convention)."""
    mask = torch.full((size, size), float("-inf"), device=device, dtype=dtype)
    return torch.triu(mask, diagonal=1)


def indices_to_multihot(indices: Tensor, num_classes: int, dtype=torch.bool, device=None) -> Tensor:
    """(..., L) int -> (..., num_classes) multi-hot with True at every listed index."""
    if device is None:
        device = indices.device
    out = torch.zeros(tuple(indices.shape[:-1]) + (num_classes,), dtype=dtype, device=device)
    src = torch.ones((), dtype=dtype, device=device).expand(indices.shape)
    out.scatter_(-1, indices.to(device=device, dtype=torch.long), src)
    return out


def repeat_interleave_nd(x: Tensor, repeats: int, dim: int = 0) -> Tensor:
    return x.repeat_interleave(repeats, dim=dim)


def tensor_to_lengths(x: Tensor, *, pad_value=None, end_value=None, dim: int = -1) -> Tensor:
    """Index of the first ``end_value`` along ``dim`` (or the full length when absent)."""
    if (pad_value is None) == (end_value is None):
        raise ValueError("exactly one of pad_value / end_value is expected")
    if end_value is not None:
        contains = x == end_value
        idx = contains.long().argmax(dim=dim)
        return torch.where(contains.any(dim=dim), idx, torch.full_like(idx, x.shape[dim]))
    non_pad = x != pad_value
    return non_pad.long().sum(dim=dim)


def lengths_to_pad_mask(lengths: Tensor, max_len: Union[int, Tensor, None] = None, include: bool = True) -> Tensor:
    """(B,) -> (B, max_len) bool, True where position >= length (fed to memory_key_padding_mask, conette.py:462)."""
    if max_len is None:
        max_len = int(lengths.max().item())
    max_len = int(max_len)
    ar = torch.arange(max_len, device=lengths.device)
    return ar.unsqueeze(0) >= lengths.unsqueeze(1)


def tensor_to_pad_mask(x: Tensor, *, pad_value=None, end_value=None) -> Tensor:
    if pad_value is not None:
        return x == pad_value
    raise NotImplementedError("oracle shim: end_value variant unused on the inference path")


def randperm_diff(n: int, device=None, generator=None) -> Tensor:
    raise NotImplementedError("oracle shim: training-only helper")


def count_parameters(model: nn.Module, *, recurse: bool = True, only_trainable: bool = False) -> int:
    return sum(p.numel() for p in model.parameters(recurse) if (not only_trainable or p.requires_grad))
