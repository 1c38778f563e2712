"""Host side of the GPU polyphase resampler.

The reference resamples with ``torchaudio.functional.resample(x, sr, 32000)`` (huggingface/preprocessor.py:139-141;
torchaudio 0.13.1 defaults: ``sinc_interp_hann``, ``lowpass_filter_width=6``, ``rolloff=0.99``).  ``filter_bank`` restates
that filter-bank formula in the same dtype (``functional.resample`` passes ``dtype=waveform.dtype``: the whole bank is
evaluated in float32) and compacts it to the support of every phase: outside 6 zero crossings the clamped Hann window is
cos(pi/2)^2, which float32 evaluates to ~1e-15 instead of 0, so those taps are ~1e-23 -- they are dropped
(``NEGLIGIBLE_TAP``; their total contribution is < 1e-9 of the input peak, far below fp32 resolution of the output).  The
convolution itself runs in ``cnb_resample`` (csrc/resample.cu).
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import torch
from torch import Tensor

LOWPASS_FILTER_WIDTH = 6
ROLLOFF = 0.99
NEGLIGIBLE_TAP = 1e-12
MAX_BANK_FLOATS = 1 << 20  # exotic ratios (e.g. 44101 -> 32000) would need a gigantic bank: rejected, not approximated


def reduced_ratio(orig_sr: int, new_sr: int) -> Tuple[int, int]:
    if int(orig_sr) != orig_sr or int(new_sr) != new_sr or orig_sr <= 0 or new_sr <= 0:
        raise ValueError(f"Sample rates must be positive integers (found {orig_sr=} and {new_sr=}).")
    g = math.gcd(int(orig_sr), int(new_sr))
    return int(orig_sr) // g, int(new_sr) // g


def resampled_length(n: int, orig: int, new: int) -> int:
    """ceil(new * n / orig) -- the ``target_length`` of torchaudio's ``_apply_sinc_resample_kernel``."""
    return (new * n + orig - 1) // orig


@lru_cache(maxsize=16)
def filter_bank(orig: int, new: int) -> Tuple[Tensor, Tensor, int]:
    """-> taps (new, n_taps) f32, tap_lo (new,) i32, width; ``orig``/``new`` already divided by their gcd."""
    base_freq = min(orig, new) * ROLLOFF
    width = math.ceil(LOWPASS_FILTER_WIDTH * orig / base_freq)
    if new * (2 * width + orig) > 64 * MAX_BANK_FLOATS:
        raise ValueError(f"Unsupported sample-rate ratio {orig}:{new} (filter bank too large).")
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None] / new + idx
    t *= base_freq
    t = t.clamp_(-LOWPASS_FILTER_WIDTH, LOWPASS_FILTER_WIDTH)
    window = torch.cos(t * math.pi / LOWPASS_FILTER_WIDTH / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale  # (new, 2 * width + orig) f32
    nz = kernels.abs() > NEGLIGIBLE_TAP
    k_dense = kernels.shape[1]
    cols = torch.arange(k_dense)
    lo = torch.where(nz, cols, k_dense).min(dim=1).values
    hi = torch.where(nz, cols, -1).max(dim=1).values
    lo = torch.minimum(lo, torch.full_like(lo, k_dense - 1))
    n_taps = int((hi - lo).max().item()) + 1
    if new * n_taps > MAX_BANK_FLOATS:
        raise ValueError(f"Unsupported sample-rate ratio {orig}:{new} (filter bank too large).")
    lo = torch.minimum(lo, torch.full_like(lo, k_dense - n_taps)).clamp_(min=0)
    gather = (lo[:, None] + torch.arange(n_taps)[None]).clamp_(max=k_dense - 1)
    taps = torch.gather(kernels, 1, gather)
    taps = torch.where(lo[:, None] + torch.arange(n_taps)[None] < k_dense, taps, torch.zeros_like(taps))
    return taps.contiguous(), lo.to(torch.int32).contiguous(), width
