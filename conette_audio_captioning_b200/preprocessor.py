"""Host mirror of the reference ``CoNeTTEPreprocessor._load_resample`` (huggingface/preprocessor.py:82-154).

Same argument forms and error behaviour: str / list[str] paths, Tensor (N,) / (C, N) [ONE clip with C channels] /
(B, C, N), list[Tensor (C, N_i)]; ``sr`` int / list / None; ``x_shapes`` overrides the inferred lengths and may not be
combined with resampling.  Output: mono waveforms right-zero-padded to the batch max (B, Nmax) f32 on the HOST in pinned
memory (the single H2D copy happens inside ``cnb_caption_host``) plus the true lengths (B,) i64.

Resampling (only when sr != 32 kHz; never in the benchmark configs) runs on the GPU (``cnb_resample``, SURVEY.md 8(f)
rank 1) when the caller passes the engine's ``resample`` method: clips are mixed down to mono on the host first (the
reference resamples every channel and then takes the mean, preprocessor.py:139-146 -- both are linear, so the order only
changes fp32 rounding; this uploads 1/C of the bytes), grouped by sample rate, right-zero-padded per group, resampled on the
device with their true lengths and written into one (B, Nmax) DEVICE batch.  Without a resampler callable the function raises
for sr != 32 kHz: the product path has no host fallback.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple, Union

import torch
from torch import Size, Tensor

TARGET_SR = 32_000


def _load(path: str) -> Tuple[Tensor, int]:
    from .audio_io import load_audio

    return load_audio(path)


def _is_iterable_str(x) -> bool:
    return isinstance(x, str) or (isinstance(x, Iterable) and not isinstance(x, Tensor) and all(isinstance(xi, str) for xi in x))


def load_resample(
    x: Union[Tensor, str, Iterable[str], Iterable[Tensor]],
    sr: Union[None, int, Iterable[int]] = None,
    x_shapes: Union[Tensor, None, Sequence[Size]] = None,
    resampler: Optional[Callable] = None,
) -> Tuple[Tensor, Tensor]:
    if _is_iterable_str(x):
        if isinstance(x, str):
            x = [x]
        loaded = [_load(xi) for xi in x]
        x = [w for w, _ in loaded]
        sr = [s for _, s in loaded]
    else:
        if isinstance(x, Tensor):
            if x.ndim == 1:
                x = x.unsqueeze(0).unsqueeze(1)
            elif x.ndim == 2:
                x = x.unsqueeze(0)  # (channels, time) = ONE clip (reference preprocessor.py:99-100)
            elif x.ndim == 3:
                pass
            else:
                raise ValueError(f"Invalid argument shape x.shape={tuple(x.shape)}.")
        else:
            x = list(x)  # type: ignore[arg-type]
        if isinstance(sr, int):
            sr = [sr]
        elif sr is None:
            sr = [TARGET_SR]
        else:
            sr = list(sr)

    if len(sr) == 1 and len(x) != len(sr):
        sr = list(sr) * len(x)
    assert len(x) == len(sr) and len(x) > 0

    if any(sri != TARGET_SR for sri in sr):
        if x_shapes is not None:
            raise ValueError(f"Invalid argument x_shapes={x_shapes}.")
        if resampler is None:
            raise RuntimeError("load_resample: sr != 32 kHz needs the engine's GPU resampler (pass resampler=engine.resample)")
        return _resample_on_device(x, [int(s) for s in sr], resampler)

    clips: List[Tensor] = [xi.float().mean(dim=0) for xi in x]  # mono mix (preprocessor.py:143-146)
    if x_shapes is None:
        lens = torch.tensor([c.shape[-1] for c in clips], dtype=torch.int64)
    else:
        xs = torch.as_tensor(x_shapes)
        lens = xs.reshape(len(clips), -1)[:, -1].to(torch.int64).cpu()  # last column = time length (convnext.py:312)
    return _pad_stack(clips), lens


def _pad_stack(clips: Sequence[Tensor]) -> Tensor:
    n_max = max(c.shape[-1] for c in clips)
    pin = torch.cuda.is_available()
    out = torch.zeros(len(clips), n_max, dtype=torch.float32, pin_memory=pin)
    for i, c in enumerate(clips):
        out[i, : c.shape[-1]] = c  # right zero-pad to the batch max (nn/functional/pad.py:11-17)
    return out


def _resample_on_device(x, sr: List[int], resampler: Callable) -> Tuple[Tensor, Tensor]:
    """Mono mix on the host, one GPU resample per distinct sample rate, one (B, Nmax) device batch."""
    from .resample import reduced_ratio, resampled_length

    clips = [xi.float().mean(dim=0) for xi in x]
    out_lens = []
    for c, s in zip(clips, sr):
        o, n = reduced_ratio(s, TARGET_SR)
        out_lens.append(c.shape[-1] if s == TARGET_SR else resampled_length(c.shape[-1], o, n))
    n_max = max(out_lens)
    batch: Optional[Tensor] = None
    for s in sorted(set(sr), key=lambda v: (v == TARGET_SR, v)):  # resampled groups first: they fix the device of the batch
        rows = [i for i, si in enumerate(sr) if si == s]
        group = _pad_stack([clips[i] for i in rows])
        if s != TARGET_SR:
            glens = torch.tensor([clips[i].shape[-1] for i in rows], dtype=torch.int64)
            dev_rows, _ = resampler(group, s, glens, TARGET_SR, n_max)
            if batch is None:
                batch = torch.zeros(len(clips), n_max, dtype=torch.float32, device=dev_rows.device)
            batch[torch.tensor(rows, device=batch.device)] = dev_rows
        else:
            batch[torch.tensor(rows, device=batch.device), : group.shape[1]] = group.to(batch.device, non_blocking=True)
    assert batch is not None
    return batch, torch.tensor(out_lens, dtype=torch.int64)
