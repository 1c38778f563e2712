"""Inference-relevant subset of the reference ``CoNeTTEConfig`` (huggingface/config.py:13-88), same field names."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

from .synth import TASK_NAMES


@dataclass
class CoNeTTEConfig:
    task_mode: str = "ds_src"
    task_names: Tuple[str, ...] = field(default_factory=lambda: tuple(TASK_NAMES))
    min_pred_size: int = 3
    max_pred_size: int = 20
    beam_size: int = 3
    nhead: int = 8
    d_model: int = 256
    num_decoder_layers: int = 6
    dim_feedforward: int = 2048
    acti_name: str = "gelu"
    proj_name: str = "lin768"
    verbose: int = 0

    def __post_init__(self) -> None:
        fixed = dict(nhead=8, d_model=256, num_decoder_layers=6, dim_feedforward=2048, acti_name="gelu", proj_name="lin768")
        for k, v in fixed.items():
            if getattr(self, k) != v:
                raise ValueError(f"the CUDA kernels are specialised on {k}={v} (found {getattr(self, k)})")
        if self.task_mode not in ("ds_src", "ds", "none"):
            raise ValueError(f"Invalid argument task_mode={self.task_mode!r}.")
