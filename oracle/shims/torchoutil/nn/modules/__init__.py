from torch import nn

from .tensor import Transpose  # noqa: F401


class CropDim(nn.Module):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__()

    def forward(self, x):
        raise NotImplementedError("oracle shim: training-only module (speed_perturb.py)")


class PadDim(CropDim):
    pass
