"""SpecAugmentation stub: only applied when ``self.training`` (reference convnext.py:294-295), identity here."""
from torch import nn


class SpecAugmentation(nn.Module):
    def __init__(self, time_drop_width, time_stripes_num, freq_drop_width, freq_stripes_num) -> None:
        super().__init__()
        self.time_drop_width = time_drop_width
        self.time_stripes_num = time_stripes_num
        self.freq_drop_width = freq_drop_width
        self.freq_stripes_num = freq_stripes_num

    def forward(self, x):
        if self.training:
            raise NotImplementedError("oracle shim: SpecAugmentation is training-only and out of scope")
        return x
