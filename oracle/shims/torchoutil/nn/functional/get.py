from typing import Optional, Union

import torch


def get_device(device: Union[str, torch.device, None] = "cuda_if_available") -> Optional[torch.device]:
    if device is None or isinstance(device, torch.device):
        return device
    if device in ("cuda_if_available", "auto"):
        return torch.device("cuda" if torch.cuda.is_available() else "cpu")
    return torch.device(device)
