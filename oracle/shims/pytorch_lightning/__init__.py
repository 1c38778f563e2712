"""Oracle shim for ``pytorch-lightning==1.9.5`` (reference requirements.txt).  TEST INFRASTRUCTURE ONLY.

The inference path only needs ``LightningModule`` as an ``nn.Module`` with ``save_hyperparameters`` / ``hparams`` /
``device`` / ``dtype`` (reference call sites: pl_modules/base.py:30,134-137; pl_modules/conette.py:93,138).
"""
import inspect

import torch
from torch import nn


class AttributeDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, val):
        self[key] = val


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs) -> None:
        super().__init__()
        self._hparams = AttributeDict()
        self._hparams_initial = AttributeDict()
        self._trainer = None

    def save_hyperparameters(self, *args, ignore=None, frame=None, logger=True) -> None:
        if frame is None:
            frame = inspect.currentframe().f_back
        init_self = frame.f_locals.get("self", None)
        cls = type(init_self) if init_self is not None else type(self)
        params = inspect.signature(cls.__init__).parameters
        local_vars = frame.f_locals
        hp = {}
        for name, p in params.items():
            if name == "self" or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
                continue
            if name in local_vars:
                hp[name] = local_vars[name]
        if ignore is not None:
            if isinstance(ignore, str):
                ignore = (ignore,)
            for k in ignore:
                hp.pop(k, None)
        self._hparams.update(hp)
        self._hparams_initial = AttributeDict(dict(hp))

    @property
    def hparams(self):
        return self._hparams

    @property
    def hparams_initial(self):
        return self._hparams_initial

    @property
    def device(self) -> torch.device:
        for t in list(self.parameters()) + list(self.buffers()):
            return t.device
        return torch.device("cpu")

    @property
    def dtype(self) -> torch.dtype:
        for t in self.parameters():
            return t.dtype
        return torch.float32

    @property
    def trainer(self):
        return self._trainer

    def log(self, *args, **kwargs) -> None:
        return None

    def log_dict(self, *args, **kwargs) -> None:
        return None


class LightningDataModule:
    pass


class Trainer:
    pass
