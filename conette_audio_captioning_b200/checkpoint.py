"""Reference checkpoints -> (state dict by the reference's tensor names, vocabulary, config)  [SURVEY.md 8f rank 2].

Mirrors what ``CoNeTTEModel.from_pretrained`` / ``_pre_hook_load_state_dict`` do to a stored state dict
(reference huggingface/model.py:125-183): non-tensor entries travel pickled inside the ``_extra_state_`` uint8 tensor, the
fitted vocabulary is the ``itos`` table of ``model.tokenizers.0._extra_state`` (tokenization/aac_tokenizer.py:819-837), and
ConvNeXt checkpoints that still call the layer scale ``gamma`` are renamed to ``scale_layer`` (nn/encoders/convnext.py:76-102).
Nothing here touches the GPU; the tensors are handed to ``Engine`` / ``cnb_load_weight`` by name."""
from __future__ import annotations

import io
import json
import os
import os.path as osp
import pickle
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from .config import CoNeTTEConfig

TOKENIZER_STATE_KEY = "model.tokenizers.0._extra_state"
WEIGHT_FILES = ("model.safetensors", "pytorch_model.bin", "model.bin", "model.pt", "state_dict.pt")


class _PlainUnpickler(pickle.Unpickler):
    """The extra state is plain data (dict / list / str / int / float / bool / None).  Refuse everything else: a checkpoint
    must not be able to run code on load."""

    _ALLOWED = {("collections", "OrderedDict"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "set"),
                ("builtins", "tuple"), ("builtins", "frozenset")}

    def find_class(self, module: str, name: str):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name} from a checkpoint's extra state")


def unpack_extra_state(sd: Dict[str, Any]) -> Dict[str, Any]:
    """``_extra_state_`` (uint8 tensor holding a pickled dict) -> entries merged back (model.py:134-138)."""
    sd = dict(sd)
    if "_extra_state_" in sd:
        raw = bytes(sd.pop("_extra_state_").to(torch.uint8).cpu().numpy().tobytes())
        extra = _PlainUnpickler(io.BytesIO(raw)).load()
        if not isinstance(extra, dict):
            raise TypeError("_extra_state_ does not hold a dict")
        sd.update(extra)
    return sd


def rename_legacy_keys(sd: Dict[str, Any]) -> Dict[str, Any]:
    """``gamma`` -> ``scale_layer`` (convnext.py:76-102); both present at once is an error, as in the reference."""
    out = dict(sd)
    for key in list(out.keys()):
        if "gamma" not in key:
            continue
        new_key = key.replace("gamma", "scale_layer")
        if new_key in out:
            raise RuntimeError(f"Invalid state_dict conversion. (found {key} and {new_key} at the same time)")
        out[new_key] = out.pop(key)
    return out


def vocabulary(sd: Dict[str, Any]) -> List[str]:
    """id -> token table of the fitted tokenizer stored with the weights."""
    state = sd.get(TOKENIZER_STATE_KEY)
    if not isinstance(state, dict) or "tokenizer" not in state:
        raise KeyError(f"the checkpoint holds no fitted tokenizer ({TOKENIZER_STATE_KEY})")
    tok = state["tokenizer"]
    itos = tok.get("itos", tok.get("_itos"))
    if isinstance(itos, dict):
        itos = [itos[i] for i in range(len(itos))]
    if not isinstance(itos, (list, tuple)) or not all(isinstance(t, str) for t in itos):
        raise TypeError("tokenizer state has no usable 'itos' table")
    return list(itos)


def read_tensors(path: str) -> Dict[str, Any]:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file

        return dict(load_file(path))
    data = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(data, dict) and "state_dict" in data and isinstance(data["state_dict"], dict):
        data = data["state_dict"]  # Lightning checkpoint (predict.py:166-169)
    return dict(data)


def load_checkpoint(path: str) -> Tuple[Dict[str, Any], List[str], CoNeTTEConfig]:
    """``path`` = a weight file, or a directory in the Hugging Face layout (weights + ``config.json``)."""
    cfg_kwargs: Dict[str, Any] = {}
    if osp.isdir(path):
        cfg_path = osp.join(path, "config.json")
        if osp.isfile(cfg_path):
            raw = json.load(open(cfg_path))
            fields = CoNeTTEConfig.__dataclass_fields__ if hasattr(CoNeTTEConfig, "__dataclass_fields__") else {}
            cfg_kwargs = {k: v for k, v in raw.items() if k in fields}
        found = [osp.join(path, f) for f in WEIGHT_FILES if osp.isfile(osp.join(path, f))]
        if not found:
            raise FileNotFoundError(f"no weight file ({', '.join(WEIGHT_FILES)}) under '{path}'")
        path = found[0]
    elif not osp.isfile(path):
        raise FileNotFoundError(path)
    sd = rename_legacy_keys(unpack_extra_state(read_tensors(path)))
    if not any(k.startswith("preprocessor.encoder.") for k in sd):
        raise KeyError("the checkpoint has no 'preprocessor.encoder.*' tensors (a CoNeTTEModel state dict is expected)")
    itos = vocabulary(sd)
    if "task_names" in cfg_kwargs and isinstance(cfg_kwargs["task_names"], list):
        cfg_kwargs["task_names"] = tuple(cfg_kwargs["task_names"])
    return sd, itos, CoNeTTEConfig(**cfg_kwargs)


def pack_for_saving(sd: Dict[str, Tensor], itos: List[str]) -> Dict[str, Tensor]:
    """The inverse, used by tests and by users who want to store synthetic weights in the reference's layout
    (model.py:165-183): tokenizer state pickled into ``_extra_state_``."""
    state = {TOKENIZER_STATE_KEY: {"_target_": "conette.tokenization.aac_tokenizer.AACTokenizer", "_version_": "2.1.0",
                                   "_type_": "txt", "tokenizer": {"itos": dict(enumerate(itos)),
                                                                  "stoi": {t: i for i, t in enumerate(itos)}}}}
    out = {k: v.contiguous() for k, v in sd.items() if isinstance(v, Tensor)}
    out["_extra_state_"] = torch.frombuffer(bytearray(pickle.dumps(state)), dtype=torch.uint8).clone()
    return out
