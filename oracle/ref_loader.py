"""Import the reference's own (unmodified) modules and build its ``CoNeTTEModel`` on CPU.  TEST INFRASTRUCTURE ONLY.

The reference package cannot be imported the normal way: ``conette/__init__.py:19-20`` pulls ``pytorch_lightning``,
``nltk``, ``spacy``, ``torchoutil`` and ``torchlibrosa``, none of which is installed (no network).  We register a bare
namespace module named ``conette`` whose ``__path__`` points at the reference sources (all sub-package ``__init__``
files are empty), put ``oracle/shims`` on ``sys.path`` and import the reference's files as they are.

Source resolution order: ``$CONETTE_REF_SRC``, ``/root/reference/src`` (dev container), ``<repo>/baseline/_ref``
(``pip install --target`` copy that travels to the GPU box).  ``available()`` tells callers whether any exists.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path
from typing import Any, Dict, Optional

import torch

_REPO = Path(__file__).resolve().parent.parent
_SHIMS = Path(__file__).resolve().parent / "shims"
_CANDIDATES = (
    os.environ.get("CONETTE_REF_SRC", ""),
    "/root/reference/src",
    str(_REPO / "baseline" / "_ref"),
)


def ref_src() -> Optional[str]:
    for c in _CANDIDATES:
        if c and (Path(c) / "conette" / "nn" / "decoding" / "beam.py").is_file():
            return c
    return None


def available() -> bool:
    return ref_src() is not None


def _install() -> None:
    src = ref_src()
    if src is None:
        raise ImportError("reference sources not found (looked in: %s)" % ", ".join(c for c in _CANDIDATES if c))
    if str(_SHIMS) not in sys.path:
        sys.path.insert(0, str(_SHIMS))
    if "conette" not in sys.modules or not getattr(sys.modules["conette"], "_oracle_namespace", False):
        pkg = types.ModuleType("conette")
        pkg.__path__ = [str(Path(src) / "conette")]  # skip conette/__init__.py
        pkg._oracle_namespace = True
        sys.modules["conette"] = pkg
    # the AudioSet class-name CSV is a network download (transforms/audioset_mapping.py:12-17)
    am = importlib.import_module("conette.transforms.audioset_mapping")
    am.load_audioset_idx_to_name = lambda offline=False, cache_path=None, verbose=0: {i: f"class{i}" for i in range(527)}


def ref_module(name: str):
    """Import ``conette.<name>`` from the reference sources."""
    _install()
    return importlib.import_module(f"conette.{name}")


def build_reference_model(state_dict: Dict[str, torch.Tensor], corpus: list, *, verbose: int = 0) -> Any:
    """Reference ``CoNeTTEModel`` (huggingface/model.py:38) on CPU, fp32, eval, loaded with ``state_dict``.

    ``corpus`` is fitted by the reference ``AACTokenizer`` (spacy shim = whitespace split) so that token ids follow the
    reference's own first-appearance order; task tokens are appended by ``CoNeTTEPLM.build_model``.
    """
    _install()
    model_mod = importlib.import_module("conette.huggingface.model")
    # `load_audioset_idx_to_name` was imported by name into model.py: patch there too
    model_mod.load_audioset_idx_to_name = lambda offline=False, verbose=0: {i: f"class{i}" for i in range(527)}
    config_mod = importlib.import_module("conette.huggingface.config")
    plm_mod = importlib.import_module("conette.pl_modules.conette")
    tok_mod = importlib.import_module("conette.tokenization.aac_tokenizer")

    tokenizer = tok_mod.AACTokenizer()
    tokenizer.fit(list(corpus))
    plm = plm_mod.CoNeTTEPLM(train_tokenizer=tokenizer, verbose=verbose)
    config = config_mod.CoNeTTEConfig()
    model = model_mod.CoNeTTEModel(config, device="cpu", offline=True, model_override=plm)

    own = model.state_dict()
    missing = [k for k in own if k not in state_dict and k != "_extra_state_"]
    extra = [k for k in state_dict if k not in own]
    if missing or extra:
        raise RuntimeError(f"synthetic state dict does not match the reference: missing={missing[:5]} extra={extra[:5]}")
    with torch.no_grad():
        params = dict(model.named_parameters())
        bufs = dict(model.named_buffers())
        for k, v in state_dict.items():
            dst = params.get(k, bufs.get(k))
            if dst is None:
                raise RuntimeError(f"no parameter/buffer named {k} in the reference model")
            if tuple(dst.shape) != tuple(v.shape):
                raise RuntimeError(f"shape mismatch for {k}: reference {tuple(dst.shape)} vs synthetic {tuple(v.shape)}")
            dst.copy_(v.to(dst.dtype))
    model.eval_and_disable_grad()
    return model


def reference_itos(model: Any) -> list:
    tok = model.model.tokenizer
    return [tok.id_to_token(i) for i in range(tok.get_vocab_size())]
