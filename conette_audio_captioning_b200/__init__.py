"""conette_audio_captioning_b200: B200-native (sm_100a) CoNeTTE inference hot path behind the reference's Python API.

Public surface: ``CoNeTTEModel`` (drop-in for the reference's ``conette.CoNeTTEModel`` call signature / output dict),
``BaselinePLM`` (the task-agnostic sibling on precomputed embeddings), ``CoNeTTEConfig``, ``Engine`` (operator-level seams over the C ABI) and ``synth`` (offline stand-in weights).  The CUDA
library is built in tree by ``conette_audio_captioning_b200.build.build()``; importing the package does not need a GPU,
calling into it does (no CPU fallback).
"""
from .config import CoNeTTEConfig  # noqa: F401
from .tokenizer import IdTokenizer  # noqa: F401


def __getattr__(name):
    if name in ("CoNeTTEModel",):
        from .model import CoNeTTEModel

        return CoNeTTEModel
    if name in ("conette", "main_predict"):
        from . import predict

        return getattr(predict, name)
    if name in ("BaselinePLM",):
        from .baseline import BaselinePLM

        return BaselinePLM
    if name in ("Engine",):
        from .engine import Engine

        return Engine
    raise AttributeError(name)
