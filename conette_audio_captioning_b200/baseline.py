"""``BaselinePLM``: the reference's task-agnostic sibling of CoNeTTE on the same CUDA decoder.

Mirrors the inference surface of reference ``pl_modules/baseline.py:35`` -- ``forward(batch, decode_method=...)`` (:310-335),
``decode_audio`` (:339-402), ``encode_audio`` (:404-418): a ``FrameIdentEncoder`` (precomputed frame embeddings pass through,
nn/encoders/ident.py:14-34), the ``lin768`` projection and the same ``AACTransformerDecoder``; every clip starts from the plain
``<bos>`` token (id 1) instead of a task token, and ``beam_size`` / ``min_pred_size`` / ``max_pred_size`` are hyper-parameters of
the module.  The three decode methods of the reference are provided:

  ``generate``  beam search (nn/decoding/beam.py:22)        -> dict with cands / preds / lprobs / mult_* like the reference
  ``forcing``   teacher forcing (nn/decoding/forcing.py:12)  -> logits (B, V, L)
  ``greedy``    greedy search (nn/decoding/greedy.py:17)     -> masked logits (B, V, pred_size)

All arithmetic runs in libconette_b200.so (``cnb_decode`` / ``cnb_decode_tap`` / ``cnb_decoder_logits``); a state dict without
audio-encoder weights is enough.  Training (mixup, label smoothing, optimiser) is out of scope.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Sequence, Union

import torch
from torch import Tensor

from .engine import Engine
from .tokenizer import IdTokenizer

PAD_ID, BOS_ID, EOS_ID = 0, 1, 2  # reference tokenization/constants.py:15, tokenizers/common.py:8-19


class BaselinePLM:
    def __init__(
        self,
        state_dict: Dict[str, Tensor],
        tokenizer: Union[Sequence[str], Any],
        min_pred_size: int = 3,
        max_pred_size: int = 20,
        beam_size: int = 3,
        device: Union[int, str, torch.device] = 0,
        precision: str = "fast",
        decoder: str = "auto",
    ) -> None:
        """``state_dict``: a reference ``BaselinePLM.state_dict()`` (keys ``projection.2.*``, ``decoder.*``, ``forbid_rep_mask``) or
        a CoNeTTE one (same keys under ``model.``; encoder tensors, if present, are loaded too but never used here)."""
        if isinstance(tokenizer, (list, tuple)):
            tokenizer = IdTokenizer(tokenizer)
        self.tokenizer = tokenizer
        vocab = tokenizer.get_vocab_size()
        sd = {(k if k.startswith(("model.", "preprocessor.")) else f"model.{k}"): v for k, v in state_dict.items()}
        if sd["model.decoder.classifier.weight"].shape[0] != vocab:
            raise ValueError("vocabulary size does not match decoder.classifier.weight")
        if isinstance(device, str) and device in ("cuda_if_available", "auto"):
            device = "cuda"
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError(f"Invalid argument device={device!r}. (conette_b200 runs on a CUDA device only)")
        self.engine = Engine(sd, vocab, dev.index or 0, precision, 0, decoder)
        fm = sd.get("model.forbid_rep_mask")
        self.forbid_rep_mask = None if fm is None else fm.to("cpu", torch.uint8)
        self.min_pred_size, self.max_pred_size, self.beam_size = min_pred_size, max_pred_size, beam_size
        self.vocab_size = vocab

    # reference properties (pl_modules/base.py)
    pad_id, bos_id, eos_id = PAD_ID, BOS_ID, EOS_ID

    def decode_text(self, preds: Tensor):
        return self.tokenizer.decode_rec(preds)

    def encode_audio(self, audio: Tensor, audio_shape: Tensor) -> Dict[str, Tensor]:
        """FrameIdentEncoder: ``audio`` (B, T', 768) are the frame embeddings, ``audio_shape[:, 1]`` their valid lengths.  The
        projection and the pad mask of reference :404-418 happen inside the decode entry points of the library."""
        return {"frame_embs": audio, "frame_embs_lens": audio_shape[:, 1].to(torch.int32)}

    def forward(self, batch: Dict[str, Any], decode_method: str = "generate", **kwargs: Any):
        enc = self.encode_audio(batch["audio"], batch["audio_shape"])
        if decode_method == "forcing" and "captions" in batch:
            kwargs["caps_in"] = batch["captions"][:, :-1]
        outs = self.decode_audio(enc, decode_method, **kwargs)
        if decode_method == "generate":
            preds, lprobs, mult_preds, mult_lprobs = outs
            return {"cands": self.decode_text(preds), "preds": preds, "lprobs": lprobs, "mult_cands": self.decode_text(mult_preds),
                    "mult_preds": mult_preds, "mult_lprobs": mult_lprobs}
        return outs

    __call__ = forward

    def decode_audio(self, encoder_outs: Dict[str, Tensor], decode_method: str, **kwargs: Any):
        fe, lens = encoder_outs["frame_embs"], encoder_outs["frame_embs_lens"]
        b = fe.shape[0]
        bos = torch.full((b,), int(kwargs.get("bos_id", BOS_ID)), dtype=torch.int64)
        forbid = kwargs.get("forbid_rep_mask", self.forbid_rep_mask)
        min_len = int(kwargs.get("min_pred_size", self.min_pred_size))
        max_len = int(kwargs.get("max_pred_size", self.max_pred_size))
        if decode_method == "forcing":
            if "caps_in" not in kwargs:
                raise ValueError(f"Please provide a 'caps_in' keyword argument with {decode_method=}. (found {tuple(kwargs.keys())})")
            caps_in = torch.as_tensor(kwargs["caps_in"])
            if caps_in.is_floating_point():
                raise ValueError("conette_b200 teacher forcing takes token ids (B, L); mixed token embeddings are a training-time input")
            # (B, L, V) -> (B, V, L) like forcing.py:75.  Positions at or after a row's first pad see the pad tokens as keys here,
            # the reference masks them; those positions are ignore_index targets in every loss of the reference.
            return self.engine.decoder_logits(fe, lens, caps_in).permute(0, 2, 1).contiguous().cpu()
        if decode_method == "generate":
            beam = int(kwargs.get("beam_size", self.beam_size))
            if beam > 8 or not 1 <= max_len <= 64:
                raise ValueError(f"conette_b200 supports beam_size in [1, 8] and max_pred_size in [1, 64]. (found {beam=} {max_len=})")
            return tuple(t.cpu() for t in self.engine.decode(fe, lens, bos, forbid, beam, min_len, max_len))
        if decode_method == "greedy":
            return self._greedy(fe, lens, bos, forbid, min_len, max_len)
        raise ValueError(f"Unknown argument {decode_method=}. (expected one of ('forcing', 'greedy', 'generate'))")

    def _greedy(self, fe: Tensor, lens: Tensor, bos: Tensor, forbid: Optional[Tensor], min_len: int, max_len: int) -> Tensor:
        """greedy.py:17-131: the decode runs on the GPU as a beam-1 search with the step logits tapped; the reference's output
        format (logits with the EOS / no-repeat masks applied, -inf columns with 0 at <pad> after a clip has finished, trimmed to
        the longest clip) is assembled from the tap on the host."""
        preds, _, _, _, info, logits = self.engine.decode_tap(fe, lens, bos, forbid, 1, min_len, max_len)
        b = fe.shape[0]
        preds, logits = preds.cpu(), logits.cpu()  # (B, max_len), (max_len, B, V)
        pred_size = int(info[0])
        out = torch.full((b, self.vocab_size, max_len), -math.inf)
        out[:, PAD_ID, :] = 0
        fm = None if forbid is None else forbid.to(torch.bool).cpu()
        for r in range(b):
            hist = [int(bos[r])]
            for i in range(pred_size):
                lg = logits[i, r].clone()
                if i < min_len:
                    lg[EOS_ID] = -math.inf
                if fm is not None and bool(fm.any()):
                    h = torch.tensor(hist)
                    lg[h[fm[h]]] = -math.inf
                out[r, :, i] = lg
                tok = int(preds[r, i])
                hist.append(tok)
                if tok == EOS_ID:
                    break
        return out[:, :, :pred_size].contiguous()

    def close(self) -> None:
        self.engine.close()
