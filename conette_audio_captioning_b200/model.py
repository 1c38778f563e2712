"""``CoNeTTEModel``: drop-in for the reference's public inference API, backed by the sm_100a CUDA library.

Mirrors ``CoNeTTEModel.forward / __call__`` of the reference (huggingface/model.py:185-289): same keyword arguments, same
task validation and error messages, same output dictionary (``cands, preds, lprobs, mult_cands, mult_preds, mult_lprobs,
tasks`` and, when preprocessing, ``tags_probs, tags``).  Waveform loading / mono-mix / padding and ids->text stay in
Python as in the reference; everything between the padded waveform batch and the token ids runs in libconette_b200.so.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Optional, Sequence, Union

import torch
from torch import Size, Tensor

from .config import CoNeTTEConfig
from .engine import Engine
from .preprocessor import load_resample
from .tokenizer import IdTokenizer, make_forbid_rep_mask


class CoNeTTEModel:
    def __init__(
        self,
        config: Optional[CoNeTTEConfig],
        state_dict: Dict[str, Tensor],
        tokenizer: Union[Sequence[str], Any],
        device: Union[int, str, torch.device] = 0,
        precision: str = "fast",
        enc_chunk: int = 0,
        decoder: str = "auto",
        audioset_idx_to_name: Optional[Dict[int, str]] = None,
    ) -> None:
        self.config = config or CoNeTTEConfig()
        if isinstance(tokenizer, (list, tuple)):
            tokenizer = IdTokenizer(tokenizer)
        self.tokenizer = tokenizer  # anything with decode_rec / get_vocab_size / has / token_to_id
        vocab = tokenizer.get_vocab_size()
        if state_dict["model.decoder.classifier.weight"].shape[0] != vocab:
            raise ValueError("vocabulary size does not match decoder.classifier.weight")
        if isinstance(device, str) and device in ("cuda_if_available", "auto"):  # torchoutil get_device (reference model.py:60)
            device = "cuda"
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise ValueError(f"Invalid argument device={device!r}. (conette_b200 runs on a CUDA device of compute capability 10.x "
                             "only; there is no CPU path)")
        self.engine = Engine(state_dict, vocab, dev.index or 0, precision, enc_chunk, decoder)
        self.task_id_to_token_id = state_dict["model.task_id_to_token_id"].to("cpu", torch.int64)
        fm = state_dict.get("model.forbid_rep_mask")
        self.forbid_rep_mask = None if fm is None else fm.to("cpu", torch.uint8)
        self._itos = [tokenizer.id_to_token(i) for i in range(vocab)]
        self.audioset_idx_to_name = audioset_idx_to_name or {i: f"class{i}" for i in range(527)}

    # ---- reference properties (model.py:109-115) -------------------------------------------------------------------
    @property
    def default_task(self) -> str:
        return next(iter(self.config.task_names))

    @property
    def tasks(self) -> List[str]:
        return list(self.config.task_names)

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _task_token_ids(self, dataset_lst: List[str], source_lst: List[Optional[str]]) -> Tensor:
        """reference ``batch_to_task_token_ids`` (pl_modules/conette.py:486-525)."""
        mode = self.config.task_mode
        if mode == "none":
            return torch.full((len(dataset_lst),), 1, dtype=torch.int64)
        names = list(self.config.task_names)
        if mode == "ds":
            idx = [names.index(ds) for ds in dataset_lst]
        else:
            idx = [names.index(ds if src is None else f"{ds}_{src}".lower()) for ds, src in zip(dataset_lst, source_lst)]
        return self.task_id_to_token_id[torch.tensor(idx, dtype=torch.int64)]

    def _forbid_mask(self, forbid_rep_mode: Optional[str]) -> Optional[Tensor]:
        if forbid_rep_mode is None:
            return self.forbid_rep_mask
        mask = make_forbid_rep_mask(self._itos, forbid_rep_mode)  # raises ValueError on unknown modes
        return None if mask is None else mask.to(torch.uint8)

    @staticmethod
    def _check_limits(beam: int, max_len: int, n_caps: int = 1) -> None:
        """The CUDA library's hard limits (include/conette_b200.h CNB_MAX_BEAM / CNB_MAX_PRED_SIZE); the reference accepts any
        value, so exceeding them is reported as a ValueError that names the supported range."""
        if beam > 8:
            raise ValueError(f"Invalid argument beam_size={beam}. (conette_b200 supports beam_size in [1, 8])")
        if not 1 <= max_len <= 64:
            raise ValueError(f"Invalid argument max_pred_size={max_len}. (conette_b200 supports max_pred_size in [1, 64])")
        if not 1 <= n_caps <= 8:
            raise ValueError(f"Invalid number of captions per clip {n_caps}. (conette_b200 scores 1 to 8 captions per clip)")

    # ---- forward (reference model.py:185-261) ----------------------------------------------------------------------------
    def __call__(
        self,
        x: Union[Tensor, str, Iterable[str], Iterable[Tensor]],
        sr: Union[None, int, Iterable[int]] = None,
        x_shapes: Union[Tensor, None, List[Size]] = None,
        preprocess: bool = True,
        threshold: Union[float, Tensor] = 0.3,
        task: Union[str, List[str], None] = None,
        beam_size: Optional[int] = None,
        min_pred_size: Optional[int] = None,
        max_pred_size: Optional[int] = None,
        forbid_rep_mode: Optional[str] = None,
    ) -> Dict[str, Any]:
        if preprocess:
            wav, x_lens = load_resample(x, sr, x_shapes, resampler=self.engine.resample)
            bsize = wav.shape[0]
        else:
            assert isinstance(x, Tensor) and isinstance(x_shapes, Tensor)
            bsize = len(x)

        if task is None:
            tasks = [self.default_task] * bsize
        elif isinstance(task, str):
            tasks = [task] * bsize
        elif len(task) != bsize:
            raise ValueError(f"Invalid number of tasks with input. (found {len(task)} tasks but {bsize} elements)")
        else:
            tasks = list(task)
        for t in tasks:
            if t not in self.config.task_names:
                raise ValueError(f"Invalid argument tasks={tasks}. (task {t} is not in {self.config.task_names})")
        dataset_lst, source_lst = [], []
        for t in tasks:
            parts = t.split("_")
            dataset_lst.append(parts[0])
            source_lst.append("_".join(parts[1:]) if len(parts) >= 2 else None)
        bos_ids = self._task_token_ids(dataset_lst, source_lst)

        beam = self.config.beam_size if beam_size is None else beam_size
        min_len = self.config.min_pred_size if min_pred_size is None else min_pred_size
        max_len = self.config.max_pred_size if max_pred_size is None else max_pred_size
        assert beam > 0  # reference beam.py:57-58
        assert min_len >= 0
        self._check_limits(beam, max_len)
        forbid = self._forbid_mask(forbid_rep_mode)

        if preprocess and wav.device.type == "cpu":
            preds, lprobs, mult_preds, mult_lprobs, clip_probs = self.engine.caption_host(
                wav, x_lens, bos_ids, forbid, beam, min_len, max_len, with_tags=True)
        elif preprocess:  # resampled on the GPU: the batch is already device-resident
            preds, lprobs, mult_preds, mult_lprobs, clip_probs = (
                t.cpu() for t in self.engine.caption(wav, x_lens, bos_ids, forbid, beam, min_len, max_len, with_tags=True))
        else:
            lens = x_shapes[:, 1].to(torch.int32)  # FrameIdentEncoder: lens = audio_shape[:, 1] (nn/encoders/ident.py:14-34)
            preds, lprobs, mult_preds, mult_lprobs = (
                t.cpu() for t in self.engine.decode(x, lens, bos_ids, forbid, beam, min_len, max_len))
            clip_probs = None

        outs: Dict[str, Any] = {
            "cands": self.tokenizer.decode_rec(preds),
            "preds": preds,
            "lprobs": lprobs,
            "mult_cands": self.tokenizer.decode_rec(mult_preds),
            "mult_preds": mult_preds,
            "mult_lprobs": mult_lprobs,
            "tasks": tasks,
        }
        if clip_probs is not None:
            outs["tags_probs"] = clip_probs
            hot = clip_probs >= threshold  # torchoutil probs_to_names (reference model.py:203-204)
            outs["tags"] = [[self.audioset_idx_to_name[int(i)] for i in row.nonzero().flatten().tolist()] for row in hot]
        return outs

    forward = __call__

    # ---- teacher-forced scoring (evaluation workflows) --------------------------------------------------------------------
    def score(
        self,
        x: Union[Tensor, str, Iterable[str], Iterable[Tensor]],
        mult_captions: Tensor,
        sr: Union[None, int, Iterable[int]] = None,
        x_shapes: Union[Tensor, None, List[Size]] = None,
        preprocess: bool = True,
        task: Union[str, List[str], None] = None,
    ) -> Dict[str, Any]:
        """Teacher-forced losses of given captions, the loss half of the reference's ``CoNeTTEPLM.test_step`` /
        ``validation_step`` (pl_modules/conette.py:293-318 and :236-256): ``mult_captions`` (B, n_caps, L+1) i64 token ids
        (or (B, L+1) for one caption per clip), 0-padded; position 0 is overwritten with the clip's task BOS id exactly like
        ``replace_first_ids_in_batch`` (conette.py:526-541) on a copy.  Returns ``losses`` (B, n_caps) =
        CrossEntropyLossMean(ignore_index=pad, dim=1) of logits(caps[:, :-1]) against caps[:, 1:], ``loss`` = their mean,
        and ``token_lprobs`` (B, n_caps, L)."""
        caps = torch.as_tensor(mult_captions)
        if caps.ndim == 2:
            caps = caps[:, None, :]
        if caps.ndim != 3 or caps.is_floating_point():
            raise ValueError(f"Invalid captions shape/dtype {tuple(caps.shape)} {caps.dtype}; expected integer ids (B, n_caps, L+1)")
        if preprocess:
            wav, x_lens = load_resample(x, sr, x_shapes, resampler=self.engine.resample)
            bsize = wav.shape[0]
        else:
            assert isinstance(x, Tensor) and isinstance(x_shapes, Tensor)
            bsize = len(x)
        if caps.shape[0] != bsize:
            raise ValueError(f"Invalid number of captions with input. (found {caps.shape[0]} caption sets but {bsize} elements)")
        tasks = [self.default_task] * bsize if task is None else ([task] * bsize if isinstance(task, str) else list(task))
        if len(tasks) != bsize:
            raise ValueError(f"Invalid number of tasks with input. (found {len(tasks)} tasks but {bsize} elements)")
        for t in tasks:
            if t not in self.config.task_names:
                raise ValueError(f"Invalid argument tasks={tasks}. (task {t} is not in {self.config.task_names})")
        parts = [t.split("_") for t in tasks]
        bos_ids = self._task_token_ids([p[0] for p in parts], ["_".join(p[1:]) if len(p) >= 2 else None for p in parts])
        self._check_limits(1, max(int(caps.shape[2]) - 1, 1), int(caps.shape[1]))
        caps = caps.to("cpu", torch.int64).clone()
        caps[:, :, 0] = bos_ids.to("cpu")[:, None]
        if preprocess:
            n = int(wav.shape[1])
            frame_embs, _ = self.engine.encoder(wav, with_tags=False)
            red = n // int(frame_embs.shape[1])  # reference convnext.py:312-315: float32 divide, round half to even
            x_lens = torch.full((bsize,), n, dtype=torch.int64) if x_lens is None else x_lens.to("cpu", torch.int64)
            lens = x_lens.to(torch.float32).div(red).round().to(torch.int32)
        else:
            frame_embs, lens = x, x_shapes[:, 1].to(torch.int32)
        tok_lp, losses = self.engine.score_captions(frame_embs, lens, caps)
        losses = losses.cpu()
        return {"losses": losses, "loss": losses.mean(), "token_lprobs": tok_lp.cpu(), "tasks": tasks}

    # ---- streaming (dataset captioning / serving) --------------------------------------------------------------------------
    def stream(self, batches: Iterable[Any], sr: Union[None, int, Iterable[int]] = None, task: Union[str, List[str], None] = None,
               threshold: Union[float, Tensor] = 0.3, **kwargs: Any) -> Iterable[Dict[str, Any]]:
        """Caption an iterable of batches (each one anything ``__call__`` accepts as ``x``) and yield one output dict per
        batch, identical to ``self(x, ...)``.  Two batches are in flight: while the GPU works on batch i the host loads /
        pads batch i+1, its H2D copy runs on the copy stream, and batch i-1 is detokenised -- the call a user makes to caption a
        dataset (reference predict.py:209 passes the whole file list at once)."""
        pending = None
        for x in batches:
            nxt = self._begin(x, sr, task, **kwargs)
            if pending is not None:
                yield self._finish(pending, threshold)
            pending = nxt
        if pending is not None:
            yield self._finish(pending, threshold)

    def _begin(self, x, sr, task, beam_size=None, min_pred_size=None, max_pred_size=None, forbid_rep_mode=None):
        wav, x_lens = load_resample(x, sr, None, resampler=self.engine.resample)
        bsize = wav.shape[0]
        tasks = [self.default_task] * bsize if task is None else ([task] * bsize if isinstance(task, str) else list(task))
        if len(tasks) != bsize:
            raise ValueError(f"Invalid number of tasks with input. (found {len(tasks)} tasks but {bsize} elements)")
        for t in tasks:
            if t not in self.config.task_names:
                raise ValueError(f"Invalid argument tasks={tasks}. (task {t} is not in {self.config.task_names})")
        parts = [t.split("_") for t in tasks]
        bos_ids = self._task_token_ids([p[0] for p in parts], ["_".join(p[1:]) if len(p) >= 2 else None for p in parts])
        beam = self.config.beam_size if beam_size is None else beam_size
        min_len = self.config.min_pred_size if min_pred_size is None else min_pred_size
        max_len = self.config.max_pred_size if max_pred_size is None else max_pred_size
        assert beam > 0 and min_len >= 0
        self._check_limits(beam, max_len)
        forbid = self._forbid_mask(forbid_rep_mode)
        if wav.device.type == "cpu":
            return ("host", self.engine.caption_host_begin(wav, x_lens, bos_ids, forbid, beam, min_len, max_len, with_tags=True), tasks)
        return ("dev", self.engine.caption(wav, x_lens, bos_ids, forbid, beam, min_len, max_len, with_tags=True), tasks)

    def _finish(self, pending, threshold) -> Dict[str, Any]:
        kind, ticket, tasks = pending
        if kind == "host":
            preds, lprobs, mult_preds, mult_lprobs, clip_probs = self.engine.caption_host_end(ticket)
        else:
            preds, lprobs, mult_preds, mult_lprobs, clip_probs = (t.cpu() for t in ticket)
        outs: Dict[str, Any] = {"cands": self.tokenizer.decode_rec(preds), "preds": preds, "lprobs": lprobs,
                                "mult_cands": self.tokenizer.decode_rec(mult_preds), "mult_preds": mult_preds,
                                "mult_lprobs": mult_lprobs, "tasks": tasks, "tags_probs": clip_probs}
        hot = clip_probs >= threshold
        outs["tags"] = [[self.audioset_idx_to_name[int(i)] for i in row.nonzero().flatten().tolist()] for row in hot]
        return outs
