"""Thin torch-facing wrapper over the C ABI: one ``Engine`` = one ``cnb_handle`` on one GPU.

PyTorch is used here only for device memory, streams and dtype bookkeeping; all arithmetic happens inside
libconette_b200.so.  The methods mirror the reference's operator seams:

  ``encoder``        ConvNeXt.forward                 reference nn/encoders/convnext.py:264-336
  ``decode``         projection + generate()          reference pl_modules/conette.py:452-467, nn/decoding/beam.py:22-227
  ``decoder_logits`` AACTransformerDecoder.forward    reference nn/decoders/aac_tfmer.py:71-118 (teacher-forced, parity only)
  ``caption``        CoNeTTEModel.forward minus text  reference huggingface/model.py:185-261
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib

N_TAGS = 527
EOS_ID = 2


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


DECODER_MODES = {"auto": 0, "graph": 1, "eager": 2, "cluster": 3}


class Engine:
    def __init__(self, state_dict: Dict[str, Tensor], vocab_size: int, device: int = 0, precision: str = "fast",
                 enc_chunk: int = 0, decoder: str = "auto") -> None:
        if not torch.cuda.is_available():
            raise _lib.CnbError("conette_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.vocab_size = int(vocab_size)
        self.precision = precision
        cfg = _lib.Config()
        cfg.abi_version = _lib.ABI_VERSION
        cfg.device = device
        cfg.vocab_size = self.vocab_size
        cfg.precision = {"fast": _lib.PRECISION_FAST, "parity": _lib.PRECISION_PARITY}[precision]
        cfg.enc_chunk = enc_chunk
        # decoder implementation: "auto" = "cluster" in precision "fast" when the shape allows it (beam <= 8, max_len <= 64,
        # T' <= 128, V <= 65535), else "graph".  "cluster" = one launch, a cluster of 8 CTAs decodes a group of 16/beam or
        # 32/beam clips start to finish on the tensor cores (fp16 hi/lo split operands: fp32-level accuracy); "graph" = CUDA-graph
        # replay of the fp32 CUDA-core step kernels (the decoder of precision "parity"); "eager" = the same launches without a graph.
        cfg.reserved[0] = DECODER_MODES[decoder]
        handle = C.c_void_p()
        _lib.check(self.lib.cnb_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        dtype_code = {torch.float32: _lib.DTYPE_F32, torch.int64: _lib.DTYPE_I64, torch.bool: _lib.DTYPE_BOOL,
                      torch.uint8: _lib.DTYPE_U8}
        for name, t in state_dict.items():
            if not isinstance(t, Tensor):
                continue
            if t.dtype in (torch.float16, torch.bfloat16, torch.float64):  # checkpoints stored in another float type
                t = t.to(torch.float32)
            if t.dtype not in dtype_code:
                continue
            t = t.detach().to("cpu").contiguous()
            shape = (C.c_int64 * max(t.ndim, 1))(*t.shape)
            _lib.check(self.lib.cnb_load_weight(self.handle, name.encode(), t.data_ptr(), dtype_code[t.dtype], t.ndim, shape))
        _lib.check(self.lib.cnb_finalize_weights(self.handle))
        self._banks: Dict[Tuple[int, int], Tuple[Tensor, Tensor, int]] = {}  # resampler filter banks on the device

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.cnb_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, t: Tensor, dtype: torch.dtype) -> Tensor:
        return t.to(device=self.device, dtype=dtype).contiguous()

    def launch_count(self) -> int:
        return int(self.lib.cnb_launch_count(self.handle))

    def device_bytes(self) -> int:
        return int(self.lib.cnb_device_bytes(self.handle))

    def profile_begin(self) -> None:
        _lib.check(self.lib.cnb_profile_begin(self.handle))

    def profile_end(self) -> Dict[str, Tuple[float, int]]:
        """{kernel class: (summed device ms, number of event brackets)} since profile_begin."""
        n = len(_lib.KERNEL_CLASSES)
        ms = (C.c_float * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(self.lib.cnb_profile_end(self.handle, ms, cnt, n))
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(_lib.KERNEL_CLASSES)}

    def profile_timeline_begin(self) -> None:
        _lib.check(self.lib.cnb_profile_timeline_begin(self.handle))

    def profile_timeline_end(self, cap: int = 1 << 16) -> List[Tuple[str, float, float]]:
        """[(kernel class, begin ms, end ms)] for every bracket since profile_timeline_begin, in issue order (streaming overlap
        stays on, so brackets of the decode stream interleave with the encoder's)."""
        cls = (C.c_int32 * cap)()
        t0 = (C.c_float * cap)()
        t1 = (C.c_float * cap)()
        n = C.c_int32(0)
        _lib.check(self.lib.cnb_profile_timeline_end(self.handle, cls, t0, t1, cap, C.byref(n)))
        return [(_lib.KERNEL_CLASSES[cls[i]], float(t0[i]), float(t1[i])) for i in range(min(n.value, cap))]

    # ---- stages -------------------------------------------------------------------------------------------------------
    def resample(self, wav: Tensor, orig_sr: int, lens: Optional[Tensor] = None, new_sr: int = 32_000,
                 n_out: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        """(B, N_in) f32 at ``orig_sr`` -> (B, N_out) f32 at ``new_sr`` on the GPU, plus the resampled lengths (B,) i64 (host).

        Mirrors ``torchaudio.functional.resample`` as the reference calls it (preprocessor.py:139-141); with ``lens`` (true
        lengths of right-zero-padded clips) every clip equals its own resample followed by right zero-padding."""
        from . import resample as rs

        wav = self._dev(wav, torch.float32)
        b, n_in = wav.shape
        orig, new = rs.reduced_ratio(orig_sr, new_sr)
        lens_h = torch.full((b,), n_in, dtype=torch.int64) if lens is None else lens.to("cpu", torch.int64)
        out_lens = torch.tensor([rs.resampled_length(int(n), orig, new) for n in lens_h.tolist()], dtype=torch.int64)
        if n_out is None:
            n_out = int(out_lens.max())
        key = (orig, new)
        if key not in self._banks:
            taps, lo, width = rs.filter_bank(orig, new)
            self._banks[key] = (taps.to(self.device), lo.to(self.device), width)
        taps, lo, width = self._banks[key]
        lens_d = None if lens is None else lens_h.to(self.device)
        out = torch.empty(b, n_out, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_resample(self.handle, wav.data_ptr(), _ptr(lens_d), b, n_in, taps.data_ptr(), lo.data_ptr(),
                                         orig, new, taps.shape[1], width, out.data_ptr(), n_out, self._stream()))
        return out, out_lens

    def frontend(self, wav: Tensor, apply_bn: bool = True) -> Tensor:
        """(B, N) f32 -> (B, T, 224) log-mel [dB], optionally through bn0."""
        wav = self._dev(wav, torch.float32)
        b, n = wav.shape
        t, _, _ = _lib.geometry(n)
        out = torch.empty(b, t, 224, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_frontend(self.handle, wav.data_ptr(), b, n, int(apply_bn), out.data_ptr(), self._stream()))
        return out

    def encoder(self, wav: Tensor, with_tags: bool = True) -> Tuple[Tensor, Optional[Tensor]]:
        """(B, N) f32 -> frame_embs (B, T', 768), clip_probs (B, 527)."""
        wav = self._dev(wav, torch.float32)
        b, n = wav.shape
        _, _, tp = _lib.geometry(n)
        fe = torch.empty(b, tp, 768, device=self.device, dtype=torch.float32)
        clip = torch.empty(b, N_TAGS, device=self.device, dtype=torch.float32) if with_tags else None
        _lib.check(self.lib.cnb_encoder(self.handle, wav.data_ptr(), b, n, fe.data_ptr(), _ptr(clip), self._stream()))
        return fe, clip

    def encoder_tap(self, wav: Tensor, kind: int, stage: int = 0, block: int = 0) -> Tensor:
        """Intermediate activation (fp32, NHWC) for stage-isolated parity tests."""
        wav = self._dev(wav, torch.float32)
        b, n = wav.shape
        t, hs, _ = _lib.geometry(n)
        dims, widths = (96, 192, 384, 768), (56, 28, 14, 7)
        if kind == _lib.TAP_LOGMEL_BN:
            shape = (b, t, 224)
        elif kind == _lib.TAP_STEM:
            shape = (b, hs[0], 56, 96)
        else:
            shape = (b, hs[stage], widths[stage], dims[stage])
        out = torch.empty(shape, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_encoder_tap(self.handle, wav.data_ptr(), b, n, kind, stage, block, out.data_ptr(),
                                            out.numel(), self._stream()))
        return out

    def debug_gemm(self, a: Tensor, w: Tensor, bias: Tensor, scale: Optional[Tensor] = None, resid: Optional[Tensor] = None,
                   epi: int = 0, use_tc: bool = True, out_bf16: bool = False) -> Tensor:
        """Test hook: out = epi(a @ w.T) through the tcgen05 (fp16 operands) or CUDA-core (fp32) GEMM kernel."""
        a, w, bias = self._dev(a, torch.float32), self._dev(w, torch.float32), self._dev(bias, torch.float32)
        scale = None if scale is None else self._dev(scale, torch.float32)
        resid = None if resid is None else self._dev(resid, torch.float32)
        m, k = a.shape
        n = w.shape[0]
        out = torch.empty(m, n, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_debug_gemm(self.handle, a.data_ptr(), w.data_ptr(), bias.data_ptr(), _ptr(scale), _ptr(resid),
                                           m, n, k, epi, int(use_tc), int(out_bf16), out.data_ptr(), self._stream()))
        return out

    def debug_mlp_fused(self, y: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, scale: Tensor, x: Tensor) -> Tensor:
        """Test hook: x + scale * (W2 . GELU(W1 . y + b1) + b2) through the fused stage-1 MLP kernel (C = 96); returns the new x."""
        y, w1, b1, w2, b2, scale = (self._dev(v, torch.float32) for v in (y, w1, b1, w2, b2, scale))
        out = self._dev(x, torch.float32).clone()
        assert y.shape[1] == 96 and w1.shape == (384, 96) and w2.shape == (96, 384) and out.shape == y.shape
        _lib.check(self.lib.cnb_debug_mlp_fused(self.handle, y.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                                b2.data_ptr(), scale.data_ptr(), out.data_ptr(), y.shape[0], self._stream()))
        return out

    def debug_mlp_fused_pair(self, y: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, scale: Tensor, x: Tensor) -> Tensor:
        """Test hook: the same through the fused stage-2 / stage-3 MLP kernel (C = 192 / 384, hidden 4C); returns the new x."""
        y, w1, b1, w2, b2, scale = (self._dev(v, torch.float32) for v in (y, w1, b1, w2, b2, scale))
        out = self._dev(x, torch.float32).clone()
        c = y.shape[1]
        assert c in (192, 384) and w1.shape == (4 * c, c) and w2.shape == (c, 4 * c) and out.shape == y.shape
        _lib.check(self.lib.cnb_debug_mlp_fused_pair(self.handle, c, y.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                                     b2.data_ptr(), scale.data_ptr(), out.data_ptr(), y.shape[0], self._stream()))
        return out

    def _alloc_outputs(self, b: int, beam: int, max_len: int):
        dev = self.device
        return (torch.empty(b, max_len, device=dev, dtype=torch.int64), torch.empty(b, device=dev, dtype=torch.float32),
                torch.empty(b, beam, max_len, device=dev, dtype=torch.int64),
                torch.empty(b, beam, device=dev, dtype=torch.float32), torch.empty(2 + b, device=dev, dtype=torch.int32))

    @staticmethod
    def _trim(preds: Tensor, lprobs: Tensor, mult_preds: Tensor, mult_lprobs: Tensor, info: Tensor):
        """Output trimming of reference beam.py:205-225 (one device->host read of the tiny info vector)."""
        info_h = info.tolist()
        pred_size = info_h[0]
        best_len = max(info_h[2:]) + 1
        return (preds[:, : min(best_len, pred_size)].contiguous(), lprobs, mult_preds[:, :, :pred_size].contiguous(),
                mult_lprobs)

    def decode(self, frame_embs: Tensor, lens: Tensor, bos_ids: Tensor, forbid_mask: Optional[Tensor], beam: int = 3,
               min_len: int = 3, max_len: int = 20, trim: bool = True):
        """frame_embs (B, T', 768) -> (preds, lprobs, mult_preds, mult_lprobs) like reference ``generate``
        (``trim=False``: the untrimmed (B, max_len) / (B, beam, max_len) buffers + the info vector)."""
        fe = self._dev(frame_embs, torch.float32)
        b, tp, _ = fe.shape
        lens = self._dev(lens, torch.int32)
        bos = self._dev(bos_ids, torch.int64)
        forbid = None if forbid_mask is None else self._dev(forbid_mask, torch.uint8)
        outs = self._alloc_outputs(b, beam, max_len)
        _lib.check(self.lib.cnb_decode(self.handle, fe.data_ptr(), lens.data_ptr(), bos.data_ptr(), _ptr(forbid), b, tp, beam,
                                       min_len, max_len, *[o.data_ptr() for o in outs], self._stream()))
        return self._trim(*outs) if trim else tuple(outs)

    def decode_tap(self, frame_embs: Tensor, lens: Tensor, bos_ids: Tensor, forbid_mask: Optional[Tensor], beam: int = 3,
                   min_len: int = 3, max_len: int = 20):
        """``decode(trim=False)`` through the cluster kernel plus its per-step raw logits (max_len, B*beam, V); steps after the
        early exit hold NaN.  Test hook for the reference seam ``AACDecoder.__call__`` (nn/decoding/common.py:9-29)."""
        fe = self._dev(frame_embs, torch.float32)
        b, tp, _ = fe.shape
        lens = self._dev(lens, torch.int32)
        bos = self._dev(bos_ids, torch.int64)
        forbid = None if forbid_mask is None else self._dev(forbid_mask, torch.uint8)
        outs = self._alloc_outputs(b, beam, max_len)
        logits = torch.full((max_len, b * beam, self.vocab_size), float("nan"), device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_decode_tap(self.handle, fe.data_ptr(), lens.data_ptr(), bos.data_ptr(), _ptr(forbid), b, tp, beam,
                                           min_len, max_len, *[o.data_ptr() for o in outs], logits.data_ptr(), self._stream()))
        return (*outs, logits)

    def decoder_logits(self, frame_embs: Tensor, lens: Tensor, tokens: Tensor) -> Tensor:
        """Teacher-forced logits (B, steps, V) for given token prefixes (B, steps)."""
        fe = self._dev(frame_embs, torch.float32)
        b, tp, _ = fe.shape
        lens = self._dev(lens, torch.int32)
        tokens = self._dev(tokens, torch.int64)
        steps = tokens.shape[1]
        out = torch.empty(b, steps, self.vocab_size, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_decoder_logits(self.handle, fe.data_ptr(), lens.data_ptr(), tokens.data_ptr(), b, tp, steps,
                                               out.data_ptr(), self._stream()))
        return out

    def score_captions(self, frame_embs: Tensor, lens: Tensor, captions: Tensor) -> Tuple[Tensor, Tensor]:
        """Teacher-forced scoring: captions (B, n_caps, L+1) i64 (position 0 = task BOS id, 0-padded) ->
        (token_lprobs (B, n_caps, L), losses (B, n_caps)); reference pl_modules/conette.py:293-318, nn/modules/ce_mean.py."""
        fe = self._dev(frame_embs, torch.float32)
        b, tp, _ = fe.shape
        lens = self._dev(lens, torch.int32)
        caps = self._dev(captions, torch.int64)
        assert caps.ndim == 3 and caps.shape[0] == b, "captions must be (B, n_caps, cap_len)"
        n_caps, cap_len = int(caps.shape[1]), int(caps.shape[2])
        tok_lp = torch.empty(b, n_caps, cap_len - 1, device=self.device, dtype=torch.float32)
        losses = torch.empty(b, n_caps, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.cnb_score_captions(self.handle, fe.data_ptr(), lens.data_ptr(), caps.data_ptr(), b, tp, n_caps,
                                               cap_len, tok_lp.data_ptr(), losses.data_ptr(), self._stream()))
        return tok_lp, losses

    def caption(self, wav: Tensor, x_lens: Optional[Tensor], bos_ids: Tensor, forbid_mask: Optional[Tensor], beam: int = 3,
                min_len: int = 3, max_len: int = 20, with_tags: bool = True, trim: bool = True):
        """Device-resident path: wav (B, N) on the GPU -> ids; returns (preds, lprobs, mult_preds, mult_lprobs, clip_probs)."""
        wav = self._dev(wav, torch.float32)
        b, n = wav.shape
        bos = self._dev(bos_ids, torch.int64)
        forbid = None if forbid_mask is None else self._dev(forbid_mask, torch.uint8)
        xl = None if x_lens is None else x_lens.to("cpu", torch.int64).contiguous()
        outs = self._alloc_outputs(b, beam, max_len)
        clip = torch.empty(b, N_TAGS, device=self.device, dtype=torch.float32) if with_tags else None
        _lib.check(self.lib.cnb_caption(self.handle, wav.data_ptr(), _ptr(xl), bos.data_ptr(), _ptr(forbid), b, n, beam,
                                        min_len, max_len, *[o.data_ptr() for o in outs], _ptr(clip), self._stream()))
        if not trim:
            return (*outs, clip)
        return (*self._trim(*outs), clip)

    def caption_host(self, wav: Tensor, x_lens: Optional[Tensor], bos_ids: Tensor, forbid_mask: Optional[Tensor],
                     beam: int = 3, min_len: int = 3, max_len: int = 20, with_tags: bool = True, out: Optional[dict] = None):
        """End-to-end path with HOST buffers: H2D of the waveforms and D2H of the ids happen inside the C call."""
        assert wav.device.type == "cpu" and wav.dtype == torch.float32 and wav.is_contiguous()
        b, n = wav.shape
        bos = bos_ids.to("cpu", torch.int64).contiguous()
        forbid = None if forbid_mask is None else forbid_mask.to("cpu", torch.uint8).contiguous()
        xl = None if x_lens is None else x_lens.to("cpu", torch.int64).contiguous()
        if out is None:
            out = self.alloc_host_outputs(b, beam, max_len, with_tags)
        clip = out.get("clip_probs")
        _lib.check(self.lib.cnb_caption_host(self.handle, wav.data_ptr(), _ptr(xl), bos.data_ptr(), _ptr(forbid), b, n, beam,
                                             min_len, max_len, out["preds"].data_ptr(), out["lprobs"].data_ptr(),
                                             out["mult_preds"].data_ptr(), out["mult_lprobs"].data_ptr(),
                                             out["info"].data_ptr(), _ptr(clip)))
        preds, lprobs, mult_preds, mult_lprobs = self._trim(out["preds"], out["lprobs"], out["mult_preds"],
                                                            out["mult_lprobs"], out["info"])
        return preds, lprobs, mult_preds, mult_lprobs, clip

    def caption_host_begin(self, wav: Tensor, x_lens: Optional[Tensor], bos_ids: Tensor, forbid_mask: Optional[Tensor],
                           beam: int = 3, min_len: int = 3, max_len: int = 20, with_tags: bool = True,
                           out: Optional[dict] = None) -> dict:
        """Split-phase ``caption_host``: enqueue one batch and return a ticket; up to two batches may be in flight, so the H2D
        copy of the next batch overlaps this batch's compute and (fast precision) this batch decodes while the next one is
        encoded.  ``wav`` may be a host tensor or a device-resident one (used in place).  ``caption_host_end(ticket)`` returns
        what ``caption_host`` does."""
        assert wav.dtype == torch.float32 and wav.is_contiguous()  # host (pinned for real overlap) or device-resident
        b, n = wav.shape
        keep = dict(wav=wav, bos=bos_ids.to("cpu", torch.int64).contiguous(),
                    forbid=None if forbid_mask is None else forbid_mask.to("cpu", torch.uint8).contiguous(),
                    xl=None if x_lens is None else x_lens.to("cpu", torch.int64).contiguous(),
                    out=out if out is not None else self.alloc_host_outputs(b, beam, max_len, with_tags))
        o = keep["out"]
        ticket = C.c_int32(-1)
        _lib.check(self.lib.cnb_caption_host_begin(
            self.handle, wav.data_ptr(), _ptr(keep["xl"]), keep["bos"].data_ptr(), _ptr(keep["forbid"]), b, n, beam, min_len,
            max_len, o["preds"].data_ptr(), o["lprobs"].data_ptr(), o["mult_preds"].data_ptr(), o["mult_lprobs"].data_ptr(),
            o["info"].data_ptr(), _ptr(o.get("clip_probs")), C.byref(ticket)))
        keep["ticket"] = ticket.value  # `keep` holds the host tensors alive until the batch has been collected
        return keep

    def caption_host_end(self, ticket: dict):
        _lib.check(self.lib.cnb_caption_host_end(self.handle, ticket["ticket"]))
        o = ticket["out"]
        preds, lprobs, mult_preds, mult_lprobs = self._trim(o["preds"], o["lprobs"], o["mult_preds"], o["mult_lprobs"], o["info"])
        return preds, lprobs, mult_preds, mult_lprobs, o.get("clip_probs")

    @staticmethod
    def alloc_host_outputs(b: int, beam: int, max_len: int, with_tags: bool = True) -> dict:
        pin = dict(pin_memory=True)
        out = {
            "preds": torch.empty(b, max_len, dtype=torch.int64, **pin),
            "lprobs": torch.empty(b, dtype=torch.float32, **pin),
            "mult_preds": torch.empty(b, beam, max_len, dtype=torch.int64, **pin),
            "mult_lprobs": torch.empty(b, beam, dtype=torch.float32, **pin),
            "info": torch.empty(2 + b, dtype=torch.int32, **pin),
        }
        if with_tags:
            out["clip_probs"] = torch.empty(b, N_TAGS, dtype=torch.float32, **pin)
        return out
