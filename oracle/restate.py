"""Plain-PyTorch fp32 CPU restatement of the CoNeTTE inference hot path.  TEST INFRASTRUCTURE ONLY.

Self-contained (needs only torch + a state dict with the reference's tensor names), so it can run on the GPU box where
``/root/reference`` does not exist.  Every function cites the reference file:line it follows.  It is pinned against the
real reference in ``tests/test_oracle_vs_reference.py`` (dev container) and against the committed fixtures under
``tests/golden`` (everywhere).  The decoder is restated in the KV-cached form and the beam search in the fixed-slot
form that the CUDA path uses, which the tests show to be output-identical to the
reference's full-recompute decoder and row-compacting ``generate()``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]
ENC = "preprocessor.encoder."
DEC = "model.decoder."
DIMS = (96, 192, 384, 768)
DEPTHS = (3, 3, 9, 3)
PAD_ID, BOS_ID, EOS_ID, UNK_ID = 0, 1, 2, 3  # reference tokenization/constants.py:15


# ---------------------------------------------------------------------------------------------------------------------
# input side: resample
# ---------------------------------------------------------------------------------------------------------------------
def resample(wav: Tensor, orig_sr: int, new_sr: int = 32_000) -> Tensor:
    """Restatement of ``torchaudio.functional.resample`` (torchaudio==0.13.1 pinned in the reference's requirements.txt:11,
    NOT under /root/reference; call site huggingface/preprocessor.py:139-141; defaults sinc_interp_hann,
    lowpass_filter_width=6, rolloff=0.99).  Published algorithm: one windowed-sinc FIR per output phase (``new`` phases after
    dividing both rates by their gcd), applied as a strided conv1d over the input zero-padded by (width, width + orig), output
    cut to ceil(new * len / orig).  Pinned against the torchaudio installed in the dev container by
    tests/test_resample.py (bit-identical filter bank, outputs to fp32 summation order).  wav: (..., N) f32."""
    g = math.gcd(int(orig_sr), int(new_sr))
    orig, new = int(orig_sr) // g, int(new_sr) // g
    if orig == new:
        return wav
    lpw, rolloff = 6, 0.99
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lpw * orig / base_freq)
    # functional.resample passes dtype=waveform.dtype, so the whole bank is evaluated in float32
    idx = torch.arange(-width, width + orig, dtype=torch.float32)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float32)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lpw, lpw)
    window = torch.cos(t * math.pi / lpw / 2) ** 2
    t *= math.pi
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * (base_freq / orig)
    shape = wav.shape
    x = wav.reshape(-1, shape[-1]).float()
    n = x.shape[-1]
    x = F.pad(x, (width, width + orig))
    y = F.conv1d(x[:, None], kernels, stride=orig).transpose(1, 2).reshape(x.shape[0], -1)
    y = y[:, : (new * n + orig - 1) // orig]
    return y.reshape(shape[:-1] + y.shape[-1:])


# ---------------------------------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------------------------------
def n_stft_frames(n_samples: int) -> int:
    return n_samples // 320 + 1  # center=True, hop 320


def stage_heights(n_samples: int) -> List[int]:
    t = n_stft_frames(n_samples)
    h1 = (t + 8 - 4) // 4 + 1  # Conv2d k=4, s=4, pad=(4,0): reference convnext.py:405-408
    return [h1, h1 // 2, h1 // 4, h1 // 8]


def n_out_frames(n_samples: int) -> int:
    return stage_heights(n_samples)[3]


def frame_lens(x_lens: Tensor, n_samples_padded: int) -> Tensor:
    """reference convnext.py:312-315: ``input_lens.div(N // T').round().int()`` (torch.round = half to even)."""
    red = n_samples_padded // n_out_frames(n_samples_padded)
    return x_lens.div(red).round().int()


# ---------------------------------------------------------------------------------------------------------------------
# front-end (torchlibrosa 0.1.0 Spectrogram + LogmelFilterBank as configured at convnext.py:144-180; Appendix A)
# ---------------------------------------------------------------------------------------------------------------------
def logmel(sd: SD, wav: Tensor) -> Tensor:
    """(B, N) -> (B, T, 224) log-mel in dB before BatchNorm (reference convnext.py:276-278)."""
    x = F.pad(wav[:, None, :], (512, 512), mode="reflect")
    real = F.conv1d(x, sd[ENC + "spectrogram_extractor.stft.conv_real.weight"], stride=320)
    imag = F.conv1d(x, sd[ENC + "spectrogram_extractor.stft.conv_imag.weight"], stride=320)
    power = (real**2 + imag**2).transpose(1, 2)  # (B, T, 513)
    mel = power @ sd[ENC + "logmel_extractor.melW"]
    out = 10.0 * torch.log10(torch.clamp(mel, min=1e-10))
    out = out - 10.0 * math.log10(max(1e-10, 1.0))
    return out


def bn0(sd: SD, lm: Tensor) -> Tensor:
    """Eval-mode BatchNorm2d over mel bins (reference convnext.py:290-292), eps 1e-5."""
    mean = sd[ENC + "bn0.running_mean"]
    var = sd[ENC + "bn0.running_var"]
    return (lm - mean) / torch.sqrt(var + 1e-5) * sd[ENC + "bn0.weight"] + sd[ENC + "bn0.bias"]


def frontend(sd: SD, wav: Tensor) -> Tensor:
    return bn0(sd, logmel(sd, wav))


# ---------------------------------------------------------------------------------------------------------------------
# ConvNeXt-Tiny encoder (reference nn/encoders/convnext.py)
# ---------------------------------------------------------------------------------------------------------------------
def ln_cf(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """channels_first LayerNorm with biased variance (reference nn/modules/norm.py:35-40)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def stem(sd: SD, lm_bn: Tensor) -> Tensor:
    """(B, T, 224) -> (B, 96, H1, 56): Conv2d(1,96,4,4,pad=(4,0)) + LN cf (reference convnext.py:405-408,207-210)."""
    x = F.conv2d(lm_bn[:, None], sd[ENC + "downsample_layers.0.0.weight"], sd[ENC + "downsample_layers.0.0.bias"],
                 stride=(4, 4), padding=(4, 0))
    return ln_cf(x, sd[ENC + "downsample_layers.0.1.weight"], sd[ENC + "downsample_layers.0.1.bias"])


def block(sd: SD, x: Tensor, s: int, b: int, taps: Optional[dict] = None) -> Tensor:
    """ConvNeXtBlock.forward (reference convnext.py:61-74)."""
    p = ENC + f"stages.{s}.{b}."
    c = x.shape[1]
    y = F.conv2d(x, sd[p + "dwconv.weight"], sd[p + "dwconv.bias"], padding=3, groups=c)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    if taps is not None:
        taps[f"dwln.{s}.{b}"] = y
    y = F.linear(y, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "pwconv2.weight"], sd[p + "pwconv2.bias"])
    y = sd[p + "scale_layer"] * y
    return x + y.permute(0, 3, 1, 2)


def downsample(sd: SD, x: Tensor, i: int) -> Tensor:
    """LN cf + Conv2d(k=2, s=2) (reference convnext.py:212-217)."""
    x = ln_cf(x, sd[ENC + f"downsample_layers.{i}.0.weight"], sd[ENC + f"downsample_layers.{i}.0.bias"])
    return F.conv2d(x, sd[ENC + f"downsample_layers.{i}.1.weight"], sd[ENC + f"downsample_layers.{i}.1.bias"], stride=2)


def encoder(sd: SD, wav: Tensor, x_lens: Optional[Tensor] = None, taps: Optional[dict] = None) -> Dict[str, Tensor]:
    """ConvNeXt.forward (reference convnext.py:264-336) -> frame_embs (B,768,T'), frame_embs_lens (B,), clipwise_output."""
    b, n = wav.shape
    lm = logmel(sd, wav)
    x = bn0(sd, lm)
    if taps is not None:
        taps["logmel"] = lm
        taps["logmel_bn"] = x
    x = stem(sd, x)
    if taps is not None:
        taps["stem"] = x
    for s in range(4):
        if s > 0:
            x = downsample(sd, x, s)
            if taps is not None:
                taps[f"down.{s}"] = x
        for blk in range(DEPTHS[s]):
            x = block(sd, x, s, blk, taps)
            if taps is not None:
                taps[f"block.{s}.{blk}"] = x
    x = torch.mean(x, dim=3)
    frame_embs = x
    if x_lens is None:
        x_lens = torch.full((b,), n, dtype=torch.long)
    red = n // frame_embs.shape[-1]
    lens = x_lens.div(red).round().int()
    x1, _ = torch.max(x, dim=2)
    x2 = torch.mean(x, dim=2)
    h = F.layer_norm(x1 + x2, (768,), sd[ENC + "norm.weight"], sd[ENC + "norm.bias"], 1e-6)
    clip = torch.sigmoid(F.linear(h, sd[ENC + "head_audioset.weight"], sd[ENC + "head_audioset.bias"]))
    return {"frame_embs": frame_embs, "frame_embs_lens": lens, "clipwise_output": clip}


# ---------------------------------------------------------------------------------------------------------------------
# projection + decoder (reference pl_modules/common.py:59-78, conette.py:452-467, nn/decoders/aac_tfmer.py:71-118)
# ---------------------------------------------------------------------------------------------------------------------
def project(sd: SD, frame_embs_btc: Tensor) -> Tensor:
    """(B, T', 768) -> (B, T', 256): Linear + ReLU; dropouts are identity in eval (reference common.py:71-78)."""
    return F.relu(F.linear(frame_embs_btc, sd["model.projection.2.weight"], sd["model.projection.2.bias"]))


def _ln(x: Tensor, sd: SD, name: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


class KVDecoder:
    """KV-cached single-position decoder step, equal to the reference's full recompute
    (``AACTransformerDecoder.forward`` + torch ``nn.TransformerDecoderLayer`` post-norm, batch_first=False).

    ``mem``: (B, T', 256) projected frames per CLIP, ``lens``: (B,) valid frames; rows = B * beam, clip-major.
    """

    def __init__(self, sd: SD, mem: Tensor, lens: Tensor, beam: int, max_len: int, n_layers: int = 6, n_head: int = 8):
        self.sd, self.beam, self.n_layers, self.h = sd, beam, n_layers, n_head
        b, tp, d = mem.shape
        self.d, self.dh = d, d // n_head
        self.rows = b * beam
        self.ck, self.cv = [], []
        for layer in range(n_layers):
            w = sd[DEC + f"layers.{layer}.multihead_attn.in_proj_weight"]
            bia = sd[DEC + f"layers.{layer}.multihead_attn.in_proj_bias"]
            self.ck.append(F.linear(mem, w[d : 2 * d], bia[d : 2 * d]))  # (B, T', 256), once per clip
            self.cv.append(F.linear(mem, w[2 * d :], bia[2 * d :]))
        self.cross_mask = torch.arange(tp)[None, :] >= lens[:, None].long()  # True = masked (conette.py:460-462)
        self.sk = torch.zeros(n_layers, self.rows, max_len, d)
        self.sv = torch.zeros(n_layers, self.rows, max_len, d)

    def reorder(self, src_rows: Tensor) -> None:
        """new row r inherits the self-attention cache of old row src_rows[r] (beam.py:167)."""
        self.sk = self.sk[:, src_rows]
        self.sv = self.sv[:, src_rows]

    def step(self, tokens: Tensor, pos: int) -> Tensor:
        """tokens (R,) int64 at position ``pos`` -> logits (R, V)."""
        sd, d, h, dh = self.sd, self.d, self.h, self.dh
        r = tokens.shape[0]
        x = sd[DEC + "emb_layer.weight"][tokens] * math.sqrt(d) + sd[DEC + "pos_encoding.pos_embedding"][pos, 0]
        clip_of_row = torch.arange(r) // self.beam
        for layer in range(self.n_layers):
            p = DEC + f"layers.{layer}."
            qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
            q, k, v = qkv.split(d, dim=-1)
            self.sk[layer, :, pos] = k
            self.sv[layer, :, pos] = v
            kk = self.sk[layer, :, : pos + 1].reshape(r, pos + 1, h, dh)
            vv = self.sv[layer, :, : pos + 1].reshape(r, pos + 1, h, dh)
            s = torch.einsum("rhd,rphd->rhp", q.reshape(r, h, dh), kk) / math.sqrt(dh)
            a = torch.einsum("rhp,rphd->rhd", torch.softmax(s, dim=-1), vv).reshape(r, d)
            a = F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
            x = _ln(x + a, sd, p + "norm1")
            w = sd[p + "multihead_attn.in_proj_weight"]
            bia = sd[p + "multihead_attn.in_proj_bias"]
            q = F.linear(x, w[:d], bia[:d]).reshape(r, h, dh)
            ck = self.ck[layer][clip_of_row].reshape(r, -1, h, dh)
            cv = self.cv[layer][clip_of_row].reshape(r, -1, h, dh)
            s = torch.einsum("rhd,rthd->rht", q, ck) / math.sqrt(dh)
            s = s.masked_fill(self.cross_mask[clip_of_row][:, None, :], float("-inf"))
            a = torch.einsum("rht,rthd->rhd", torch.softmax(s, dim=-1), cv).reshape(r, d)
            a = F.linear(a, sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"])
            x = _ln(x + a, sd, p + "norm2")
            ff = F.linear(F.gelu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                          sd[p + "linear2.weight"], sd[p + "linear2.bias"])
            x = _ln(x + ff, sd, p + "norm3")
        return F.linear(x, sd[DEC + "classifier.weight"], sd[DEC + "classifier.bias"])  # no final norm (aac_tfmer.py:58)


# ---------------------------------------------------------------------------------------------------------------------
# beam search, fixed-slot formulation (reference nn/decoding/beam.py:22-269)
# ---------------------------------------------------------------------------------------------------------------------
def beam_search(
    sd: SD,
    mem: Tensor,
    lens: Tensor,
    bos_ids: Tensor,
    beam: int = 3,
    min_len: int = 3,
    max_len: int = 20,
    forbid_mask: Optional[Tensor] = None,
    trace: Optional[list] = None,
) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Same 4-tuple as reference ``generate`` (beam.py:227).

    Physical row = clip * beam + label.  Labels stick to row *positions* (beam.py:179-187): the r-th best candidate of a
    clip goes to the clip's r-th live label, labels of finished rows leave the live set, the rest keep their order.
    """
    b = mem.shape[0]
    vocab = sd[DEC + "classifier.weight"].shape[0]
    rows = b * beam
    dec = KVDecoder(sd, mem, lens, beam, max_len)
    toks = torch.full((rows, max_len + 1), PAD_ID, dtype=torch.long)
    toks[:, 0] = bos_ids.repeat_interleave(beam)
    sum_lp = torch.zeros(rows)
    live = torch.ones(rows, dtype=torch.bool)
    out_preds = torch.full((rows, max_len), PAD_ID, dtype=torch.long)
    out_lp = torch.zeros(rows)
    use_forbid = forbid_mask is not None and bool(forbid_mask.any())
    pred_size = max_len
    for i in range(max_len):
        logits = dec.step(toks[:, i], i)  # dead rows are computed and ignored
        if trace is not None:
            trace.append({"logits": logits.clone(), "live": live.clone(), "toks": toks[:, : i + 1].clone()})
        if i < min_len:
            logits[:, EOS_ID] = -math.inf  # beam.py:129-130
        if use_forbid:  # beam.py:146-156
            hot = torch.zeros(rows, vocab, dtype=torch.bool)
            hot.scatter_(1, toks[:, : i + 1], True)
            logits = logits.masked_fill(hot & forbid_mask[None, :], -math.inf)
        new_toks = toks.clone()
        new_sum = sum_lp.clone()
        src_rows = torch.arange(rows)
        for j in range(b):
            labels = [l for l in range(beam) if live[j * beam + l]]
            if not labels:
                continue
            rws = torch.tensor([j * beam + l for l in labels])
            if i == 0:  # beam.py:243-246: only the first row, k = beam distinct first tokens
                cand = torch.log_softmax(logits[rws[:1]], dim=1)
            else:
                cand = sum_lp[rws][:, None] + torch.log_softmax(logits[rws], dim=1)
            top, idx = torch.topk(cand.reshape(-1), len(labels))  # beam.py:256-257
            if trace is not None:  # smallest score gap that decides this selection (rank order and the k / k+1 cut)
                top1 = torch.topk(cand.reshape(-1), min(len(labels) + 1, cand.numel()))[0]
                gaps = (top1[:-1] - top1[1:])
                trace[-1].setdefault("margin", {})[j] = float(gaps.min()) if gaps.numel() else float("inf")
            prev = idx // vocab
            word = idx % vocab
            for r, l in enumerate(labels):
                row = j * beam + l
                src = int(rws[int(prev[r])])
                new_toks[row, : i + 1] = toks[src, : i + 1]
                new_toks[row, i + 1] = word[r]
                new_sum[row] = top[r]
                src_rows[row] = src
                if int(word[r]) == EOS_ID or i == max_len - 1:  # beam.py:173-176
                    out_preds[row, : i + 1] = new_toks[row, 1 : i + 2]
                    out_lp[row] = top[r] / (i + 1)
                    live[row] = False
        toks, sum_lp = new_toks, new_sum
        dec.reorder(src_rows)
        if not live.any():
            pred_size = i + 1  # beam.py:192-194
            break
    g_preds = out_preds.reshape(b, beam, max_len)[:, :, :pred_size].contiguous()
    g_lp = out_lp.reshape(b, beam)
    best_lp, best = g_lp.max(dim=1)
    best_preds = g_preds[torch.arange(b), best]
    has_eos = best_preds == EOS_ID
    first = torch.where(has_eos.any(1), has_eos.long().argmax(1), torch.full((b,), best_preds.shape[1]))
    best_preds = best_preds[:, : int(first.max()) + 1].contiguous()  # beam.py:223-225
    return best_preds, best_lp, g_preds, g_lp


def score_captions(sd: SD, frame_embs_btc: Tensor, lens: Tensor, captions: Tensor) -> Tuple[Tensor, Tensor]:
    """Teacher-forced scoring (reference pl_modules/conette.py:293-318: ``decode_audio(enc, "forcing", caps_in=caps[:, :-1])``
    through nn/decoding/forcing.py:12-76, then ``CrossEntropyLossMean(ignore_index=pad_id, dim=1)`` nn/loss/ce_mean.py:10-40
    against ``caps[:, 1:]``).  captions (B, n_caps, L+1) i64, position 0 = task BOS id, 0-padded.
    Returns token_lprobs (B, n_caps, L) (0 at pad targets) and losses (B, n_caps).  Trailing pads never influence a non-pad
    position under the causal mask, so the key-padding mask of forcing.py:49 needs no restatement."""
    b, n_caps, cap_len = captions.shape
    steps = cap_len - 1
    rows = captions.reshape(b * n_caps, cap_len)
    dec = KVDecoder(sd, project(sd, frame_embs_btc), lens, beam=n_caps, max_len=steps)
    tok_lp = torch.zeros(b * n_caps, steps)
    for i in range(steps):
        lp = torch.log_softmax(dec.step(rows[:, i], i), dim=-1)
        tgt = rows[:, i + 1]
        tok_lp[:, i] = torch.where(tgt != 0, lp.gather(1, tgt[:, None])[:, 0], torch.zeros(()))
    non_pad = rows[:, 1:] != 0
    losses = -(tok_lp * non_pad).sum(1) / non_pad.sum(1).clamp(min=1)
    return tok_lp.reshape(b, n_caps, steps), losses.reshape(b, n_caps)


def caption(sd: SD, wav: Tensor, x_lens: Optional[Tensor], bos_ids: Tensor, beam: int = 3, min_len: int = 3,
            max_len: int = 20, forbid_mask: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """waveform (B, N) -> dict like reference ``CoNeTTEPLM.forward`` minus the detokenised strings."""
    enc = encoder(sd, wav, x_lens)
    mem = project(sd, enc["frame_embs"].transpose(1, 2))
    preds, lprobs, mult_preds, mult_lprobs = beam_search(sd, mem, enc["frame_embs_lens"], bos_ids, beam, min_len,
                                                         max_len, forbid_mask)
    return {"preds": preds, "lprobs": lprobs, "mult_preds": mult_preds, "mult_lprobs": mult_lprobs,
            "frame_embs": enc["frame_embs"], "frame_embs_lens": enc["frame_embs_lens"],
            "clip_probs": enc["clipwise_output"]}
