"""Synthetic (offline) stand-ins for the released CoNeTTE checkpoint: weights, vocabulary and audio.

There is no network in the build/bench environment, so the ``Labbeti/conette`` checkpoint is unavailable.  This module
produces a *state dict with exactly the reference's tensor names and shapes* (reference
``CoNeTTEModel.state_dict()``, huggingface/model.py:165-183) filled with seeded random values, a synthetic word-level
vocabulary in the order the reference tokenizer would assign ids (tokenization/tokenizers/common.py:8-19: specials
first, then first-appearance order; task tokens appended by pl_modules/conette.py:114-123), and seeded waveforms.

The same dict is loaded into the CUDA engine and (by ``oracle/ref_loader.py``) into the reference model, so both sides
compute with identical parameters.  Default PyTorch/ConvNeXt init is degenerate for parity purposes (layer scale 1e-6,
BN stats (0, 1), zero biases: SURVEY.md Appendix F.1), so every parameter is re-randomised at O(1) effect size.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

TASK_NAMES: Tuple[str, ...] = (
    "clotho",
    "audiocaps",
    "macs",
    "wavcaps_audioset_sl",
    "wavcaps_bbc_sound_effects",
    "wavcaps_freesound",
    "wavcaps_soundbible",
)  # reference huggingface/config.py:17-25
SPECIAL_TOKENS: Tuple[str, ...] = ("<pad>", "<bos>", "<eos>", "<unk>")  # reference tokenization/constants.py:15
STOPWORDS_IN_VOCAB: Tuple[str, ...] = ("a", "the", "is", "of", "and", "in", "on")

SAMPLE_RATE = 32000
N_FFT = 1024
HOP = 320
N_MELS = 224
N_BINS = N_FFT // 2 + 1
FMIN, FMAX = 50.0, 14000.0
DIMS = (96, 192, 384, 768)
DEPTHS = (3, 3, 9, 3)
D_MODEL, N_HEAD, D_FF, N_LAYERS = 256, 8, 2048, 6
N_TAGS = 527


# ---------------------------------------------------------------------------------------------------------------------
# vocabulary
# ---------------------------------------------------------------------------------------------------------------------
def make_corpus(n_words: int = 4000) -> List[str]:
    """Sentences that, once fitted by the reference tokenizer, give ids 4.. in the order returned by make_itos."""
    words = list(STOPWORDS_IN_VOCAB) + [f"w{i}" for i in range(n_words)]
    return [" ".join(words[i : i + 10]) for i in range(0, len(words), 10)]


def make_itos(n_words: int = 4000, task_names: Sequence[str] = TASK_NAMES) -> List[str]:
    words = list(STOPWORDS_IN_VOCAB) + [f"w{i}" for i in range(n_words)]
    return list(SPECIAL_TOKENS) + words + [f"<bos_{t}>" for t in task_names]


from .tokenizer import make_forbid_rep_mask  # noqa: E402,F401  (product code lives in tokenizer.py; kept here for the tests)


# ---------------------------------------------------------------------------------------------------------------------
# analytic (frozen) front-end parameters
# ---------------------------------------------------------------------------------------------------------------------
def hann_window(n: int = N_FFT) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def stft_basis() -> Tuple[Tensor, Tensor]:
    """Hann-windowed DFT basis as stored in ``spectrogram_extractor.stft.conv_{real,imag}.weight`` (513,1,1024)."""
    n = np.arange(N_FFT, dtype=np.float64)
    ang = 2.0 * np.pi * np.outer(np.arange(N_BINS, dtype=np.float64), n) / N_FFT
    w = hann_window()[None, :]
    real = torch.tensor(np.cos(ang) * w, dtype=torch.float32)[:, None, :]
    imag = torch.tensor(-np.sin(ang) * w, dtype=torch.float32)[:, None, :]
    return real, imag


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.atleast_1d(np.asarray(f, dtype=np.float64))
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    out = f / f_sp
    hi = f >= min_log_hz
    out[hi] = min_log_mel + np.log(f[hi] / min_log_hz) / logstep
    return out


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    out = f_sp * m
    hi = m >= min_log_mel
    out[hi] = min_log_hz * np.exp(logstep * (m[hi] - min_log_mel))
    return out


def mel_matrix() -> Tensor:
    """Slaney-scale, slaney-normalised triangular filter bank (513, 224) as in ``logmel_extractor.melW``."""
    fftfreqs = np.fft.rfftfreq(N_FFT, d=1.0 / SAMPLE_RATE)
    lo, hi = _hz_to_mel(np.array([FMIN, FMAX]))
    mel_f = _mel_to_hz(np.linspace(lo, hi, N_MELS + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((N_MELS, N_BINS), dtype=np.float32)
    for i in range(N_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]
    return torch.tensor(np.ascontiguousarray(w.T))


# ---------------------------------------------------------------------------------------------------------------------
# state dict
# ---------------------------------------------------------------------------------------------------------------------
def make_state_dict(seed: int = 1234, n_words: int = 4000, eos_bias: float = 0.0) -> Dict[str, Tensor]:
    """Seeded random weights under the reference's state-dict names (fp32 / int64 / bool, CPU).

    ``eos_bias`` is added to the classifier bias of ``<eos>`` (id 2) so tests can make beams finish at different steps.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def randn(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    def uniform(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    # ---- encoder -----------------------------------------------------------------------------------------------
    e = "preprocessor.encoder."
    real, imag = stft_basis()
    sd[e + "spectrogram_extractor.stft.conv_real.weight"] = real
    sd[e + "spectrogram_extractor.stft.conv_imag.weight"] = imag
    sd[e + "logmel_extractor.melW"] = mel_matrix()
    sd[e + "bn0.weight"] = randn(N_MELS, std=0.02, mean=1.0)
    sd[e + "bn0.bias"] = randn(N_MELS, std=0.02)
    sd[e + "bn0.running_mean"] = randn(N_MELS, std=5.0, mean=-20.0)
    sd[e + "bn0.running_var"] = uniform(N_MELS, lo=50.0, hi=150.0)
    sd[e + "bn0.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    # stem: Conv2d(1, 96, (4,4), stride (4,4), padding (4,0)) + LayerNorm(channels_first)
    sd[e + "downsample_layers.0.0.weight"] = randn(DIMS[0], 1, 4, 4, std=0.25)
    sd[e + "downsample_layers.0.0.bias"] = randn(DIMS[0], std=0.02)
    sd[e + "downsample_layers.0.1.weight"] = randn(DIMS[0], std=0.02, mean=1.0)
    sd[e + "downsample_layers.0.1.bias"] = randn(DIMS[0], std=0.02)
    for i in range(1, 4):
        cin, cout = DIMS[i - 1], DIMS[i]
        sd[e + f"downsample_layers.{i}.0.weight"] = randn(cin, std=0.02, mean=1.0)
        sd[e + f"downsample_layers.{i}.0.bias"] = randn(cin, std=0.02)
        sd[e + f"downsample_layers.{i}.1.weight"] = randn(cout, cin, 2, 2, std=1.0 / math.sqrt(4 * cin))
        sd[e + f"downsample_layers.{i}.1.bias"] = randn(cout, std=0.02)
    for s, (c, depth) in enumerate(zip(DIMS, DEPTHS)):
        for b in range(depth):
            p = e + f"stages.{s}.{b}."
            sd[p + "scale_layer"] = uniform(c, lo=0.2, hi=1.0)
            sd[p + "dwconv.weight"] = randn(c, 1, 7, 7, std=1.0 / 7.0)
            sd[p + "dwconv.bias"] = randn(c, std=0.02)
            sd[p + "norm.weight"] = randn(c, std=0.02, mean=1.0)
            sd[p + "norm.bias"] = randn(c, std=0.02)
            sd[p + "pwconv1.weight"] = randn(4 * c, c, std=1.0 / math.sqrt(c))
            sd[p + "pwconv1.bias"] = randn(4 * c, std=0.02)
            sd[p + "pwconv2.weight"] = randn(c, 4 * c, std=1.0 / math.sqrt(4 * c))
            sd[p + "pwconv2.bias"] = randn(c, std=0.02)
    sd[e + "norm.weight"] = randn(DIMS[-1], std=0.02, mean=1.0)
    sd[e + "norm.bias"] = randn(DIMS[-1], std=0.02)
    sd[e + "head_audioset.weight"] = randn(N_TAGS, DIMS[-1], std=0.05)
    sd[e + "head_audioset.bias"] = randn(N_TAGS, std=0.02)

    # ---- projection + decoder ------------------------------------------------------------------------------------
    m = "model."
    itos = make_itos(n_words)
    vocab = len(itos)
    n_fit = vocab - len(TASK_NAMES)
    sd[m + "task_id_to_token_id"] = torch.arange(n_fit, vocab, dtype=torch.long)
    sd[m + "forbid_rep_mask"] = torch.tensor([tok not in STOPWORDS_IN_VOCAB for tok in itos], dtype=torch.bool)
    sd[m + "projection.2.weight"] = randn(D_MODEL, DIMS[-1], std=1.0 / math.sqrt(DIMS[-1]))
    sd[m + "projection.2.bias"] = randn(D_MODEL, std=0.02)
    d = m + "decoder."
    emb = randn(vocab, D_MODEL, std=1.0)
    emb[0].zero_()  # padding_idx = pad_id = 0 (reference aac_tfmer.py:39-44)
    sd[d + "emb_layer.weight"] = emb
    sd[d + "pos_encoding.pos_embedding"] = positional_table(5000)
    for layer in range(N_LAYERS):
        p = d + f"layers.{layer}."
        for attn in ("self_attn", "multihead_attn"):
            lim = math.sqrt(6.0 / (D_MODEL + 3 * D_MODEL))  # xavier_uniform on (768, 256)
            sd[p + f"{attn}.in_proj_weight"] = uniform(3 * D_MODEL, D_MODEL, lo=-lim, hi=lim)
            sd[p + f"{attn}.in_proj_bias"] = randn(3 * D_MODEL, std=0.02)
            lim = 1.0 / math.sqrt(D_MODEL)
            sd[p + f"{attn}.out_proj.weight"] = uniform(D_MODEL, D_MODEL, lo=-lim, hi=lim)
            sd[p + f"{attn}.out_proj.bias"] = randn(D_MODEL, std=0.02)
        lim = 1.0 / math.sqrt(D_MODEL)
        sd[p + "linear1.weight"] = uniform(D_FF, D_MODEL, lo=-lim, hi=lim)
        sd[p + "linear1.bias"] = uniform(D_FF, lo=-lim, hi=lim)
        lim = 1.0 / math.sqrt(D_FF)
        sd[p + "linear2.weight"] = uniform(D_MODEL, D_FF, lo=-lim, hi=lim)
        sd[p + "linear2.bias"] = uniform(D_MODEL, lo=-lim, hi=lim)
        for n in ("norm1", "norm2", "norm3"):
            sd[p + f"{n}.weight"] = randn(D_MODEL, std=0.02, mean=1.0)
            sd[p + f"{n}.bias"] = randn(D_MODEL, std=0.02)
    lim = 1.0 / math.sqrt(D_MODEL)
    sd[d + "classifier.weight"] = uniform(vocab, D_MODEL, lo=-lim, hi=lim)
    bias = uniform(vocab, lo=-lim, hi=lim)
    bias[2] += eos_bias
    sd[d + "classifier.bias"] = bias
    return sd


def positional_table(maxlen: int = 5000, emb_size: int = D_MODEL) -> Tensor:
    """Sinusoidal table exactly as the reference builds it (nn/modules/positional_encoding.py:22-27)."""
    den = torch.exp(-torch.arange(0, emb_size, 2) * math.log(10000) / emb_size)
    pos = torch.arange(0, maxlen).reshape(maxlen, 1)
    pe = torch.zeros((maxlen, emb_size))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.unsqueeze(-2)


# ---------------------------------------------------------------------------------------------------------------------
# audio
# ---------------------------------------------------------------------------------------------------------------------
def make_audio(batch: int, n_samples: int, seed: int = 1234, tones: bool = True) -> Tensor:
    """(B, 1, N) float32: 0.1*randn (+ three per-clip sinusoids so the log-mel is not flat) -- SURVEY.md §8(d)."""
    g = torch.Generator().manual_seed(seed)
    x = 0.1 * torch.randn(batch, 1, n_samples, generator=g)
    if tones:
        t = torch.arange(n_samples, dtype=torch.float32) / SAMPLE_RATE
        freqs = 100.0 + 6000.0 * torch.rand(batch, 3, generator=g)
        amps = 0.05 + 0.2 * torch.rand(batch, 3, generator=g)
        for k in range(3):
            x[:, 0, :] += amps[:, k : k + 1] * torch.sin(2 * math.pi * freqs[:, k : k + 1] * t[None, :])
    return x
