"""Restatement of torchlibrosa 0.1.0 ``stft.py`` (Spectrogram / LogmelFilterBank) as configured by the reference at
``src/conette/nn/encoders/convnext.py:144-180``.  TEST INFRASTRUCTURE ONLY (see package docstring).

Algorithm:
  * STFT = two frozen ``Conv1d(1, n_fft//2+1, n_fft, stride=hop, bias=False)`` whose weights are the Hann-windowed real /
    imaginary DFT basis (cast f64 -> f32); ``center=True`` reflect-pads ``n_fft//2`` samples on both sides.
  * power spectrogram = real**2 + imag**2  -> (B, 1, T, n_fft//2+1).
  * mel = power @ melW, melW = librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax).T (Slaney scale, slaney norm).
  * log = 10*log10(clamp(mel, amin)) - 10*log10(max(amin, ref)); optional top_db clamp.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


def hann_window_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True) in float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        log_t = f >= min_log_hz
        mels[log_t] = min_log_mel + np.log(f[log_t] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    log_t = m >= min_log_mel
    freqs[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    return freqs


def librosa_mel(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) -> (n_mels, 1+n_fft//2)."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


class STFT(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", freeze_parameters=True) -> None:
        super().__init__()
        assert pad_mode in ("constant", "reflect")
        assert window == "hann", "oracle shim supports the reference's 'hann' window only"
        self.n_fft = n_fft
        self.win_length = n_fft if win_length is None else win_length
        self.hop_length = self.win_length // 4 if hop_length is None else hop_length
        self.center = center
        self.pad_mode = pad_mode

        fft_window = hann_window_periodic(self.win_length)
        lpad = (n_fft - self.win_length) // 2
        fft_window = np.pad(fft_window, (lpad, n_fft - self.win_length - lpad))
        out_channels = n_fft // 2 + 1
        # DFT matrix W[n, k] = exp(-2*pi*j*n*k/n_fft)
        n = np.arange(n_fft, dtype=np.float64)
        ang = -2.0 * np.pi * np.outer(n, n[:out_channels]) / n_fft
        basis = np.exp(1j * ang) * fft_window[:, None]

        self.conv_real = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, padding=0, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, n_fft, stride=self.hop_length, padding=0, bias=False)
        self.conv_real.weight.data = torch.tensor(np.real(basis).T, dtype=torch.float32)[:, None, :]
        self.conv_imag.weight.data = torch.tensor(np.imag(basis).T, dtype=torch.float32)[:, None, :]
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        x = input[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)
        imag = self.conv_imag(x)
        real = real[:, None, :, :].transpose(2, 3)
        imag = imag[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
                 pad_mode="reflect", power=2.0, freeze_parameters=True) -> None:
        super().__init__()
        self.power = power
        self.stft = STFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length, window=window,
                         center=center, pad_mode=pad_mode, freeze_parameters=True)

    def forward(self, input):
        real, imag = self.stft.forward(input)
        spectrogram = real ** 2 + imag ** 2
        if self.power == 2.0:
            pass
        else:
            spectrogram = spectrogram ** (self.power / 2.0)
        return spectrogram


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True, ref=1.0,
                 amin=1e-10, top_db=80.0, freeze_parameters=True) -> None:
        super().__init__()
        self.is_log = is_log
        self.ref = ref
        self.amin = amin
        self.top_db = top_db
        if fmax is None:
            fmax = sr // 2
        self.melW = nn.Parameter(torch.tensor(librosa_mel(sr, n_fft, n_mels, fmin, fmax).T.copy()))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        mel_spectrogram = torch.matmul(input, self.melW)
        if self.is_log:
            return self.power_to_db(mel_spectrogram)
        return mel_spectrogram

    def power_to_db(self, input):
        log_spec = 10.0 * torch.log10(torch.clamp(input, min=self.amin, max=np.inf))
        log_spec -= 10.0 * np.log10(np.maximum(self.amin, self.ref))
        if self.top_db is not None:
            if self.top_db < 0:
                raise ValueError("top_db must be non-negative")
            log_spec = torch.clamp(log_spec, min=log_spec.max().item() - self.top_db, max=np.inf)
        return log_spec
