"""Audio file loading for the caller side (reference: ``torchaudio.load`` in huggingface/preprocessor.py:90-93).

``torchaudio.load`` is used when its decoding backend is importable; plain PCM / IEEE-float WAV files are otherwise read
with the standard library, with torchaudio's normalisation (integer PCM scaled by 2^-(bits-1) to [-1, 1))."""
from __future__ import annotations

import struct
import wave
from typing import Tuple

import torch
from torch import Tensor


def _load_wav_stdlib(path: str) -> Tuple[Tensor, int]:
    try:
        with wave.open(path, "rb") as f:
            n_ch, width, sr, n_frames = f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()
            raw = f.readframes(n_frames)
        is_float = False
    except wave.Error:
        raw, n_ch, width, sr, is_float = _read_riff(path)  # WAVE_FORMAT_IEEE_FLOAT / EXTENSIBLE are not handled by `wave`
    buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    if is_float and width == 4:
        x = buf.view(torch.float32).clone()
    elif is_float and width == 8:
        x = buf.view(torch.float64).to(torch.float32)
    elif width == 1:
        x = (buf.to(torch.float32) - 128.0) / 128.0
    elif width == 2:
        x = buf.view(torch.int16).to(torch.float32) / 32768.0
    elif width == 3:
        b = buf.view(-1, 3).to(torch.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = torch.where(v >= 1 << 23, v - (1 << 24), v)
        x = v.to(torch.float32) / float(1 << 23)
    elif width == 4:
        x = buf.view(torch.int32).to(torch.float32) / float(1 << 31)
    else:
        raise ValueError(f"Unsupported WAV sample width {width} in '{path}'.")
    return x.view(-1, n_ch).t().contiguous(), sr


def _read_riff(path: str):
    data = open(path, "rb").read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"'{path}' is not a RIFF/WAVE file.")
    pos, fmt, payload = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            if fmt[0] == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                fmt = (struct.unpack("<H", body[24:26])[0],) + fmt[1:]
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None or fmt[0] not in (1, 3):
        raise ValueError(f"Unsupported WAV encoding in '{path}'.")
    return payload, fmt[1], fmt[5] // 8, fmt[2], fmt[0] == 3


def load_audio(path: str) -> Tuple[Tensor, int]:
    """-> waveform (channels, time) f32 in [-1, 1], sample rate."""
    try:
        import torchaudio

        wav, sr = torchaudio.load(path)
        return wav.to(torch.float32), int(sr)
    except (ImportError, RuntimeError, OSError):
        return _load_wav_stdlib(path)
