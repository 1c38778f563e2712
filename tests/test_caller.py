"""the caller side -- checkpoint layout of the reference (extra state, gamma rename), WAV loading,
the conette() factory and the conette-predict command line (reference predict.py:181-233, __init__.py:25-49)."""
import csv
import io
import math
import pickle
import struct
import wave

import pytest
import torch

from conette_audio_captioning_b200 import checkpoint, synth
from conette_audio_captioning_b200.audio_io import load_audio


def _write_wav(path, x, sr, width=2):
    with wave.open(str(path), "wb") as f:
        f.setnchannels(x.shape[0])
        f.setsampwidth(width)
        f.setframerate(sr)
        q = (x.t().contiguous() * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
        f.writeframes(q.numpy().tobytes())


def _tone(n, sr, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / sr
    return (0.2 * torch.sin(2 * math.pi * 440.0 * t) + 0.05 * torch.randn(n, generator=g))[None]


def test_wav_reader_pcm16_and_float32(tmp_path):
    x = _tone(3200, 32000, 0).repeat(2, 1)
    x[1] *= 0.5
    _write_wav(tmp_path / "a.wav", x, 32000)
    w, sr = load_audio(str(tmp_path / "a.wav"))
    assert sr == 32000 and w.shape == (2, 3200)
    torch.testing.assert_close(w, (x * 32768).round().clamp(-32768, 32767) / 32768.0, rtol=0, atol=1e-7)
    # IEEE float WAV (format tag 3), which the stdlib `wave` module refuses
    body = x.t().contiguous().numpy().astype("<f4").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 3, 2, 44100, 44100 * 8, 8, 32)
    (tmp_path / "f.wav").write_bytes(hdr + b"data" + struct.pack("<I", len(body)) + body)
    w, sr = load_audio(str(tmp_path / "f.wav"))
    assert sr == 44100 and torch.equal(w, x)


def test_checkpoint_layout_roundtrip(tmp_path):
    sd = synth.make_state_dict(seed=7, n_words=50)
    itos = synth.make_itos(50)
    packed = checkpoint.pack_for_saving(sd, itos)
    # an older ConvNeXt checkpoint: layer scale stored as "gamma" (reference convnext.py:76-102)
    packed["preprocessor.encoder.stages.0.0.gamma"] = packed.pop("preprocessor.encoder.stages.0.0.scale_layer")
    torch.save(packed, tmp_path / "pytorch_model.bin")
    (tmp_path / "config.json").write_text('{"beam_size": 2, "max_pred_size": 12, "model_type": "conette", "task_names": '
                                          '["clotho", "audiocaps", "macs", "wavcaps_audioset_sl", "wavcaps_bbc_sound_effects", '
                                          '"wavcaps_freesound", "wavcaps_soundbible"]}')
    got, vocab, cfg = checkpoint.load_checkpoint(str(tmp_path))
    assert vocab == itos and cfg.beam_size == 2 and cfg.max_pred_size == 12
    assert "preprocessor.encoder.stages.0.0.scale_layer" in got and not any("gamma" in k for k in got)
    for k, v in sd.items():
        if isinstance(v, torch.Tensor):
            assert torch.equal(got[k], v), k
    both = dict(packed)
    both["preprocessor.encoder.stages.0.0.scale_layer"] = both["preprocessor.encoder.stages.0.0.gamma"]
    with pytest.raises(RuntimeError):
        checkpoint.rename_legacy_keys(both)
    with pytest.raises(FileNotFoundError):
        checkpoint.load_checkpoint(str(tmp_path / "missing"))


def test_extra_state_refuses_code_execution():
    class Evil:
        def __reduce__(self):
            return (print, ("pwned",))

    raw = torch.frombuffer(bytearray(pickle.dumps({"k": Evil()})), dtype=torch.uint8)
    with pytest.raises(pickle.UnpicklingError):
        checkpoint.unpack_extra_state({"_extra_state_": raw})


def test_predict_cli_arguments_mirror_the_reference():
    from conette_audio_captioning_b200.predict import get_predict_args

    a = get_predict_args(["--audio", "x.wav", "y.wav", "--task", "clotho", "audiocaps", "--csv_export", "o.csv"])
    assert list(a.audio) == ["x.wav", "y.wav"] and a.task == ["clotho", "audiocaps"] and a.model_name == "Labbeti/conette"
    assert a.seed == 1234 and a.verbose == 1 and a.device == "cuda_if_available" and a.model_path is None


@pytest.mark.gpu
def test_factory_and_predict_cli_end_to_end(tmp_path):
    """conette(path) and `conette-predict --audio ... --csv_export ...` give the captions of the directly built model; the
    44.1 kHz file goes through the GPU resampler."""
    from conette_audio_captioning_b200 import CoNeTTEModel, conette
    from conette_audio_captioning_b200.predict import main_predict

    sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
    itos = synth.make_itos(300)
    torch.save(checkpoint.pack_for_saving(sd, itos), tmp_path / "model.bin")
    _write_wav(tmp_path / "a.wav", _tone(32000, 32000, 1), 32000)
    _write_wav(tmp_path / "b.wav", _tone(44100, 44100, 2).repeat(2, 1), 44100)
    files = [str(tmp_path / "a.wav"), str(tmp_path / "b.wav")]

    direct = CoNeTTEModel(None, sd, itos, precision="parity")
    want = direct(files, task=["clotho", "audiocaps"])
    direct.engine.close()

    model = conette(str(tmp_path / "model.bin"), model_kwds=dict(precision="parity"))
    got = model(files, task=["clotho", "audiocaps"])
    model.engine.close()
    assert got["cands"] == want["cands"] and torch.equal(got["preds"], want["preds"])

    rows = main_predict(["--audio", *files, "--task", "clotho", "audiocaps", "--model_path", str(tmp_path / "model.bin"),
                         "--csv_export", str(tmp_path / "out.csv"), "--precision", "parity", "--verbose", "0"])
    assert [r["candidate"] for r in rows] == want["cands"] and [r["task"] for r in rows] == ["clotho", "audiocaps"]
    with open(tmp_path / "out.csv") as f:
        table = list(csv.DictReader(f))
    assert [r["audio"] for r in table] == ["a.wav", "b.wav"] and [r["candidate"] for r in table] == want["cands"]
