"""the GPU polyphase resampler that replaces torchaudio.functional.resample on the input side
(reference huggingface/preprocessor.py:139-141).  CPU tests pin the oracle restatement and the product's filter bank to the
torchaudio installed here; GPU tests compare cnb_resample (through the C ABI) with the oracle."""
import math

import pytest
import torch

from oracle import restate

RATES = [44100, 48000, 16000, 22050, 8000, 24000, 96000]


def _wave(b, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 1000.0
    return 0.1 * torch.randn(b, n, generator=g) + 0.3 * torch.sin(2 * math.pi * 3.7 * t)[None]


@pytest.mark.parametrize("sr", RATES)
def test_oracle_resample_matches_torchaudio(sr):
    ta = pytest.importorskip("torchaudio")
    x = _wave(3, 4001, sr)
    ref = ta.functional.resample(x, sr, 32000)
    got = restate.resample(x, sr)
    assert got.shape == ref.shape
    torch.testing.assert_close(got, ref, rtol=0, atol=2e-6)  # fp32, |x| < 1: summation order of the conv only


@pytest.mark.parametrize("sr", RATES + [11025])
def test_filter_bank_is_torchaudios_bank(sr):
    ta = pytest.importorskip("torchaudio")
    from torchaudio.functional.functional import _get_sinc_resample_kernel

    from conette_audio_captioning_b200.resample import filter_bank, reduced_ratio

    orig, new = reduced_ratio(sr, 32000)
    taps, lo, width = filter_bank(orig, new)
    dense, w2 = _get_sinc_resample_kernel(sr, 32000, math.gcd(sr, 32000), dtype=torch.float32)  # as functional.resample calls it
    dense = dense[:, 0]
    assert width == w2 and taps.shape[0] == new
    rebuilt = torch.zeros_like(dense)
    for p in range(new):
        k = min(taps.shape[1], dense.shape[1] - int(lo[p]))
        rebuilt[p, int(lo[p]): int(lo[p]) + k] = taps[p, :k]
    kept = rebuilt != 0
    assert torch.equal(rebuilt[kept], dense[kept])  # bit-identical where kept: same formula, same dtype
    assert float(dense[~kept].abs().max()) <= 1e-12  # dropped: the float32 residue of the clamped window (~1e-23)
    assert taps.shape[1] <= 2 * width + 2 * math.ceil(orig / new) + 2


def test_bad_ratios_are_rejected():
    from conette_audio_captioning_b200.resample import filter_bank, reduced_ratio

    with pytest.raises(ValueError):
        reduced_ratio(44100.5, 32000)
    with pytest.raises(ValueError):
        filter_bank(*reduced_ratio(44101, 32000))


def test_host_grouping_with_a_stub_resampler():
    """load_resample: mono mix, grouping by rate, per-clip lengths and right zero-padding (host logic; the stub stands in
    for Engine.resample and is the oracle itself)."""
    from conette_audio_captioning_b200.preprocessor import load_resample

    def stub(group, sr, lens, new_sr, n_out):
        lens = torch.full((len(group),), group.shape[1]) if lens is None else lens
        rows = [restate.resample(group[i, : int(lens[i])], sr, new_sr) for i in range(len(group))]
        n_out = n_out or max(r.shape[-1] for r in rows)
        out = torch.zeros(len(rows), n_out)
        for i, r in enumerate(rows):
            out[i, : r.shape[-1]] = r
        return out, torch.tensor([r.shape[-1] for r in rows])

    clips = [_wave(2, 4410, 1), _wave(1, 3200, 2), _wave(1, 2400, 3), _wave(2, 2205, 4)]
    srs = [44100, 32000, 48000, 44100]
    wav, lens = load_resample(clips, srs, resampler=stub)
    assert lens.tolist() == [3200, 3200, 1600, 1600] and wav.shape == (4, 3200)
    for i, (c, s) in enumerate(zip(clips, srs)):
        ref = restate.resample(c, s).mean(dim=0)  # reference order: resample every channel, then the mean
        torch.testing.assert_close(wav[i, : ref.shape[-1]], ref, rtol=0, atol=2e-6)
        assert float(wav[i, ref.shape[-1]:].abs().sum()) == 0.0
    with pytest.raises(RuntimeError):
        load_resample(clips, srs)  # no host fallback
    with pytest.raises(ValueError):
        load_resample(torch.ones(1, 1, 1000), sr=16000, x_shapes=torch.tensor([[1000]]), resampler=stub)


# ----------------------------------------------------------------------------------------------------------------------
# GPU: cnb_resample through the C ABI vs the oracle
# ----------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def engine():
    from conette_audio_captioning_b200 import synth
    from conette_audio_captioning_b200.engine import Engine

    sd = synth.make_state_dict(seed=1234, n_words=300)
    eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision="parity")
    yield eng
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sr", RATES + [11025])
def test_gpu_resample_matches_oracle(engine, sr):
    x = _wave(3, 50021, sr)
    got, lens = engine.resample(x, sr)
    ref = restate.resample(x, sr)
    assert got.shape == ref.shape and lens.tolist() == [ref.shape[-1]] * 3
    torch.testing.assert_close(got.cpu(), ref, rtol=0, atol=2e-6)


@pytest.mark.gpu
def test_gpu_resample_ragged_batch_equals_per_clip_resample_then_pad(engine):
    lens = torch.tensor([44100, 1, 30001, 7])
    x = _wave(4, 44100, 9)
    for i, n in enumerate(lens.tolist()):
        x[i, n:] = 0
    got, out_lens = engine.resample(x, 44100, lens)
    assert out_lens.tolist() == [32000, 1, 21770, 6] and got.shape == (4, 32000)
    for i, n in enumerate(lens.tolist()):
        ref = restate.resample(x[i, :n], 44100)
        torch.testing.assert_close(got[i, : ref.shape[-1]].cpu(), ref, rtol=0, atol=2e-6)
        assert float(got[i, ref.shape[-1]:].abs().sum()) == 0.0  # exact zeros where the reference pads


@pytest.mark.gpu
def test_gpu_resample_linearity_and_full_size(engine):
    """Size-independent property at a BASELINE-sized batch (64 x 10 s at 44.1 kHz): resample(a x + b y) = a R(x) + b R(y)."""
    g = torch.Generator().manual_seed(3)
    x = 0.1 * torch.randn(64, 441000, generator=g)
    y = 0.1 * torch.randn(64, 441000, generator=g)
    rx, _ = engine.resample(x, 44100)
    ry, _ = engine.resample(y, 44100)
    rz, _ = engine.resample(0.5 * x - 2.0 * y, 44100)
    assert rx.shape == (64, 320000)
    torch.testing.assert_close(rz, 0.5 * rx - 2.0 * ry, rtol=0, atol=5e-6)


@pytest.mark.gpu
def test_model_resamples_on_the_gpu_like_the_reference_preprocessor():
    """CoNeTTEModel(x, sr=44100): same ids as feeding the oracle-resampled 32 kHz audio (parity mode)."""
    from conette_audio_captioning_b200 import CoNeTTEModel, synth

    sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
    model = CoNeTTEModel(None, sd, synth.make_itos(300), precision="parity")
    try:
        x44 = _wave(2, 44100 * 2, 11)[:, None]  # (B, C=1, N)
        out = model(x44, sr=44100, task="clotho")
        x32 = restate.resample(x44, 44100)
        ref = model(x32, sr=32000, task="clotho")
        assert torch.equal(out["preds"], ref["preds"]) and out["cands"] == ref["cands"]
        torch.testing.assert_close(out["lprobs"], ref["lprobs"], rtol=0, atol=1e-3)
    finally:
        model.engine.close()
