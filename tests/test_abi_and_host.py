"""CPU-side checks: the C-ABI library builds/loads and exports every declared symbol; host mirrors behave like the reference."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from conette_audio_captioning_b200 import _lib, build

    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from conette_audio_captioning_b200 import _lib

    header = (ROOT / "include" / "conette_b200.h").read_text()
    declared = set(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/conette_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes prototypes out of sync with the header"
    assert lib.cnb_abi_version() == 1


def test_geometry_closed_form(lib):
    from conette_audio_captioning_b200 import _lib
    from oracle import restate

    for n in (8639, 8640, 32000, 50000, 100001, 160000, 320000, 333333, 960000):
        t, hs, tp = _lib.geometry(n)
        assert t == restate.n_stft_frames(n) and hs == restate.stage_heights(n) and tp == restate.n_out_frames(n)
    assert _lib.geometry(320000) == (1001, [252, 126, 63, 31], 31)
    assert _lib.geometry(960000) == (3001, [752, 376, 188, 94], 94)


def test_no_device_means_loud_failure_not_fallback(lib):
    """On a machine without a GPU the product path must raise, never fall back to the oracle / PyTorch."""
    from conette_audio_captioning_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = _lib.Config()
    cfg.abi_version, cfg.device, cfg.vocab_size = 1, 0, 318
    h = ctypes.c_void_p()
    rc = lib.cnb_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and lib.cnb_last_error()
    from conette_audio_captioning_b200.engine import Engine

    with pytest.raises(_lib.CnbError):
        Engine({}, 318)


def test_product_code_never_imports_the_oracle():
    for f in (ROOT / "conette_audio_captioning_b200").rglob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle/"


def test_preprocessor_input_forms():
    from conette_audio_captioning_b200.preprocessor import load_resample

    n = 1000
    w, l = load_resample(torch.ones(n))
    assert w.shape == (1, n) and l.tolist() == [n]
    w, l = load_resample(torch.stack([torch.ones(n), 3 * torch.ones(n)]))  # (C, N): one clip, channel mean
    assert w.shape == (1, n) and torch.allclose(w, torch.full((1, n), 2.0))
    w, l = load_resample(torch.ones(4, 2, n), sr=32000)
    assert w.shape == (4, n) and l.tolist() == [n] * 4
    w, l = load_resample([torch.ones(1, n), torch.ones(2, 600)], sr=[32000, 32000])
    assert w.shape == (2, n) and l.tolist() == [n, 600] and float(w[1, 600:].abs().sum()) == 0.0
    w, l = load_resample(torch.ones(2, 1, n), x_shapes=torch.tensor([[n], [400]]))
    assert l.tolist() == [n, 400]
    with pytest.raises(RuntimeError):  # sr != 32 kHz needs the engine's GPU resampler: no host fallback (tests/test_resample.py)
        load_resample(torch.ones(1, 1, 1600), sr=16000)
    with pytest.raises(ValueError):
        load_resample(torch.ones(1, 1, 1, n))
    with pytest.raises(ValueError):
        load_resample(torch.ones(1, 1, n), sr=16000, x_shapes=torch.tensor([[n]]), resampler=lambda *a: None)


def test_id_tokenizer_decode_rules():
    from conette_audio_captioning_b200.tokenizer import IdTokenizer

    tok = IdTokenizer(["<pad>", "<bos>", "<eos>", "<unk>", "rain", "is", "pouring", ",", "<bos_clotho>"])
    assert tok.decode_rec(torch.tensor([[4, 5, 6, 2, 0, 0], [4, 7, 4, 2, 0, 0]])) == ["rain is pouring", "rain, rain"]
    assert tok.decode_rec(torch.tensor([[[4, 2], [5, 2]]])) == [["rain", "is"]]
    assert tok.decode_rec([4, 3, 5]) == "rain is"  # <unk> is stripped like the other specials


def test_forbid_mask_modes():
    from conette_audio_captioning_b200 import synth

    itos = synth.make_itos(10)
    assert synth.make_forbid_rep_mask(itos, "none") is None
    assert synth.make_forbid_rep_mask(itos, "all").all()
    m = synth.make_forbid_rep_mask(itos, "content_words")
    assert not m[itos.index("the")] and m[itos.index("w3")] and m[itos.index("<bos_clotho>")]
    with pytest.raises(ValueError):
        synth.make_forbid_rep_mask(itos, "bogus")


def test_batched_detokenise_equals_the_regex_pipeline():
    """the vectorised id -> text path must give exactly what the normaliser chain gives
    (reference aac_tokenizer.py:197-209, :327-388; normalizers.py:160-188), for plain and for awkward vocabularies."""
    from conette_audio_captioning_b200 import synth
    from conette_audio_captioning_b200.tokenizer import IdTokenizer

    g = torch.Generator().manual_seed(0)
    tok = IdTokenizer(synth.make_itos(500))
    ids = torch.randint(0, tok.get_vocab_size(), (64, 3, 20), generator=g)
    ids[:, :, 15:] = 0
    ids[:, :, 14] = 2
    assert tok.decode_rec(ids) == tok.decode_rec(ids.tolist())
    assert tok.decode_rec(ids[:, 0]) == tok.decode_rec(ids[:, 0].tolist())
    assert tok.decode_rec(ids[3, 1]) == tok.decode_rec(ids[3, 1].tolist())
    odd = IdTokenizer(["<pad>", "<bos>", "<eos>", "<unk>", "rain", ",", ".", "well-known", "-", "it's", "'", "Dog", "a", "!",
                       "<bos_clotho>", "x y"])
    ids = torch.randint(0, odd.get_vocab_size(), (400, 12), generator=g)
    assert odd.decode_rec(ids) == odd.decode_rec(ids.tolist())
    assert odd.decode_rec(torch.tensor([[4, 2, 0, 0]])) == ["rain"] and odd.decode_rec(torch.tensor([[0, 0]])) == [""]


def test_wrapper_rejects_unsupported_arguments_before_touching_the_gpu():
    """a CPU device must not silently become cuda:0, and values beyond the CUDA library's limits (beam <= 8,
    max_pred_size <= 64, <= 8 captions per clip) must surface as a ValueError naming the range, not as a generic C-ABI error."""
    import pytest

    from conette_audio_captioning_b200 import CoNeTTEModel, synth

    sd = {"model.decoder.classifier.weight": __import__("torch").zeros(318, 256)}
    with pytest.raises(ValueError, match="CUDA device"):
        CoNeTTEModel(None, sd, synth.make_itos(300), device="cpu")
    CoNeTTEModel._check_limits(8, 64, 8)
    for bad in ((9, 20, 1), (3, 65, 1), (3, 0, 1), (3, 20, 9)):
        with pytest.raises(ValueError, match="supports|scores"):
            CoNeTTEModel._check_limits(*bad)
