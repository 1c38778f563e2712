"""Pin the oracle restatement (oracle/restate.py) against the REAL reference code run on CPU under the shims.

Skipped where the reference sources are absent (neither /root/reference nor baseline/_ref).  The committed fixtures in
tests/golden (tests/test_golden.py) carry the same pin to machines without the reference.
"""
import math

import pytest
import torch

from conette_audio_captioning_b200 import synth
from oracle import ref_loader, restate

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not available")


@pytest.fixture(scope="module")
def ref_small(small_sd):
    return ref_loader.build_reference_model(small_sd, synth.make_corpus(300))


def test_vocab_order_matches_reference_tokenizer(ref_small):
    assert ref_loader.reference_itos(ref_small) == synth.make_itos(300)
    assert torch.equal(ref_small.model.forbid_rep_mask, synth.make_forbid_rep_mask(synth.make_itos(300)))


def test_frozen_frontend_parameters_match_shim(small_sd):
    """synth's analytic DFT basis / mel matrix == the torchlibrosa restatement the reference is run with."""
    ref_loader.ref_module("nn.encoders.convnext")
    from torchlibrosa.stft import LogmelFilterBank, Spectrogram

    spec = Spectrogram(n_fft=1024, hop_length=320, win_length=1024)
    mel = LogmelFilterBank(sr=32000, n_fft=1024, n_mels=224, fmin=50, fmax=14000, top_db=None)
    e = restate.ENC
    assert torch.equal(spec.stft.conv_real.weight, small_sd[e + "spectrogram_extractor.stft.conv_real.weight"])
    assert torch.equal(spec.stft.conv_imag.weight, small_sd[e + "spectrogram_extractor.stft.conv_imag.weight"])
    assert torch.equal(mel.melW, small_sd[e + "logmel_extractor.melW"])
    # torchaudio's independent Slaney implementation agrees to f32 rounding
    import torchaudio

    fb = torchaudio.functional.melscale_fbanks(513, 50.0, 14000.0, 224, 32000, norm="slaney", mel_scale="slaney")
    assert torch.allclose(fb, mel.melW, atol=2e-6)


@pytest.mark.parametrize("n_samples", [32000, 50000, 100001])
def test_encoder_matches_reference(ref_small, small_sd, n_samples):
    wav = synth.make_audio(2, n_samples, seed=7)[:, 0]
    wav[1, n_samples // 2 :] = 0.0  # digital silence tail (zero-padded shorter clip)
    x_lens = torch.tensor([[n_samples], [n_samples // 2]])
    with torch.no_grad():
        ref = ref_small.preprocessor.encoder(wav, x_lens)
        mine = restate.encoder(small_sd, wav, x_lens[:, 0])
    assert torch.equal(ref["frame_embs_lens"], mine["frame_embs_lens"])
    assert ref["frame_embs"].shape == (2, 768, restate.n_out_frames(n_samples))
    torch.testing.assert_close(mine["frame_embs"], ref["frame_embs"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mine["clipwise_output"], ref["clipwise_output"], rtol=1e-4, atol=1e-5)


def test_frame_lens_rounding_depends_on_padded_length():
    # 5 s clip: 15 frames alone, 16 inside a 10 s-padded batch
    assert restate.frame_lens(torch.tensor([160000]), 160000).tolist() == [15]
    assert restate.frame_lens(torch.tensor([160000]), 320000).tolist() == [16]


def _ref_generate(ref_model, mem, lens, bos_ids, beam, min_len, max_len, forbid):
    beam_mod = ref_loader.ref_module("nn.decoding.beam")
    dec = ref_model.model.decoder
    mask = torch.arange(mem.shape[1])[None, :] >= lens[:, None]
    return beam_mod.generate(
        decoder=dec, pad_id=0, bos_id=bos_ids, eos_id=2, vocab_size=dec.vocab_size,
        frame_embs=mem.transpose(1, 2).contiguous(), frame_embs_pad_mask=mask,
        beam_size=beam, min_pred_size=min_len, max_pred_size=max_len, forbid_rep_mask=forbid,
    )


def test_kv_decoder_step_matches_full_recompute(ref_small, small_sd):
    g = torch.Generator().manual_seed(0)
    b, tp, steps = 3, 9, 8
    mem = torch.relu(torch.randn(b, tp, 256, generator=g))
    lens = torch.tensor([9, 4, 7])
    toks = torch.randint(4, 300, (b, steps), generator=g)
    dec = restate.KVDecoder(small_sd, mem, lens, beam=1, max_len=steps)
    ref_dec = ref_small.model.decoder
    causal = torch.triu(torch.full((steps, steps), float("-inf")), diagonal=1)
    mask = torch.arange(tp)[None, :] >= lens[:, None]
    with torch.no_grad():
        full = ref_dec(mem.permute(1, 0, 2).contiguous(), mask, toks.T.contiguous(), None, causal)
    for i in range(steps):
        logits = dec.step(toks[:, i], i)
        torch.testing.assert_close(logits, full[i], rtol=1e-4, atol=5e-5)
        assert torch.equal(logits.argmax(-1), full[i].argmax(-1))


CASES = [
    # beam, min_len, max_len, forbid mode, eos_bias
    (1, 3, 20, "content_words", 3.0),
    (2, 0, 20, "content_words", 4.0),
    (3, 3, 20, "content_words", 3.0),
    (3, 3, 20, "none", 5.0),
    (3, 0, 5, "all", 2.0),
    (5, 3, 20, "content_words", 3.5),
    (5, 3, 30, "all", 2.5),
    (3, 3, 20, "content_words", -20.0),  # nothing finishes: all forced at max-1
    (3, 2, 12, "content_words", 6.0),
]


@pytest.mark.parametrize("beam,min_len,max_len,mode,eos_bias", CASES)
def test_beam_search_matches_reference_generate(ref_small, small_sd, beam, min_len, max_len, mode, eos_bias):
    sd = dict(small_sd)
    bias = small_sd[restate.DEC + "classifier.bias"].clone()
    bias[2] += eos_bias - 3.0
    sd[restate.DEC + "classifier.bias"] = bias
    ref_small.model.decoder.classifier.bias.copy_(bias)
    try:
        g = torch.Generator().manual_seed(100 * beam + max_len)
        b, tp = 5, 7
        mem = torch.relu(torch.randn(b, tp, 256, generator=g))
        lens = torch.randint(1, tp + 1, (b,), generator=g)
        bos_ids = small_sd["model.task_id_to_token_id"][torch.randint(0, 7, (b,), generator=g)]
        forbid = synth.make_forbid_rep_mask(synth.make_itos(300), mode)
        ref = _ref_generate(ref_small, mem, lens, bos_ids, beam, min_len, max_len, forbid)
        mine = restate.beam_search(sd, mem, lens, bos_ids, beam, min_len, max_len, forbid)
    finally:
        ref_small.model.decoder.classifier.bias.copy_(small_sd[restate.DEC + "classifier.bias"])
    for name, r, m in zip(("preds", "lprobs", "mult_preds", "mult_lprobs"), ref, mine):
        assert r.shape == m.shape, name
        if r.dtype == torch.long:
            assert torch.equal(r, m), name
        else:
            torch.testing.assert_close(m, r, rtol=1e-5, atol=1e-5)
    if eos_bias > 0:
        # the sweep really exercises early finishing / shrinking beams
        assert (ref[2] == 2).any()


def _random_captions(b, n_caps, cap_len, bos_ids, seed):
    """(B, n_caps, cap_len) ids: task BOS, words, EOS, then pads -- ragged lengths, one caption filling the whole row."""
    g = torch.Generator().manual_seed(seed)
    caps = torch.zeros(b, n_caps, cap_len, dtype=torch.long)
    for i in range(b):
        for j in range(n_caps):
            n_words = cap_len - 2 if (i, j) == (0, 0) else int(torch.randint(1, cap_len - 2, (1,), generator=g))
            caps[i, j, 0] = bos_ids[i]
            caps[i, j, 1 : 1 + n_words] = torch.randint(4, 300, (n_words,), generator=g)
            caps[i, j, 1 + n_words] = 2
    return caps


def test_teacher_forced_scoring_matches_reference(ref_small, small_sd):
    """restate.score_captions == the reference's test_step loss loop (conette.py:307-313) run with the real modules."""
    g = torch.Generator().manual_seed(1)
    b, tp, n_caps, cap_len = 4, 9, 3, 12
    fe = torch.randn(b, tp, 768, generator=g)
    lens = torch.tensor([9, 4, 7, 1])
    bos_ids = small_sd["model.task_id_to_token_id"][torch.tensor([0, 1, 2, 0])]
    caps = _random_captions(b, n_caps, cap_len, bos_ids, seed=2)
    tok_lp, losses = restate.score_captions(small_sd, fe, lens, caps)

    plm = ref_small.model
    crit = ref_loader.ref_module("nn.loss.ce_mean").CrossEntropyLossMean(ignore_index=0, dim=1)
    mem = restate.project(small_sd, fe).transpose(1, 2).contiguous()
    enc_outs = {"frame_embs": mem, "frame_embs_pad_mask": torch.arange(tp)[None, :] >= lens[:, None]}
    with torch.no_grad():
        for i in range(n_caps):
            logits = plm.decode_audio(enc_outs, "forcing", caps_in=caps[:, i, :-1])  # (B, V, L)
            ref_loss = crit(logits, caps[:, i, 1:])
            torch.testing.assert_close(losses[:, i], ref_loss, rtol=1e-5, atol=1e-5)
            ref_lp = torch.log_softmax(logits, dim=1).gather(1, caps[:, i, 1:][:, None, :])[:, 0]
            ref_lp = torch.where(caps[:, i, 1:] != 0, ref_lp, torch.zeros(()))
            torch.testing.assert_close(tok_lp[:, i], ref_lp, rtol=1e-4, atol=1e-4)


def test_precomputed_embeddings_path_matches_reference(ref_small, small_sd):
    """``CoNeTTEModel(x=(B, T', 768), x_shapes=[[768, len]], preprocess=False)`` (reference huggingface/model.py:205-212,
    FrameIdentEncoder nn/encoders/ident.py:14-34) == projection + fixed-slot beam search of the restatement."""
    g = torch.Generator().manual_seed(4)
    b, tp = 4, 11
    fe = torch.randn(b, tp, 768, generator=g)
    lens = torch.tensor([11, 3, 8, 1])
    x_shapes = torch.stack([torch.full((b,), 768), lens], dim=1)
    tasks = ["clotho", "audiocaps", "macs", "clotho"]
    with torch.no_grad():
        ref = ref_small(fe, x_shapes=x_shapes, preprocess=False, task=tasks)
    bos_ids = small_sd["model.task_id_to_token_id"][torch.tensor([synth.TASK_NAMES.index(t) for t in tasks])]
    preds, lprobs, mult_preds, mult_lprobs = restate.beam_search(
        small_sd, restate.project(small_sd, fe), lens, bos_ids, 3, 3, 20, small_sd["model.forbid_rep_mask"])
    assert torch.equal(ref["preds"], preds) and torch.equal(ref["mult_preds"], mult_preds)
    torch.testing.assert_close(lprobs, ref["lprobs"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mult_lprobs, ref["mult_lprobs"], rtol=1e-4, atol=1e-4)
    assert "tags" not in ref and "tags_probs" not in ref


def test_end_to_end_matches_reference_model(ref_small, small_sd):
    wav = synth.make_audio(3, 64000, seed=3)
    wav[2, :, 40000:] = 0
    x_shapes = torch.tensor([[64000], [64000], [40000]])
    tasks = ["clotho", "audiocaps", "wavcaps_audioset_sl"]
    with torch.no_grad():
        ref = ref_small(wav, sr=32000, x_shapes=x_shapes, task=tasks)
    task_idx = torch.tensor([synth.TASK_NAMES.index(t) for t in tasks])
    bos_ids = small_sd["model.task_id_to_token_id"][task_idx]
    mine = restate.caption(small_sd, wav[:, 0], x_shapes[:, 0], bos_ids, 3, 3, 20, small_sd["model.forbid_rep_mask"])
    assert torch.equal(ref["preds"], mine["preds"])
    assert torch.equal(ref["mult_preds"], mine["mult_preds"])
    torch.testing.assert_close(mine["lprobs"], ref["lprobs"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mine["clip_probs"], ref["tags_probs"], rtol=1e-4, atol=1e-5)
