"""oracle/restate.py against the committed golden vectors (made by oracle/make_golden.py from the real reference)."""
import numpy as np
import pytest
import torch

from conette_audio_captioning_b200 import synth
from conette_audio_captioning_b200.tokenizer import IdTokenizer
from oracle import restate

from golden_util import assert_weights_match, load, t


def test_encoder_golden(small_sd):
    fx = load("encoder.npz")
    assert_weights_match(small_sd, fx)
    wav, x_lens = t(fx["wav"]), t(fx["x_lens"])
    taps = {}
    out = restate.encoder(small_sd, wav, x_lens, taps)
    torch.testing.assert_close(taps["logmel"], t(fx["logmel"]), rtol=0, atol=2e-3)  # dB
    assert (taps["logmel"][1, -10:] == -100.0).all()  # digital silence -> exactly -100 dB
    assert np.array_equal(out["frame_embs_lens"].numpy(), fx["frame_embs_lens"])
    torch.testing.assert_close(out["frame_embs"], t(fx["frame_embs"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["clipwise_output"], t(fx["clip_probs"]), rtol=1e-4, atol=1e-5)


def test_decoder_logits_golden(small_sd):
    fx = load("decode.npz")
    assert_weights_match(small_sd, fx)
    mem, lens, toks = t(fx["mem"]), t(fx["lens"]), t(fx["tf_tokens"])
    dec = restate.KVDecoder(small_sd, mem, lens, beam=1, max_len=toks.shape[1])
    for i in range(toks.shape[1]):
        logits = dec.step(toks[:, i], i)
        torch.testing.assert_close(logits, t(fx["tf_logits"][:, i]), rtol=1e-4, atol=5e-5)


def test_teacher_forced_scoring_golden(small_sd):
    fx = load("score.npz")
    assert_weights_match(small_sd, fx)
    tok_lp, losses = restate.score_captions(small_sd, t(fx["frame_embs"]), t(fx["lens"]), t(fx["captions"]))
    torch.testing.assert_close(losses, t(fx["losses"]), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(tok_lp, t(fx["token_lprobs"]), rtol=1e-4, atol=1e-4)


def test_beam_search_golden(small_sd):
    fx = load("decode.npz")
    mem, lens, bos_ids = t(fx["mem"]), t(fx["lens"]), t(fx["bos_ids"])
    itos = synth.make_itos(300)
    for ci, case in enumerate(fx["cases"]):
        k, mn, mx, mode = str(case).split("|")
        forbid = synth.make_forbid_rep_mask(itos, mode)
        out = restate.beam_search(small_sd, mem, lens, bos_ids, int(k), int(mn), int(mx), forbid)
        for name, mine in zip(("preds", "lprobs", "mult_preds", "mult_lprobs"), out):
            ref = fx[f"c{ci}_{name}"]
            assert tuple(mine.shape) == ref.shape, (case, name)
            if mine.dtype == torch.long:
                assert np.array_equal(mine.numpy(), ref), (case, name)
            else:
                np.testing.assert_allclose(mine.numpy(), ref, rtol=1e-5, atol=1e-5)


def test_e2e_golden(small_sd):
    fx = load("e2e.npz")
    assert_weights_match(small_sd, fx)
    wav, x_shapes = t(fx["wav"]), t(fx["x_shapes"])
    tasks = [str(s) for s in fx["tasks"]]
    bos = small_sd["model.task_id_to_token_id"][torch.tensor([synth.TASK_NAMES.index(s) for s in tasks])]
    out = restate.caption(small_sd, wav[:, 0], x_shapes[:, 0], bos, 3, 3, 20, small_sd["model.forbid_rep_mask"])
    assert np.array_equal(out["preds"].numpy(), fx["preds"])
    assert np.array_equal(out["mult_preds"].numpy(), fx["mult_preds"])
    np.testing.assert_allclose(out["lprobs"].numpy(), fx["lprobs"], rtol=1e-4, atol=1e-4)
    tok = IdTokenizer(synth.make_itos(300))
    assert tok.decode_rec(out["preds"]) == [str(s) for s in fx["cands"]]
    assert tok.decode_rec(out["mult_preds"]) == [[str(s) for s in row] for row in fx["mult_cands"]]
