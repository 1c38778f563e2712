"""``BaselinePLM`` (reference pl_modules/baseline.py:35, forward :310, decode_audio :339) on the CUDA decoder
against outputs of the REAL reference module (tests/golden/baseline.npz, made by oracle/make_golden.py --only-baseline): plain
``<bos>`` start token, a vocabulary without task tokens (V = 311), the three decode methods generate / greedy / forcing."""
import numpy as np
import pytest
import torch

from conette_audio_captioning_b200 import synth

from golden_util import assert_weights_match, load, t


def _baseline_state_dict(small_sd, v):
    sd = {k[len("model."):]: w for k, w in small_sd.items() if k.startswith(("model.decoder.", "model.projection."))}
    for k in ("decoder.emb_layer.weight", "decoder.classifier.weight", "decoder.classifier.bias"):
        sd[k] = sd[k][:v].contiguous()  # BaselinePLM has no <bos_task> rows
    return sd


def test_baseline_fixture_is_self_consistent(small_sd):
    """CPU: the oracle's beam search with a plain <bos> reproduces the real BaselinePLM's beams (pins restate.py for this caller)."""
    from oracle import restate

    fx = load("baseline.npz")
    assert_weights_match(small_sd, fx)
    v = int(fx["vocab"])
    sd = {"model." + k: w for k, w in _baseline_state_dict(small_sd, v).items()}
    fe, lens = t(fx["frame_embs"]), t(fx["lens"])
    ref = restate.beam_search(sd, restate.project(sd, fe), lens, torch.ones(5, dtype=torch.long), 3, 3, 12, t(fx["forbid_rep_mask"]))
    assert np.array_equal(ref[2].numpy(), fx["mult_preds"]) and np.array_equal(ref[0].numpy(), fx["preds"])
    np.testing.assert_allclose(ref[3].numpy(), fx["mult_lprobs"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fast", "parity"])
def test_baseline_plm_vs_reference(small_sd, precision):
    from conette_audio_captioning_b200.baseline import BaselinePLM

    fx = load("baseline.npz")
    assert_weights_match(small_sd, fx)
    v = int(fx["vocab"])
    sd = _baseline_state_dict(small_sd, v)
    sd["forbid_rep_mask"] = t(fx["forbid_rep_mask"])
    itos = synth.make_itos(300, task_names=())
    assert len(itos) == v
    plm = BaselinePLM(sd, itos, min_pred_size=3, max_pred_size=12, beam_size=3, precision=precision)
    try:
        fe, lens, caps = t(fx["frame_embs"]), t(fx["lens"]), t(fx["captions"])
        batch = {"audio": fe, "audio_shape": torch.stack([torch.full_like(lens, 768), lens], dim=1), "captions": caps}
        gen = plm(batch, "generate")
        assert np.array_equal(gen["preds"].numpy(), fx["preds"]) and np.array_equal(gen["mult_preds"].numpy(), fx["mult_preds"])
        np.testing.assert_allclose(gen["lprobs"].numpy(), fx["lprobs"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(gen["mult_lprobs"].numpy(), fx["mult_lprobs"], rtol=1e-4, atol=1e-4)
        assert gen["cands"] == [str(s) for s in fx["cands"]]
        forcing = plm(batch, "forcing").numpy()  # (B, V, L)
        assert forcing.shape == fx["forcing_logits"].shape
        live = (caps[:, :-1] != 0).numpy()  # rows are causal: positions before the first pad do not see the pads
        diff = np.abs(forcing - fx["forcing_logits"]).max(axis=1)
        assert float(diff[live].max()) < 2e-4, float(diff[live].max())
        if precision == "fast":  # the tap lives in the cluster kernel
            greedy = plm(batch, "greedy").numpy()
            want = fx["greedy_logits"]
            assert greedy.shape == want.shape
            assert np.array_equal(np.isinf(greedy), np.isinf(want))
            fin = np.isfinite(want)
            assert float(np.abs(greedy[fin] - want[fin]).max()) < 2e-4
            assert np.array_equal(greedy.argmax(1), want.argmax(1))
        with pytest.raises(ValueError):
            plm(batch, "sampling")
        with pytest.raises(ValueError):
            plm.decode_audio(plm.encode_audio(batch["audio"], batch["audio_shape"]), "forcing")
    finally:
        plm.close()
