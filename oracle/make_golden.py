"""Generate tests/golden/*.npz from the REAL reference code (run on CPU under the shims).  TEST INFRASTRUCTURE ONLY.

Run in the dev container (where /root/reference exists):   python -m oracle.make_golden
The fixtures pin oracle/restate.py and the CUDA path on machines where the reference cannot be imported.
Weights are not stored: they are regenerated from ``synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)`` and
guarded by the float64 checksums stored in ``weights_checksum``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from conette_audio_captioning_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SD_KW = dict(seed=1234, n_words=300, eos_bias=3.0)
CHECK_KEYS = (
    "preprocessor.encoder.stages.2.4.pwconv1.weight",
    "preprocessor.encoder.bn0.running_mean",
    "model.decoder.layers.3.linear2.weight",
    "model.decoder.classifier.bias",
)


def checksum(sd) -> np.ndarray:
    return np.array([float(sd[k].double().sum()) for k in CHECK_KEYS] + [float(sd[k].double().abs().sum()) for k in CHECK_KEYS])


def make_score_fixture(sd, model) -> None:
    """score.npz: teacher-forced losses of the reference's test_step loop (pl_modules/conette.py:307-313) on given frame
    embeddings, through the REAL ``CoNeTTEPLM.encode_audio`` / ``decode_audio("forcing")`` / ``CrossEntropyLossMean``."""
    g = torch.Generator().manual_seed(91)
    b, tp, n_caps, cap_len = 4, 9, 3, 12
    fe = torch.randn(b, tp, 768, generator=g)
    lens = torch.tensor([9, 4, 7, 1])
    bos_ids = sd["model.task_id_to_token_id"][torch.tensor([0, 1, 2, 0])]
    caps = torch.zeros(b, n_caps, cap_len, dtype=torch.long)
    for i in range(b):
        for j in range(n_caps):
            n_words = cap_len - 2 if (i, j) == (0, 0) else int(torch.randint(1, cap_len - 2, (1,), generator=g))
            caps[i, j, 0] = bos_ids[i]
            caps[i, j, 1 : 1 + n_words] = torch.randint(4, 300, (n_words,), generator=g)
            caps[i, j, 1 + n_words] = 2
    plm = model.model
    crit = ref_loader.ref_module("nn.loss.ce_mean").CrossEntropyLossMean(ignore_index=0, dim=1)
    audio_shape = torch.stack([torch.full((b,), 768), lens], dim=1)
    losses = torch.empty(b, n_caps)
    tok_lp = torch.empty(b, n_caps, cap_len - 1)
    with torch.no_grad():
        enc_outs = plm.encode_audio(fe, audio_shape)  # FrameIdentEncoder: (B, T', 768), lens = audio_shape[:, 1]
        for i in range(n_caps):
            logits = plm.decode_audio(enc_outs, "forcing", caps_in=caps[:, i, :-1])
            losses[:, i] = crit(logits, caps[:, i, 1:])
            lp = torch.log_softmax(logits, dim=1).gather(1, caps[:, i, 1:][:, None, :])[:, 0]
            tok_lp[:, i] = torch.where(caps[:, i, 1:] != 0, lp, torch.zeros(()))
    np.savez_compressed(os.path.join(GOLDEN, "score.npz"), frame_embs=fe.numpy(), lens=lens.numpy(), captions=caps.numpy(),
                        losses=losses.numpy(), token_lprobs=tok_lp.numpy(), weights_checksum=checksum(sd))


def make_baseline_fixture(sd) -> None:
    """baseline.npz: the REAL ``BaselinePLM`` (pl_modules/baseline.py:35; beam 3, min 3, max 12) loaded with the projection +
    decoder of the synthetic state dict, on given frame embeddings: ``forward(batch, "generate")``, ``"greedy"`` and
    ``"forcing"`` outputs.  Note BaselinePLM's vocabulary has no task tokens: V = 311 = rows 0..310 of the CoNeTTE tensors."""
    tok_mod = ref_loader.ref_module("tokenization.aac_tokenizer")
    base_mod = ref_loader.ref_module("pl_modules.baseline")
    tokenizer = tok_mod.AACTokenizer()
    tokenizer.fit(list(synth.make_corpus(300)))
    plm = base_mod.BaselinePLM(train_tokenizer=tokenizer, beam_size=3, min_pred_size=3, max_pred_size=12)
    v = tokenizer.get_vocab_size()
    own = plm.state_dict()
    with torch.no_grad():
        for k, dst in list(plm.named_parameters()) + list(plm.named_buffers()):
            if k == "forbid_rep_mask":
                continue
            src = sd["model." + k]
            if src.shape != dst.shape:  # vocabulary-sized tensors: BaselinePLM has no <bos_task> rows
                src = src[:v]
            dst.copy_(src)
    plm.eval()
    assert set(own) >= {"decoder.classifier.weight", "projection.2.weight"}
    g = torch.Generator().manual_seed(123)
    b, tp, cap_len = 5, 9, 10
    fe = torch.randn(b, tp, 768, generator=g)
    lens = torch.tensor([9, 4, 7, 1, 9])
    audio_shape = torch.stack([torch.full((b,), 768), lens], dim=1)
    caps = torch.zeros(b, cap_len, dtype=torch.long)
    for i in range(b):
        n_words = cap_len - 2 if i == 0 else int(torch.randint(1, cap_len - 2, (1,), generator=g))
        caps[i, 0] = 1
        caps[i, 1 : 1 + n_words] = torch.randint(4, 300, (n_words,), generator=g)
        caps[i, 1 + n_words] = 2
    batch = {"audio": fe, "audio_shape": audio_shape, "captions": caps}
    with torch.no_grad():
        gen = plm(batch, "generate")
        greedy = plm(batch, "greedy")
        forcing = plm(batch, "forcing")
    np.savez_compressed(
        os.path.join(GOLDEN, "baseline.npz"), frame_embs=fe.numpy(), lens=lens.numpy(), captions=caps.numpy(), vocab=np.array(v),
        forbid_rep_mask=plm.forbid_rep_mask.numpy(), cands=np.array(gen["cands"]), preds=gen["preds"].numpy(),
        lprobs=gen["lprobs"].numpy(), mult_preds=gen["mult_preds"].numpy(), mult_lprobs=gen["mult_lprobs"].numpy(),
        greedy_logits=greedy.numpy().astype(np.float32), forcing_logits=forcing.numpy().astype(np.float32),
        weights_checksum=checksum(sd))


def main() -> None:
    if "--only-baseline" in sys.argv:  # added in round 2: leaves the other fixtures untouched
        torch.manual_seed(0)
        make_baseline_fixture(synth.make_state_dict(**SD_KW))
        return
    if "--only-score" in sys.argv:  # added after the other fixtures were committed: leaves them untouched
        torch.manual_seed(0)
        sd = synth.make_state_dict(**SD_KW)
        make_score_fixture(sd, ref_loader.build_reference_model(sd, synth.make_corpus(300)))
        return
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = synth.make_state_dict(**SD_KW)
    model = ref_loader.build_reference_model(sd, synth.make_corpus(300))
    beam_mod = ref_loader.ref_module("nn.decoding.beam")
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- encoder: 2 clips, second one zero-padded (digital silence tail) -------------------------------------
    n = 24000
    wav = synth.make_audio(2, n, seed=11)[:, 0].contiguous()
    wav[1, 15000:] = 0.0
    x_lens = torch.tensor([[n], [15000]])
    enc = model.preprocessor.encoder
    with torch.no_grad():
        lm = enc.logmel_extractor(enc.spectrogram_extractor(wav))[:, 0]
        out = enc(wav, x_lens)
    np.savez_compressed(
        os.path.join(GOLDEN, "encoder.npz"),
        wav=wav.numpy(), x_lens=x_lens[:, 0].numpy(), logmel=lm.numpy(),
        frame_embs=out["frame_embs"].numpy(), frame_embs_lens=out["frame_embs_lens"].numpy(),
        clip_probs=out["clipwise_output"].numpy(), weights_checksum=checksum(sd),
    )

    # ---- beam search on given projected frames --------------------------------------------------------------
    cases = [(1, 3, 20, "content_words"), (2, 0, 20, "content_words"), (3, 3, 20, "content_words"),
             (3, 3, 20, "none"), (3, 0, 5, "all"), (5, 3, 20, "content_words"), (5, 3, 30, "all")]
    g = torch.Generator().manual_seed(77)
    b, tp = 6, 9
    mem = torch.relu(torch.randn(b, tp, 256, generator=g))
    lens = torch.randint(1, tp + 1, (b,), generator=g)
    bos_ids = sd["model.task_id_to_token_id"][torch.randint(0, 7, (b,), generator=g)]
    mask = torch.arange(tp)[None, :] >= lens[:, None]
    store = dict(mem=mem.numpy(), lens=lens.numpy(), bos_ids=bos_ids.numpy(),
                 cases=np.array([f"{k}|{mn}|{mx}|{mode}" for k, mn, mx, mode in cases]), weights_checksum=checksum(sd))
    dec = model.model.decoder
    for ci, (k, mn, mx, mode) in enumerate(cases):
        forbid = synth.make_forbid_rep_mask(synth.make_itos(300), mode)
        preds, lprobs, mpreds, mlprobs = beam_mod.generate(
            decoder=dec, pad_id=0, bos_id=bos_ids, eos_id=2, vocab_size=dec.vocab_size,
            frame_embs=mem.transpose(1, 2).contiguous(), frame_embs_pad_mask=mask, beam_size=k,
            min_pred_size=mn, max_pred_size=mx, forbid_rep_mask=forbid)
        store[f"c{ci}_preds"] = preds.numpy()
        store[f"c{ci}_lprobs"] = lprobs.numpy()
        store[f"c{ci}_mult_preds"] = mpreds.numpy()
        store[f"c{ci}_mult_lprobs"] = mlprobs.numpy()
    # decoder logits for fixed token prefixes (full recompute of the reference decoder)
    steps = 6
    toks = torch.randint(4, 300, (b, steps), generator=g)
    causal = torch.triu(torch.full((steps, steps), float("-inf")), diagonal=1)
    with torch.no_grad():
        full = dec(mem.permute(1, 0, 2).contiguous(), mask, toks.T.contiguous(), None, causal)
    store["tf_tokens"] = toks.numpy()
    store["tf_logits"] = full.permute(1, 0, 2).contiguous().numpy().astype(np.float32)  # (B, steps, V)
    np.savez_compressed(os.path.join(GOLDEN, "decode.npz"), **store)

    # ---- end to end through CoNeTTEModel -----------------------------------------------------------------------
    wav3 = synth.make_audio(3, 32000, seed=5)
    wav3[2, :, 20000:] = 0.0
    x_shapes = torch.tensor([[32000], [32000], [20000]])
    tasks = ["clotho", "audiocaps", "wavcaps_audioset_sl"]
    with torch.no_grad():
        out = model(wav3, sr=32000, x_shapes=x_shapes, task=tasks)
    np.savez_compressed(
        os.path.join(GOLDEN, "e2e.npz"),
        wav=wav3.numpy(), x_shapes=x_shapes.numpy(), tasks=np.array(tasks), cands=np.array(out["cands"]),
        mult_cands=np.array(out["mult_cands"]), preds=out["preds"].numpy(), lprobs=out["lprobs"].numpy(),
        mult_preds=out["mult_preds"].numpy(), mult_lprobs=out["mult_lprobs"].numpy(),
        tags_probs=out["tags_probs"].numpy(), weights_checksum=checksum(sd),
    )
    make_score_fixture(sd, model)
    make_baseline_fixture(sd)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
