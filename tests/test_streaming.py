"""Split-phase host API (cnb_caption_host_begin/_end) and CoNeTTEModel.stream: two batches in flight must give exactly what
one blocking call per batch gives."""
import pytest
import torch

from conette_audio_captioning_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)


def test_begin_end_equals_blocking_call(sd):
    from conette_audio_captioning_b200.engine import Engine

    eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision="fast")
    try:
        forbid = sd["model.forbid_rep_mask"]
        batches = []
        for i, (b, n) in enumerate([(16, 48000), (20, 64000), (3, 32000), (16, 48000), (17, 40000)]):
            wav = synth.make_audio(b, n, seed=50 + i)[:, 0].contiguous().pin_memory()
            lens = torch.randint(n // 2, n + 1, (b,), generator=torch.Generator().manual_seed(i))
            bos = sd["model.task_id_to_token_id"][torch.randint(0, 7, (b,), generator=torch.Generator().manual_seed(i))]
            batches.append((wav, lens, bos))
        want = [eng.caption_host(w, l, bos, forbid, 3, 3, 20) for w, l, bos in batches]
        got, ticket = [], None
        for w, l, bos in batches:  # begin(i+1) before end(i)
            nxt = eng.caption_host_begin(w, l, bos, forbid, 3, 3, 20)
            if ticket is not None:
                got.append(eng.caption_host_end(ticket))
            ticket = nxt
        got.append(eng.caption_host_end(ticket))
        for a, c in zip(want, got):
            for x, y in zip(a, c):
                assert torch.equal(x, y)
        with pytest.raises(Exception):
            eng.caption_host_end(ticket)  # already collected
        # device-resident waveforms are used in place by the same entry point
        t0 = eng.caption_host_begin(batches[0][0].cuda(), batches[0][1], batches[0][2], forbid, 3, 3, 20)
        t1 = eng.caption_host_begin(batches[1][0].cuda(), batches[1][1], batches[1][2], forbid, 3, 3, 20)
        for i, tk in ((0, t0), (1, t1)):
            for x, y in zip(want[i], eng.caption_host_end(tk)):
                assert torch.equal(x, y)
        # three begins without an end: the oldest batch is waited for internally, results stay right
        t = [eng.caption_host_begin(*batches[i][:3], forbid, 3, 3, 20) for i in range(3)]
        for i in (1, 2):
            for x, y in zip(want[i], eng.caption_host_end(t[i])):
                assert torch.equal(x, y)
    finally:
        eng.close()


def test_model_stream_equals_calls(sd):
    from conette_audio_captioning_b200 import CoNeTTEModel

    model = CoNeTTEModel(None, sd, synth.make_itos(300), precision="parity")
    try:
        xs = [synth.make_audio(b, n, seed=70 + i) for i, (b, n) in enumerate([(2, 32000), (16, 40000), (1, 50000)])]
        want = [model(x, sr=32000, task="audiocaps") for x in xs]
        got = list(model.stream(xs, sr=32000, task="audiocaps"))
        assert len(got) == len(want)
        for a, c in zip(want, got):
            assert a["cands"] == c["cands"] and a["mult_cands"] == c["mult_cands"] and a["tasks"] == c["tasks"]
            assert torch.equal(a["preds"], c["preds"]) and torch.equal(a["mult_preds"], c["mult_preds"])
            assert torch.equal(a["lprobs"], c["lprobs"]) and torch.equal(a["tags_probs"], c["tags_probs"]) and a["tags"] == c["tags"]
        assert list(model.stream([])) == []
    finally:
        model.engine.close()


def test_stream_with_mixed_sample_rates(sd):
    """A 32 kHz batch goes through the split-phase host path (the handle's own streams), a 44.1 kHz batch is resampled on the
    GPU and goes through the device-buffer path (torch's stream); both use the same workspaces, so the library orders them
    (dev_enter / dev_leave in csrc/api.cu).  Alternating the two -- and interleaving unrelated engine calls between a yield and
    the next collect -- must give exactly what one call per batch gives."""
    from conette_audio_captioning_b200 import CoNeTTEModel

    model = CoNeTTEModel(None, sd, synth.make_itos(300), precision="fast")
    try:
        specs = [(16, 48000, 32000), (16, 66150, 44100), (16, 48000, 32000), (3, 44100, 44100), (17, 40000, 32000),
                 (16, 48000, 32000), (2, 70000, 44100)]
        xs = [synth.make_audio(b, n, seed=90 + i) for i, (b, n, _) in enumerate(specs)]
        srs = [s for _, _, s in specs]
        want = [model(x, sr=s, task="clotho") for x, s in zip(xs, srs)]
        for rep in range(3):  # the race, when present, is timing dependent
            got, pending = [], None
            for i, (x, s) in enumerate(zip(xs, srs)):
                nxt = model._begin(x, s, "clotho")
                if i % 2 == 0:  # unrelated device-path work while a host batch may be in flight
                    model.engine.encoder(xs[0][:2, 0].cuda(), with_tags=False)
                if pending is not None:
                    got.append(model._finish(pending, 0.3))
                pending = nxt
            got.append(model._finish(pending, 0.3))
            for a, c in zip(want, got):
                assert torch.equal(a["preds"], c["preds"]) and torch.equal(a["mult_preds"], c["mult_preds"]), rep
                assert torch.equal(a["lprobs"], c["lprobs"]) and torch.equal(a["tags_probs"], c["tags_probs"]), rep
    finally:
        model.engine.close()


def test_timeline_brackets_keep_the_overlap(sd):
    """cnb_profile_timeline_begin/_end: every launch group is bracketed in issue order while the decode of batch i still runs
    on its own stream next to the encoder of batch i+1, and the results are unchanged."""
    from conette_audio_captioning_b200.engine import Engine

    eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision="fast")
    try:
        forbid = sd["model.forbid_rep_mask"]
        b, n = 16, 64000
        wavs = [synth.make_audio(b, n, seed=70 + i)[:, 0].contiguous().pin_memory() for i in range(3)]
        bos = sd["model.task_id_to_token_id"][torch.zeros(b, dtype=torch.long)]
        want = [eng.caption_host(w, None, bos, forbid, 3, 3, 20) for w in wavs]
        eng.profile_timeline_begin()
        got, ticket = [], None
        for w in wavs:
            nxt = eng.caption_host_begin(w, None, bos, forbid, 3, 3, 20)
            if ticket is not None:
                got.append(eng.caption_host_end(ticket))
            ticket = nxt
        got.append(eng.caption_host_end(ticket))
        tl = eng.profile_timeline_end()
        for a, c in zip(want, got):
            for x, y in zip(a, c):
                assert torch.equal(x, y)
        names = [name for name, _, _ in tl]
        assert names.count("dec_gemm") == 3 and names.count("dwconv_ln.s1") == 9 and "frontend" in names
        assert all(t1 >= t0 >= 0.0 for _, t0, t1 in tl)
        dec = [(t0, t1) for name, t0, t1 in tl if name == "dec_gemm"]
        enc = [(t0, t1) for name, t0, t1 in tl if name == "dwconv_ln.s1"]
        # the second batch's first ConvNeXt block starts before the first batch's decode has finished
        assert enc[3][0] < dec[0][1]
        assert eng.profile_timeline_end() == []  # nothing recorded outside a begin/end pair
    finally:
        eng.close()
