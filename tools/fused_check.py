#!/usr/bin/env python
"""Fast-mode stage-1 block outputs and frame embeddings against the fp32 oracle (run with / without CNB_NO_MLP_FUSED=1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import _lib, synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402
from oracle import restate  # noqa: E402

sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
for b, n in ((2, 48000), (3, 100000)):
    wav = synth.make_audio(b, n, seed=5 + b)[:, 0].contiguous()
    taps = {}
    ref = restate.encoder(sd, wav, None, taps)["frame_embs"].transpose(1, 2)
    eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision="fast", enc_chunk=8)
    for blk in (0, 1, 2):
        got = eng.encoder_tap(wav, _lib.TAP_BLOCK, 0, blk).cpu()
        want = taps[f"block.0.{blk}"].permute(0, 2, 3, 1)
        print(f"b={b} n={n} block 0.{blk}: rel-L2 {float((got - want).norm() / want.norm()):.3e}  max|d| {float((got - want).abs().max()):.3e}")
    fe, _ = eng.encoder(wav)
    print(f"b={b} n={n} frame_embs rel-L2 {float((fe.cpu() - ref).norm() / ref.norm()):.3e}")
    eng.close()
