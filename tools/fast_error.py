#!/usr/bin/env python
"""Relative L2 error of the fast (fp16-operand GEMM) and parity (fp32) encoders against the fp32 CPU oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402
from oracle import restate  # noqa: E402

sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
for n, b, seed in ((32000, 2, 5), (160000, 2, 7)):
    wav = synth.make_audio(b, n, seed=seed)[:, 0]
    ref = restate.encoder(sd, wav)["frame_embs"].transpose(1, 2)
    for prec in ("parity", "fast"):
        eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision=prec)
        fe, _ = eng.encoder(wav)
        err = float((fe.cpu() - ref).norm() / ref.norm())
        print(f"n={n} {prec}: frame_embs rel-L2 {err:.3e}")
        eng.close()
