#!/usr/bin/env python
"""Decode-only timing at several batch sizes (is the beam-search graph latency- or throughput-bound?), plus two sub-batches
decoded concurrently from two engines/streams: python tools/decode_scaling.py"""
import os
import sys
import threading

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

sd = synth.make_state_dict(seed=1234, n_words=4000)
V = sd["model.decoder.classifier.weight"].shape[0]
forbid = sd["model.forbid_rep_mask"].cuda().to(torch.uint8)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


eng = Engine(sd, V, precision="fast")
for b in (1, 2, 4, 8, 16, 32, 64, 128):
    fe = torch.randn(b, 31, 768, device="cuda")
    bos = sd["model.task_id_to_token_id"][torch.zeros(b, dtype=torch.long)].cuda()
    lens = torch.full((b,), 31)
    ms = timed(lambda: eng.decode(fe, lens, bos, forbid))
    print(f"decode B={b}: {ms:.3f} ms", flush=True)
