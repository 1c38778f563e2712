#!/usr/bin/env python
"""Run one pass of a part of the hot path (for ncu captures): python tools/run_once.py [encoder|decode|caption] --batch 8"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["encoder", "decode", "caption"])
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--chunk", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--precision", default="fast")
ap.add_argument("--no-graphs", action="store_true")
a = ap.parse_args()
sd = synth.make_state_dict(seed=1234, n_words=4000)
eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision=a.precision, enc_chunk=a.chunk)
n = int(a.seconds * 32000)
wav = synth.make_audio(a.batch, n, seed=1)[:, 0].cuda()
bos = sd["model.task_id_to_token_id"][torch.zeros(a.batch, dtype=torch.long)].cuda()
forbid = sd["model.forbid_rep_mask"].cuda().to(torch.uint8)
for _ in range(a.reps):
    if a.what == "encoder":
        eng.encoder(wav)
    elif a.what == "decode":
        fe = torch.randn(a.batch, 31, 768, device="cuda")
        eng.decode(fe, torch.full((a.batch,), 31), bos, forbid)
    else:
        eng.caption(wav, None, bos, forbid)
torch.cuda.synchronize()
print("done")
