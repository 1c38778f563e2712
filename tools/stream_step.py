#!/usr/bin/env python
"""ms per step of the streaming host API (two batches in flight) under the current environment (CNB_* switches are read once per
process, so sweeps run one process per setting): python tools/stream_step.py [--batch 64] [--steps 30] [--device-inputs]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--device-inputs", action="store_true")
ap.add_argument("--tag", default="")
a = ap.parse_args()
sd = synth.make_state_dict(seed=1234, n_words=4000)
V = sd["model.decoder.classifier.weight"].shape[0]
forbid = sd["model.forbid_rep_mask"].to(torch.uint8)
eng = Engine(sd, V, precision="fast")
b, n = a.batch, int(a.seconds * 32000)
g = torch.Generator().manual_seed(0)
wavs = [(torch.randn(b, n, generator=g) * 0.1).pin_memory() for _ in range(2)]
if a.device_inputs:
    wavs = [w.cuda() for w in wavs]
bos = sd["model.task_id_to_token_id"][torch.zeros(b, dtype=torch.long)]
outs = [eng.alloc_host_outputs(b, 3, 20) for _ in range(2)]


def run(steps):
    prev = None
    for i in range(steps):
        t = eng.caption_host_begin(wavs[i & 1], None, bos, forbid, out=outs[i & 1])
        if prev is not None:
            eng.caption_host_end(prev)
        prev = t
    return eng.caption_host_end(prev)


run(5)
torch.cuda.synchronize()
t0 = time.perf_counter()  # the library runs on its own streams and caption_host_end waits for the batch: wall time is exact here
res = run(a.steps)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / a.steps
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CNB_"))
print(f"{a.tag or env or 'default':60s} {ms:7.3f} ms/step  {b * a.seconds / ms:7.1f} k audio-s/s  ids[0]={res[0][0][:6].tolist()}")
