#!/usr/bin/env python
"""Kernel timeline of the streaming host API (two batches in flight): python tools/stream_timeline.py [--batch 64] [--steps 6]

Brackets every launch group with CUDA events (cnb_profile_timeline_begin/_end) while the decoder of batch i runs on its own stream
next to the encoder of batch i+1, then prints, for the last steady-state step, each bracket's begin / end relative to the step's
first bracket, plus per-class sums to set against the stand-alone (sequential) kernel times of bench.py."""
import argparse
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--device-inputs", action="store_true")
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
    eng.caption_host_end(prev)


run(4)
torch.cuda.synchronize()
eng.profile_timeline_begin()
run(a.steps)
tl = eng.profile_timeline_end()
# split into steps at every "frontend" bracket that follows a non-frontend/stem one
steps, cur = [], []
for name, t0, t1 in tl:
    if name == "frontend" and cur and cur[-1][0] not in ("frontend", "stem"):
        steps.append(cur)
        cur = []
    cur.append((name, t0, t1))
steps.append(cur)
print(f"{len(tl)} brackets, {len(steps)} issue groups; total {tl[-1][2]:.3f} ms for {a.steps} steps")
starts = [s[0][1] for s in steps]
print("step starts (ms):", " ".join(f"{x:.3f}" for x in starts))
# decoder brackets: class dec_gemm
dec = [(t0, t1) for name, t0, t1 in tl if name == "dec_gemm"]
print("decoder kernel (begin, end, dur):", " | ".join(f"{t0:.3f} {t1:.3f} {t1 - t0:.3f}" for t0, t1 in dec))
k = len(steps) - 2 if len(steps) >= 3 else len(steps) - 1
base = steps[k][0][1]
print(f"--- issue group {k} (relative to its first bracket at {base:.3f} ms) ---")
for name, t0, t1 in steps[k]:
    print(f"{name:16s} {t0 - base:8.3f} -> {t1 - base:8.3f}   {t1 - t0:7.3f}")
sums = defaultdict(float)
for name, t0, t1 in steps[k]:
    sums[name] += t1 - t0
print("--- per-class sums in that group ---")
for name, v in sums.items():
    print(f"{name:16s} {v:7.3f}")
