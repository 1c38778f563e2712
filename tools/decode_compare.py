#!/usr/bin/env python
"""Cluster (fp16 hi/lo split tensor-core) decoder vs fp32 graph decoder on the same inputs: agreement + timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = synth.make_state_dict(seed=1234, n_words=4000)
V = sd["model.decoder.classifier.weight"].shape[0]
forbid = sd["model.forbid_rep_mask"].cuda().to(torch.uint8)
g = torch.Generator().manual_seed(5)
fe = torch.randn(b, 31, 768, generator=g).cuda()
bos = sd["model.task_id_to_token_id"][torch.zeros(b, dtype=torch.long)].cuda()
lens = torch.full((b,), 31)
outs = {}
for mode in ("graph", "cluster"):
    eng = Engine(sd, V, precision="fast", decoder=mode)
    for _ in range(3):
        o = eng.decode(fe, lens, bos, forbid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = eng.decode(fe, lens, bos, forbid)
    e1.record()
    torch.cuda.synchronize()
    outs[mode] = [t.cpu() for t in o]
    print(f"{mode}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    eng.close()
gp, gl, gmp, gml = outs["graph"]
cp, cl, cmp_, cml = outs["cluster"]
print("shapes", gp.shape, cp.shape, gmp.shape, cmp_.shape)
if gmp.shape == cmp_.shape:
    same = (gmp == cmp_).flatten(1).all(1)
    print(f"clips with identical beams: {int(same.sum())}/{b}")
    print("max |lprob diff| on identical clips:", float((gml[same] - cml[same]).abs().max()) if same.any() else None)
    if gp.shape == cp.shape:
        print(f"identical best captions: {int((gp == cp).all(1).sum())}/{b}")
    print("sorted beam score diff (all clips) max:", float((gml.sort(1).values - cml.sort(1).values).abs().max()))
print("graph best[0]:", gp[0].tolist())
print("clust best[0]:", cp[0].tolist())
