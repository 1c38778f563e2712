#!/usr/bin/env python
"""Time the tcgen05 GEMM on ConvNeXt-shaped problems with different epilogues (run under ncu -k regex:gemm_tc)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

sd = synth.make_state_dict(seed=1234, n_words=300)
eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], precision="fast")
g = torch.Generator().manual_seed(0)
for (m, n, k) in [(903168, 384, 96), (225792, 768, 192), (56448, 1536, 384), (903168, 96, 384)]:
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / k**0.5).cuda()
    bias = torch.randn(n, generator=g).cuda()
    scale = torch.rand(n, generator=g).cuda()
    resid = torch.randn(m, n, generator=g).cuda()
    for epi, bf16 in ((0, True), (1, True), (0, False), (3, False)):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.debug_gemm(a, w, bias, scale, resid, epi=epi, use_tc=True, out_bf16=bf16)
        torch.cuda.synchronize()
        print(f"M={m} N={n} K={k} epi={epi} bf16_out={bf16}: ok", flush=True)
    del a, w, resid
print("done")
