#!/usr/bin/env python
"""Phase trace of the cluster decoder (CNB_DEC_TRACE=1): python tools/decode_trace.py [--batch 8]"""
import argparse
import os
import sys

os.environ.setdefault("CNB_DEC_TRACE", "1")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conette_audio_captioning_b200 import synth  # noqa: E402
from conette_audio_captioning_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--decoder", default="cluster")
a = ap.parse_args()
sd = synth.make_state_dict(seed=1234, n_words=4000)
V = sd["model.decoder.classifier.weight"].shape[0]
forbid = sd["model.forbid_rep_mask"].cuda().to(torch.uint8)
eng = Engine(sd, V, precision="fast", decoder=a.decoder)
if os.environ.get("DUMMY_MB"):  # experiment: shift the addresses of the workspaces the first decode allocates
    import ctypes
    _rt = ctypes.CDLL("libcudart.so.12")
    _p = ctypes.c_void_p()
    _rt.cudaMalloc(ctypes.byref(_p), ctypes.c_size_t(int(os.environ["DUMMY_MB"]) << 20))
b = a.batch
fe = torch.randn(b, 31, 768, device="cuda")
bos = sd["model.task_id_to_token_id"][torch.zeros(b, dtype=torch.long)].cuda()
for _ in range(2):
    out = eng.decode(fe, torch.full((b,), 31), bos, forbid)
torch.cuda.synchronize()
print("pred_size", out[0].shape, out[2].shape)
eng.profile_begin()
for _ in range(3):
    eng.decode(fe, torch.full((b,), 31), bos, forbid)
prof = eng.profile_end()
print("event-timed: decoder kernel %.3f ms/launch, projection + cross K/V %.3f ms/decode" % (
    prof["dec_gemm"][0] / 3, prof["proj_crosskv"][0] / 3))
