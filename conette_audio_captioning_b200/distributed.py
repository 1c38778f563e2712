"""Clip-sharded multi-GPU captioning: one process per GPU (``torchrun``), contiguous split of the batch, ONE collective.

The path is embarrassingly data-parallel: eval-mode BatchNorm, per-sample LayerNorm / attention and per-clip
beam search mean no cross-clip arithmetic.  Two batch-level couplings of the reference are preserved so that the gathered
result equals the single-process call on the whole batch:
  * every shard is padded to the GLOBAL maximum length (host-known, no collective): ``frame_embs_lens`` is
    ``round(len / (Nmax // T'))`` (reference convnext.py:312-315) and the zero-padded tail is encoded too (Appendix F.3);
  * output trimming uses the global ``pred_size`` / longest best caption (reference beam.py:205-225), applied after the gather.
The only collective is an ``all_gather`` of the fixed-size id / score buffers (<= ~120 KB at B = 1024) over NCCL (gloo in the
CPU tests); nothing is fused with it because nothing follows it on the device.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def trim_like_reference(preds: Tensor, mult_preds: Tensor, pred_size: int, best_first_eos: Tensor):
    """beam.py:205-225: mult_preds -> [:, :, :pred_size]; preds -> [:, :max(first EOS index or pred_size) + 1]."""
    best_len = int(best_first_eos.clamp(max=pred_size).max()) + 1
    return preds[:, : min(best_len, pred_size)].contiguous(), mult_preds[:, :, :pred_size].contiguous()


def _pack(preds: Tensor, lprobs: Tensor, mult_preds: Tensor, mult_lprobs: Tensor, first_eos: Tensor, pred_size: int, cap: int,
          beam: int, max_len: int) -> Tensor:
    """One int64 slab per rank: word 0 = pred_size, then per clip [preds (max_len) | mult_preds (beam*max_len) | first-EOS index |
    lprob bits | mult_lprob bits (beam)]; float32 values travel as their bit patterns, unused clip slots are zero."""
    per = max_len + beam * max_len + 2 + beam
    slab = torch.zeros(1 + cap * per, dtype=torch.int64)
    slab[0] = pred_size
    b = preds.shape[0]
    if b:
        body = torch.cat([
            preds.cpu().reshape(b, max_len), mult_preds.cpu().reshape(b, beam * max_len), first_eos.cpu().reshape(b, 1).long(),
            lprobs.cpu().float().reshape(b, 1).view(torch.int32).long(),
            mult_lprobs.cpu().float().reshape(b, beam).view(torch.int32).long()], dim=1)
        slab[1 : 1 + b * per] = body.reshape(-1)
    return slab


def _unpack(slab: Tensor, n: int, beam: int, max_len: int):
    per = max_len + beam * max_len + 2 + beam
    body = slab[1 : 1 + n * per].reshape(n, per)
    o = 0
    preds = body[:, o : o + max_len]; o += max_len
    mult_preds = body[:, o : o + beam * max_len].reshape(n, beam, max_len); o += beam * max_len
    first_eos = body[:, o]; o += 1
    lprobs = body[:, o].to(torch.int32).view(torch.float32); o += 1
    mult_lprobs = body[:, o : o + beam].to(torch.int32).contiguous().view(torch.float32)
    return int(slab[0]), preds, lprobs, mult_preds, mult_lprobs, first_eos


def caption_sharded(
    run_shard: Callable[[Tensor, Tensor, Tensor], Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]],
    wav: Tensor,
    x_lens: Tensor,
    bos_ids: Tensor,
    beam: int,
    max_len: int,
    group: Optional[dist.ProcessGroup] = None,
    device: Optional[torch.device] = None,
):
    """Run ``run_shard(wav_shard, x_lens_shard, bos_shard) -> (preds (b,max_len) i64, lprobs (b,), mult_preds (b,beam,max_len),
    mult_lprobs (b,beam), info (2+b) i32 = [pred_size, _, first-EOS index per clip])`` on this rank's contiguous slice of the
    globally padded batch and gather the untrimmed buffers from every rank with ONE ``all_gather_into_tensor`` of a packed int64
    slab.  Returns the reference-shaped 4-tuple (host tensors, on every rank).  A rank whose slice is empty (fewer clips than
    ranks) skips ``run_shard`` and contributes an empty slab, so the collective never hangs.

    ``wav`` is the full (B, Nmax) batch (already right-zero-padded to the global max) or this rank's slice of it; pass the full
    ``x_lens`` / ``bos_ids`` (B,) either way so the split is reproducible.  ``device``: where the collective runs (the rank's GPU
    under NCCL, the host under gloo).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_total = x_lens.shape[0]
    lo, hi = shard_bounds(n_total, rank, world)
    wav_shard = wav[lo:hi] if wav.shape[0] == n_total else wav
    assert wav_shard.shape[0] == hi - lo, "wav must be the full batch or exactly this rank's slice"
    if hi > lo:
        preds, lprobs, mult_preds, mult_lprobs, info = run_shard(wav_shard, x_lens[lo:hi], bos_ids[lo:hi])
        info = info.cpu()
        pred_size, first_eos = int(info[0]), info[2:].long()
    else:
        preds = torch.zeros(0, max_len, dtype=torch.int64)
        mult_preds = torch.zeros(0, beam, max_len, dtype=torch.int64)
        lprobs, mult_lprobs = torch.zeros(0), torch.zeros(0, beam)
        pred_size, first_eos = 0, torch.zeros(0, dtype=torch.int64)
    if world == 1:
        p, mp = trim_like_reference(preds.cpu(), mult_preds.cpu(), pred_size, first_eos)
        return p, lprobs.cpu(), mp, mult_lprobs.cpu()
    cap = -(-n_total // world)  # fixed-size per-rank slabs (ranks may own one clip more or less)
    dev = device or (preds.device if hi > lo else torch.device("cpu"))
    mine = _pack(preds, lprobs, mult_preds, mult_lprobs, first_eos, pred_size, cap, beam, max_len).to(dev)
    everyone = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everyone, mine, group=group)  # the only collective of the path
    everyone = everyone.cpu().reshape(world, -1)
    parts = [_unpack(everyone[r], b - a, beam, max_len) for r, (a, b) in enumerate(shard_bounds(n_total, r, world) for r in range(world))]
    pred_size = max(p[0] for p in parts)  # a rank that finished early has only pad ids beyond its own pred_size
    cat = [torch.cat([p[k] for p in parts]) for k in range(1, 6)]
    p, mp = trim_like_reference(cat[0], cat[2], pred_size, cat[4])
    return p, cat[1], mp, cat[3]


def engine_shard_runner(engine, forbid_mask: Optional[Tensor], beam: int = 3, min_len: int = 3, max_len: int = 20):
    """``run_shard`` for ``caption_sharded`` on top of a real ``Engine``: waveform shard -> untrimmed fixed-size buffers through
    ``cnb_caption`` (the shard is used in place when it is already on the engine's device)."""

    def run(wav_shard: Tensor, x_lens_shard: Optional[Tensor], bos_shard: Tensor):
        outs = engine.caption(wav_shard, x_lens_shard, bos_shard, forbid_mask, beam, min_len, max_len, with_tags=False, trim=False)
        return outs[0], outs[1], outs[2], outs[3], outs[4]

    return run


def global_pad(clips: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """Right-zero-pad mono clips (N_i,) to the global maximum -> (B, Nmax) f32, lens (B,) i64 (reference pad.py:11-17)."""
    lens = torch.tensor([c.shape[-1] for c in clips], dtype=torch.int64)
    out = torch.zeros(len(clips), int(lens.max()), dtype=torch.float32)
    for i, c in enumerate(clips):
        out[i, : c.shape[-1]] = c
    return out, lens
