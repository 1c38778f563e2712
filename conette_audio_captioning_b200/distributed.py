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
    globally padded batch and gather the untrimmed buffers from every rank.  Returns the reference-shaped 4-tuple (on every rank).

    ``wav`` is the full (B, Nmax) batch (already right-zero-padded to the global max) or this rank's slice of it; pass the full
    ``x_lens`` / ``bos_ids`` (B,) either way so the split is reproducible.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_total = x_lens.shape[0]
    lo, hi = shard_bounds(n_total, rank, world)
    wav_shard = wav[lo:hi] if wav.shape[0] == n_total else wav
    assert wav_shard.shape[0] == hi - lo, "wav must be the full batch or exactly this rank's slice"
    preds, lprobs, mult_preds, mult_lprobs, info = run_shard(wav_shard, x_lens[lo:hi], bos_ids[lo:hi])
    if world == 1:
        p, mp = trim_like_reference(preds, mult_preds, int(info[0]), info[2:].long())
        return p, lprobs, mp, mult_lprobs
    dev = device or preds.device
    # fixed-size per-rank slabs (ranks may own one clip more or less): pad to the largest shard
    cap = -(-n_total // world)
    def slab(t: Tensor, fill=0) -> Tensor:
        out = torch.full((cap,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)
        out[: t.shape[0]] = t.to(dev)
        return out
    first_eos = info[2:].to(torch.int32)
    meta = torch.tensor([int(info[0])], dtype=torch.int32, device=dev)
    parts = [slab(preds), slab(lprobs), slab(mult_preds), slab(mult_lprobs), slab(first_eos)]
    gathered = []
    for t in parts:
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t, group=group)
        gathered.append(buf)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    cat = [torch.cat([g[r][: b - a] for r, (a, b) in enumerate(sizes)]) for g in gathered]
    pred_size = max(int(m[0]) for m in metas)  # a rank that finished early has only pad ids beyond its own pred_size
    p, mp = trim_like_reference(cat[0], cat[2], pred_size, cat[4].long())
    return p, cat[1], mp, cat[3]


def global_pad(clips: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """Right-zero-pad mono clips (N_i,) to the global maximum -> (B, Nmax) f32, lens (B,) i64 (reference pad.py:11-17)."""
    lens = torch.tensor([c.shape[-1] for c in clips], dtype=torch.int64)
    out = torch.zeros(len(clips), int(lens.max()), dtype=torch.float32)
    for i, c in enumerate(clips):
        out[i, : c.shape[-1]] = c
    return out, lens
