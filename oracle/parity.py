"""Margin-classified comparison of a CUDA caption run with the CPU oracle.  TEST INFRASTRUCTURE ONLY (imported by tests/,
__graft_entry__.smoke() and the cpu_baseline leg of bench.py -- never by the product package).

`north_star`: "greedy token ids bit-exact, beam outputs identical except on documented score ties".  A "documented tie" is
defined by a margin: reference beam.py:256-257 picks ``topk`` of cumulative log-probabilities, and ``oracle.restate.beam_search``
records, per clip and step, the smallest score gap that decided that step's selection (rank order of the kept candidates
and the k / k+1 cut).  A clip whose smallest gap over the whole decode is >= eps is FIRM: the CUDA path must reproduce all
of its beams bit-for-bit.  Clips below eps are near-ties and are reported, not asserted.

eps must dominate the score error of the path under test: two candidates' cumulative scores each carry at most
``score_err``, so a selection can only flip when the oracle gap is < 2 * score_err.  ``compare`` measures score_err on the
clips whose beams agree (|sum log-prob difference|, i.e. avg-lprob difference x length) so that a test can assert
``2 * score_err <= eps`` next to the identity claim.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import Tensor

from . import restate


def oracle_run(sd: Dict[str, Tensor], wav: Tensor, x_lens: Optional[Tensor], bos_ids: Tensor, beam: int, min_len: int,
               max_len: int, forbid: Optional[Tensor], frame_embs: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """CPU fp32 oracle chain (restate.encoder -> project -> beam_search) with per-clip selection margins.

    ``frame_embs`` (B, T', 768), if given, replaces the oracle encoder's output (decoder-isolated comparison)."""
    enc = restate.encoder(sd, wav, x_lens)
    fe = enc["frame_embs"].transpose(1, 2) if frame_embs is None else frame_embs
    trace: List[dict] = []
    preds, lprobs, mult_preds, mult_lprobs = restate.beam_search(sd, restate.project(sd, fe), enc["frame_embs_lens"], bos_ids,
                                                                 beam, min_len, max_len, forbid, trace=trace)
    b = wav.shape[0]
    margin = torch.full((b,), float("inf"))
    for tr in trace:
        for j, mg in tr.get("margin", {}).items():
            margin[j] = min(float(margin[j]), mg)
    return {"preds": preds, "lprobs": lprobs, "mult_preds": mult_preds, "mult_lprobs": mult_lprobs, "margin": margin,
            "frame_embs": enc["frame_embs"].transpose(1, 2), "lens": enc["frame_embs_lens"],
            "logits": [tr["logits"] for tr in trace], "live": [tr["live"] for tr in trace], "toks": [tr["toks"] for tr in trace]}


def _seq_lens(mult_preds: Tensor) -> Tensor:
    """Tokens per beam that entered its average log-prob (everything up to and including the first EOS, or all)."""
    eos = mult_preds == restate.EOS_ID
    full = torch.full(mult_preds.shape[:2], mult_preds.shape[2], dtype=torch.long)
    return torch.where(eos.any(2), eos.long().argmax(2) + 1, full)


def compare(ours_mult_preds: Tensor, ours_mult_lprobs: Tensor, ref: Dict[str, Tensor], eps: float) -> Dict[str, object]:
    """ours: (B, k, L') i64 / (B, k) f32 (L' >= the oracle's pred_size is allowed: extra columns must be padding)."""
    rp, rl, margin = ref["mult_preds"], ref["mult_lprobs"], ref["margin"]
    b, k, size = rp.shape
    op = ours_mult_preds.cpu()[:, :, :size]
    extra = ours_mult_preds.cpu()[:, :, size:]
    ol = ours_mult_lprobs.cpu()
    same = (op == rp).flatten(1).all(1) & (extra == 0).flatten(1).all(1)
    firm = margin >= eps
    lens = _seq_lens(rp).float()
    score_err = ((ol - rl).abs() * lens)[same]
    best_same = (op[torch.arange(b), ol.argmax(1)] == rp[torch.arange(b), rl.argmax(1)]).all(1)
    return {
        "clips": int(b), "beam": int(k), "eps": float(eps),
        "identical": int(same.sum()), "firm": int(firm.sum()),
        "identical_where_margin_ge_eps": int((same & firm).sum()),
        "best_caption_identical": int(best_same.sum()),
        "max_lprob_err": float((ol - rl).abs()[same].max()) if bool(same.any()) else None,
        "max_score_err": float(score_err.max()) if score_err.numel() else None,
        "min_margin": float(margin.min()),
        "mismatched_firm_clips": [int(i) for i in torch.nonzero(firm & ~same).flatten()],
        "margins_of_mismatches": [float(margin[i]) for i in torch.nonzero(~same).flatten()],
    }
