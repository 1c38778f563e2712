"""world_size-2 gloo test of the clip-sharding + gather logic (host side of the multi-GPU path), with the CPU oracle
standing in for the per-rank engine: gathered result == single-process result on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conette_audio_captioning_b200 import synth
from conette_audio_captioning_b200.distributed import caption_sharded, global_pad, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (1, 5, 64, 1024, 7):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_global_pad_matches_reference_padding():
    clips = [torch.ones(5), torch.ones(9), torch.ones(2)]
    wav, lens = global_pad(clips)
    assert wav.shape == (3, 9) and lens.tolist() == [5, 9, 2] and float(wav[2, 2:].abs().sum()) == 0


def _oracle_shard_runner(sd, mem_all, lens_all, beam, max_len, lo_hi):
    """Fake engine: beam-search this rank's clips with the oracle and return UNtrimmed fixed-size buffers like cnb_decode."""
    from oracle import restate

    def run(wav_shard, x_lens_shard, bos_shard):
        lo, hi = lo_hi
        b = hi - lo
        preds, lprobs, mpreds, mlprobs = restate.beam_search(sd, mem_all[lo:hi], lens_all[lo:hi], bos_shard, beam, 3, max_len,
                                                             sd["model.forbid_rep_mask"])
        pred_size = mpreds.shape[2]
        P = torch.zeros(b, max_len, dtype=torch.long)
        MP = torch.zeros(b, beam, max_len, dtype=torch.long)
        MP[:, :, :pred_size] = mpreds
        best = mlprobs.argmax(1)
        P[:, :pred_size] = mpreds[torch.arange(b), best]
        first = torch.full((b,), max_len, dtype=torch.int32)
        for i in range(b):
            hit = (P[i] == 2).nonzero()
            if len(hit):
                first[i] = int(hit[0])
        info = torch.cat([torch.tensor([pred_size, 0], dtype=torch.int32), first])
        return P, lprobs, MP, mlprobs, info

    return run


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
        g = torch.Generator().manual_seed(5)
        b, tp, beam, max_len = 5, 4, 3, 20
        mem = torch.relu(torch.randn(b, tp, 256, generator=g))
        lens = torch.randint(1, tp + 1, (b,), generator=g)
        bos = sd["model.task_id_to_token_id"][torch.randint(0, 7, (b,), generator=g)]
        wav = torch.zeros(b, 8)  # placeholder: the fake runner works from `mem`
        lo_hi = shard_bounds(b, rank, world)
        out = caption_sharded(_oracle_shard_runner(sd, mem, lens, beam, max_len, lo_hi), wav, lens, bos, beam, max_len)
        # fewer clips than ranks: rank 1's slice is empty, it must still join the (single) collective
        one = caption_sharded(_oracle_shard_runner(sd, mem[:1], lens[:1], beam, max_len, shard_bounds(1, rank, world)), wav[:1],
                              lens[:1], bos[:1], beam, max_len)
        if rank == 0:
            from oracle import restate

            ref = restate.beam_search(sd, mem, lens, bos, beam, 3, max_len, sd["model.forbid_rep_mask"])
            ok = all(torch.equal(a, r) if a.dtype == torch.long else torch.allclose(a, r, atol=1e-6) for a, r in zip(out, ref))
            ref1 = restate.beam_search(sd, mem[:1], lens[:1], bos[:1], beam, 3, max_len, sd["model.forbid_rep_mask"])
            ok1 = all(torch.equal(a, r) if a.dtype == torch.long else torch.allclose(a, r, atol=1e-6) for a, r in zip(one, ref1))
            ret.put(bool(ok) and bool(ok1) and all(a.shape == r.shape for a, r in zip(out, ref)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True
