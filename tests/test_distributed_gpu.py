"""The real Engine behind distributed.caption_sharded, one process per rank: G = 2 ranks return exactly what G = 1 returns on the
whole batch (ids, scores, shapes) -- SURVEY.md 7.3 "Multi-GPU" row, BASELINE configs[3]/[4] at a reduced batch: a mixed-length
(1-30 s) batch padded to the GLOBAL maximum with x_lens, task=audiocaps, beam 5.  With two or more GPUs the ranks use one GPU each
and gather over NCCL; on a single-GPU box both ranks share cuda:0 and gather over gloo (NCCL refuses two ranks per device)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conette_audio_captioning_b200 import synth

pytestmark = pytest.mark.gpu

N = 960000
SECS = [1, 30, 9, 17, 4, 22, 13]  # 7 clips over 2 ranks: shards of 4 and 3
BEAM, MAX_LEN = 5, 20


def _batch(sd):
    wav = synth.make_audio(len(SECS), N, seed=41)[:, 0].contiguous()
    x_lens = torch.tensor([s * 32000 for s in SECS])
    for i, ln in enumerate(x_lens.tolist()):
        wav[i, ln:] = 0
    bos = sd["model.task_id_to_token_id"][torch.full((len(SECS),), synth.TASK_NAMES.index("audiocaps"))]
    return wav, x_lens, bos


def _worker(rank, world, port, n_gpus, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from conette_audio_captioning_b200.distributed import caption_sharded, engine_shard_runner
    from conette_audio_captioning_b200.engine import Engine

    dev_idx = rank if n_gpus >= world else 0
    torch.cuda.set_device(dev_idx)
    use_nccl = n_gpus >= world
    dist.init_process_group("nccl" if use_nccl else "gloo", rank=rank, world_size=world)
    try:
        sd = synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
        wav, x_lens, bos = _batch(sd)
        eng = Engine(sd, sd["model.decoder.classifier.weight"].shape[0], device=dev_idx, precision="fast", enc_chunk=4)
        run = engine_shard_runner(eng, sd["model.forbid_rep_mask"], BEAM, 3, MAX_LEN)
        out = caption_sharded(run, wav, x_lens, bos, BEAM, MAX_LEN,
                              device=torch.device("cuda", dev_idx) if use_nccl else torch.device("cpu"))
        if rank == 0:
            whole = eng.caption(wav, x_lens, bos, sd["model.forbid_rep_mask"], BEAM, 3, MAX_LEN, with_tags=False)[:4]
            ok = all(a.shape == b.shape and torch.equal(a.cpu(), b.cpu()) for a, b in zip(out, whole))
            ret.put((bool(ok), [tuple(t.shape) for t in out]))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_the_gpu():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    n_gpus = torch.cuda.device_count()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_gpus, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    ok, shapes = ret.get(timeout=10)
    print(f"[distributed] 2 ranks on {min(n_gpus, 2)} GPU(s): gathered shapes {shapes}")
    assert ok
