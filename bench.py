#!/usr/bin/env python
"""Benchmark of the CoNeTTE inference hot path (BASELINE.json metric: captioned audio-seconds per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the whole path (waveform -> log-mel -> ConvNeXt-Tiny -> projection -> beam search -> token ids) over one
batch of synthetic clips.  `--config` selects the BASELINE.json workload (random-init weights of the checkpoint architecture,
V = 4018, seed 1234):
  1 (default) configs[1]: 64 x 10 s clips PER GPU, task=clotho, beam 3 -- weak scaling, ids all-gathered over NCCL every step
  2           configs[2]: encoder only (log-mel + ConvNeXt-Tiny frame embeddings), 512 x 10 s clips per GPU (roofline study)
  3           configs[3]: 1024 x 30 s clips IN TOTAL, beam 3, sharded over the N GPUs (strong scaling) through
              conette_audio_captioning_b200.distributed.caption_sharded (global padding, one packed all_gather)
  4           configs[4]: 256 mixed-length clips (1-30 s, zero-padded to 30 s, x_lens), task=audiocaps, beam 5, sharded likewise

Reported on one JSON line (rank 0):
  value     device-resident throughput: inputs already in HBM, K steps, CUDA-event timed, max over ranks (config 1: through the
            streaming API, two batches in flight: batch i decodes on a high-priority stream while batch i+1 is encoded)
  value_sequential  (config 1) the same K steps as one blocking cnb_caption call after the other (per-kernel times add up to this)
  e2e       the same through the host-buffer API (H2D of every step's waveforms + D2H of its results inside the timed region)
  roofline  dominant kernel class: algorithmic FLOPs or bytes / CUDA-event time, against MEASURED_PEAKS.json; `traffic` = DRAM
            bytes per launch from the ncu capture recorded in profiles/ncu_traffic.json (null when none matches the shape)
  kernels   the same for every kernel class (share of the step, achieved, fraction of peak)
  parity    (config 1, N = 1) the benchmarked path against the reference on the cpu_baseline sample: ids / scores of the same
            clips, classified by the oracle's selection margin (oracle/parity.py)
  cpu_baseline  the reference's own CPU code (baseline/_ref, kind "reference") or the oracle port, on a bounded sample
`--impl reference` times that CPU path alone and prints the same line shape with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "audio-s/s"
SR = 32000
DIMS, DEPTHS, WIDTHS = (96, 192, 384, 768), (3, 3, 9, 3), (56, 28, 14, 7)
MAX_LEN, MIN_LEN = 20, 3

# BASELINE.json configs: name, clips (per GPU for weak scaling / in total for strong), seconds, beam, task, scaling
CONFIGS = {
    1: dict(name="configs[1]", batch=64, seconds=10.0, beam=3, task="clotho", scaling="weak", mixed=False, encoder_only=False),
    2: dict(name="configs[2]", batch=512, seconds=10.0, beam=3, task="clotho", scaling="weak", mixed=False, encoder_only=True),
    3: dict(name="configs[3]", batch=1024, seconds=30.0, beam=3, task="clotho", scaling="strong", mixed=False, encoder_only=False),
    4: dict(name="configs[4]", batch=256, seconds=30.0, beam=5, task="audiocaps", scaling="strong", mixed=True, encoder_only=False),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4], help="BASELINE.json configs[i]")
    ap.add_argument("--encoder-only", action="store_true", help="alias of --config 2")
    ap.add_argument("--batch", type=int, default=0, help="override the config's clip count")
    ap.add_argument("--seconds", type=float, default=0.0, help="override the config's clip length")
    ap.add_argument("--beam", type=int, default=0, help="override the config's beam size")
    ap.add_argument("--vocab-words", type=int, default=4000, help="synthetic vocabulary words (V = words + 18); 8174 -> V = 8192")
    ap.add_argument("--precision", default="fast", choices=["fast", "parity"])
    ap.add_argument("--enc-chunk", type=int, default=0)
    ap.add_argument("--decoder", default="auto", choices=["auto", "cluster", "graph", "eager"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.encoder_only:
        args.config = 2
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    if args.seconds:
        cfg["seconds"] = args.seconds
    if args.beam:
        cfg["beam"] = args.beam
    args.cfg = cfg
    return args


def metric_name(cfg):
    if cfg["encoder_only"]:
        return f"encoded audio-seconds per second ({cfg['seconds']:g} s clips, log-mel + ConvNeXt-Tiny frame embeddings)"
    return f"captioned audio-seconds per second ({cfg['seconds']:g} s clips, beam {cfg['beam']})"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


def workload_config(args, world, vocab):
    cfg = args.cfg
    per_gpu = cfg["batch"] if cfg["scaling"] == "weak" else -(-cfg["batch"] // world)
    what = "encoder only (frame embeddings)" if cfg["encoder_only"] else f"task={cfg['task']}, beam {cfg['beam']}, min 3 / max 20 tokens"
    lens = "lengths 1-30 s zero-padded to 30 s (x_lens)" if cfg["mixed"] else f"{cfg['seconds']:g} s"
    total = cfg["batch"] * (world if cfg["scaling"] == "weak" else 1)
    return {
        "workload": f"{cfg['name']}: {total} synthetic clips @ 32 kHz ({per_gpu} per GPU), {lens}, {what}, V={vocab}, "
                    f"random-init weights (seed 1234)",
        "clips_per_gpu": per_gpu, "clips_total": total, "n_samples": int(cfg["seconds"] * SR), "beam": cfg["beam"],
        "precision": args.precision,
        "parallelism": f"clip-sharded x{world} (no data-path collective; one all_gather of the ids per step)",
        "l2": "inputs alternate between two batches / exceed the 126 MB L2, and GBs of activations stream through L2 per step",
    }


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic work per kernel class for one step (DESIGN.md "Algorithmic work")
# ---------------------------------------------------------------------------------------------------------------------
def algorithmic_work(batch: int, n_samples: int, act_bytes: int = 2, fused=(0, 1, 2)):
    t = n_samples // 320 + 1
    h1 = (t + 4) // 4 + 1
    hs = [h1, h1 // 2, h1 // 4, h1 // 8]
    work = {}
    work["frontend"] = ("hbm", batch * (4 * n_samples + 4 * t * 224))
    work["stem"] = ("hbm", batch * (4 * t * 224 + 4 * hs[0] * 56 * 96))
    ds_f = pack_b = 0
    for s in range(4):
        m = batch * hs[s] * WIDTHS[s]
        c = DIMS[s]
        # dw+LN reads the fp32 residual stream and writes the GEMM operand; flops 98 per element (informational)
        work[f"dwconv_ln.s{s + 1}"] = ("hbm", DEPTHS[s] * m * c * (4 + act_bytes))
        pw = DEPTHS[s] * 2 * m * c * 4 * c
        # stages 1-3, fast precision: ONE fused kernel does pw1 + GELU + pw2 and is bracketed under the pw1 class
        work[f"gemm_pw1_gelu.s{s + 1}"] = ("tensor", 2 * pw if s in fused else pw)
        work[f"gemm_pw2_resid.s{s + 1}"] = ("tensor", pw)
        if s > 0:
            cin = DIMS[s - 1]
            ds_f += 2 * m * 4 * cin * c
            pack_b += m * 4 * cin * (4 + act_bytes)
    work["ds_gemm"] = ("tensor", ds_f)
    work["ds_ln_pack"] = ("hbm", pack_b)
    return work


def secondary_bounds(batch: int, n_samples: int, act_bytes: int = 2, fused=(0, 1, 2)):
    """The other roofline of the classes whose nominal bound (north_star) is not the physical one:
    pointwise GEMMs -> minimum HBM bytes (operand in, result out, fp32 residual in+out, weights once);
    depthwise conv + LN -> FP32 FMAs (49 per output element) against the CUDA-core peak."""
    t = n_samples // 320 + 1
    h1 = (t + 4) // 4 + 1
    hs = [h1, h1 // 2, h1 // 4, h1 // 8]
    out = {}
    for s in range(4):
        m = batch * hs[s] * WIDTHS[s]
        c = DIMS[s]
        if s in fused:  # fused MLP: operand in, fp32 residual in + out, weights once
            out[f"gemm_pw1_gelu.s{s + 1}"] = ("hbm", DEPTHS[s] * (m * c * act_bytes + 2 * m * c * 4 + 8 * c * c * 2))
        else:
            out[f"gemm_pw1_gelu.s{s + 1}"] = ("hbm", DEPTHS[s] * (m * c * act_bytes + m * 4 * c * act_bytes + 4 * c * c * 2))
        out[f"gemm_pw2_resid.s{s + 1}"] = ("hbm", DEPTHS[s] * (m * 4 * c * act_bytes + 2 * m * c * 4 + 4 * c * c * 2))
        out[f"dwconv_ln.s{s + 1}"] = ("fp32", DEPTHS[s] * m * c * 49 * 2)
    return out


def decoder_work(batch: int, beam: int, steps: int, vocab: int, tp: int):
    """Decode = `steps` dependent passes over R = batch x beam rows.  Algorithmic bytes: every pass has to read the decoder's
    weights once (6 layers x 1.44 M + 256 V parameters at 4 B: fp32, or the fp16 hi/lo pair of the cluster decoder; they stay in
    L2, so this is an L2->SM stream, bounded above by the HBM figure) plus the cross-attention K|V of every clip; algorithmic
    flops: SURVEY.md 8d (19.9 MFLOP / row / step)."""
    params = 6 * (3 * 256 * 256 + 3 * 256 * 256 + 2 * 256 * 2048) + 256 * vocab
    byts = steps * (4 * params + batch * tp * 6 * 512 * 4)
    flops = steps * batch * beam * 2 * params
    return byts, flops


def ncu_traffic(kernel_class: str, batch: int, n_samples: int, vocab: int = 4018):
    """DRAM bytes per launch of a kernel class from the committed ncu capture summary (profiles/ncu_traffic.json), when one
    exists for exactly this shape; None otherwise (never a guess)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    rec = json.load(open(path)).get(f"{batch}x{n_samples}", {}).get(kernel_class)
    if rec is None or rec.get("vocab", vocab) != vocab:
        return None
    return rec["dram_bytes_per_launch"]


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        mine = [r for r in rows if len(r) >= 8 and r[0].strip() == str(self.gpu)]
        if not mine:
            return out
        sm = [float(r[1]) for r in mine]
        busy = [x for x in sm if x > 0.5 * max(sm)] or sm
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in mine for n, v in zip(names, r[4:8]) if v.strip() == "Active"})
        out.update(sm_mhz=statistics.median(busy), sm_max_mhz=float(mine[0][2]), reasons=reasons, samples=len(mine),
                   power_w_max=max(float(r[3]) for r in mine))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------------------
def make_batch(cfg, n_clips: int, seed: int):
    """(wav (B, N) f32 pinned host, x_lens (B,) i64 or None, audio seconds) of one batch of the config."""
    import torch

    from conette_audio_captioning_b200 import synth

    n = int(cfg["seconds"] * SR)
    wav = synth.make_audio(n_clips, n, seed=seed)[:, 0].contiguous()
    x_lens = None
    secs = n_clips * n / SR
    if cfg["mixed"]:  # durations randint(1, 31) s, seed 1234, right-zero-padded to 30 s
        dur = torch.randint(1, 31, (n_clips,), generator=torch.Generator().manual_seed(1234))
        x_lens = dur * SR
        for i, ln in enumerate(x_lens.tolist()):
            wav[i, ln:] = 0
        secs = float(dur.sum())
    return wav, x_lens, secs


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own code when importable (baseline/_ref or /root/reference), else the oracle port
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_throughput(args, n_clips: int, steps: int, warmup: int):
    """Times the reference (or the oracle port) on the first `n_clips` clips of the workload's first batch; returns the baseline
    record, the per-call times and the outputs of the last call (kept for the parity record)."""
    import torch

    from conette_audio_captioning_b200 import synth
    from oracle import ref_loader, restate

    cfg = args.cfg
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sd = synth.make_state_dict(seed=1234, n_words=args.vocab_words)
    full = cfg["batch"] if cfg["scaling"] == "weak" else min(cfg["batch"], 64)
    wav_all, x_lens_all, _ = make_batch(cfg, full, 1234)
    wav = wav_all[:n_clips].contiguous()
    x_lens = None if x_lens_all is None else x_lens_all[:n_clips]
    secs = n_clips * wav.shape[1] / SR if x_lens is None else float(x_lens.sum()) / SR
    task_idx = synth.TASK_NAMES.index(cfg["task"])
    bos = sd["model.task_id_to_token_id"][torch.full((n_clips,), task_idx)]
    if cfg["encoder_only"]:
        kind = "port"
        if ref_loader.available():
            kind = "reference"
            model = ref_loader.build_reference_model(sd, synth.make_corpus(args.vocab_words))

            def run():
                with torch.no_grad():
                    return model.preprocessor(wav[:, None, :], SR)
        else:
            def run():
                with torch.no_grad():
                    return restate.encoder(sd, wav, None)
    elif ref_loader.available():
        kind = "reference"
        model = ref_loader.build_reference_model(sd, synth.make_corpus(args.vocab_words))
        xs = None if x_lens is None else x_lens[:, None]

        def run():
            with torch.no_grad():
                return model(wav[:, None, :], sr=SR, x_shapes=xs, task=cfg["task"], beam_size=cfg["beam"])
    else:
        kind = "port"

        def run():
            with torch.no_grad():
                return restate.caption(sd, wav, x_lens, bos, cfg["beam"], MIN_LEN, MAX_LEN, sd["model.forbid_rep_mask"])
    out = None
    for _ in range(warmup):
        out = run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = run()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {
        "value": secs / best, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{n_clips} of the workload's clips ({secs:g} audio-s) per call, {warmup} warm-up + best of {steps} calls, "
                  f"fp32 CPU, torch threads={cores}",
        "s_per_call": best,
    }, times, out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from conette_audio_captioning_b200 import synth

    base, times, _ = cpu_reference_throughput(args, args.cpu_sample, max(1, args.steps), max(1, min(args.warmup, 2)))
    ms = 1e3 * statistics.mean(times)
    vocab = args.vocab_words + 18
    cfg = workload_config(args, world, vocab)
    cfg["reference_batch_per_call"] = args.cpu_sample  # the CPU arm times a bounded sample of the workload (throughput-normalised)
    line = {
        "impl": "reference", "metric": metric_name(args.cfg), "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": args.cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": base, "gpu_launches": 0,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def kernel_table(prof, steps, work, second, pk, clocks):
    total_ms = sum(ms for ms, _ in prof.values()) or 1.0
    kernels = {}
    for name, (ms, cnt) in prof.items():
        ent = {"ms_per_step": ms / steps, "share": ms / total_ms, "brackets_per_step": cnt / steps}
        if name in work and ms > 0:
            bound, amount = work[name]
            rate = amount / (ms / steps * 1e-3)
            if bound == "tensor":
                ent.update(bound="tensor", achieved=rate / 1e12, unit="TFLOP/s", frac=rate / 1e12 / pk["bf16_tflops"])
            else:
                ent.update(bound="hbm", achieved=rate / 1e9, unit="GB/s", frac=rate / 1e9 / pk["hbm_gbs"])
        kernels[name] = ent
    kernels = {k: v for k, v in kernels.items() if v["brackets_per_step"] > 0}
    # the physical bound next to the nominal one (HBM for the pointwise GEMMs, FP32 FMA pipe for the depthwise conv)
    fp32_peak_tflops = 148 * 128 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12
    for name, (kind, amount) in second.items():
        if name in kernels and kernels[name]["ms_per_step"] > 0:
            rate = amount / (kernels[name]["ms_per_step"] * 1e-3)
            if kind == "hbm":
                kernels[name].update(hbm_gbs=rate / 1e9, hbm_frac=rate / 1e9 / pk["hbm_gbs"])
            else:
                kernels[name].update(fp32_tflops=rate / 1e12, fp32_frac=rate / 1e12 / fp32_peak_tflops)
    return kernels, total_ms / steps


def roofline_of(kernels, work, pk, b, n, vocab):
    dec_parts = ("dec_gemm", "dec_attn_ln", "dec_classifier", "beam")
    classified = [k for k in kernels if "bound" in kernels[k] and k not in dec_parts]
    top = max(classified, key=lambda k: kernels[k]["ms_per_step"])
    kt = kernels[top]
    launches_per_step = kt["brackets_per_step"]
    roofline = {"kernel": top, "bound": kt["bound"], "achieved": kt["achieved"],
                "peak": pk["bf16_tflops"] if kt["bound"] == "tensor" else pk["hbm_gbs"], "unit": kt["unit"],
                "frac": kt["frac"], "traffic": ncu_traffic(top, b, n, vocab),
                "algorithmic_per_launch": work[top][1] / launches_per_step, "launches_per_step": launches_per_step,
                "avg_launch_ms": kt["ms_per_step"] / launches_per_step, "peak_source": pk["source"] + " (sustained)",
                "share_of_step": kt["share"]}
    if "note" in kt:
        roofline["note"] = kt["note"]
    enc = [k for k in classified if k != "decoder"]
    if enc and top == "decoder":  # runner-up: the largest throughput-bound kernel class (what the encoder work is judged by)
        enc_top = max(enc, key=lambda k: kernels[k]["ms_per_step"])
        roofline["next"] = {"kernel": enc_top, **{k: v for k, v in kernels[enc_top].items() if k != "brackets_per_step"},
                            "avg_launch_ms": kernels[enc_top]["ms_per_step"] / kernels[enc_top]["brackets_per_step"],
                            "traffic": ncu_traffic(enc_top, b, n, vocab)}
    return roofline


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from conette_audio_captioning_b200 import _lib, synth
    from conette_audio_captioning_b200.distributed import caption_sharded, engine_shard_runner, shard_bounds
    from conette_audio_captioning_b200.engine import Engine

    if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout and would precede the JSON line
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = args.cfg
    n = int(cfg["seconds"] * SR)
    beam = cfg["beam"]
    sd = synth.make_state_dict(seed=1234, n_words=args.vocab_words)
    vocab = sd["model.decoder.classifier.weight"].shape[0]
    eng = Engine(sd, vocab, device=local_rank, precision=args.precision, enc_chunk=args.enc_chunk, decoder=args.decoder)
    forbid = sd["model.forbid_rep_mask"]
    task_idx = synth.TASK_NAMES.index(cfg["task"])
    tp = _lib.geometry(n)[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- this rank's clips -------------------------------------------------------------------------------------------------
    if cfg["scaling"] == "weak":  # every rank its own batches (two alternate so that inputs do not sit in L2)
        b = cfg["batch"]
        batches = [make_batch(cfg, b, 1234 + 2 * rank + i) for i in range(2)]
        audio_s = batches[0][2] * world
        lo, hi, b_total = 0, b, b * world
    else:  # one global batch, contiguous shards (global padding: every clip is already padded to the same 30 s)
        b_total = cfg["batch"]
        lo, hi = shard_bounds(b_total, rank, world)
        b = hi - lo
        wav_all, xl_all, audio_s = make_batch(cfg, b_total, 1234)
        batches = [(wav_all[lo:hi].contiguous(), None if xl_all is None else xl_all[lo:hi], audio_s)]
        x_lens_all = torch.full((b_total,), n, dtype=torch.int64) if xl_all is None else xl_all
        bos_all = sd["model.task_id_to_token_id"][torch.full((b_total,), task_idx)]
        del wav_all
    host_wavs = [w.pin_memory() for w, _, _ in batches]
    x_lens = [xl for _, xl, _ in batches]
    dev_wavs = [w.to(dev) for w in host_wavs]
    bos = sd["model.task_id_to_token_id"][torch.full((b,), task_idx)]
    bos_dev, forbid_dev = bos.to(dev), forbid.to(dev, torch.uint8)
    nb = len(batches)
    sampler = ClockSampler(local_rank)
    pk = peaks()
    act_bytes = 2 if args.precision == "fast" else 4
    fused = ()
    if args.precision == "fast" and os.environ.get("CNB_NO_MLP_FUSED") is None:
        fused = (0,) + ((1,) if os.environ.get("CNB_NO_MLP_FUSED192") is None else ()) + (
            (2,) if os.environ.get("CNB_NO_MLP_FUSED384") is None else ())
    work = algorithmic_work(b, n, act_bytes, fused)
    second = secondary_bounds(b, n, act_bytes, fused)
    line_extra = {}

    if cfg["encoder_only"]:
        # ---- configs[2]: waveform -> frame embeddings ------------------------------------------------------------------------
        def step_device(i):
            return eng.encoder(dev_wavs[i % nb], with_tags=False)[0]

        fe_host = torch.empty(b, tp, 768, dtype=torch.float32).pin_memory()

        def step_host(i):
            fe_host.copy_(eng.encoder(host_wavs[i % nb].to(dev, non_blocking=True), with_tags=False)[0], non_blocking=True)
            torch.cuda.synchronize()
            return fe_host

        h2d, d2h = b * n * 4, fe_host.numel() * 4
        sample_out = lambda r: {"frame_embs_mean_abs": float(r.abs().mean())}  # noqa: E731
    elif cfg["scaling"] == "weak":
        # ---- configs[1]: blocking device-resident calls (value_sequential), streaming API below --------------------------------
        host_out = eng.alloc_host_outputs(b, beam, MAX_LEN, with_tags=False)
        gathered = None
        if world > 1:
            gathered = (torch.empty(world * b, MAX_LEN, device=dev, dtype=torch.int64), torch.empty(world * b, device=dev))

        def step_device(i):
            outs = eng.caption(dev_wavs[i % nb], x_lens[i % nb], bos_dev, forbid_dev, beam, MIN_LEN, MAX_LEN, with_tags=False, trim=False)
            if world > 1:  # the only collective of the path: gather token ids + scores
                dist.all_gather_into_tensor(gathered[0], outs[0])
                dist.all_gather_into_tensor(gathered[1], outs[1])
            return outs

        def step_host(i):
            return eng.caption_host(host_wavs[i % nb], x_lens[i % nb], bos, forbid, beam, MIN_LEN, MAX_LEN, with_tags=False, out=host_out)

        h2d = b * n * 4 + b * 8 + vocab
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())
        sample_out = lambda r: {"preds0": r[0][0].tolist(), "lprob0": float(r[1][0])}  # noqa: E731
    else:
        # ---- configs[3] / [4]: one global batch sharded over the ranks through distributed.caption_sharded ---------------------
        host_out = eng.alloc_host_outputs(b, beam, MAX_LEN, with_tags=False)
        run_dev = engine_shard_runner(eng, forbid_dev, beam, MIN_LEN, MAX_LEN)

        def run_host(wav_shard, xl_shard, bos_shard):
            eng.caption_host(wav_shard, xl_shard, bos_shard, forbid, beam, MIN_LEN, MAX_LEN, with_tags=False, out=host_out)
            return host_out["preds"], host_out["lprobs"], host_out["mult_preds"], host_out["mult_lprobs"], host_out["info"]

        def step_device(i):
            return caption_sharded(run_dev, dev_wavs[0], x_lens_all, bos_all, beam, MAX_LEN, device=dev)

        def step_host(i):
            return caption_sharded(run_host, host_wavs[0], x_lens_all, bos_all, beam, MAX_LEN, device=dev)

        h2d = b * n * 4 + b * 8 + vocab
        d2h = sum(v.numel() * v.element_size() for v in host_out.values())
        sample_out = lambda r: {"preds0": r[0][0].tolist(), "lprob0": float(r[1][0]), "gathered_clips": int(r[0].shape[0])}  # noqa: E731

    # ---- device-resident, one blocking call after the other --------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        res = step_device(i)
    ev1.record()
    barrier()
    launches = eng.launch_count() - launches0
    seq_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    dev_ms = seq_ms
    streaming = cfg["scaling"] == "weak" and not cfg["encoder_only"]
    if streaming:
        # the same K steps through the streaming form of the API with DEVICE-resident inputs (two batches in flight: batch i
        # decodes on a high-priority stream while batch i+1 is encoded).  The torch stream is idle, so the two events are device
        # timestamps taken right before the first enqueue and right after the last batch has been collected.
        dev_outs = [eng.alloc_host_outputs(b, beam, MAX_LEN, with_tags=False) for _ in range(2)]

        def run_stream_dev(k):
            ticket = eng.caption_host_begin(dev_wavs[0], x_lens[0], bos, forbid, beam, MIN_LEN, MAX_LEN, with_tags=False, out=dev_outs[0])
            outs = None
            for i in range(k):
                nxt = (eng.caption_host_begin(dev_wavs[(i + 1) % nb], x_lens[(i + 1) % nb], bos, forbid, beam, MIN_LEN, MAX_LEN,
                                              with_tags=False, out=dev_outs[(i + 1) & 1]) if i + 1 < k else None)
                outs = eng.caption_host_end(ticket)
                if world > 1:
                    dist.all_gather_into_tensor(gathered[0], ticket["out"]["preds"].to(dev, non_blocking=True))
                    dist.all_gather_into_tensor(gathered[1], ticket["out"]["lprobs"].to(dev, non_blocking=True))
                ticket = nxt
            return outs

        run_stream_dev(args.warmup)
        barrier()
        launches0 = eng.launch_count()
        ev0.record()
        run_stream_dev(args.steps)
        ev1.record()
        barrier()
        launches = eng.launch_count() - launches0
        dev_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    # ---- end to end from host buffers ---------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = step_host(i)
    torch.cuda.synchronize()
    sync_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    e2e_ms, e2e_api = sync_ms, "one blocking host-buffer call per step"
    if streaming:
        # the streaming form of the host API (caption_host_begin / _end, two batches in flight): every step still copies its own
        # waveforms from pinned host memory and reads its own ids back inside the timed region
        host_outs = [host_out, eng.alloc_host_outputs(b, beam, MAX_LEN, with_tags=False)]

        def run_stream(k):
            ticket = eng.caption_host_begin(host_wavs[0], x_lens[0], bos, forbid, beam, MIN_LEN, MAX_LEN, with_tags=False, out=host_outs[0])
            r = None
            for i in range(k):
                nxt = (eng.caption_host_begin(host_wavs[(i + 1) % nb], x_lens[(i + 1) % nb], bos, forbid, beam, MIN_LEN, MAX_LEN,
                                              with_tags=False, out=host_outs[(i + 1) & 1]) if i + 1 < k else None)
                r = eng.caption_host_end(ticket)
                ticket = nxt
            return r

        run_stream(args.warmup)
        barrier()
        t0 = time.perf_counter()
        res = run_stream(args.steps)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
        e2e_api = ("cnb_caption_host_begin/_end (Engine.caption_host_begin/_end, CoNeTTEModel.stream): host buffers, two batches "
                   "in flight, every step's H2D + D2H inside the timed region")
    clocks = sampler.stop()
    # ---- per-kernel-class device time (event brackets inside the library) ----------------------------------------------------
    eng.profile_begin()
    for i in range(args.steps):
        step_device(i)
    prof = eng.profile_end()
    barrier()
    kernels, total_ms = kernel_table(prof, args.steps, work, second, pk, clocks)
    dec_ms = sum(kernels[k]["ms_per_step"] for k in ("dec_gemm", "dec_attn_ln", "dec_classifier", "beam") if k in kernels)
    if dec_ms > 0:
        pred_steps = int(res[0].shape[1]) if not cfg["encoder_only"] else MAX_LEN
        dbytes, dflops = decoder_work(b, beam, pred_steps, vocab, tp)
        rows_per_cluster = 16 if -(-b // max(1, 16 // beam)) <= 15 else 32
        n_groups = -(-b // max(1, rows_per_cluster // beam))
        kernels["decoder"] = {
            "ms_per_step": dec_ms, "share": dec_ms / total_ms,
            "brackets_per_step": sum(kernels[k]["brackets_per_step"] for k in ("dec_gemm", "dec_attn_ln", "dec_classifier", "beam") if k in kernels),
            "bound": "hbm", "achieved": dbytes / (dec_ms * 1e-3) / 1e9, "unit": "GB/s",
            "frac": dbytes / (dec_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "tensor_tflops": dflops / (dec_ms * 1e-3) / 1e12, "tensor_frac": dflops / (dec_ms * 1e-3) / 1e12 / pk["bf16_tflops"],
            "decode_steps": pred_steps, "us_per_decode_step": 1e3 * dec_ms / pred_steps,
            # what the cluster decoder actually moves from L2 into shared memory: every group of rows_per_cluster / beam clips
            # streams all decoder weights once per step (weights-as-M formulation)
            "row_groups": n_groups, "rows_per_cluster": rows_per_cluster,
            "l2_to_smem_gbs": n_groups * dbytes / (dec_ms * 1e-3) / 1e9,
            "note": "one launch; latency-bound chain of ~45 dependent phases per step (6 layers x {4 tensor-core GEMM phases, 2 attention "
                    "phases, 3 reduce-scatter + all-gather + LayerNorm exchanges} + classifier rounds + beam merge) on 16-row operands, "
                    "fp16 hi/lo split MMAs (fp32-level accuracy); reported against the weight bytes one step has to stream.  "
                    "`traffic` exceeds that figure because the weight ring loads with L2 evict-first priority: the stream is "
                    "fetched from DRAM about twice per step (0.4 TB/s) so that the latency-critical attention data stays in L2 "
                    "(with the default policy 0.85 GB per launch, once per step, but the kernel is 6 % slower: "
                    "profiles/r2_decoder_l2_policy.txt)"}
        work["decoder"] = ("hbm", dbytes)
    roofline = roofline_of(kernels, work, pk, b, n, vocab)

    if rank == 0:
        cpu = parity_rec = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _, ref_out = cpu_reference_throughput(args, args.cpu_sample, 2, 1)
            if args.config == 1 and args.precision == "fast":
                parity_rec = parity_record(args, eng, sd, dev_wavs[0], bos_dev, forbid_dev, ref_out)
        throughput = lambda ms: audio_s / (ms * 1e-3)  # noqa: E731
        line = {
            "metric": metric_name(cfg), "value": throughput(dev_ms), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "fp16-operand tcgen05 GEMMs (f32 accumulate, f32 residual stream); decoder GEMMs on fp16 hi/lo split operands "
                     "(22-bit significands, f32 accumulate); f32 front-end, depthwise conv, LayerNorm, softmax, beam scores"
                     if args.precision == "fast" else "f32 CUDA-core GEMMs everywhere",
            "data": "synthetic", "config": workload_config(args, world, vocab),
            "e2e": {"value": throughput(e2e_ms), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": e2e_api},
            "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks,
            "sample_output": sample_out(res),
        }
        if streaming:
            line["value_sequential"] = {"value": throughput(seq_ms), "ms_per_step": seq_ms,
                                        "note": "one cnb_caption call after the other on one stream (no overlap between batches); "
                                                "the per-kernel times under `kernels` add up to this step"}
            line["overlap"] = ("value and e2e use the streaming API (cnb_caption_host_begin/_end): the latency-bound decoder of batch i "
                               "runs on a high-priority stream while batch i+1 is encoded; every step's work is inside the timed region")
            line["e2e_sync"] = {"value": throughput(sync_ms), "unit": UNIT, "ms_per_step": sync_ms,
                                "api": "cnb_caption_host (one blocking call per batch)"}
        if world > 1:
            line["note_scaling"] = ("N GPUs against the reference's ONE CPU host is not a speed-up figure; read the N-GPU values as "
                                    "scaling efficiency against the 1-GPU value of the same build")
        if parity_rec is not None:
            line["parity"] = parity_rec
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def parity_record(args, eng, sd, dev_wav, bos_dev, forbid_dev, ref_out):
    """The benchmarked path (whole batch, fast precision) against the outputs the cpu_baseline run just produced for the first
    `cpu_sample` clips (the reference's own CoNeTTEModel when importable), classified by the oracle's selection margins."""
    import torch

    from oracle import parity

    k = args.cpu_sample
    cfg = args.cfg
    eps = 5e-3  # tests/test_bench_parity.py: twice the measured end-to-end cumulative-score error (<= 2e-3) with headroom
    outs = eng.caption(dev_wav, None, bos_dev, forbid_dev, cfg["beam"], MIN_LEN, MAX_LEN, with_tags=False, trim=False)
    orc = parity.oracle_run(sd, dev_wav[:k].cpu(), None, bos_dev[:k].cpu(), cfg["beam"], MIN_LEN, MAX_LEN, sd["model.forbid_rep_mask"])
    ref = orc
    source = "oracle/restate.py (CPU fp32 restatement pinned to the reference)"
    if isinstance(ref_out, dict) and "mult_preds" in ref_out and torch.is_tensor(ref_out["mult_preds"]):
        ref = {"mult_preds": ref_out["mult_preds"], "mult_lprobs": ref_out["mult_lprobs"], "margin": orc["margin"]}
        source = "the reference's own CoNeTTEModel (baseline/_ref) on the CPU, margins from oracle/restate.py"
    rec = parity.compare(outs[2][:k], outs[3][:k], ref, eps)
    rec["against"] = source
    rec["rule"] = ("ids of every beam bit-identical on each clip whose smallest oracle selection gap is >= eps; eps = 2 x the measured "
                   "cumulative-score error of the path with headroom (tests/test_bench_parity.py asserts both halves)")
    return rec


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    from conette_audio_captioning_b200 import build

    if rank == 0:
        build.build()
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
