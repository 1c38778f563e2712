"""``conette-predict`` on the B200 path (reference src/conette/predict.py:27-233) and the ``conette()`` factory
(reference src/conette/__init__.py:25-49, hubconf.py:7)  [SURVEY.md 8f rank 2].

Same command line (``--audio``, ``--task``, ``--model_name``, ``--model_path``, ``--device``, ``--csv_export``, ``--seed``,
``--verbose``; ``--token`` is accepted and unused: there is no network here) and the same CSV columns.  ``--model_name`` /
``--model_path`` name a local weight file or a Hugging Face style directory holding a ``CoNeTTEModel`` state dict
(``checkpoint.load_checkpoint``); downloading ``Labbeti/conette`` from the hub is the reference's job, not this library's.

    python -m conette_audio_captioning_b200.predict --audio a.wav b.wav --task clotho --model_path /ckpt/conette --csv_export out.csv
"""
from __future__ import annotations

import csv
import logging
import os.path as osp
import random
from argparse import ArgumentParser, Namespace
from typing import Any, Dict, List, Optional, Sequence

import torch

pylog = logging.getLogger("conette")
DEFAULT_MODEL_NAME = "Labbeti/conette"


def conette(pretrained_model_name_or_path: Optional[str] = DEFAULT_MODEL_NAME, config_kwds: Optional[Dict[str, Any]] = None,
            model_kwds: Optional[Dict[str, Any]] = None):
    """Create a CoNeTTEModel for inference from a local checkpoint (reference ``conette.conette``)."""
    from .checkpoint import load_checkpoint
    from .config import CoNeTTEConfig
    from .model import CoNeTTEModel

    if pretrained_model_name_or_path is None:
        raise ValueError("conette(None): the B200 path has no randomly initialised default model; pass a checkpoint path "
                         "(synthetic weights: conette_audio_captioning_b200.synth)")
    if not osp.exists(pretrained_model_name_or_path):
        raise FileNotFoundError(
            f"'{pretrained_model_name_or_path}' is not a local file or directory (this build has no hub download: fetch the "
            "checkpoint with the reference tooling and pass its path)")
    sd, itos, config = load_checkpoint(pretrained_model_name_or_path)
    if config_kwds:
        config = CoNeTTEConfig(**{**config.__dict__, **config_kwds})
    return CoNeTTEModel(config, sd, itos, **(model_kwds or {}))


def get_predict_args(argv: Optional[Sequence[str]] = None) -> Namespace:
    parser = ArgumentParser(description="CoNeTTE predict (B200-native path).")
    parser.add_argument("--audio", type=str, nargs="+", default=(), help="Audio file path(s).")
    parser.add_argument("--task", type=str, nargs="+", default=None, help="CoNeTTE task embedding input(s).")
    parser.add_argument("--model_name", type=str, default=DEFAULT_MODEL_NAME, help="Model name or local path.")
    parser.add_argument("--model_path", type=str, default=None, help="Local checkpoint file / directory.")
    parser.add_argument("--device", type=str, default="cuda_if_available", help="cuda device (there is no CPU path).")
    parser.add_argument("--csv_export", type=str, default=None, help="Path to the CSV output file.")
    parser.add_argument("--seed", type=int, default=1234)
    parser.add_argument("--token", type=str, default=None, help="Accepted for CLI compatibility; unused (no hub access).")
    parser.add_argument("--verbose", type=int, default=1)
    parser.add_argument("--precision", type=str, default="fast", choices=("fast", "parity"))
    return parser.parse_args(argv)


def main_predict(argv: Optional[Sequence[str]] = None) -> List[Dict[str, str]]:
    args = get_predict_args(argv)
    logging.basicConfig(level=logging.INFO if args.verbose >= 1 else logging.WARNING, format="%(message)s")
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    fpaths = list(args.audio)
    if len(fpaths) == 0:
        raise ValueError("Invalid argument --audio. (expected at least one file)")
    path = args.model_path if args.model_path is not None else args.model_name
    if path is None:
        raise ValueError(f"Invalid arguments {args.model_name=} and {args.model_path=}. (expected at one str value)")
    device = args.device  # "cuda_if_available" (the reference default) resolves to cuda:0; anything but CUDA raises ValueError
    model = conette(path, model_kwds=dict(device=device, precision=args.precision))
    tasks = args.task
    if tasks is not None and len(tasks) == 1:
        tasks = tasks[0]
    outs = model(fpaths, task=tasks)
    results = [{"audio": osp.basename(f), "task": t, "candidate": c} for f, t, c in zip(fpaths, outs["tasks"], outs["cands"])]
    for r in results:
        pylog.info(f"File '{r['audio']}' with task '{r['task']}':\n - '{r['candidate']}'")
    if args.csv_export is not None:
        with open(args.csv_export, "w") as file:
            writer = csv.DictWriter(file, fieldnames=["audio", "task", "candidate"])
            writer.writeheader()
            writer.writerows(results)
    model.engine.close()
    return results


if __name__ == "__main__":
    main_predict()
