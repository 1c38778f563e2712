"""Build recipe for libconette_b200.so (hand-written sm_100a CUDA + C ABI), in tree, with plain nvcc.

The shared library lands in ``conette_audio_captioning_b200/lib/`` (git-ignored, but it travels to the GPU box with the
gpurun snapshot).  nvcc cross-compiles sm_100a without a GPU, so this also is the "does it build" check.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libconette_b200.so"
SOURCES = ("api.cu", "frontend.cu", "encoder.cu", "gemm_simt.cu", "gemm_tc.cu", "decoder.cu", "decoder_persistent.cu", "decoder_cluster.cu", "beam.cu")
NVCC_FLAGS = (
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def _fingerprint() -> str:
    hsh = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                    [PKG.parent / "include" / "conette_b200.h"]):
        hsh.update(f.name.encode())
        hsh.update(f.read_bytes())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into one shared library; no-op when sources are unchanged."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.sha256"
    fp = _fingerprint()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr, file=sys.stderr)
    stamp.write_text(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
