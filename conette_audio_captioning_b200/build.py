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
SOURCES = ("api.cu", "frontend.cu", "encoder.cu", "gemm_simt.cu", "gemm_tc.cu", "decoder.cu", "decoder_cluster.cu", "beam.cu", "dwconv_ring.cu", "resample.cu", "mlp_fused.cu", "mlp_fused_pair.cu")
NVCC_FLAGS = (
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def _headers_digest() -> bytes:
    hsh = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "conette_b200.h"]):
        hsh.update(f.name.encode())
        hsh.update(f.read_bytes())
    hsh.update(" ".join(NVCC_FLAGS).encode())
    return hsh.digest()


def _compile_one(nvcc: str, src: Path, obj: Path, hdr: bytes, force: bool, verbose: bool) -> str:
    """One translation unit -> one object file; skipped when neither the source nor any header changed."""
    fp = hashlib.sha256(hdr + src.read_bytes()).hexdigest()
    stamp = obj.with_suffix(".sha256")
    if not force and obj.exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return ""
    cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "-shared"], "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name} ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    stamp.write_text(fp)
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a (one object per source, in parallel, cached) and link one shared library."""
    from concurrent.futures import ThreadPoolExecutor

    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    nvcc, hdr = _nvcc(), _headers_digest()
    objs = [obj_dir / (Path(s).stem + ".o") for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        logs = list(ex.map(lambda so: _compile_one(nvcc, CSRC / so[0], so[1], hdr, force, verbose), zip(SOURCES, objs)))
    if verbose:
        print("\n".join(l for l in logs if l), file=sys.stderr)
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        res = subprocess.run([nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs)], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
