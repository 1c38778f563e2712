"""ctypes binding of libconette_b200.so (the C ABI declared in include/conette_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing this binding raises, and
every product entry point fails loudly (the oracle under oracle/ is test infrastructure and is never used here).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

ABI_VERSION = 1
PRECISION_FAST, PRECISION_PARITY = 0, 1
DTYPE_F32, DTYPE_I64, DTYPE_BOOL, DTYPE_U8 = 0, 1, 2, 3
TAP_LOGMEL_BN, TAP_STEM, TAP_BLOCK, TAP_DOWN, TAP_DWLN = 0, 1, 2, 3, 4

KERNEL_CLASSES = ("frontend", "stem", "dwconv_ln", "gemm_pw1_gelu", "gemm_pw2_resid", "ds_ln_pack", "ds_gemm", "head",
                  "proj_crosskv", "dec_gemm", "dec_attn_ln", "dec_classifier", "beam",
                  "dwconv_ln.s1", "dwconv_ln.s2", "dwconv_ln.s3", "dwconv_ln.s4",
                  "gemm_pw1_gelu.s1", "gemm_pw1_gelu.s2", "gemm_pw1_gelu.s3", "gemm_pw1_gelu.s4",
                  "gemm_pw2_resid.s1", "gemm_pw2_resid.s2", "gemm_pw2_resid.s3", "gemm_pw2_resid.s4")

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libconette_b200.so"


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("vocab_size", C.c_int32),
        ("precision", C.c_int32),
        ("enc_chunk", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class CnbError(RuntimeError):
    pass


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes); must list every symbol include/conette_b200.h declares (tests check this)
SIGNATURES = {
    "cnb_last_error": (C.c_char_p, []),
    "cnb_abi_version": (C.c_int, []),
    "cnb_create": (C.c_int, [C.POINTER(Config), C.POINTER(_vp)]),
    "cnb_destroy": (C.c_int, [_vp]),
    "cnb_load_weight": (C.c_int, [_vp, C.c_char_p, _vp, _i32, _i32, C.POINTER(_i64)]),
    "cnb_finalize_weights": (C.c_int, [_vp]),
    "cnb_geometry": (C.c_int, [_i64, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "cnb_resample": (C.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "cnb_frontend": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "cnb_encoder": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "cnb_encoder_tap": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _i64, _vp]),
    "cnb_decode": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cnb_decode_tap": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cnb_decoder_logits": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "cnb_score_captions": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "cnb_caption": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cnb_caption_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cnb_caption_host_begin": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                         C.POINTER(_i32)]),
    "cnb_caption_host_end": (C.c_int, [_vp, _i32]),
    "cnb_debug_gemm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "cnb_debug_mlp_fused": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "cnb_debug_mlp_fused_pair": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "cnb_profile_begin": (C.c_int, [_vp]),
    "cnb_profile_end": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(_i64), _i32]),
    "cnb_profile_timeline_begin": (C.c_int, [_vp]),
    "cnb_profile_timeline_end": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(C.c_float), C.POINTER(C.c_float), _i32, C.POINTER(_i32)]),
    "cnb_launch_count": (_i64, [_vp]),
    "cnb_device_bytes": (_i64, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CnbError(
            f"{LIB_PATH} not found: build it with `python -m conette_audio_captioning_b200.build` "
            "(or __graft_entry__.build()); there is no CPU / PyTorch fallback for the CUDA path"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.cnb_abi_version() != ABI_VERSION:
        raise CnbError(f"ABI mismatch: library {lib.cnb_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().cnb_last_error()
        raise CnbError(f"conette_b200 error {rc}: {msg.decode() if msg else '?'}")


def geometry(n_samples: int):
    """(T stft frames, [H1..H4], T' output frames) for a padded batch of n_samples."""
    t, tp = _i32(), _i32()
    hs = (_i32 * 4)()
    check(load().cnb_geometry(int(n_samples), C.byref(t), hs, C.byref(tp)))
    return t.value, list(hs), tp.value
