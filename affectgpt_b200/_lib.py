"""ctypes binding of libagb.so (the C ABI declared in include/agb.h).

There is no CPU or eager-PyTorch fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference's error convention is Python exceptions; its extractor catches Exception per sample,
A/extract_multimodal_features_precompute.py:227-229).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libagb.so"
_lib: Optional[C.CDLL] = None


class AgbCfg(C.Structure):
    _fields_ = [
        ("vit_layers", C.c_int32), ("hubert_layers", C.c_int32), ("hubert_conv_norm", C.c_int32),
        ("hubert_do_normalize", C.c_int32), ("llm_dim", C.c_int32), ("max_frames_per_chunk", C.c_int32),
        ("max_segments_per_chunk", C.c_int32), ("qf_video_queries", C.c_int32), ("qf_audio_queries", C.c_int32),
        ("qf_multi_queries", C.c_int32), ("vit_dense_last_block", C.c_int32), ("hubert_post_layer_norm", C.c_int32),
        ("hubert_no_feat_proj_ln", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/agb.h declares (checked by tests)
SIGNATURES = {
    "agb_version": (_i, []),
    "agb_default_cfg": (None, [C.POINTER(AgbCfg)]),
    "agb_create": (_i, [C.POINTER(_vp), _i, C.POINTER(AgbCfg)]),
    "agb_destroy": (None, [_vp]),
    "agb_last_error": (C.c_char_p, [_vp]),
    "agb_load_tensor": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i64), _i, _vp]),
    "agb_has_tensor": (_i, [_vp, C.c_char_p]),
    "agb_missing_count": (_i, [_vp, C.c_char_p]),
    "agb_missing_name": (C.c_char_p, [_vp, C.c_char_p, _i]),
    "agb_clip_forward": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "agb_hubert_forward": (_i, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "agb_video_fuse": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "agb_audio_fuse": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "agb_multi_fuse": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "agb_splice_tokens": (_i, [_vp, _i, _vp, _i, _i, _i, _i, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                C.POINTER(_vp), _vp, _vp]),
    "agb_stage_faces": (_i, [_vp, _i64, _i, _i, _vp, _i, _vp, _vp]),
    "agb_stage_audio": (_i, [_vp, _i, _i64, _vp, _i, _i, _vp, _vp]),
    "agb_resample_audio": (_i, [_vp, _i, _i64, _vp, _i, _i, _i, _i, _vp, _i64, _vp]),
    "agb_qformer_fuse": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "agb_set_tap": (_i, [_vp, C.c_char_p, _vp, _i64]),
    "agb_launch_count": (_i64, []),
    "agb_profile_begin": (_i, []),
    "agb_profile_end": (_i, [_vp, _vp, _vp, _vp]),
    "agb_op_gemm": (_i, [_i, _i, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _vp]),
    "agb_op_layernorm": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _f, _vp, _i64, _vp, _i64, _vp]),
    "agb_op_layernorm512_bf16": (_i, [_vp, _i64, _vp, _vp, _f, _i, _vp, _vp]),
    "agb_op_attention": (_i, [_vp, _vp, _i, _i, _i, _f, _vp]),
    "agb_op_patchify": (_i, [_vp, _i, _i64, _vp, _vp]),
    "agb_op_conv0": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "agb_op_seq_mean_pool": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _f, _i, _vp, _vp]),
    "agb_op_linear_f32": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _vp]),
    "agb_op_f32_to_bf16": (_i, [_vp, _vp, _i64, _vp]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libagb.so; raises RuntimeError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    from .build import build_lib, is_current
    if not is_current():
        # missing, or built from other sources than the ones on disk (a stale library would be used silently otherwise)
        if os.environ.get("AGB_AUTOBUILD", "1") == "1":
            build_lib()                                       # serialised across ranks by a file lock, atomic rename
        elif _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is stale (csrc/ changed since it was built) and AGB_AUTOBUILD=0: "
                               "rebuild with `python -m affectgpt_b200.build`")
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m affectgpt_b200.build` "
                "(affectgpt_b200 has no CPU / eager fallback)")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError here == ABI drift; fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(ctx=None) -> str:
    msg = load().agb_last_error(ctx)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, ctx=None, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libagb {what} failed (code {rc}): {last_error(ctx)}")
