"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/agb.h declares.
No compute calls here (CPU box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "agb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for must in ("agb_create", "agb_destroy", "agb_last_error", "agb_load_tensor", "agb_clip_forward",
                 "agb_hubert_forward", "agb_video_fuse", "agb_audio_fuse", "agb_multi_fuse", "agb_op_gemm"):
        assert must in syms


def test_library_builds_and_exports_every_declared_symbol():
    from affectgpt_b200.build import build_lib
    path = build_lib()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    for sym in _declared_symbols():
        assert hasattr(lib, sym), f"libagb.so does not export {sym}"


def test_ctypes_signatures_cover_the_header():
    from affectgpt_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.agb_version() == 100


def test_cfg_struct_layout_matches_header():
    from affectgpt_b200 import _lib
    assert ctypes.sizeof(_lib.AgbCfg) == 16 * 4
    cfg = _lib.AgbCfg()
    _lib.load().agb_default_cfg(ctypes.byref(cfg))
    assert (cfg.vit_layers, cfg.hubert_layers, cfg.hubert_conv_norm, cfg.hubert_do_normalize, cfg.llm_dim) == (24, 24, 0, 1, 3584)


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md 'What proves ...')."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from affectgpt_b200.build import build_lib
    sass = subprocess.run([cuobjdump, "-sass", str(build_lib())], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from affectgpt_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine("cuda:0")
    from affectgpt_b200 import _lib
    ctx = ctypes.c_void_p()
    rc = _lib.load().agb_create(ctypes.byref(ctx), 0, None)
    assert rc != 0 and "no CUDA device" in _lib.last_error(None)
