"""Engine: one libagb context on one GPU + torch-tensor plumbing (device memory, streams).

PyTorch is used only to own device buffers and the CUDA stream; all arithmetic of the encode-and-fuse
path runs in libagb.so's sm_100a kernels.  No fallback: without a CUDA device or the built library every
entry point raises RuntimeError.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Tuple

import torch

from . import _lib

FUSE_ATTENTION, FUSE_MEAN = 0, 1
_FUSION_CODES = {"attention": FUSE_ATTENTION, "mean": FUSE_MEAN}


def _stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else int(t.data_ptr())


class Engine:
    """Owns an ``agb_ctx``.  One per (process, device, model configuration); not thread-safe."""

    def __init__(self, device="cuda:0", vit_layers: int = 24, hubert_layers: int = 24,
                 feat_extract_norm: str = "layer", do_normalize: bool = True, llm_dim: int = 3584,
                 max_frames_per_chunk: int = 512, max_segments_per_chunk: int = 512, qf_video_queries: int = 0,
                 qf_audio_queries: int = 0, qf_multi_queries: int = 0, vit_dense_last_block: bool = False,
                 do_stable_layer_norm: bool = True, feat_proj_layer_norm: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("affectgpt_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if feat_extract_norm not in ("layer", "group"):
            raise ValueError(f"feat_extract_norm must be 'layer' or 'group', got {feat_extract_norm!r}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"affectgpt_b200 runs on CUDA devices only, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lib = _lib.load()
        cfg = _lib.AgbCfg()
        self.lib.agb_default_cfg(C.byref(cfg))
        cfg.vit_layers = vit_layers
        cfg.hubert_layers = hubert_layers
        cfg.hubert_conv_norm = 0 if feat_extract_norm == "layer" else 1
        cfg.hubert_do_normalize = 1 if do_normalize else 0
        cfg.llm_dim = llm_dim
        cfg.vit_dense_last_block = 1 if vit_dense_last_block else 0
        cfg.hubert_post_layer_norm = 0 if do_stable_layer_norm else 1     # HubertConfig.do_stable_layer_norm
        cfg.hubert_no_feat_proj_ln = 0 if feat_proj_layer_norm else 1     # HubertConfig.feat_proj_layer_norm
        cfg.max_frames_per_chunk = max_frames_per_chunk
        cfg.max_segments_per_chunk = max_segments_per_chunk
        cfg.qf_video_queries, cfg.qf_audio_queries, cfg.qf_multi_queries = qf_video_queries, qf_audio_queries, qf_multi_queries
        self.qf_queries = {"video": qf_video_queries, "audio": qf_audio_queries, "multi": qf_multi_queries}
        self.cfg = cfg
        self.llm_dim = llm_dim
        self._ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device)  # make sure the primary context exists
            _lib.check(self.lib.agb_create(C.byref(self._ctx), self.device.index, C.byref(cfg)), None, "agb_create")
        self._taps: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self.lib.agb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> None:
        _lib.check(rc, self._ctx, what)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, family: str, state_dict: Dict[str, torch.Tensor], strict: bool = True) -> Tuple[list, list]:
        """Load checkpoint tensors of one family ("clip", "hubert", "fusion", "qf_video", "qf_audio", "qf_multi") by
        their checkpoint keys.  Returns (missing, unexpected) like nn.Module.load_state_dict;
        with ``strict`` a missing tensor raises RuntimeError."""
        assert family in ("clip", "hubert", "fusion", "qf_video", "qf_audio", "qf_multi")
        unexpected = []
        stream = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            for key, value in state_dict.items():
                name = f"{family}.{key}".encode()
                alias = key.replace("conv.weight_g", "conv.parametrizations.weight.original0").replace(
                    "conv.weight_v", "conv.parametrizations.weight.original1")
                if not self.lib.agb_has_tensor(self._ctx, f"{family}.{alias}".encode()):
                    unexpected.append(key)
                    continue
                t = value.detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
                self._check(self.lib.agb_load_tensor(self._ctx, name, _ptr(t), 0, shape, max(t.dim(), 1), stream),
                            f"agb_load_tensor({key})")
            torch.cuda.current_stream(self.device).synchronize()   # staging tensors die after this scope
        missing = self.missing(family)
        if strict and missing:
            raise RuntimeError(f"missing {family} tensors: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
        return missing, unexpected

    def missing(self, family: str) -> list:
        n = self.lib.agb_missing_count(self._ctx, family.encode())
        out = []
        for i in range(max(n, 0)):
            s = self.lib.agb_missing_name(self._ctx, family.encode(), i)
            if s:
                out.append(s.decode()[len(family) + 1:])
        return out

    # ------------------------------------------------------------------ encoders
    def clip_forward(self, raw_frames: torch.Tensor) -> torch.Tensor:
        """raw_frames [N,3,224,224] float32 (0..255) or uint8, on this device -> [N,768] float32."""
        if raw_frames.dim() != 4 or tuple(raw_frames.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected [N,3,224,224] frames, got {tuple(raw_frames.shape)}")
        if raw_frames.dtype == torch.uint8:
            code = 1
        else:
            raw_frames = raw_frames.to(torch.float32)
            code = 0
        x = raw_frames.to(self.device).contiguous()
        out = torch.empty((x.shape[0], 768), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_clip_forward(self._ctx, _ptr(x), code, x.shape[0], _ptr(out),
                                                  _stream_ptr(self.device)), "agb_clip_forward")
        return out

    def hubert_forward(self, wav: torch.Tensor, seg_per_norm_group: int = 0) -> torch.Tensor:
        """wav [N,32000] float32 -> [N,1024] float32.  ``seg_per_norm_group`` <= 0: one normalisation
        group for the whole call (what the reference does for one call)."""
        if wav.dim() != 2 or wav.shape[1] != 32000:
            raise ValueError(f"expected [N,32000] waveform segments, got {tuple(wav.shape)}")
        x = wav.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((x.shape[0], 1024), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_hubert_forward(self._ctx, _ptr(x), x.shape[0], seg_per_norm_group, _ptr(out),
                                                    _stream_ptr(self.device)), "agb_hubert_forward")
        return out

    # ------------------------------------------------------------------ fusion heads + projectors
    def video_fuse(self, hidden: torch.Tensor, fusion_type: str = "attention", bypass_t1: bool = False) -> torch.Tensor:
        h = hidden.to(device=self.device, dtype=torch.float32).contiguous()
        b, t, d = h.shape
        assert d == 768
        llm = torch.empty((b, self.llm_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_video_fuse(self._ctx, _ptr(h), b, t, _FUSION_CODES[fusion_type], int(bypass_t1),
                                                _ptr(llm), _stream_ptr(self.device)), "agb_video_fuse")
        return llm

    def audio_fuse(self, hidden: torch.Tensor, fusion_type: str = "attention") -> torch.Tensor:
        h = hidden.to(device=self.device, dtype=torch.float32).contiguous()
        b, t, d = h.shape
        assert d == 1024
        llm = torch.empty((b, self.llm_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_audio_fuse(self._ctx, _ptr(h), b, t, _FUSION_CODES[fusion_type], _ptr(llm),
                                                _stream_ptr(self.device)), "agb_audio_fuse")
        return llm

    def multi_fuse(self, video_hidden: torch.Tensor, audio_hidden: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        hv = video_hidden.to(device=self.device, dtype=torch.float32).contiguous()
        ha = audio_hidden.to(device=self.device, dtype=torch.float32).contiguous()
        b, tv, dv = hv.shape
        b2, ta, da = ha.shape
        assert b == b2 and dv == 768 and da == 1024
        multi = torch.empty((b, 1024), dtype=torch.float32, device=self.device)
        llm = torch.empty((b, self.llm_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_multi_fuse(self._ctx, _ptr(hv), tv, _ptr(ha), ta, b, _ptr(multi), _ptr(llm),
                                                _stream_ptr(self.device)), "agb_multi_fuse")
        return multi, llm

    def qformer_fuse(self, which: str, video_hidden: Optional[torch.Tensor] = None,
                     audio_hidden: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """``*_fusion_type == 'qformer'``: -> (query hidden [b,n_q,768], llm [b,n_q,llm_dim]).  which: 'video'
        (video_hidden [b,t,768]), 'audio' (audio_hidden [b,t,1024]) or 'multi' (both)."""
        code = {"video": 0, "audio": 1, "multi": 2}[which]
        nq = self.qf_queries[which]
        if nq <= 0:
            raise RuntimeError(f"engine was created without a {which} Q-Former (qf_{which}_queries == 0)")
        hv = video_hidden.to(device=self.device, dtype=torch.float32).contiguous() if video_hidden is not None else None
        ha = audio_hidden.to(device=self.device, dtype=torch.float32).contiguous() if audio_hidden is not None else None
        ref = hv if hv is not None else ha
        if ref is None:
            raise ValueError("qformer_fuse needs at least one input")
        b = ref.shape[0]
        if hv is not None:
            assert hv.dim() == 3 and hv.shape[2] == 768 and hv.shape[0] == b
        if ha is not None:
            assert ha.dim() == 3 and ha.shape[2] == 1024 and ha.shape[0] == b
        hidden = torch.empty((b, nq, 768), dtype=torch.float32, device=self.device)
        llm = torch.empty((b, nq, self.llm_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.agb_qformer_fuse(self._ctx, code, _ptr(hv), hv.shape[1] if hv is not None else 0,
                                                  _ptr(ha), ha.shape[1] if ha is not None else 0, b, _ptr(hidden),
                                                  _ptr(llm), _stream_ptr(self.device)), "agb_qformer_fuse")
        return hidden, llm

    # ------------------------------------------------------------------ debug taps
    def tap(self, name: str, shape: Iterable[int], dtype=torch.float32) -> torch.Tensor:
        """Register a destination for a named intermediate; filled by the next forward."""
        buf = torch.zeros(tuple(shape), dtype=dtype, device=self.device)
        self._taps[name] = buf
        self._check(self.lib.agb_set_tap(self._ctx, name.encode(), _ptr(buf), buf.numel() * buf.element_size()),
                    "agb_set_tap")
        return buf

    def clear_taps(self) -> None:
        for name in list(self._taps):
            self.lib.agb_set_tap(self._ctx, name.encode(), None, 0)
        self._taps.clear()


_ENGINES: Dict[tuple, Engine] = {}


def get_engine(device="cuda:0", **cfg) -> Engine:
    """Process-wide engine cache: encoders and fusion heads on one device share one context."""
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    key = (str(dev), tuple(sorted(cfg.items())))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = Engine(dev, **cfg)
        _ENGINES[key] = eng
    return eng


def launch_count() -> int:
    return int(_lib.load().agb_launch_count())


PROFILE_CATEGORIES = ("gemm", "attention", "layernorm", "frontend", "other")


def profile_begin() -> None:
    _lib.check(_lib.load().agb_profile_begin(), None, "agb_profile_begin")


def profile_end() -> dict:
    """-> {category: {"ms", "launches", "flops", "bytes"}} summed over the profiled region (device-synchronises)."""
    n = len(PROFILE_CATEGORIES)
    ms, cnt, fl, by = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
    _lib.check(_lib.load().agb_profile_end(ms, cnt, fl, by), None, "agb_profile_end")
    return {name: {"ms": ms[i], "launches": int(cnt[i]), "flops": fl[i], "bytes": by[i]}
            for i, name in enumerate(PROFILE_CATEGORIES)}
