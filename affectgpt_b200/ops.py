"""Kernel-level entry points of libagb.so on torch tensors (unit tests, micro-benchmarks).

Every function launches the hand-written sm_100a kernel through the C ABI on the current stream; none has
an eager fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

ACT = {"none": 0, None: 0, "quick_gelu": 1, "gelu": 2}
A_PLAIN, A_CONV3, A_POSCONV = 0, 1, 2


def _s(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else int(t.data_ptr())


def _chk(rc: int, what: str):
    _lib.check(rc, None, what)


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
         res: Optional[torch.Tensor] = None, out_f32: bool = False, out_bf16: bool = True):
    """act(a @ w.T + bias) (+ res).  a [M,K] bf16 (row stride may exceed K), w [N,K] bf16.
    Returns (f32 or None, bf16 or None)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_cuda
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.is_contiguous()
    of = ob = None
    lib = _lib.load()
    with torch.cuda.device(a.device):
        if out_f32 or res is not None:       # one output per launch: the fp32 result (optionally += into a copy of res)
            if res is not None:
                assert res.dtype == torch.float32 and res.shape == (M, N)
                of = res.clone().contiguous()
            else:
                of = torch.empty((M, N), dtype=torch.float32, device=a.device)
            _chk(lib.agb_op_gemm(A_PLAIN, ACT[act], _p(a), a.stride(0), _p(w), M, N, K, _p(bias),
                                 _p(of) if res is not None else None, N, _p(of), N, None, 0, 0, 0, 0, _s(a)), "agb_op_gemm")
        if out_bf16 and res is None:
            ob = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
            _chk(lib.agb_op_gemm(A_PLAIN, ACT[act], _p(a), a.stride(0), _p(w), M, N, K, _p(bias), None, 0, None, 0,
                                 _p(ob), N, 0, 0, 0, _s(a)), "agb_op_gemm")
    return of, ob


def gemm_inplace_residual(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, resid: torch.Tensor, act=None):
    """resid += act(a @ w.T + bias), fp32 in place (the out-proj / fc2 epilogue of the encoder blocks)."""
    M, K = a.shape
    N = w.shape[0]
    with torch.cuda.device(a.device):
        _chk(_lib.load().agb_op_gemm(A_PLAIN, ACT[act], _p(a), a.stride(0), _p(w), M, N, K, _p(bias), _p(resid), N,
                                     _p(resid), N, None, 0, 0, 0, 0, _s(a)), "agb_op_gemm")
    return resid


def conv3_s2(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, m_out: int, act=None) -> torch.Tensor:
    """k=3, s=2 conv over channels-last rows: x [rows, 512] bf16 (rows even, >= 2*m_out), w_packed
    [512, 1536] (K index = tap*512 + c_in).  out[t] = sum_j W_j x[2t+j]; row 2*m_out (third tap of the last
    output row) reads as zero."""
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[1] == 512
    ob = torch.empty((m_out, 512), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_gemm(A_CONV3, ACT[act], _p(x), 0, _p(w_packed), m_out, 512, 1536, _p(bias), None, 0,
                                     None, 0, _p(ob), 512, 0, 0, 0, _s(x)), "agb_op_gemm(conv3)")
    return ob


def posconv_residual(x_bf16: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, resid: torch.Tensor,
                     nseq: int, seq_stride: int, seq_valid: int) -> torch.Tensor:
    """resid[seq, t, :] += gelu(grouped_conv_k128(x)[seq, t, :] + bias) for t < seq_valid (in place)."""
    with torch.cuda.device(x_bf16.device):
        _chk(_lib.load().agb_op_gemm(A_POSCONV, ACT["gelu"], _p(x_bf16), 0, _p(w_packed), nseq * seq_stride, 1024, 8192,
                                     _p(bias), _p(resid), 1024, _p(resid), 1024, None, 0, nseq, seq_stride, seq_valid,
                                     _s(x_bf16)), "agb_op_gemm(posconv)")
    return resid


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, want_f32: bool = False,
              row_stride: Optional[int] = None, rows: Optional[int] = None):
    """LayerNorm over the last dim (512/768/1024) of fp32 rows -> bf16 (and optionally fp32)."""
    assert x.dtype == torch.float32 and x.is_cuda
    D = gamma.numel()
    if rows is None:
        x2 = x.reshape(-1, D)
        rows, row_stride = x2.shape[0], D
    ob = torch.empty((rows, D), dtype=torch.bfloat16, device=x.device)
    of = torch.empty((rows, D), dtype=torch.float32, device=x.device) if want_f32 else None
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_layernorm(_p(x), row_stride, rows, D, _p(gamma), _p(beta), eps, _p(ob), D, _p(of), D,
                                          _s(x)), "agb_op_layernorm")
    return ob, of


def layernorm512_bf16(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, gelu: bool = True):
    assert x.dtype == torch.bfloat16 and x.shape[-1] == 512 and x.is_contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_layernorm512_bf16(_p(x), x.numel() // 512, _p(gamma), _p(beta), eps, int(gelu), _p(out),
                                                  _s(x)), "agb_op_layernorm512_bf16")
    return out


def attention(qkv: torch.Tensor, nseq: int, seq_stride: int, seq_valid: int, scale: float = 0.125) -> torch.Tensor:
    """qkv [nseq*seq_stride, 3072] bf16 -> context [nseq*seq_stride, 1024] bf16 (rows >= seq_valid untouched)."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (nseq * seq_stride, 3072)
    out = torch.zeros((nseq * seq_stride, 1024), dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _chk(_lib.load().agb_op_attention(_p(qkv), _p(out), nseq, seq_stride, seq_valid, scale, _s(qkv)),
             "agb_op_attention")
    return out


def patchify(frames: torch.Tensor) -> torch.Tensor:
    """frames [N,3,224,224] f32 (0..255) or u8 -> CLIP-normalised im2col matrix [N*256, 640] bf16."""
    code = 1 if frames.dtype == torch.uint8 else 0
    x = frames.contiguous() if code else frames.float().contiguous()
    out = torch.empty((x.shape[0] * 256, 640), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_patchify(_p(x), code, x.shape[0], _p(out), _s(x)), "agb_op_patchify")
    return out


def conv0(wav: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
          norm: str = "layer", seg_per_norm_group: int = 0, do_normalize: bool = True) -> torch.Tensor:
    """HuBERT conv0 (+LayerNorm/GroupNorm +GELU): wav [n,32000] f32 -> [n, 6400, 512] bf16 (row 6399 zero)."""
    n = wav.shape[0]
    out = torch.empty((n, 6400, 512), dtype=torch.bfloat16, device=wav.device)
    with torch.cuda.device(wav.device):
        _chk(_lib.load().agb_op_conv0(_p(wav.contiguous()), n, seg_per_norm_group, int(do_normalize),
                                      _p(w.contiguous()), _p(bias), _p(gamma), _p(beta), 0 if norm == "layer" else 1,
                                      _p(out), _s(wav)), "agb_op_conv0")
    return out


def seq_mean_pool(x: torch.Tensor, nseq: int, seq_stride: int, seq_valid: int, gamma=None, beta=None, eps: float = 1e-5,
                  scale: float = 1.0, accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    out = accumulate_into if accumulate_into is not None else torch.empty((nseq, 1024), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_seq_mean_pool(_p(x), nseq, seq_stride, seq_valid, _p(gamma), _p(beta), eps, scale,
                                              int(accumulate_into is not None), _p(out), _s(x)), "agb_op_seq_mean_pool")
    return out


def linear_f32(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_linear_f32(_p(x.contiguous()), K, _p(w.contiguous()), _p(b), _p(out), N, M, N, K, _s(x)),
             "agb_op_linear_f32")
    return out


def f32_to_bf16(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _chk(_lib.load().agb_op_f32_to_bf16(_p(x.contiguous()), _p(out), x.numel(), _s(x)), "agb_op_f32_to_bf16")
    return out
