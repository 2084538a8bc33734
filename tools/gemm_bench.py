"""Micro-benchmark of the tcgen05 GEMM on the path's shapes (CUDA events, L2 flushed between iterations by cycling
through operand sets larger than the 126 MB L2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from affectgpt_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
shapes = [("vit_qkv", 131584, 3072, 1024, None, False), ("vit_out", 131584, 1024, 1024, None, True),
          ("vit_fc1", 131584, 4096, 1024, "quick_gelu", False), ("vit_fc2", 131584, 1024, 4096, None, True),
          ("hub_qkv", 51200, 3072, 1024, None, False), ("hub_ffn1", 51200, 4096, 1024, "gelu", False),
          ("hub_ffn2", 51200, 1024, 4096, None, True), ("patch", 131072, 1024, 640, None, False)]
res = {}
lib = _lib.load()
for name, M, N, K, act, resid in shapes:
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if resid else None
    ob = None if resid else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    s = int(torch.cuda.current_stream().cuda_stream)

    def run():
        rc = lib.agb_op_gemm(0, ops.ACT[act], a.data_ptr(), K, w.data_ptr(), M, N, K, bias.data_ptr(),
                             r.data_ptr() if resid else None, N, r.data_ptr() if resid else None, N,
                             None if resid else ob.data_ptr(), N, 0, 0, 0, s)
        assert rc == 0, _lib.last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # the library on the same shape (cuBLAS through torch: bare A @ W^T into bf16, no bias / activation / residual)
    wt = w.t()
    oc = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, wt, out=oc)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(a, wt, out=oc)
    e1.record()
    torch.cuda.synchronize()
    ms_lib = e0.elapsed_time(e1) / iters
    del oc
    res[name] = {"ms": ms, "tflops": tf, "cublas_ms": ms_lib, "cublas_tflops": 2.0 * M * N * K / ms_lib / 1e9}
    print(f"{name:9s} M={M} N={N} K={K}: {ms:.3f} ms  {tf:.0f} TFLOP/s   | cuBLAS bare GEMM {ms_lib:.3f} ms "
          f"{2.0 * M * N * K / ms_lib / 1e9:.0f} TFLOP/s", flush=True)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(out, exist_ok=True)
json.dump(res, open(os.path.join(out, "gemm_vs_cublas.json"), "w"), indent=1)

# grouped positional conv (k = 128, 16 groups of 64): 512 sequences of 99 frames, resid += gelu(conv)
nseq, stride, valid = 512, 100, 99
x16 = torch.randn(nseq * stride, 1024, device=dev).to(torch.bfloat16)
wp = (torch.randn(1024, 8192, device=dev) * 0.01).to(torch.bfloat16)
pb = torch.randn(1024, device=dev)
resid = torch.zeros(nseq * stride, 1024, device=dev)
for _ in range(3):
    ops.posconv_residual(x16, wp, pb, resid, nseq, stride, valid)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.posconv_residual(x16, wp, pb, resid, nseq, stride, valid)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"posconv   512 x 99 x 1024, k=128 g=16: {ms:.3f} ms  {2.0 * nseq * valid * 1024 * 64 * 128 / ms / 1e9:.0f} TFLOP/s", flush=True)
