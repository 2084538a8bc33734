"""Micro-benchmark of the attention kernels on the path's shapes (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from affectgpt_b200 import ops  # noqa: E402

dev = "cuda:0"
for name, nseq, stride, valid in [("vit", 512, 257, 257), ("hubert", 512, 100, 99)]:
    qkv = torch.randn(nseq * stride, 3072, device=dev).to(torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, nseq, stride, valid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 10
    e0.record()
    for _ in range(it):
        ops.attention(qkv, nseq, stride, valid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    fl = 4.0 * nseq * 16 * valid * valid * 64
    print(f"{name}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (includes a {nseq*stride*2048/1e6:.0f} MB output memset)", flush=True)
