"""Micro-benchmark of the HBM / ALU-bound front-end kernels on the C2 shapes (CUDA events, inputs larger than L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from affectgpt_b200 import ops  # noqa: E402

dev = "cuda:0"


def timeit(name, fn, bytes_moved, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms:8.3f} ms  {bytes_moved / ms / 1e6:8.0f} GB/s", flush=True)


n = 512
wav = torch.randn(n, 32000, device=dev) * 0.1
w = torch.randn(512, 10, device=dev) * 0.3
b = torch.randn(512, device=dev) * 0.1
g = torch.rand(512, device=dev) + 0.5
be = torch.randn(512, device=dev) * 0.1
timeit("conv0+LN+GELU (512 seg)", lambda: ops.conv0(wav, w, b, g, be, "layer", 8), n * (32000 * 4 + 6400 * 512 * 2))
x = torch.randn(n * 3200, 512, device=dev).to(torch.bfloat16)
timeit("ln512+GELU (512x3200 rows)", lambda: ops.layernorm512_bf16(x, g, be, 1e-5, True), x.numel() * 4)
frames = torch.randint(0, 256, (512, 3, 224, 224), device=dev, dtype=torch.uint8)
timeit("patchify u8 (512 frames)", lambda: ops.patchify(frames), 512 * (3 * 224 * 224 + 256 * 640 * 2))
ff = frames.float()
timeit("patchify f32 (512 frames)", lambda: ops.patchify(ff), 512 * (3 * 224 * 224 * 4 + 256 * 640 * 2))
