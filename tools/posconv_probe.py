"""Runs the grouped positional conv a few times (ncu target: tools/posconv_ab.sh / manual captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from affectgpt_b200 import ops  # noqa: E402

dev = "cuda:0"
nseq, stride, valid = 512, 100, 99
x16 = torch.randn(nseq * stride, 1024, device=dev).to(torch.bfloat16)
wp = (torch.randn(1024, 8192, device=dev) * 0.01).to(torch.bfloat16)
pb = torch.randn(1024, device=dev)
resid = torch.zeros(nseq * stride, 1024, device=dev)
for _ in range(4):
    ops.posconv_residual(x16, wp, pb, resid, nseq, stride, valid)
torch.cuda.synchronize()
