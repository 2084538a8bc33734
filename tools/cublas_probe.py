"""What does the library do on this path's GEMM shapes?  Runs torch.matmul (cuBLAS) on the FC2 / QKV shapes so that ncu can
report the kernel it picks (name = tile / cluster shape), its launch configuration and its L2 / DRAM traffic."""
import torch
dev = "cuda:0"
for M, N, K in ((131584, 1024, 4096), (131584, 3072, 1024)):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a, w.t(), out=o)
    torch.cuda.synchronize()
