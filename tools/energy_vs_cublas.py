import json, os, sys, time
import torch
sys.path.insert(0, '.')
import pynvml
from affectgpt_b200 import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
DEV = torch.device("cuda:0")
def probe(name, fn, flops, secs=1.5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    n = max(int(secs * 1e3 / (e0.elapsed_time(e1) / 5)), 10)
    time.sleep(0.3); torch.cuda.synchronize()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t0 = time.perf_counter()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter(); j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    ms = e0.elapsed_time(e1) / n; J = (j1 - j0) / 1e3 / n
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "J": round(J, 4), "W": round((j1 - j0) / 1e3 / (t1 - t0), 1), "TFLOPs": round(flops / ms / 1e9, 1), "pJ_per_FLOP": round(J / flops * 1e12, 3)}), flush=True)
g = torch.Generator(device=DEV).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=DEV)
M = 512 * 257
L = ops._lib.load()
for name, N, K in (("QKV", 3072, 1024), ("FC1", 4096, 1024), ("FC2", 1024, 4096), ("out", 1024, 1024)):
    x = (rn(M, K) * 0.5).bfloat16(); w = (rn(N, K) * 0.02).bfloat16(); b = rn(N) * 0.05
    o = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    wt = w.t()
    def ours():
        ops._chk(L.agb_op_gemm(0, 0, ops._p(x), K, ops._p(w), M, N, K, ops._p(b), None, 0, None, 0, ops._p(o), N, 0, 0, 0, ops._s(x)), "gemm")
    for rep in range(2):
        probe(f"ours  {name} -> bf16 (+bias)", ours, 2.0 * M * N * K)
        probe(f"cuBLAS {name} -> bf16 (bare)", lambda: torch.matmul(x, wt, out=o), 2.0 * M * N * K)
    del x, w, o
