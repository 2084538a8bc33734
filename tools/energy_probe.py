"""Energy per launch of the step's kernel classes (NVML total-energy counter, mJ resolution).

The C2 step runs against the 1 kW power cap (every nvidia-smi sample shows sw_power_cap), so the step time is set by
joules, not by pipe occupancy: time ~ energy / 1 kW.  This probe runs each kernel class back to back for >= 1.5 s,
reads nvmlDeviceGetTotalEnergyConsumption on both sides and reports J / launch, average W, ms / launch and pJ / FLOP
(or pJ / byte), next to torch.matmul 8192^3 (the cuBLAS kernel behind MEASURED_PEAKS' sustained figure).

    python tools/energy_probe.py [--seconds 1.5] > profiles/r02_energy_probe.json
"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
from affectgpt_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=1.5)
args = ap.parse_args()
DEV = torch.device("cuda:0")
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


def energy_mj():
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(h)


def probe(name, fn, flops=0.0, bytes_=0.0, step_launches=0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # calibrate the launch count for ~args.seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 5
    n = max(int(args.seconds * 1e3 / per), 10)
    time.sleep(0.3)
    torch.cuda.synchronize()
    j0 = energy_mj(); t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter(); j1 = energy_mj()
    ms = e0.elapsed_time(e1) / n
    joule = (j1 - j0) / 1e3 / n
    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    r = {"kernel": name, "launches": n, "ms_per_launch": round(ms, 4), "J_per_launch": round(joule, 4),
         "avg_W": round((j1 - j0) / 1e3 / (t1 - t0), 1), "sm_mhz_after": sm}
    if flops:
        r["TFLOPs"] = round(flops / ms / 1e9, 1); r["pJ_per_FLOP"] = round(joule / flops * 1e12, 3)
    if bytes_:
        r["GBs"] = round(bytes_ / ms / 1e6, 1); r["pJ_per_byte"] = round(joule / bytes_ * 1e12, 2)
    if step_launches:
        r["launches_per_C2_step"] = step_launches; r["J_per_C2_step"] = round(joule * step_launches, 2)
    print(json.dumps(r), flush=True)
    return r


g = torch.Generator(device=DEV).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device=DEV)
out = []
# ---- cuBLAS reference point
a = rn(8192, 8192).bfloat16(); b = rn(8192, 8192).bfloat16()
out.append(probe("torch.matmul bf16 8192^3 (cuBLAS)", lambda: torch.matmul(a, b), flops=2 * 8192 ** 3))
del a, b
# ---- ViT block kernels at C2 (M = 512 * 257)
M = 512 * 257
x = (rn(M, 1024) * 0.5).bfloat16()
wqkv = (rn(3072, 1024) * 0.02).bfloat16(); bqkv = rn(3072) * 0.05
wo = (rn(1024, 1024) * 0.02).bfloat16(); bo = rn(1024) * 0.05
w1 = (rn(4096, 1024) * 0.02).bfloat16(); b1 = rn(4096) * 0.05
w2 = (rn(1024, 4096) * 0.01).bfloat16(); b2 = rn(1024) * 0.05
resid = rn(M, 1024)
h1 = (rn(M, 4096) * 0.5).bfloat16()
qkv = (rn(M, 3072) * 0.5).bfloat16()
gam, bet = 1 + 0.1 * rn(1024), 0.1 * rn(1024)
L = ops._lib.load()


def gemm_bf16_out(a_, w_, b_, act, ob):
    Mm, K = a_.shape
    ops._chk(L.agb_op_gemm(0, ops.ACT[act], ops._p(a_), a_.stride(0), ops._p(w_), Mm, w_.shape[0], K, ops._p(b_), None, 0, None, 0,
                           ops._p(ob), w_.shape[0], 0, 0, 0, ops._s(a_)), "gemm")


o_qkv = torch.empty(M, 3072, dtype=torch.bfloat16, device=DEV)
o_h1 = torch.empty(M, 4096, dtype=torch.bfloat16, device=DEV)
out.append(probe("gemm QKV  M=131584 N=3072 K=1024 -> bf16", lambda: gemm_bf16_out(x, wqkv, bqkv, None, o_qkv), flops=2.0 * M * 3072 * 1024, step_launches=23))
out.append(probe("gemm FC1  M=131584 N=4096 K=1024 quick_gelu -> bf16", lambda: gemm_bf16_out(x, w1, b1, "quick_gelu", o_h1), flops=2.0 * M * 4096 * 1024, step_launches=23))
out.append(probe("gemm out-proj M=131584 N=1024 K=1024 += fp32", lambda: ops.gemm_inplace_residual(x, wo, bo, resid), flops=2.0 * M * 1024 * 1024, step_launches=23))
out.append(probe("gemm FC2  M=131584 N=1024 K=4096 += fp32", lambda: ops.gemm_inplace_residual(h1, w2, b2, resid), flops=2.0 * M * 1024 * 4096, step_launches=23))
resid.normal_(generator=g)
out.append(probe("layernorm 131584 x 1024 fp32 -> bf16", lambda: ops.layernorm(resid, gam, bet), bytes_=M * 1024 * 6.0, step_launches=47))
out.append(probe("attention 512 x 16 heads x 257 tokens", lambda: ops.attention(qkv, 512, 257, 257), flops=4.0 * 512 * 16 * 257 * 257 * 64, step_launches=23))
# ---- HuBERT-side kernels at C2 (512 segments)
del x, h1, qkv, o_qkv, o_h1, resid
torch.cuda.empty_cache()
Mh = 512 * 100
qkv_h = (rn(Mh, 3072) * 0.5).bfloat16()
out.append(probe("attention 512 x 16 heads x 99 frames (HuBERT)", lambda: ops.attention(qkv_h, 512, 100, 99), flops=4.0 * 512 * 16 * 99 * 99 * 64, step_launches=24))
wav = rn(512, 32000) * 0.1
c0w, c0b = rn(512, 1, 10) * 0.3, rn(512) * 0.05
g512, b512 = 1 + 0.1 * rn(512), 0.1 * rn(512)
out.append(probe("conv0 + LayerNorm(512) + GELU, 512 segments (tcgen05)", lambda: ops.conv0(wav, c0w, c0b, g512, b512, "layer", 8),
                 bytes_=512 * (32000 * 4.0 + 6400 * 512 * 2.0), step_launches=1))
xc = (rn(512 * 3200, 512) * 0.5).bfloat16()
out.append(probe("ln512 + GELU over 1.64 M rows bf16 (conv stack layer 1)", lambda: ops.layernorm512_bf16(xc, g512, b512), bytes_=512 * 3200 * 512 * 4.0, step_launches=1))
x16 = (rn(Mh, 1024) * 0.5).bfloat16(); wp = (rn(1024, 8192) * 0.01).bfloat16(); pb = rn(1024) * 0.05; rs = rn(Mh, 1024)
out.append(probe("positional conv k=128 g=16 as GEMM, 512 x 99 (+= fp32)", lambda: ops.posconv_residual(x16, wp, pb, rs, 512, 100, 99),
                 flops=2.0 * 512 * 99 * 1024 * 64 * 128, step_launches=1))
del qkv_h, xc, x16, wp, rs
torch.cuda.empty_cache()
# ---- copy kernel as the HBM reference point
src = torch.empty(1 << 30, dtype=torch.bfloat16, device=DEV); dst = torch.empty_like(src)
out.append(probe("torch copy 2 GiB bf16 (HBM read+write)", lambda: dst.copy_(src), bytes_=4.0 * (1 << 30)))
# ---- idle floor
time.sleep(0.5)
j0 = energy_mj(); time.sleep(1.0); j1 = energy_mj()
print(json.dumps({"kernel": "idle (no work, context alive)", "avg_W": round((j1 - j0) / 1e3, 1)}), flush=True)
