"""Per-kernel parity on the GPU: each hand-written sm_100a kernel (called through the C ABI) against a plain
PyTorch fp32 reference of the same op on the same seeded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(DEV)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 128), (514, 3072, 1024), (514, 1024, 4096), (200, 768, 1024),
                                   (1028, 4096, 1024), (8, 256, 64), (257 * 8, 1024, 1024), (512, 1024, 640),
                                   (4096, 3072, 1024), (2300, 1024, 4096), (5000, 768, 1024), (2049, 512, 1536)])
def test_gemm_plain_bias_shapes(M, N, K):
    from affectgpt_b200 import ops
    a = _rand((M, K), 1, 1.0, torch.bfloat16)
    w = _rand((N, K), 2, 1.0 / math.sqrt(K), torch.bfloat16)
    bias = _rand((N,), 3)
    of, ob = ops.gemm(a, w, bias, out_f32=True, out_bf16=True)
    want = a.float() @ w.float().T + bias
    e32, e16 = rel_err(of, want), rel_err(ob, want)
    print(f"gemm {M}x{N}x{K}: rel f32 {e32:.2e} bf16 {e16:.2e}")
    assert e32 < 1e-4 and e16 < 8e-3


@pytest.mark.parametrize("act", ["quick_gelu", "gelu"])
def test_gemm_activation_epilogues(act):
    from affectgpt_b200 import ops
    M, N, K = 2570, 4096, 1024
    a = _rand((M, K), 4, 1.0, torch.bfloat16)
    w = _rand((N, K), 5, 2.0 / math.sqrt(K), torch.bfloat16)
    bias = _rand((N,), 6)
    of, ob = ops.gemm(a, w, bias, act=act, out_f32=True)
    z = a.float() @ w.float().T + bias
    want = z * torch.sigmoid(1.702 * z) if act == "quick_gelu" else F.gelu(z)
    assert rel_err(of, want) < 2e-3 and rel_err(ob, want) < 8e-3


def test_gemm_inplace_residual_and_no_bias():
    from affectgpt_b200 import ops
    M, N, K = 2313, 1024, 1024
    a = _rand((M, K), 7, 1.0, torch.bfloat16)
    w = _rand((N, K), 8, 1.0 / math.sqrt(K), torch.bfloat16)
    bias = _rand((N,), 9)
    resid = _rand((M, N), 10)
    want = resid + a.float() @ w.float().T + bias
    got = ops.gemm_inplace_residual(a, w, bias, resid.clone())
    assert rel_err(got, want) < 1e-4
    of, _ = ops.gemm(a, w, None, out_f32=True, out_bf16=False)
    assert rel_err(of, a.float() @ w.float().T) < 1e-4


def test_gemm_strided_a_rows():
    from affectgpt_b200 import ops
    big = _rand((260, 1024), 11, 1.0, torch.bfloat16)
    a = big[:, :512]                                     # row stride 1024, K = 512 (the k=2 conv "pair" view)
    w = _rand((512, 512), 12, 0.05, torch.bfloat16)
    of, _ = ops.gemm(a, w, None, out_f32=True, out_bf16=False)
    assert rel_err(of, a.float() @ w.float().T) < 1e-4


@pytest.mark.parametrize("nseg,lp", [(1, 400), (3, 200), (2, 3200)])
def test_conv3_s2_as_gemm(nseg, lp):
    """k=3, s=2 Conv1d over channels-last rows == GEMM over the pair view (no im2col buffer)."""
    from affectgpt_b200 import ops
    x = _rand((nseg * lp, 512), 13, 1.0, torch.bfloat16)
    w = _rand((512, 512, 3), 14, 1.0 / math.sqrt(1536))
    bias = _rand((512,), 15)
    wp = w.permute(0, 2, 1).reshape(512, 1536).contiguous().to(torch.bfloat16)      # K index = tap*512 + c_in
    m_out = nseg * lp // 2
    got = ops.conv3_s2(x, wp, bias, m_out).float()
    xf = torch.cat([x.float(), torch.zeros(2, 512, device=DEV)], 0)                 # rows past the end read as zero
    want = F.conv1d(xf.T.unsqueeze(0), wp.float().reshape(512, 3, 512).permute(0, 2, 1), bias, stride=2)[0].T[:m_out]
    assert rel_err(got, want) < 8e-3


def test_posconv_grouped_k128_residual():
    from affectgpt_b200 import ops
    nseq, stride, valid = 3, 100, 99
    e = _rand((nseq * stride, 1024), 16)
    w = _rand((1024, 64, 128), 17, 1.0 / math.sqrt(64 * 128))
    bias = _rand((1024,), 18)
    wp = w.permute(0, 2, 1).reshape(1024, 8192).contiguous().to(torch.bfloat16)     # K index = tap*64 + c_in
    x16 = e.to(torch.bfloat16)
    resid = e.clone()
    ops.posconv_residual(x16, wp, bias, resid, nseq, stride, valid)
    xs = x16.float().reshape(nseq, stride, 1024)[:, :valid]                         # [n, 99, 1024]
    wq = wp.float().reshape(1024, 128, 64).permute(0, 2, 1)
    conv = F.conv1d(xs.transpose(1, 2), wq, bias, padding=64, groups=16)[:, :, :-1]  # HubertSamePadLayer drops the last
    want = e.reshape(nseq, stride, 1024).clone()
    want[:, :valid] += F.gelu(conv).transpose(1, 2)
    got = resid.reshape(nseq, stride, 1024)
    assert rel_err(got[:, :valid], want[:, :valid]) < 5e-3
    assert torch.equal(got[:, valid:], e.reshape(nseq, stride, 1024)[:, valid:])     # padded row untouched


@pytest.mark.parametrize("D", [512, 768, 1024])
def test_layernorm_rows(D):
    from affectgpt_b200 import ops
    x = _rand((777, D), 19, 3.0) + 1.5
    g, b = _rand((D,), 20) * 0.1 + 1, _rand((D,), 21) * 0.1
    ob, of = ops.layernorm(x, g, b, 1e-5, want_f32=True)
    want = F.layer_norm(x, (D,), g, b, 1e-5)
    assert rel_err(of, want) < 1e-5 and rel_err(ob, want) < 5e-3


def test_layernorm_strided_cls_rows():
    from affectgpt_b200 import ops
    x = _rand((6 * 257, 1024), 22, 2.0)
    g, b = _rand((1024,), 23) * 0.1 + 1, _rand((1024,), 24) * 0.1
    ob, _ = ops.layernorm(x, g, b, 1e-5, row_stride=257 * 1024, rows=6)
    want = F.layer_norm(x.reshape(6, 257, 1024)[:, 0], (1024,), g, b, 1e-5)
    assert rel_err(ob, want) < 5e-3


@pytest.mark.parametrize("gelu", [True, False])
def test_layernorm512_bf16(gelu):
    from affectgpt_b200 import ops
    x = _rand((1001, 512), 25, 2.0, torch.bfloat16)
    g, b = _rand((512,), 26) * 0.1 + 1, _rand((512,), 27) * 0.1
    got = ops.layernorm512_bf16(x, g, b, 1e-5, gelu)
    want = F.layer_norm(x.float(), (512,), g, b, 1e-5)
    if gelu:
        want = F.gelu(want)
    assert rel_err(got, want) < 6e-3


@pytest.mark.parametrize("nseq,stride,valid", [(3, 257, 257), (5, 100, 99), (2, 16, 16), (2, 40, 33), (1, 272, 272),
                                               (4, 130, 120), (24, 257, 257), (45, 100, 99), (2, 150, 129), (80, 257, 257),
                                               (3, 260, 256), (100, 100, 99), (2, 112, 112)])
def test_attention_vs_sdpa(nseq, stride, valid):
    from affectgpt_b200 import ops
    qkv = _rand((nseq * stride, 3072), 28, 1.0, torch.bfloat16)
    got = ops.attention(qkv, nseq, stride, valid).float().reshape(nseq, stride, 16, 64)[:, :valid]
    x = qkv.float().reshape(nseq, stride, 3, 16, 64)[:, :valid]
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    want = F.scaled_dot_product_attention(q, k, v, scale=0.125).permute(0, 2, 1, 3)
    assert rel_err(got, want) < 1e-2


@pytest.mark.parametrize("nseq,stride,valid", [(20, 257, 257), (40, 100, 99)])
def test_attention_peaked_rows(nseq, stride, valid):
    """Large logits (|s| up to ~40) and one dominant key per row: the row-max subtraction and the stray 257th key
    / 257th query handling must hold when a single probability carries the row."""
    from affectgpt_b200 import ops
    qkv = _rand((nseq * stride, 3072), 31, 1.0, torch.bfloat16)
    qkv[:, :1024] *= 6.0
    x = qkv.float().reshape(nseq, stride, 3, 16, 64)
    x[:, valid - 1, 1] = x[:, 0, 0] * 0.5            # last key aligned with query 0 of every head
    x[:, valid - 30, 1] = x[:, 5, 0] * 0.5           # a late key dominates query 5 (second key block when valid = 257)
    x[:, 17, 1] = x[:, 9, 0] * 0.5                   # an early key dominates query 9
    x[:, 3, 1] = x[:, valid - 1, 0] * 0.5            # ... and the last query (the odd 257th row)
    qkv = x.reshape(nseq * stride, 3072).to(torch.bfloat16).contiguous()
    got = ops.attention(qkv, nseq, stride, valid).float().reshape(nseq, stride, 16, 64)[:, :valid]
    x = qkv.float().reshape(nseq, stride, 3, 16, 64)[:, :valid]
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    want = F.scaled_dot_product_attention(q, k, v, scale=0.125).permute(0, 2, 1, 3)
    assert torch.isfinite(got).all()
    assert rel_err(got, want) < 1.5e-2


def test_patchify_matches_clip_preprocess_unfold():
    from affectgpt_b200 import ops
    from oracle import reference_path as R
    raw = R.synth_frames(1, 3, seed=5) + 0.7                                        # truncation toward zero matters
    frames = raw.permute(0, 2, 1, 3, 4).reshape(3, 3, 224, 224)
    got = ops.patchify(frames.to(DEV)).float().cpu()
    px = R.clip_preprocess(raw)                                                     # [(b t),3,224,224]
    want = F.unfold(px, kernel_size=14, stride=14).transpose(1, 2).reshape(3 * 256, 588)
    assert got.shape == (768, 640)
    assert float((got[:, :588] - want).abs().max()) < 2e-2                          # bf16 rounding of values in [-2, 2.2]
    assert float(got[:, 588:].abs().max()) == 0.0
    got_u8 = ops.patchify(frames.to(torch.uint8).to(DEV)).float().cpu()
    assert torch.equal(got_u8, got)


@pytest.mark.parametrize("norm", ["layer", "group"])
def test_conv0_norm_gelu(norm):
    from affectgpt_b200 import ops
    n = 3
    wav = _rand((n, 32000), 29, 0.3) + 0.05
    w, bias = _rand((512, 10), 30, 0.3), _rand((512,), 31, 0.1)
    g, b = _rand((512,), 32) * 0.1 + 1, _rand((512,), 33) * 0.1
    got = ops.conv0(wav, w, bias, g, b, norm=norm, seg_per_norm_group=0).float()
    xn = (wav - wav.mean()) / torch.sqrt(wav.var(unbiased=False) + 1e-7)            # joint statistics (F3)
    c = F.conv1d(xn.unsqueeze(1), w.unsqueeze(1), bias, stride=5)                   # [n,512,6399]
    if norm == "layer":
        c = F.layer_norm(c.transpose(1, 2), (512,), g, b, 1e-5)
    else:
        c = F.group_norm(c, 512, g, b, 1e-5).transpose(1, 2)
    want = F.gelu(c)
    assert got.shape == (n, 6400, 512)
    assert rel_err(got[:, :6399], want) < 8e-3
    assert float(got[:, 6399].abs().max()) == 0.0
    per_clip = ops.conv0(wav, w, bias, g, b, norm=norm, seg_per_norm_group=1).float()
    xn1 = (wav - wav.mean(1, keepdim=True)) / torch.sqrt(wav.var(1, unbiased=False, keepdim=True) + 1e-7)
    c1 = F.conv1d(xn1.unsqueeze(1), w.unsqueeze(1), bias, stride=5)
    c1 = F.layer_norm(c1.transpose(1, 2), (512,), g, b, 1e-5) if norm == "layer" else F.group_norm(c1, 512, g, b, 1e-5).transpose(1, 2)
    assert rel_err(per_clip[:, :6399], F.gelu(c1)) < 8e-3


def test_seq_mean_pool_plain_ln_and_accumulate():
    from affectgpt_b200 import ops
    x = _rand((4 * 100, 1024), 34, 2.0)
    g, b = _rand((1024,), 35) * 0.1 + 1, _rand((1024,), 36) * 0.1
    xs = x.reshape(4, 100, 1024)[:, :99]
    p = ops.seq_mean_pool(x, 4, 100, 99, scale=0.25)
    assert rel_err(p, 0.25 * xs.mean(1)) < 1e-5
    ops.seq_mean_pool(x, 4, 100, 99, g, b, 1e-5, scale=0.25, accumulate_into=p)
    want = 0.25 * xs.mean(1) + 0.25 * F.layer_norm(xs, (1024,), g, b, 1e-5).mean(1)
    assert rel_err(p, want) < 1e-5


def test_linear_f32_ragged():
    from affectgpt_b200 import ops
    x, w, b = _rand((5, 770), 37), _rand((3584, 770), 38, 0.03), _rand((3584,), 39)
    assert rel_err(ops.linear_f32(x, w, b), x @ w.T + b) < 1e-5
    x, w = _rand((67, 1024), 40), _rand((2, 1024), 41, 0.03)
    assert rel_err(ops.linear_f32(x, w, None), x @ w.T) < 1e-5
