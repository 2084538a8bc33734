// ViT front-end kernels: CLIP preprocessing fused with im2col, and CLS/position-embedding assembly
// fused with pre_layrnorm.
//
// Replaces (reference): the CPU round trip of CLIP_VIT_LARGE.forward -- per-frame .cpu() -> PIL ->
// HF CLIP processor -> .to(device) (A/my_affectgpt/models/encoder.py:34-39, 200-204) -- and
// CLIPVisionEmbeddings.forward + pre_layrnorm (HF modeling_clip.py:148-154, 209-217, 677).
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

__constant__ float kClipMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
__constant__ float kClipStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};

// A_patch[(n*256 + py*16 + px), c*196 + ky*14 + kx] = ((u8)frame[n,c,py*14+ky,px*14+kx] / 255 - mean[c]) / std[c]
// columns 588..639 are zero padding so that K is a multiple of 64 (the packed conv weight is padded too).
template <typename TIn>
__global__ void __launch_bounds__(256) patchify_kernel(const TIn* __restrict__ frames, long long nframes,
                                                      bf16* __restrict__ a_patch) {
  // one thread -> 2 consecutive K columns of one patch row
  const long long total = nframes * 256 * 320;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kp = (int)(idx % 320);
    const long long prow = idx / 320;
    const int patch = (int)(prow & 255);
    const long long n = prow >> 8;
    const int py = patch >> 4, px = patch & 15;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = kp * 2 + e;
      if (k < 588) {
        const int c = k / 196, r = k % 196;
        const int ky = r / 14, kx = r % 14;
        const long long off = ((n * 3 + c) * 224 + (py * 14 + ky)) * 224 + (px * 14 + kx);
        float x = (float)frames[off];
        // Image.fromarray(frame.astype(np.uint8)): truncation toward zero (encoder.py:37)
        x = truncf(x);
        x = fminf(fmaxf(x, 0.f), 255.f);
        v[e] = (x * (1.0f / 255.0f) - kClipMean[c]) / kClipStd[c];
      } else {
        v[e] = 0.f;
      }
    }
    reinterpret_cast<uint32_t*>(a_patch)[idx] = pack_bf16x2(v[0], v[1]);
  }
}

// resid[n*257 + tok, :] = LN_pre( (tok == 0 ? cls : patch_out[n*256 + tok - 1, :]) + pos[tok, :] )
__global__ void __launch_bounds__(256) embed_preln_kernel(const float* __restrict__ patch_out,
                                                         const float* __restrict__ cls,
                                                         const float* __restrict__ pos, const float* __restrict__ g,
                                                         const float* __restrict__ b, float eps, long long nframes,
                                                         float* __restrict__ resid) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= nframes * 257) return;
  const int lane = threadIdx.x & 31;
  const long long n = row / 257;
  const int tok = (int)(row % 257);
  const float4* src = reinterpret_cast<const float4*>(tok == 0 ? cls : patch_out + (n * 256 + tok - 1) * 1024);
  const float4* pr = reinterpret_cast<const float4*>(pos + (long long)tok * 1024);
  float4 v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a = src[lane + 32 * i], p = __ldg(pr + lane + 32 * i);
    v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / 1024.f);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 1024.f) + eps);
  float4* orow = reinterpret_cast<float4*>(resid + row * 1024);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    orow[lane + 32 * i] = y;
  }
}

// CLS-query attention of the LAST encoder block: CLIPVisionTransformer pools hidden_state[:, 0] only
// (modeling_clip.py:685-686), so after the last block's K/V projection every other query row is dead work.
// One warp per (frame, head): lanes split the 257 keys for the scores (each lane reads whole 128-byte K rows), exact
// fp32 softmax across the warp, then lanes split the 64 output dims for P.V (coalesced 128-byte V rows).
__global__ void __launch_bounds__(512) cls_attention_kernel(const bf16* __restrict__ q_cls /*[F,1024]*/,
                                                           const bf16* __restrict__ qkv /*[F*257,3072]*/, int nframes,
                                                           float scale, bf16* __restrict__ out /*[F,1024]*/) {
  constexpr int T = 257, KPL = 9;                       // keys per lane: ceil(257 / 32)
  __shared__ float sp[16][T + 3];
  const int frame = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (frame >= nframes) return;
  const bf16* kbase = qkv + (size_t)frame * T * 3072 + 1024 + head * 64;
  const bf16* vbase = kbase + 1024;
  float q[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q_cls + (size_t)frame * 1024 + head * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = __ldg(qp + i);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        q[i * 8 + e * 2] = f.x * scale;
        q[i * 8 + e * 2 + 1] = f.y * scale;
      }
    }
  }
  float sc[KPL];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int key = lane + 32 * j;
    float acc = -INFINITY;
    if (key < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(kbase + (size_t)key * 3072);
      acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 u = __ldg(kp + i);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h[e]);
          acc = fmaf(q[i * 8 + e * 2], f.x, acc);
          acc = fmaf(q[i * 8 + e * 2 + 1], f.y, acc);
        }
      }
    }
    sc[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int key = lane + 32 * j;
    const float e = key < T ? __expf(sc[j] - mx) : 0.f;
    if (key < T) sp[head][key] = e;
    sum += e;
  }
  const float inv = 1.0f / warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  const __nv_bfloat162* vp = reinterpret_cast<const __nv_bfloat162*>(vbase) + lane;
#pragma unroll 4
  for (int key = 0; key < T; ++key) {
    const float2 f = __bfloat1622float2(__ldg(vp + (size_t)key * (3072 / 2)));
    const float pk = sp[head][key];
    o0 = fmaf(pk, f.x, o0);
    o1 = fmaf(pk, f.y, o1);
  }
  reinterpret_cast<__nv_bfloat162*>(out + (size_t)frame * 1024 + head * 64)[lane] = __floats2bfloat162_rn(o0 * inv, o1 * inv);
}

}  // namespace

int vit_patchify(const void* frames, int in_dtype, int64_t nframes, bf16* a_patch, cudaStream_t s) {
  AGB_REQUIRE(nframes > 0, "patchify: no frames");
  AGB_REQUIRE(in_dtype == 0 || in_dtype == 1, "patchify: input dtype must be f32 (0) or u8 (1)");
  LaunchTimer timer(PROF_FRONTEND, 0.0, (double)nframes * (150528.0 * (in_dtype == 0 ? 4 : 1) + 256.0 * 640 * 2), s);
  const long long total = nframes * 256 * 320;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  if (in_dtype == 0)
    patchify_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const float*>(frames), nframes, a_patch);
  else
    patchify_kernel<uint8_t><<<(unsigned)blocks, 256, 0, s>>>(static_cast<const uint8_t*>(frames), nframes, a_patch);
  AGB_CHECK_LAUNCH();
  return 0;
}

int vit_embed_preln(const float* patch_out, const float* cls, const float* pos, const float* g, const float* b,
                    float eps, int64_t nframes, float* resid, cudaStream_t s) {
  AGB_REQUIRE(nframes > 0, "embed: no frames");
  const long long rows = nframes * 257;
  LaunchTimer timer(PROF_LAYERNORM, 8.0 * rows * 1024, (double)rows * 1024 * 8.0, s);
  embed_preln_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(patch_out, cls, pos, g, b, eps, nframes, resid);
  AGB_CHECK_LAUNCH();
  return 0;
}

int vit_cls_attention(const bf16* q_cls, const bf16* qkv, int64_t nframes, float scale, bf16* out, cudaStream_t s) {
  AGB_REQUIRE(nframes > 0, "cls attention: no frames");
  LaunchTimer timer(PROF_ATTENTION, 4.0 * nframes * 16 * 257 * 64, (double)nframes * 257 * 2048 * 2.0, s);
  cls_attention_kernel<<<(unsigned)nframes, 512, 0, s>>>(q_cls, qkv, (int)nframes, scale, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
