// Q-Former fusion variants (`*_fusion_type: qformer`) of AffectGPT: 2-layer BERT-style query transformer with
// cross-attention to the encoder tokens, followed by the projector to the LLM width.
// No shipped config selects them and they are ~0.5 GFLOP per clip, so they are kept in fp32 SIMT for
// exactness (parity to 1e-4 against the reference's own Qformer.py, tests/golden/reference_fusion.npz).
//
// Replaces (reference): A/my_affectgpt/models/Qformer.py BertEmbeddings.forward :81-109 (query_embeds only),
// BertSelfAttention.forward :173-276, BertSelfOutput :279-290, BertLayer.forward :411-485 (self-attention ->
// cross-attention in every layer -> query FFN), and the callers AffectGPT.encode_video_qformer :420-468 /
// :563-580 (pre-extracted twin), encode_audio_qformer :628-664, encode_multi_qformer :848-881.
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

// y[r, :] = LayerNorm(x[r % x_rows, :] + res[r, :]) * g + b      (one block per row, any D)
__global__ void __launch_bounds__(256) ln_res_kernel(const float* __restrict__ x, int x_rows, const float* __restrict__ res,
                                                    int D, const float* __restrict__ g, const float* __restrict__ b,
                                                    float eps, float* __restrict__ y) {
  __shared__ float red[2][8];
  const long long r = blockIdx.x;
  const float* xr = x + (r % x_rows) * (long long)D;
  const float* rr = res ? res + r * (long long)D : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) s += xr[d] + (rr ? rr[d] : 0.f);
  s = warp_sum(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float mean = 0.f;
  for (int i = 0; i < 8; ++i) mean += red[0][i];
  mean /= (float)D;
  float v = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float t = xr[d] + (rr ? rr[d] : 0.f) - mean;
    v = fmaf(t, t, v);
  }
  v = warp_sum(v);
  if (lane == 0) red[1][warp] = v;
  __syncthreads();
  float var = 0.f;
  for (int i = 0; i < 8; ++i) var += red[1][i];
  const float rstd = rsqrtf(var / (float)D + eps);
  for (int d = threadIdx.x; d < D; d += 256) {
    const float t = xr[d] + (rr ? rr[d] : 0.f);
    y[r * (long long)D + d] = (t - mean) * rstd * g[d] + b[d];
  }
}

// out[b, t0 + t, :] = in[b, t, :] + pos[t0 + t, :]   (writes a slice of a [B, T_total, D] buffer)
__global__ void __launch_bounds__(256) add_pos_kernel(const float* __restrict__ in, const float* __restrict__ pos, int T,
                                                     int D, int t0, int T_total, float* __restrict__ out) {
  const long long bt = blockIdx.x;
  const long long b = bt / T, t = bt % T;
  const float* src = in + bt * D;
  const float* p = pos + (t0 + t) * (long long)D;
  float* dst = out + (b * T_total + t0 + t) * (long long)D;
  for (int d = threadIdx.x; d < D; d += 256) dst[d] = src[d] + p[d];
}

__global__ void __launch_bounds__(256) gelu_erf_kernel(float* __restrict__ x, long long n) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i < n) x[i] = 0.5f * x[i] * (1.0f + erff(x[i] * 0.70710678118654752f));
}

// softmax(q k^T / 8) v for 12 heads x 64: grid (B, 12), one warp per query at a time, fp32 throughout
__global__ void __launch_bounds__(128) mha_small_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                       const float* __restrict__ v, int Tq, int Tk,
                                                       float* __restrict__ out) {
  extern __shared__ float probs[];   // [4 warps][Tk]
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* p = probs + warp * Tk;
  const float* kb = k + ((long long)b * Tk) * 768 + h * 64;
  const float* vb = v + ((long long)b * Tk) * 768 + h * 64;
  for (int i = warp; i < Tq; i += 4) {
    const float* qi = q + ((long long)b * Tq + i) * 768 + h * 64;
    float m = -INFINITY;
    for (int j = lane; j < Tk; j += 32) {
      const float* kj = kb + (long long)j * 768;
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) s = fmaf(qi[d], kj[d], s);
      s *= 0.125f;
      p[j] = s;
      m = fmaxf(m, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float l = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(p[j] - m);
      p[j] = e;
      l += e;
    }
    l = warp_sum(l);
    __syncwarp();
    const float inv = 1.0f / l;
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Tk; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, vb[(long long)j * 768 + lane], o0);
      o1 = fmaf(pj, vb[(long long)j * 768 + 32 + lane], o1);
    }
    float* oi = out + ((long long)b * Tq + i) * 768 + h * 64;
    oi[lane] = o0 * inv;
    oi[32 + lane] = o1 * inv;
    __syncwarp();
  }
}

}  // namespace

int ln_residual_f32(const float* x, int x_rows, const float* res, int64_t rows, int D, const float* g, const float* b,
                    float eps, float* y, cudaStream_t s) {
  AGB_REQUIRE(rows > 0 && D > 0 && x_rows > 0, "ln_residual: bad shape");
  ln_res_kernel<<<(unsigned)rows, 256, 0, s>>>(x, x_rows, res, D, g, b, eps, y);
  AGB_CHECK_LAUNCH();
  return 0;
}

int add_position_f32(const float* in, const float* pos, int B, int T, int D, int t0, int T_total, float* out,
                     cudaStream_t s) {
  AGB_REQUIRE(B > 0 && T > 0 && D > 0 && t0 + T <= T_total, "add_position: bad shape");
  add_pos_kernel<<<(unsigned)(B * T), 256, 0, s>>>(in, pos, T, D, t0, T_total, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

int gelu_erf_f32_inplace(float* x, int64_t n, cudaStream_t s) {
  gelu_erf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n);
  AGB_CHECK_LAUNCH();
  return 0;
}

int mha12x64_f32(const float* q, const float* k, const float* v, int B, int Tq, int Tk, float* out, cudaStream_t s) {
  AGB_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && Tk <= 8192, "mha12x64: bad shape");
  mha_small_kernel<<<dim3(B, 12), 128, sizeof(float) * 4 * Tk, s>>>(q, k, v, Tq, Tk, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

// ---- one Q-Former forward: enc [B, T, We] (position embedding already added) -> x [B, nq, 768]
// ws: x, t0, t1, t2 [B*nq*768] each; kx, vx [B*max(T,nq)*768]; ff [B*nq*3072]
int qformer_forward_f32(const QFormerW& W, const float* enc, int B, int T, QFormerWs& ws, cudaStream_t s) {
  const int nq = W.nq, M = B * nq;
  int r;
  if ((r = ln_residual_f32(W.query, nq, nullptr, M, 768, W.emb_g, W.emb_b, 1e-12f, ws.x, s))) return r;   // :107
  for (int l = 0; l < 2; ++l) {
    const QFLayerW& L = W.layer[l];
    // self-attention over the queries (:425-437)
    if ((r = linear_f32(ws.x, 768, L.sq_w, L.sq_b, ws.t0, 768, M, 768, 768, s))) return r;
    if ((r = linear_f32(ws.x, 768, L.sk_w, L.sk_b, ws.kx, 768, M, 768, 768, s))) return r;
    if ((r = linear_f32(ws.x, 768, L.sv_w, L.sv_b, ws.vx, 768, M, 768, 768, s))) return r;
    if ((r = mha12x64_f32(ws.t0, ws.kx, ws.vx, B, nq, nq, ws.t1, s))) return r;
    if ((r = linear_f32(ws.t1, 768, L.so_w, L.so_b, ws.t2, 768, M, 768, 768, s))) return r;
    if ((r = ln_residual_f32(ws.t2, M, ws.x, M, 768, L.sln_g, L.sln_b, 1e-12f, ws.x, s))) return r;
    // cross-attention to the encoder tokens (:441-456)
    if ((r = linear_f32(ws.x, 768, L.cq_w, L.cq_b, ws.t0, 768, M, 768, 768, s))) return r;
    if ((r = linear_f32(enc, W.enc_width, L.ck_w, L.ck_b, ws.kx, 768, B * T, 768, W.enc_width, s))) return r;
    if ((r = linear_f32(enc, W.enc_width, L.cv_w, L.cv_b, ws.vx, 768, B * T, 768, W.enc_width, s))) return r;
    if ((r = mha12x64_f32(ws.t0, ws.kx, ws.vx, B, nq, T, ws.t1, s))) return r;
    if ((r = linear_f32(ws.t1, 768, L.co_w, L.co_b, ws.t2, 768, M, 768, 768, s))) return r;
    if ((r = ln_residual_f32(ws.t2, M, ws.x, M, 768, L.cln_g, L.cln_b, 1e-12f, ws.x, s))) return r;
    // query feed-forward (:482-485)
    if ((r = linear_f32(ws.x, 768, L.fi_w, L.fi_b, ws.ff, 3072, M, 3072, 768, s))) return r;
    if ((r = gelu_erf_f32_inplace(ws.ff, (int64_t)M * 3072, s))) return r;
    if ((r = linear_f32(ws.ff, 3072, L.fo_w, L.fo_b, ws.t2, 768, M, 768, 3072, s))) return r;
    if ((r = ln_residual_f32(ws.t2, M, ws.x, M, 768, L.fln_g, L.fln_b, 1e-12f, ws.x, s))) return r;
  }
  return 0;
}

}  // namespace agb
