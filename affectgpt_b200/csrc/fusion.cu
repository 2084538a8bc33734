// Pre-fusion heads and projectors of AffectGPT (shipping "attention" variant: linear, NO softmax).
// ~0.034 GFLOP per clip, kept in fp32 SIMT for exactness.
// Replaces: AffectGPT.encode_video_attention / encode_audio_attention / encode_multi_attention
// (A/my_affectgpt/models/affectgpt.py:505-535, 690-716, 883-908) and their nn.Linear projectors
// (:150, 190, 195, 251, 255, 312-313, 344-345, 349).
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

// out[m, n] = sum_k in[m,k] * W[n,k] + bias[n].  Block: 32 (n) x 64 (m) outputs - a whole batch of clips per block, so W
// is read once - K tiles of 32 in smem, the next tile prefetched into registers while this one is multiplied.  One fp32
// accumulator per output, k ascending (the order of a plain dot product).
constexpr int LT_N = 32, LT_M = 64, LT_K = 32;
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, long long ld_in,
                                                    const float* __restrict__ W, const float* __restrict__ bias,
                                                    float* __restrict__ out, long long ld_out, int M, int N, int K) {
  __shared__ float sW[LT_N][LT_K + 1];
  __shared__ float sI[LT_M][LT_K + 1];
  const int n0 = blockIdx.x * LT_N, m0 = blockIdx.y * LT_M;
  const int tn = threadIdx.x & 31, tm = threadIdx.x >> 5;  // thread -> n = tn, m = tm*8 .. tm*8+7 (a warp shares its rows)
  const int lc = threadIdx.x & 31, lr = threadIdx.x >> 5;  // loader: column lc, rows lr + 8*i
  float acc[8], pw[4], pi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  auto fetch = [&](int k0) {
    const int k = k0 + lc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + lr + 8 * i;
      pw[i] = (n < N && k < K) ? __ldg(W + (long long)n * K + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + lr + 8 * i;
      pi[i] = (m < M && k < K) ? __ldg(in + (long long)m * ld_in + k) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += LT_K) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sW[lr + 8 * i][lc] = pw[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) sI[lr + 8 * i][lc] = pi[i];
    __syncthreads();
    if (k0 + LT_K < K) fetch(k0 + LT_K);
#pragma unroll
    for (int k = 0; k < LT_K; ++k) {
      const float w = sW[tn][k];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(sI[tm * 8 + j][k], w, acc[j]);
    }
    __syncthreads();
  }
  const int n = n0 + tn;
  if (n < N) {
    const float b = bias != nullptr ? bias[n] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = m0 + tm * 8 + j;
      if (m < M) out[(long long)m * ld_out + n] = acc[j] + b;
    }
  }
}

// one block per clip: a_t = H[b,t,:].w + bias ; fused[b,:] = sum_t a_t * H[b,t,:]
__global__ void __launch_bounds__(256) attn_pool_kernel(const float* __restrict__ H, int T, int D,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       int bypass_if_t1, float* __restrict__ fused) {
  extern __shared__ float att[];  // [T]
  const float* Hb = H + (long long)blockIdx.x * T * D;
  float* fb = fused + (long long)blockIdx.x * D;
  if (bypass_if_t1 && T == 1) {  // affectgpt.py:587-589 (pre-extracted single vector)
    for (int d = threadIdx.x; d < D; d += 256) fb[d] = Hb[d];
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < T; t += 8) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(Hb[(long long)t * D + d], w[d], s);
    s = warp_sum(s);
    if (lane == 0) att[t] = s + bias[0];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += 256) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s = fmaf(Hb[(long long)t * D + d], att[t], s);
    fb[d] = s;
  }
}

__global__ void __launch_bounds__(256) mean_t_kernel(const float* __restrict__ H, int T, int D, float* __restrict__ out) {
  const float* Hb = H + (long long)blockIdx.x * T * D;
  for (int d = threadIdx.x; d < D; d += 256) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += Hb[(long long)t * D + d];
    out[(long long)blockIdx.x * D + d] = s / (float)T;
  }
}

__global__ void __launch_bounds__(256) pair_sum_kernel(const float* __restrict__ v, const float* __restrict__ a,
                                                      const float* __restrict__ att, int D, float* __restrict__ out) {
  const long long b = blockIdx.x;
  const float a0 = att[2 * b], a1 = att[2 * b + 1];
  for (int d = threadIdx.x; d < D; d += 256) out[b * D + d] = a0 * v[b * D + d] + a1 * a[b * D + d];
}

__global__ void __launch_bounds__(256) concat2_kernel(const float* __restrict__ a, const float* __restrict__ b, int Da,
                                                     int Db, float* __restrict__ out) {
  const long long r = blockIdx.x;
  for (int d = threadIdx.x; d < Da + Db; d += 256) out[r * (Da + Db) + d] = d < Da ? a[r * Da + d] : b[r * Db + d - Da];
}

}  // namespace

int linear_f32(const float* in, int64_t ld_in, const float* W, const float* bias, float* out, int64_t ld_out, int M, int N,
               int K, cudaStream_t s) {
  AGB_REQUIRE(M > 0 && N > 0 && K > 0, "linear: empty problem");
  LaunchTimer timer(PROF_OTHER, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)N * K + (double)M * N), s);
  dim3 grid((N + LT_N - 1) / LT_N, (M + LT_M - 1) / LT_M);
  linear_kernel<<<grid, 256, 0, s>>>(in, ld_in, W, bias, out, ld_out, M, N, K);
  AGB_CHECK_LAUNCH();
  return 0;
}

int attn_pool_nosoftmax(const float* H, int B, int T, int D, const float* w, const float* bias, int bypass_if_t1,
                        float* fused, cudaStream_t s) {
  AGB_REQUIRE(B > 0 && T > 0 && D > 0 && T <= 4096, "attn_pool: bad shape");
  attn_pool_kernel<<<B, 256, sizeof(float) * T, s>>>(H, T, D, w, bias, bypass_if_t1, fused);
  AGB_CHECK_LAUNCH();
  return 0;
}

int mean_over_t(const float* H, int B, int T, int D, float* out, cudaStream_t s) {
  AGB_REQUIRE(B > 0 && T > 0 && D > 0, "mean_over_t: bad shape");
  mean_t_kernel<<<B, 256, 0, s>>>(H, T, D, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

int weighted_pair_sum(const float* v, const float* a, const float* att, int B, int D, float* out, cudaStream_t s) {
  pair_sum_kernel<<<B, 256, 0, s>>>(v, a, att, D, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

int concat2_f32(const float* a, const float* b, int B, int Da, int Db, float* out, cudaStream_t s) {
  concat2_kernel<<<B, 256, 0, s>>>(a, b, Da, Db, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
