// Fused LayerNorm kernels (warp per row, warp-shuffle reductions, 16-byte vector I/O; HBM-bound).
// Replaces nn.LayerNorm calls in HF CLIPEncoderLayer (modeling_clip.py:359-380), pre/post layernorm
// (:677, :686), HubertEncoderLayerStableLayerNorm (modeling_hubert.py:505-548), the HuBERT conv-stack
// LayerNorm + GELU (:150-175) and feature projection LayerNorm (:216-231).
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

// NV = D / 128 float4 vectors per lane
template <int NV>
__global__ void __launch_bounds__(256) ln_f32_kernel(const float* __restrict__ x, long long ldx, long long rows,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float eps, bf16* __restrict__ ob, long long ldb,
                                                    float* __restrict__ of, long long ldf, int reverse) {
  const long long r0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r0 >= rows) return;
  const long long row = reverse ? rows - 1 - r0 : r0;       // last rows first: they are what the producer left in L2
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  ln_row_apply<NV>(v, lane, gamma, beta, eps, ob != nullptr ? ob + row * ldb : nullptr, of != nullptr ? of + row * ldf : nullptr);
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// rows of 512 bf16 channels: y = [gelu](LN(x)); warp per row, lane owns 2 x 8 channels
// Two rows per warp: both rows' loads are in flight before the first reduction (a row is only 1 KB, so one row per warp
// leaves too few bytes in flight per SM to cover the HBM latency: 4.5 -> 6.1 TB/s, tools/frontend_bench.py).
__global__ void __launch_bounds__(256) ln512_bf16_kernel(const bf16* x, long long rows,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int apply_gelu,
                                                        bf16* out) {  // x may alias out
  const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2;
  if (row0 >= rows) return;
  const bool two = row0 + 1 < rows;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row0 * 512);
  uint4 u[4];
  u[0] = xr[lane];
  u[1] = xr[lane + 32];
  u[2] = two ? xr[lane + 64] : make_uint4(0u, 0u, 0u, 0u);
  u[3] = two ? xr[lane + 96] : make_uint4(0u, 0u, 0u, 0u);
  float f[2][16];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    unpack8(u[2 * r], f[r]);
    unpack8(u[2 * r + 1], f[r] + 8);
  }
  float s[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 16; ++i) s[r] += f[r][i];
  float mean[2], rstd[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) mean[r] = warp_sum(s[r]) * (1.0f / 512.f);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = f[r][i] - mean[r];
      ss += d * d;
    }
    rstd[r] = rsqrtf(warp_sum(ss) * (1.0f / 512.f) + eps);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c0 = (lane + 32 * h) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        y[i] = (f[r][h * 8 + i] - mean[r]) * rstd[r] * gg[i] + bb[i];
        if (apply_gelu) y[i] = gelu_erf_tanhform(y[i]);
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]);
      o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]);
      o.w = pack_bf16x2(y[6], y[7]);
      reinterpret_cast<uint4*>(out + (row0 + r) * 512)[lane + 32 * h] = o;
    }
  }
}

}  // namespace

int layernorm_f32_in(const float* x, int64_t ldx, int64_t rows, int D, const float* gamma, const float* beta, float eps,
                     bf16* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32, cudaStream_t s, bool reverse) {
  AGB_REQUIRE(rows > 0, "layernorm: no rows");
  AGB_REQUIRE(D == 512 || D == 768 || D == 1024, "layernorm: D must be 512, 768 or 1024");
  AGB_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "layernorm: input alignment");
  AGB_REQUIRE(out_bf16 != nullptr || out_f32 != nullptr, "layernorm: no output");
  if (out_bf16) AGB_REQUIRE(ld_bf16 % 4 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0, "layernorm: bf16 out alignment");
  if (out_f32) AGB_REQUIRE(ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0, "layernorm: f32 out alignment");
  LaunchTimer timer(PROF_LAYERNORM, 8.0 * rows * D, (double)rows * D * (4.0 + (out_bf16 ? 2.0 : 0.0) + (out_f32 ? 4.0 : 0.0)), s);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (D == 1024)
    ln_f32_kernel<8><<<grid, 256, 0, s>>>(x, ldx, rows, gamma, beta, eps, out_bf16, ld_bf16, out_f32, ld_f32, reverse ? 1 : 0);
  else if (D == 768)
    ln_f32_kernel<6><<<grid, 256, 0, s>>>(x, ldx, rows, gamma, beta, eps, out_bf16, ld_bf16, out_f32, ld_f32, reverse ? 1 : 0);
  else
    ln_f32_kernel<4><<<grid, 256, 0, s>>>(x, ldx, rows, gamma, beta, eps, out_bf16, ld_bf16, out_f32, ld_f32, reverse ? 1 : 0);
  AGB_CHECK_LAUNCH();
  return 0;
}

int layernorm512_bf16_gelu(bf16* x, int64_t rows, const float* gamma, const float* beta, float eps, bool apply_gelu,
                           bf16* out, cudaStream_t s) {
  AGB_REQUIRE(rows > 0, "layernorm512: no rows");
  LaunchTimer timer(PROF_LAYERNORM, 8.0 * rows * 512, (double)rows * 512 * 4.0, s);
  const unsigned grid = (unsigned)((rows + 15) / 16);
  ln512_bf16_kernel<<<grid, 256, 0, s>>>(x, rows, gamma, beta, eps, apply_gelu ? 1 : 0, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
