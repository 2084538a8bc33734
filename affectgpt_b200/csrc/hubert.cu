// HuBERT front-end / tail kernels (HBM-bound, vectorised, shared-memory staged):
//   wave_stats     : zero-mean/unit-variance statistics over a normalisation group of segments
//                    (HF Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm applied to the whole 2-D array,
//                     reference quirk SURVEY F3; A/my_affectgpt/models/encoder.py:425)
//   hubert_conv0   : Conv1d(1->512, k=10, s=5) + bias, fused with LayerNorm(512)+GELU ("layer" variant)
//                    or GroupNorm(512,512)+GELU ("group" variant)  (HF modeling_hubert.py:127-175)
//   seq_mean_pool  : mean over time of a residual stream (optionally through the encoder's final
//                    LayerNorm), accumulated: the reference's stack(hidden_states)[-4:].mean(0).mean(1)
//                    (encoder.py:428-429) without materialising 25 hidden states.
#include <cstdlib>

#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

constexpr int kL0 = 6399;       // conv0 output length for 32000 samples
constexpr int kL0Pad = 6400;    // padded row count per segment (row 6399 is zero)

// ------------------------------------------------------------------------------------------ stats
__global__ void __launch_bounds__(256) wave_partial_kernel(const float* __restrict__ wav, long long per_group,
                                                          double* __restrict__ acc /*[ngroups,2]*/) {
  const int group = blockIdx.y;
  const float4* p = reinterpret_cast<const float4*>(wav + (long long)group * per_group);
  const long long n4 = per_group >> 2;
  double s = 0.0, ss = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p + i);
    const float a = (v.x + v.y) + (v.z + v.w);
    const float b = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    s += (double)a;
    ss += (double)b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  __shared__ double sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; b += sh[1][i]; }
    atomicAdd(&acc[group * 2 + 0], a);
    atomicAdd(&acc[group * 2 + 1], b);
  }
}

__global__ void wave_finalize_kernel(const double* __restrict__ acc, long long per_group, int ngroups,
                                     int do_normalize, float* __restrict__ stats) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  if (!do_normalize) { stats[2 * g] = 0.f; stats[2 * g + 1] = 1.f; return; }
  const double mean = acc[2 * g] / (double)per_group;
  double var = acc[2 * g + 1] / (double)per_group - mean * mean;  // population variance
  if (var < 0) var = 0;
  stats[2 * g] = (float)mean;
  stats[2 * g + 1] = (float)(1.0 / sqrt(var + 1e-7));
}

// ------------------------------------------------------------------------------------------ conv0
// Warp handles 4 consecutive output rows; lane owns channels [lane*8, +8) and [256 + lane*8, +8) so that
// every store is a 16-byte vector and a warp writes two 512-byte contiguous runs per row.
constexpr int kRowsPerWarp = 4;
constexpr int kWarpsPerBlock = 8;
constexpr int kRowsPerBlock = kRowsPerWarp * kWarpsPerBlock;

// MODE 0: LayerNorm over channels + GELU -> out.  MODE 1: accumulate per-(segment, channel) sum / sumsq
// (GroupNorm statistics pass).  MODE 2: apply GroupNorm (per-channel over time) + GELU -> out.
template <int MODE>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ wav, const float* __restrict__ stats,
                                                   long long seg_per_group, long long seg_offset,
                                                   const float* __restrict__ w,
                                                   const float* __restrict__ bias, const float* __restrict__ gam,
                                                   const float* __restrict__ bet, double* __restrict__ gn_acc,
                                                   bf16* __restrict__ out) {
  __shared__ __align__(16) float sw[10][512];
  __shared__ __align__(16) float sx[kRowsPerBlock * 5 + 8];
  const long long seg = blockIdx.y;
  const int t0 = blockIdx.x * kRowsPerBlock;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 5120; i += 256) {
    const int c = i / 10, j = i % 10;
    sw[j][c] = w[i];  // w is [512][10]
  }
  const long long grp = (seg_offset + seg) / seg_per_group;  // stats are indexed by the call-global group
  const float mean = stats[2 * grp], rstd = stats[2 * grp + 1];
  const float* xs = wav + seg * 32000;
  for (int i = tid; i < kRowsPerBlock * 5 + 5; i += 256) {
    const int pos = t0 * 5 + i;
    sx[i] = pos < 32000 ? (xs[pos] - mean) * rstd : 0.f;
  }
  __syncthreads();

  const int tw = t0 + warp * kRowsPerWarp;
  const int cA = lane * 8, cB = 256 + lane * 8;
  float acc[kRowsPerWarp][16];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cA));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cA + 4));
    const float4 b2 = __ldg(reinterpret_cast<const float4*>(bias + cB));
    const float4 b3 = __ldg(reinterpret_cast<const float4*>(bias + cB + 4));
    const float bb[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[r][i] = bb[i];
  }
  const float* xw = sx + warp * kRowsPerWarp * 5;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    const float4 w0 = *reinterpret_cast<const float4*>(&sw[j][cA]);
    const float4 w1 = *reinterpret_cast<const float4*>(&sw[j][cA + 4]);
    const float4 w2 = *reinterpret_cast<const float4*>(&sw[j][cB]);
    const float4 w3 = *reinterpret_cast<const float4*>(&sw[j][cB + 4]);
    const float ww[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const float x = xw[r * 5 + j];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[r][i] = fmaf(ww[i], x, acc[r][i]);
    }
  }

  if constexpr (MODE == 1) {
    float s[16], ss[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s[i] = 0.f; ss[i] = 0.f; }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      if (tw + r < kL0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { s[i] += acc[r][i]; ss[i] += acc[r][i] * acc[r][i]; }
      }
    }
    double* ga = gn_acc + seg * 1024;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = (i < 8 ? cA : cB - 8) + i;
      atomicAdd(&ga[2 * c], (double)s[i]);
      atomicAdd(&ga[2 * c + 1], (double)ss[i]);
    }
    return;
  }

  float gg[16], bb[16];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gam + cA)), g1 = __ldg(reinterpret_cast<const float4*>(gam + cA + 4));
    const float4 g2 = __ldg(reinterpret_cast<const float4*>(gam + cB)), g3 = __ldg(reinterpret_cast<const float4*>(gam + cB + 4));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(bet + cA)), e1 = __ldg(reinterpret_cast<const float4*>(bet + cA + 4));
    const float4 e2 = __ldg(reinterpret_cast<const float4*>(bet + cB)), e3 = __ldg(reinterpret_cast<const float4*>(bet + cB + 4));
    const float tg[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
    const float tb[16] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z, e2.w, e3.x, e3.y, e3.z, e3.w};
#pragma unroll
    for (int i = 0; i < 16; ++i) { gg[i] = tg[i]; bb[i] = tb[i]; }
  }
  float cm[16], cr[16];
  if constexpr (MODE == 2) {
    const double* ga = gn_acc + seg * 1024;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = (i < 8 ? cA : cB - 8) + i;
      const double m = ga[2 * c] / (double)kL0;
      double var = ga[2 * c + 1] / (double)kL0 - m * m;
      if (var < 0) var = 0;
      cm[i] = (float)m;
      cr[i] = (float)(1.0 / sqrt(var + 1e-5));
    }
  }
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const int t = tw + r;
    if (t >= kL0Pad) continue;
    float y[16];
    if (t >= kL0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = 0.f;   // padded row of the segment
    } else if constexpr (MODE == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s += acc[r][i];
      const float m = warp_sum(s) * (1.0f / 512.f);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float d = acc[r][i] - m; ss += d * d; }
      const float rs = rsqrtf(warp_sum(ss) * (1.0f / 512.f) + 1e-5f);
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = AGB_GELU_ERF((acc[r][i] - m) * rs * gg[i] + bb[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = AGB_GELU_ERF((acc[r][i] - cm[i]) * cr[i] * gg[i] + bb[i]);
    }
    bf16* orow = out + (seg * kL0Pad + t) * 512;
    uint4 o0, o1;
    o0.x = pack_bf16x2(y[0], y[1]); o0.y = pack_bf16x2(y[2], y[3]); o0.z = pack_bf16x2(y[4], y[5]); o0.w = pack_bf16x2(y[6], y[7]);
    o1.x = pack_bf16x2(y[8], y[9]); o1.y = pack_bf16x2(y[10], y[11]); o1.z = pack_bf16x2(y[12], y[13]); o1.w = pack_bf16x2(y[14], y[15]);
    *reinterpret_cast<uint4*>(orow + cA) = o0;
    *reinterpret_cast<uint4*>(orow + cB) = o1;
  }
}

__global__ void __launch_bounds__(256) gelu_bf16_kernel(bf16* x, long long n8) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    uint4 u = reinterpret_cast<uint4*>(x)[i];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h[k]);
      h[k] = __floats2bfloat162_rn(AGB_GELU_ERF(f.x), AGB_GELU_ERF(f.y));
    }
    reinterpret_cast<uint4*>(x)[i] = u;
  }
}

// ------------------------------------------------------------------------------------------ pooling
// grid (nseq, 4): block handles 256 channels of one sequence; warp w sums rows w, w+8, ...; when LN is
// requested each row's statistics are computed by a dedicated pass (block-wide) first.
__global__ void __launch_bounds__(256) seq_pool_kernel(const float* __restrict__ x, int seq_stride, int seq_valid,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float eps, float scale, int accumulate,
                                                      float* __restrict__ pooled) {
  extern __shared__ float sh[];  // [seq_valid][2] mean, rstd (LN mode) then [8][256] partials
  float* stat = sh;
  float* part = sh + 2 * seq_valid;
  const long long seq = blockIdx.x;
  const int cblk = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xs = x + seq * seq_stride * 1024ll;
  const bool ln = gamma != nullptr;
  if (ln) {
    for (int r = warp; r < seq_valid; r += 8) {
      const float4* xr = reinterpret_cast<const float4*>(xs + (long long)r * 1024);
      float4 v[8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = xr[lane + 32 * i]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
      const float m = warp_sum(s) * (1.0f / 1024.f);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = v[i].x - m, b = v[i].y - m, c = v[i].z - m, d = v[i].w - m;
        ss += (a * a + b * b) + (c * c + d * d);
      }
      const float rs = rsqrtf(warp_sum(ss) * (1.0f / 1024.f) + eps);
      if (lane == 0) { stat[2 * r] = m; stat[2 * r + 1] = rs; }
    }
    __syncthreads();
  }
  // each warp: 32 lanes x 8 channels = 256 channels of this block
  const int c0 = cblk * 256 + lane * 8;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int r = warp; r < seq_valid; r += 8) {
    const float4 v0 = *reinterpret_cast<const float4*>(xs + (long long)r * 1024 + c0);
    const float4 v1 = *reinterpret_cast<const float4*>(xs + (long long)r * 1024 + c0 + 4);
    float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (ln) {
      const float m = stat[2 * r], rs = stat[2 * r + 1];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (f[i] - m) * rs;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += f[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[warp * 256 + lane * 8 + i] = a[i];
  __syncthreads();
  {
    const int c = tid;  // 256 threads -> 256 channels
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w * 256 + c];
    float y = s / (float)seq_valid;
    const int gc = cblk * 256 + c;
    if (ln) y = y * gamma[gc] + beta[gc];   // mean_t(LN(x)) = gamma * mean_t(xhat) + beta
    y *= scale;
    float* o = pooled + seq * 1024 + gc;
    *o = accumulate ? (*o + y) : y;
  }
}

}  // namespace

int wave_stats(const float* wav, int64_t nseg, int64_t samples, int64_t seg_per_group, float* stats, int do_normalize,
               double* acc /*[ngroups,2] scratch*/, cudaStream_t s) {
  AGB_REQUIRE(nseg > 0 && samples == 32000, "wave_stats: segments must be 32000 samples");
  if (seg_per_group <= 0) seg_per_group = nseg;
  AGB_REQUIRE(nseg % seg_per_group == 0, "wave_stats: segment count must be a multiple of the normalisation group");
  const int ngroups = (int)(nseg / seg_per_group);
  AGB_REQUIRE(ngroups <= 65535, "wave_stats: too many groups");
  AGB_REQUIRE(acc != nullptr, "wave_stats: scratch missing");
  AGB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * ngroups, s));
  const long long per_group = seg_per_group * samples;
  int bx = (int)((per_group / 4 + 255) / 256);
  if (bx > 64) bx = 64;
  if (do_normalize) {
    wave_partial_kernel<<<dim3(bx, ngroups), 256, 0, s>>>(wav, per_group, acc);
    AGB_CHECK_LAUNCH();
  }
  wave_finalize_kernel<<<(ngroups + 127) / 128, 128, 0, s>>>(acc, per_group, ngroups, do_normalize, stats);
  AGB_CHECK_LAUNCH();
  return 0;
}

int hubert_conv0(const float* wav, const float* stats, int64_t seg_per_group, int64_t seg_offset, int64_t nseg,
                 const float* w,
                 const float* bias, const float* g, const float* b, int norm_mode, float* gn_scratch, bf16* out,
                 cudaStream_t s) {
  AGB_REQUIRE(nseg > 0 && nseg <= 65535, "conv0: bad segment count");
  if (seg_per_group <= 0) seg_per_group = seg_offset + nseg;
  LaunchTimer timer(PROF_FRONTEND, 2.0 * nseg * 6399.0 * 512 * 10, (double)nseg * (32000.0 * 4 + 6400.0 * 512 * 2), s);
  dim3 grid((kL0Pad + kRowsPerBlock - 1) / kRowsPerBlock, (unsigned)nseg);
  if (norm_mode == 0) {
    // "layer" variant: tensor-core kernel (conv0_tc.cu); gn_scratch doubles as its 32 KB packed-filter buffer
    AGB_REQUIRE(gn_scratch != nullptr, "conv0: scratch missing");
    return hubert_conv0_layer_tc(wav, stats, seg_per_group, seg_offset, nseg, w, bias, g, b,
                                 reinterpret_cast<bf16*>(gn_scratch), out, s);
  } else {
    AGB_REQUIRE(gn_scratch != nullptr, "conv0: group-norm scratch missing");
    double* acc = reinterpret_cast<double*>(gn_scratch);
    AGB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 1024 * nseg, s));
    conv0_kernel<1><<<grid, 256, 0, s>>>(wav, stats, seg_per_group, seg_offset, w, bias, g, b, acc, out);
    AGB_CHECK_LAUNCH();
    conv0_kernel<2><<<grid, 256, 0, s>>>(wav, stats, seg_per_group, seg_offset, w, bias, g, b, acc, out);
  }
  AGB_CHECK_LAUNCH();
  return 0;
}

int gelu_bf16_inplace(bf16* x, int64_t n, cudaStream_t s) {
  AGB_REQUIRE(n % 8 == 0, "gelu: element count must be a multiple of 8");
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  gelu_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n / 8);
  AGB_CHECK_LAUNCH();
  return 0;
}

int seq_mean_pool(const float* x, int nseq, int seq_stride, int seq_valid, const float* gamma, const float* beta,
                  float eps, float scale, int accumulate, float* pooled, cudaStream_t s) {
  AGB_REQUIRE(nseq > 0 && seq_valid > 0 && seq_valid <= seq_stride, "pool: bad shape");
  LaunchTimer timer(PROF_FRONTEND, 0.0, (double)nseq * seq_valid * 1024 * 4.0 * (gamma ? 2 : 1), s);
  const size_t smem = sizeof(float) * (2 * seq_valid + 8 * 256);
  AGB_REQUIRE(smem <= 48 * 1024, "pool: sequence too long");
  seq_pool_kernel<<<dim3(nseq, 4), 256, smem, s>>>(x, seq_stride, seq_valid, gamma, beta, eps, scale, accumulate, pooled);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
