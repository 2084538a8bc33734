// HuBERT conv0 + LayerNorm(512) + GELU on the tensor cores ("layer" feature-extractor variant).
//
//   y[t, c] = b[c] + sum_{j<10} w[c, j] * x[5 t + j]        Conv1d(1 -> 512, k = 10, s = 5), x = normalised waveform
//   out[t, :] = gelu(LayerNorm_512(y[t, :]))                 (HF modeling_hubert.py:127-175, feat_extract_norm="layer")
//
// The SIMT kernel (hubert.cu conv0_kernel<0>) spent 4.3 ms per 512 segments against 0.52 ms of HBM time: 10 FMAs +
// two warp reductions + GELU per output on 171 registers (ncu, profiles/).  Here the 10-tap conv is one tcgen05 MMA per
// 128 output rows - a 128 x 32 im2col tile against the 512 x 32 filter matrix, fp32 accumulators for the whole
// 128 x 512 tile in TMEM (all 512 columns) - and the SM's ALUs only do what tensor cores cannot: LayerNorm and GELU,
// with each thread owning 128 channels of one row so the statistics need no shuffles.
//
// Precision: operands are bf16, so x and the bias travel as hi + lo pairs (K columns 0-9 x_hi, 10 / 11 = 1 against
// b_hi / b_lo, 16-25 x_lo against the same taps): x and b are effectively 16-bit-mantissa exact, only the filter taps
// are rounded to bf16 - as every other weight of the path is.
#include <cuda.h>

#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

constexpr int kL0 = 6399;       // conv0 output length for 32000 samples
constexpr int kL0Pad = 6400;    // padded row count per segment (row 6399 is zero)
constexpr int kTileRows = 128;
constexpr int kThreads = 512;   // 16 warps: warp w -> TMEM lane quarter w & 3, channel group w >> 2 (128 channels)
constexpr int kSmemW = 512 * 128;         // filter matrix, 128-byte rows (K padded to 64 bf16, 32 used), SW128
constexpr int kSmemA = kTileRows * 128;   // im2col tile, same layout
constexpr int kOffA = kSmemW, kOffX = kOffA + kSmemA, kOffG = kOffX + 2688, kOffStat = kOffG + 4096, kOffBar = kOffStat + 4096;
constexpr int kSmem = kOffBar + 64 + 1024;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// filter matrix for the MMA: wp[c][k] bf16, k < 32 (see the header for the column assignment)
__global__ void conv0_pack_kernel(const float* __restrict__ w /*[512,10]*/, const float* __restrict__ bias,
                                  bf16* __restrict__ wp /*[512,32]*/) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= 512) return;
  bf16* o = wp + c * 32;
  for (int k = 0; k < 32; ++k) o[k] = __float2bfloat16_rn(0.f);
  for (int j = 0; j < 10; ++j) {
    const bf16 t = __float2bfloat16_rn(w[c * 10 + j]);
    o[j] = t;
    o[16 + j] = t;
  }
  const float b = bias[c];
  const bf16 bh = __float2bfloat16_rn(b);
  o[10] = bh;
  o[11] = __float2bfloat16_rn(b - __bfloat162float(bh));
}

__global__ void __launch_bounds__(kThreads, 1)
conv0_tc_kernel(const float* __restrict__ wav, const float* __restrict__ stats, long long seg_per_group,
                long long seg_offset, const bf16* __restrict__ wp, const float* __restrict__ gam,
                const float* __restrict__ bet, bf16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smW = smem;
  uint8_t* smA = smem + kOffA;
  float* sx = reinterpret_cast<float*>(smem + kOffX);          // 645 normalised samples of this tile (+ padding)
  float* sg = reinterpret_cast<float*>(smem + kOffG);          // gamma[512] | beta[512]
  float* sstat = reinterpret_cast<float*>(smem + kOffStat);    // [4 channel groups][128 rows][2] partial mean / M2
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
  const long long seg = blockIdx.y;
  const int t0 = blockIdx.x * kTileRows;

  // ---- prologue (no TMEM yet: a co-resident CTA may still own it)
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  {  // filter matrix -> smem, 128B-swizzled rows: 16-byte chunk k of row c lands at chunk (k ^ (c & 7))
    const uint4* src = reinterpret_cast<const uint4*>(wp);
    for (int i = tid; i < 512 * 4; i += kThreads) {
      const int c = i >> 2, k = i & 3;
      *reinterpret_cast<uint4*>(smW + c * 128 + ((k ^ (c & 7)) << 4)) = __ldg(src + i);
    }
  }
  for (int i = tid; i < 1024; i += kThreads) sg[i] = i < 512 ? gam[i] : bet[i - 512];
  {
    const long long grp = (seg_offset + seg) / seg_per_group;    // stats are indexed by the call-global group
    const float mean = stats[2 * grp], rstd = stats[2 * grp + 1];
    const float* xs = wav + seg * 32000;
    for (int i = tid; i < kTileRows * 5 + 8; i += kThreads) {
      const int pos = t0 * 5 + i;
      sx[i] = pos < 32000 ? (xs[pos] - mean) * rstd : 0.f;
    }
  }
  __syncthreads();
  if (tid < kTileRows) {  // im2col row tid: [x_hi(10) | 1 1 | 0 0 0 0 | x_lo(10) | 0 ...] as four swizzled 16-byte chunks
    const int r = tid;
    float x[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) x[j] = (t0 + r < kL0) ? sx[r * 5 + j] : 0.f;
    float hi[10], lo[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      hi[j] = __bfloat162float(__float2bfloat16_rn(x[j]));
      lo[j] = x[j] - hi[j];
    }
    const float one = (t0 + r < kL0) ? 1.f : 0.f;                  // rows past the segment end stay exactly zero
    uint4 c0, c1, c2, c3;
    c0.x = pack_bf16x2(hi[0], hi[1]); c0.y = pack_bf16x2(hi[2], hi[3]); c0.z = pack_bf16x2(hi[4], hi[5]); c0.w = pack_bf16x2(hi[6], hi[7]);
    c1.x = pack_bf16x2(hi[8], hi[9]); c1.y = pack_bf16x2(one, one); c1.z = 0u; c1.w = 0u;
    c2.x = pack_bf16x2(lo[0], lo[1]); c2.y = pack_bf16x2(lo[2], lo[3]); c2.z = pack_bf16x2(lo[4], lo[5]); c2.w = pack_bf16x2(lo[6], lo[7]);
    c3.x = pack_bf16x2(lo[8], lo[9]); c3.y = 0u; c3.z = 0u; c3.w = 0u;
    uint8_t* row = smA + r * 128;
    const int s = r & 7;
    *reinterpret_cast<uint4*>(row + ((0 ^ s) << 4)) = c0;
    *reinterpret_cast<uint4*>(row + ((1 ^ s) << 4)) = c1;
    *reinterpret_cast<uint4*>(row + ((2 ^ s) << 4)) = c2;
    *reinterpret_cast<uint4*>(row + ((3 ^ s) << 4)) = c3;
  }
  fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
  __syncthreads();

  // ---- TMEM (all 512 columns) and the MMAs: D[128, 512] = A[128, 32] . W[512, 32]^T
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {                     // operands computed warp-wide (uniform registers), one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 256);
    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smA));
    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smW));
    if (elect_one()) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16_ss(tmem_base + h * 256, da + 2 * k, db + h * ((256 * 128) >> 4) + 2 * k, idesc, k != 0);
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0, 41);
  tc_fence_after();

  // ---- LayerNorm + GELU: thread = (row, 128-channel group)
  const int q = warp & 3, cg = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + cg * 128;
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    if (c == 0) shift = __uint_as_float(v[0]);
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float d0 = __uint_as_float(v[i]) - shift, d1 = __uint_as_float(v[i + 1]) - shift;
      a1 += d0; a2 = fmaf(d0, d0, a2);
      b1 += d1; b2 = fmaf(d1, d1, b2);
    }
    s1 += a1 + b1;
    s2 += a2 + b2;
  }
  {  // per-thread mean / M2 over its 128 channels, combined over the 4 channel groups (Chan et al.)
    const float mean_i = shift + s1 * (1.f / 128.f);
    const float m2_i = fmaxf(s2 - s1 * s1 * (1.f / 128.f), 0.f);
    sstat[(cg * 128 + r) * 2] = mean_i;
    sstat[(cg * 128 + r) * 2 + 1] = m2_i;
  }
  __syncthreads();
  float mean = 0.f, m2 = 0.f;
#pragma unroll
  for (int g = 0; g < 4; ++g) mean += sstat[(g * 128 + r) * 2];
  mean *= 0.25f;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float d = sstat[(g * 128 + r) * 2] - mean;
    m2 += sstat[(g * 128 + r) * 2 + 1] + 128.f * d * d;
  }
  const float rs = rsqrtf(m2 * (1.f / 512.f) + 1e-5f);
  const float nm = -mean * rs;
  const int t = t0 + r;
  bf16* orow = out + (seg * kL0Pad + t) * 512 + cg * 128;
  const float* g4 = sg + cg * 128;
  const float* b4 = sg + 512 + cg * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    if (t < kL0Pad) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const float4 ga = *reinterpret_cast<const float4*>(g4 + c * 32 + i), gb = *reinterpret_cast<const float4*>(g4 + c * 32 + i + 4);
        const float4 ba = *reinterpret_cast<const float4*>(b4 + c * 32 + i), bb = *reinterpret_cast<const float4*>(b4 + c * 32 + i + 4);
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          y[e] = t < kL0 ? gelu_erf_tanhform(fmaf(fmaf(__uint_as_float(v[i + e]), rs, nm), gg[e], be[e])) : 0.f;
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]); o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(orow + c * 32 + i) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// wp_scratch: [512*32] bf16, rebuilt per call from the fp32 filter + bias (3 us)
int hubert_conv0_layer_tc(const float* wav, const float* stats, int64_t seg_per_group, int64_t seg_offset, int64_t nseg,
                          const float* w, const float* bias, const float* g, const float* b, bf16* wp_scratch, bf16* out,
                          cudaStream_t s) {
  AGB_REQUIRE(nseg > 0 && nseg <= 65535 && wp_scratch != nullptr, "conv0_tc: bad arguments");
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(conv0_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  conv0_pack_kernel<<<4, 128, 0, s>>>(w, bias, wp_scratch);
  AGB_CHECK_LAUNCH();
  dim3 grid(kL0Pad / kTileRows, (unsigned)nseg);
  conv0_tc_kernel<<<grid, kThreads, kSmem, s>>>(wav, stats, seg_per_group, seg_offset, wp_scratch, g, b, out);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
