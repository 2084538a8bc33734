// Weight packing kernels: fp32 checkpoint tensors -> contiguous bf16 operand layouts the GEMM consumes.
// (QKV concatenation is done by writing three row blocks of one buffer; the positional conv's
// weight-norm  w = g * v / ||v||  over dims (0,1) per tap -- HF modeling_hubert.py:60-78 -- is folded here.)
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

__global__ void __launch_bounds__(256) cvt_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ src, long long rows, long long cols,
                                                       bf16* __restrict__ dst, long long ld_dst, long long cols_pad) {
  const long long total = rows * cols_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols_pad, c = i % cols_pad;
    dst[r * ld_dst + c] = c < cols ? __float2bfloat16_rn(src[r * cols + c]) : __float2bfloat16_rn(0.f);
  }
}

// src [512 (out), 512 (in), k] -> dst [512, k*512]: dst[o, tap*512 + i] = src[o, i, tap]
__global__ void __launch_bounds__(256) pack_conv_kernel(const float* __restrict__ src, int k, bf16* __restrict__ dst) {
  const long long total = 512ll * 512 * k;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long o = idx / (512 * k);
    const int rem = (int)(idx % (512 * k));
    const int tap = rem / 512, i = rem % 512;
    dst[idx] = __float2bfloat16_rn(src[(o * 512 + i) * k + tap]);
  }
}

// norm[tap] = sqrt(sum_{o,i} v[o,i,tap]^2); one block per tap
__global__ void __launch_bounds__(256) posconv_norm_kernel(const float* __restrict__ v, float* __restrict__ norm) {
  const int tap = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < 1024 * 64; i += 256) {
    const float x = v[(long long)i * 128 + tap];
    s += (double)x * (double)x;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[tap] = (float)sqrt(sh[0]);
}

// dst[o, tap*64 + i] = g[tap] * v[o, i, tap] / norm[tap]
__global__ void __launch_bounds__(256) posconv_pack_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                          const float* __restrict__ norm, bf16* __restrict__ dst) {
  const long long total = 1024ll * 8192;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long o = idx >> 13;
    const int rem = (int)(idx & 8191);
    const int tap = rem >> 6, i = rem & 63;
    dst[idx] = __float2bfloat16_rn(g[tap] * v[(o * 64 + i) * 128 + tap] / norm[tap]);
  }
}

inline unsigned grid_for(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

int convert_f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s) {
  cvt_kernel<<<grid_for(n), 256, 0, s>>>(src, dst, n);
  AGB_CHECK_LAUNCH();
  return 0;
}

int pack_rows_f32_to_bf16(const float* src, int64_t rows, int64_t cols, bf16* dst, int64_t ld_dst, int64_t cols_pad,
                          cudaStream_t s) {
  pack_rows_kernel<<<grid_for(rows * cols_pad), 256, 0, s>>>(src, rows, cols, dst, ld_dst, cols_pad);
  AGB_CHECK_LAUNCH();
  return 0;
}

int pack_conv_weight(const float* src, int k, bf16* dst, cudaStream_t s) {
  pack_conv_kernel<<<grid_for(512ll * 512 * k), 256, 0, s>>>(src, k, dst);
  AGB_CHECK_LAUNCH();
  return 0;
}

int pack_posconv_weight(const float* v, const float* g, bf16* dst, float* norm_scratch, cudaStream_t s) {
  posconv_norm_kernel<<<128, 256, 0, s>>>(v, norm_scratch);
  AGB_CHECK_LAUNCH();
  posconv_pack_kernel<<<grid_for(1024ll * 8192), 256, 0, s>>>(v, g, norm_scratch, dst);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
