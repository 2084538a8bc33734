// Fused multi-head attention for short, non-causal sequences (ViT: 257 tokens, HuBERT: 99 frames),
// 16 heads x 64, reading the packed QKV GEMM output and writing the head-concatenated context.
//
// One CTA per (sequence, head): K and V of that head live in 128B-XOR-swizzled shared memory for the
// whole CTA, each warp takes 16-query slices, holds the full score row block (16 x S_pad, fp32) in
// registers, does an exact fp32 softmax (no online rescaling needed at these lengths), converts P to
// bf16 in place (accumulator layout == A-operand layout) and multiplies by V.
// mma.sync.m16n8k16 bf16 (HMMA).  NOT on the path's shapes: ViT (257 tokens) and HuBERT (99) run the tcgen05 kernel in
// attention_tc2.cu; this kernel is the general fallback for the lengths that one has no instantiation for (113..128 and
// 258..272 tokens), exercised by tests/test_gpu_kernels.py::test_attention_vs_sdpa.
//
// Replaces: CLIPAttention.forward (modeling_clip.py:262-332, scale 1/8, fp32 softmax) and
// HubertAttention.forward (modeling_hubert.py:262-370).
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: rows of 64 bf16 (128 B = 8 chunks of 16 B); physical chunk = chunk ^ (row & 7)
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// KT16 = number of 16-key blocks held in registers (S_pad = 16 * KT16 >= seq_valid)
template <int KT16>
__global__ void __launch_bounds__(128, 2)
mha_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int seq_stride, int seq_valid, float scale_log2e) {
  constexpr int SP = KT16 * 16;
  constexpr int NT = KT16 * 2;  // 8-key score tiles
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + SP * 128;

  const int head = blockIdx.x;
  const long long seq = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bf16* base = qkv + seq * seq_stride * 3072ll + head * 64;

  // ---- stage K, V of this (sequence, head); rows >= seq_valid are zero (masked keys, 0 * P safe)
  for (int idx = tid; idx < SP * 8; idx += 128) {
    const int row = idx >> 3, ch = idx & 7;
    uint4 k = make_uint4(0, 0, 0, 0), v = make_uint4(0, 0, 0, 0);
    if (row < seq_valid) {
      const bf16* p = base + (long long)row * 3072 + ch * 8;
      k = *reinterpret_cast<const uint4*>(p + 1024);
      v = *reinterpret_cast<const uint4*>(p + 2048);
    }
    *reinterpret_cast<uint4*>(sK + swz(row, ch)) = k;
    *reinterpret_cast<uint4*>(sV + swz(row, ch)) = v;
  }
  __syncthreads();

  const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const int g = lane >> 2, t = lane & 3;
  const int nslices = (seq_valid + 15) >> 4;

  for (int slice = warp; slice < nslices; slice += 4) {
    const int q0 = slice * 16;
    // ---- Q fragments (A operand, 16 x 64): 4 k-steps x 4 regs, straight from global
    uint32_t qa[4][4];
    {
      const int r0 = q0 + g, r1 = q0 + g + 8;
      const bf16* p0 = base + (long long)r0 * 3072;
      const bf16* p1 = base + (long long)r1 * 3072;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c = ks * 16 + 2 * t;
        qa[ks][0] = r0 < seq_valid ? *reinterpret_cast<const uint32_t*>(p0 + c) : 0u;
        qa[ks][1] = r1 < seq_valid ? *reinterpret_cast<const uint32_t*>(p1 + c) : 0u;
        qa[ks][2] = r0 < seq_valid ? *reinterpret_cast<const uint32_t*>(p0 + c + 8) : 0u;
        qa[ks][3] = r1 < seq_valid ? *reinterpret_cast<const uint32_t*>(p1 + c + 8) : 0u;
      }
    }
    // ---- S = Q K^T
    float sc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < KT16; ++kb) {
      const int krow = kb * 16 + (lane & 7) + ((lane >> 4) << 3);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int ch = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sK_u + swz(krow, ch), b0, b1, b2, b3);
        mma_bf16_16816(sc[2 * kb], qa[ks], b0, b1);
        mma_bf16_16816(sc[2 * kb + 1], qa[ks], b2, b3);
      }
    }
    // ---- exact softmax over the row (rows g and g+8 of the slice), fp32
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int key = j * 8 + 2 * t;
      if (key >= seq_valid) { sc[j][0] = -INFINITY; sc[j][2] = -INFINITY; }
      if (key + 1 >= seq_valid) { sc[j][1] = -INFINITY; sc[j][3] = -INFINITY; }
      m0 = fmaxf(m0, fmaxf(sc[j][0], sc[j][1]));
      m1 = fmaxf(m1, fmaxf(sc[j][2], sc[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float o0 = m0 * scale_log2e, o1 = m1 * scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      sc[j][0] = exp2f(fmaf(sc[j][0], scale_log2e, -o0));
      sc[j][1] = exp2f(fmaf(sc[j][1], scale_log2e, -o0));
      sc[j][2] = exp2f(fmaf(sc[j][2], scale_log2e, -o1));
      sc[j][3] = exp2f(fmaf(sc[j][3], scale_log2e, -o1));
      l0 += sc[j][0] + sc[j][1];
      l1 += sc[j][2] + sc[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- O = P V
    float oc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { oc[j][0] = oc[j][1] = oc[j][2] = oc[j][3] = 0.f; }
#pragma unroll
    for (int kb = 0; kb < KT16; ++kb) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(sc[2 * kb][0], sc[2 * kb][1]);
      pa[1] = pack_bf16x2(sc[2 * kb][2], sc[2 * kb][3]);
      pa[2] = pack_bf16x2(sc[2 * kb + 1][0], sc[2 * kb + 1][1]);
      pa[3] = pack_bf16x2(sc[2 * kb + 1][2], sc[2 * kb + 1][3]);
      const int vrow = kb * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        const int ch = dp * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(sV_u + swz(vrow, ch), b0, b1, b2, b3);
        mma_bf16_16816(oc[2 * dp], pa, b0, b1);
        mma_bf16_16816(oc[2 * dp + 1], pa, b2, b3);
      }
    }
    // ---- normalise and store (head-concatenated [row, head*64 + d])
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    bf16* ob = out + seq * seq_stride * 1024ll + head * 64;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j * 8 + 2 * t;
      if (r0 < seq_valid)
        *reinterpret_cast<uint32_t*>(ob + (long long)r0 * 1024 + c) = pack_bf16x2(oc[j][0] * i0, oc[j][1] * i0);
      if (r1 < seq_valid)
        *reinterpret_cast<uint32_t*>(ob + (long long)r1 * 1024 + c) = pack_bf16x2(oc[j][2] * i1, oc[j][3] * i1);
    }
  }
}

template <int KT16>
int launch_mha(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s) {
  constexpr int smem_bytes = KT16 * 16 * 128 * 2;
  auto kern = mha_kernel<KT16>;
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  dim3 grid(16, nseq);
  kern<<<grid, 128, smem_bytes, s>>>(qkv, out, seq_stride, seq_valid, scale * 1.4426950408889634f);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int attention_mha16x64(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale,
                       cudaStream_t s, bool reverse) {
  AGB_REQUIRE(nseq > 0 && seq_valid > 0 && seq_stride >= seq_valid, "attention: bad sequence shape");
  AGB_REQUIRE(nseq <= 65535, "attention: too many sequences for one launch");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "attention: alignment");
  LaunchTimer timer(PROF_ATTENTION, 4.0 * nseq * 16.0 * (double)seq_valid * seq_valid * 64.0,
                    (double)nseq * seq_valid * 4096.0 * 2.0, s);
  // tcgen05 kernel for the shapes of the path (<= 112 and 129..257 tokens: HuBERT 99, ViT 257); the register-resident
  // mma.sync kernel below covers the remaining lengths up to 272 (113..128, 258..272) so that the op stays general
  {
    const int r = attention_tc2(qkv, out, nseq, seq_stride, seq_valid, scale, s, reverse);
    if (r >= 0) return r;
  }
  if (seq_valid <= 272) return launch_mha<17>(qkv, out, nseq, seq_stride, seq_valid, scale, s);
  AGB_REQUIRE(false, "attention: sequence length > 272 not supported by the register-resident softmax kernel");
}

}  // namespace agb
