// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM).
// Hand-written for this repo; bit layouts follow the PTX ISA "tcgen05" chapter (shared-memory
// matrix descriptor, instruction descriptor for .kind::f16).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>

namespace agb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
// Warp index as a value ptxas can PROVE warp-uniform (a shuffle broadcast), unlike threadIdx.x >> 5.  It matters for every
// instruction that takes uniform-register operands (tcgen05.mma, tcgen05.commit, cp.async.bulk.tensor, mbarrier addresses):
// when their operands are computed under control flow that ptxas must assume divergent (`if (threadIdx.x == 32)`), each
// issue is wrapped in a "waterfall" loop (ELECT + R2UR.BROADCAST per operand + BRA.U.ANY), ~160 clk per tcgen05.mma on
// B200 - more than a 128x256x16 MMA takes to execute.  Role code therefore runs warp-wide on uniform values and only
// the issuing instruction itself sits under the elected lane's predicate.
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
// one lane of the (converged) warp; the same lane every time it is called with the full mask
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must surface as a trapped launch (error code), never as a hung GPU.
// Polling cadence matters more than it looks: the waiters share issue slots with the warps that do the work.  Measured on
// one box with alternating rebuilds (tools/ab_build.sh, C2 step / ViT attention op): this form 119.2 ms / 0.431 ms; a lean
// loop that retries try_wait back to back and reads the watchdog timer only after 32 failures 120.4 ms / 0.490 ms; a 1 ms
// suspend-time hint on try_wait 120.8 ms / 0.483 ms; a nanosleep of 20 or 100 ns between retries (AGB_MBAR_BACKOFF=<ns>, A/B
// builds only) 121.6 / 121.3 ms and 0.47 ms.  The form below stays.
#ifndef AGB_MBAR_BACKOFF
#define AGB_MBAR_BACKOFF 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (AGB_MBAR_BACKOFF > 0) __nanosleep(AGB_MBAR_BACKOFF);
    if (((++spins) & 0xFFu) == 0u && (globaltimer_ns() - t0) > 4000000000ull) {
      printf("agb: mbarrier wait timeout (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------- TMA
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                            int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address (lane<<16 | column) to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major operand tile in shared memory, rows of 64 bf16 (128 B), 128-byte swizzle as written by TMA
// (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major.
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// .kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, no negate / sparsity.
//   [4,6) D fmt (1=f32) | [7,10) A fmt (1=bf16) | [10,13) B fmt | bit15/16 majors (0=K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every tcgen05.mma previously issued by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// warp-collective: lane i of the warp's TMEM lane quarter, 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// ----------------------------------------------------------------------------- clusters / CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// In the shared::cluster window bit 24 selects the CTA of a pair; clearing it addresses CTA 0's copy, so both
// CTAs' TMA loads complete_tx on the LEADER's mbarrier.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1,
                                                 uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B over a CTA pair: M = 256 (128 rows per CTA), B's N rows split across the two
// CTAs' shared memory.  Issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit arriving on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}


// ----------------------------------------------------------------------------- TMA stores (epilogue)
// smem (128B-swizzled box) -> global tile; out-of-bounds rows/cols of the box are clipped by the tensor map
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// global tile += smem tile (element-wise add performed at L2; dtype from the tensor map)
__device__ __forceinline__ void tma_reduce_add_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores committed by this thread have finished READING shared memory (buffer reusable)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk stores / reduce-adds of this thread have been PERFORMED in global memory (not merely read from smem)
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's async-proxy (TMA) global writes before its later generic-proxy operations
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int atom_add_release_gpu(int* p, int v) {
  int old;
  asm volatile("atom.add.release.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
// named barrier over a subset of the CTA's warps (id 1..15; nthreads a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float tanh_approx_fwd(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact-GELU with erf from Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution): 2 MUFU + ~12 FMA
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // MUFU.RCP, 1 ulp
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// exact-GELU without the SFU: erf(z) = z * P(z^2) on |z| <= 2.9 (degree-7 minimax in z^2, |err| < 6e-5, an order of
// magnitude below bf16 resolution of the activations it feeds), clamped beyond (erf(2.9) = 0.99996).  ~14 FMA/ALU
// instructions and no MUFU.  Measured on B200 (tools/gelu_ab.sh): the SFU-bound element-wise kernels gain (conv0 with
// LayerNorm + GELU 2.2x, ln512 + GELU 5 %), the issue-bound GEMM epilogue loses 7 % - so the GEMM keeps gelu_erf_fast.
__device__ __forceinline__ float gelu_erf_poly(float x) {
  const float z = fminf(fmaxf(x * 0.70710678118654752f, -2.9f), 2.9f);
  const float t = z * z;
  float p = fmaf(t, -4.961727313457232e-07f, 1.9877472368534654e-05f);
  p = fmaf(t, p, -0.00034695438807830215f);
  p = fmaf(t, p, 0.0035132169723510742f);
  p = fmaf(t, p, -0.023305216804146767f);
  p = fmaf(t, p, 0.10871681571006775f);
  p = fmaf(t, p, -0.373929500579834f);
  p = fmaf(t, p, 1.1280336380004883f);
  const float h = 0.5f * x;
  return fmaf(h, z * p, h);
}
// erf-GELU through the tanh identity with the inner polynomial refitted to the ERF form:
//   0.5 x (1 + tanh(x (0.797508 + 0.0370056 u - 3.51517e-4 u^2))),  u = min(x^2, 49)
// max |difference| to 0.5 x (1 + erf(x / sqrt 2)) = 2.5e-5 over all x (the classic 0.044715 form is off by 4.7e-4),
// i.e. 150x below the bf16 resolution of the FFN activations it feeds and below tanh.approx's own 2^-11.
// 1 MUFU + 7 FMA/ALU instead of 2 MUFU + 15: the GEMM epilogue warps are issue-bound (tools/gelu_ab.sh on B200:
// HuBERT FFN1 983 -> 1246 TFLOP/s).
__device__ __forceinline__ float gelu_erf_tanhform(float x) {
  const float u = fminf(x * x, 49.0f);
  const float inner = x * fmaf(u, fmaf(u, -3.51516789e-4f, 3.70056460e-2f), 7.97507884e-1f);
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx_fwd(inner), h);
}
#ifndef AGB_GELU_MUFU
#define AGB_GELU_ERF(x) ::agb::gelu_erf_poly(x)
#else
#define AGB_GELU_ERF(x) ::agb::gelu_erf_fast(x)
#endif

// ----------------------------------------------------------------------------- small math helpers
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// CLIP "quick_gelu": x * sigmoid(1.702 x) == x * (0.5 * tanh(0.851 x) + 0.5)
__device__ __forceinline__ float quick_gelu(float x) { return x * fmaf(0.5f, tanh_approx(0.851f * x), 0.5f); }
// exact (erf) GELU as used by HuBERT ("gelu")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// LayerNorm of one 1024-wide fp32 row held by a warp: lane owns float4 v[i] = row[4 * (lane + 32 * i) .. +3], i < NV.
// Two-pass in registers (mean, then centred sum of squares), bf16 and / or fp32 output.  ONE definition on purpose: the
// stand-alone kernels (layernorm.cu) and the LayerNorm tail fused into the residual GEMM (gemm_tcgen05.cu) must produce
// the same bits, so that a batch routed through either path equals its clips one at a time.
template <int NV>
__device__ __forceinline__ void ln_row_apply(const float4 (&v)[NV], int lane, const float* __restrict__ gamma,
                                             const float* __restrict__ beta, float eps, __nv_bfloat16* ob_row,
                                             float* of_row) {
  constexpr int D = NV * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (ob_row != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(y.x, y.y);
      o.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(ob_row)[lane + 32 * i] = o;
    }
    if (of_row != nullptr) reinterpret_cast<float4*>(of_row)[lane + 32 * i] = y;
  }
}

}  // namespace agb
