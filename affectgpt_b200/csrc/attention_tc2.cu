// tcgen05 multi-head attention: S, P and O live in TMEM, two co-resident CTAs per SM, eight softmax warps per CTA.
//
// Shapes on the path: ViT 257 tokens (16 heads x 64) and HuBERT 99 frames; non-causal, K/V of a head fit in smem.
// Per (frame, head) the tensor cores need ~2064 clk; the softmax needs 257^2 exponentials plus ~3 more instructions
// per score.  What the measurements say (profiles/r01_ncu_mha_tc2_summary.txt): the kernel is
// NOT bound by the SFU - routing half of the exponentials through an FMA-pipe polynomial made it 30 % slower - but by
// how many softmax warps are runnable per SM sub-partition while the tile's serial chain
// Q.K^T -> max -> exp -> P.V -> drain is in flight.  Hence:
//
//   * TMEM budget 256 columns per CTA -> 2 CTAs per SM: while one CTA's tensor work runs the other one exponentiates,
//     and every sub-partition has 4 softmax warps to pick from.
//   * softmax group A (warps 4-7) and group B (warps 8-11) split the key columns of the same 128 query rows
//     (a warp may only touch TMEM lanes 32*(warp%4)..+31, so A/B warps with the same warp%4 share rows);
//     row max and row sum are exchanged through shared memory.
//   * 257 = 2*128 + 1.  The odd KEY is a 16-row "stray" block: its scores S_C = Q K_C^T (N = 16) are produced one tile
//     ahead into spare TMEM columns and held in a register, its probability goes through one extra K=16 PV step.
//     The odd QUERY (row 256 of a 257-token sequence) gets no tile of its own.  A clock64 trace (AGB_ATTN_TRACE=1) showed
//     a third tile with one valid row costs as much as a full one (6.9 K of an item's 20 K clk): the one warp that owns
//     the row still walks all 272 columns at single-warp latency.  Instead the 256 softmax threads compute that row on
//     the CUDA cores from the K / V tiles already in shared memory, inside the two windows where they would otherwise
//     wait for the tensor cores: one key per thread (64-wide dot product) + block softmax while Q.K^T of the first tile
//     runs, and the 257 x 64 weighted sum of V while its P.V runs.  (A first attempt that gave the whole row to the two
//     idle control warps was correct but no faster: 64 threads were too slow and held K / V back.)
//     Sequences of 129..256 tokens have no odd query; other odd lengths would use the rotating third tile below.
//
//   warp 0 : TMA producer (K, V single-buffered per head; Q ring of 2)       warp 1 : tcgen05.mma issuer
//   warp 2 : TMEM allocator                                                   warps 4-11 : softmax + epilogue
//
// TMEM columns (ViT, NMAIN = 256):  S [0,256)  ->  P_A [0,64) | O [64,128) | P_B [128,192) | P_C [192,200) | S_C(next) [208,224)
//              (HuBERT, NMAIN = 112): S [0,112) ->  P_A [0,32) | P_B [64,88) ;  O [128,192)
// P (bf16 pairs) overwrites the first half of the S columns its own group has already consumed; O reuses the second
// half of group A's S columns, which are dead once P_A is complete.
//
// Exact softmax: global row max (fp32) before the first exponential, exp2 with the scale folded in, fp32 row sum of
// the unrounded probabilities, P rounded to bf16 for the PV product (same as every flash-attention implementation).
//
// Replaces CLIPAttention.forward (HF modeling_clip.py:262-332, scale 1/8, fp32 softmax) and
// HubertAttention.forward (HF modeling_hubert.py:262-370).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

int make_tmap_generic(CUtensorMap* out, int dtype_is_f32, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, uint32_t box0, uint32_t box1);

namespace {

constexpr int kThreads = 384;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// barrier 1: the 256 softmax threads only (warps 4-11)
__device__ __forceinline__ void softmax_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// V tile [keys][64 d] as the B operand of P.V: N (= d) contiguous -> MN-major, 128B swizzle, 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t m, uint32_t n, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct AParams {
  int nseq, seq_stride, seq_valid, qtiles;
  int reverse;             // items (sequence, head) from the last to the first
  int stray_q;             // 1: query row NMAIN of every sequence is computed by the softmax warps on the CUDA cores
  float scale_log2e;
  const bf16* qkv;
  bf16* out;
  long long* trace;        // AGB_ATTN_TRACE=1: clock64 stamps of the first 64 tiles of CTA 0 (tools/attn_trace.py), else null
};

template <int NMAIN, bool STRAY>
struct Cfg {
  static constexpr int CA = NMAIN == 256 ? 128 : 64;       // key columns of softmax group A
  static constexpr int CB = NMAIN - CA;                     // ... of group B
  static constexpr int NKROWS = NMAIN + (STRAY ? 16 : 0);   // K / V rows staged per head
  static constexpr int kKVBytes = NKROWS * 128;
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kColPA = 0, kColPB = CA;
  static constexpr int kColO = NMAIN == 256 ? 64 : 128;
  static constexpr int kColPC = 192, kColSC = 208;
  static constexpr int kTmemCols = 256;
  static constexpr int kOffV = kKVBytes, kOffQ = 2 * kKVBytes, kOffX = kOffQ + 2 * kQBytes;
  static constexpr int kOffY = kOffX + 2048;                // odd query row: p[NKROWS] | 16 reduction slots | q[64] | 8 x 64 partial outputs
  static constexpr int kYBytes = STRAY ? ((NKROWS + 16 + 64 + 512) * 4 + 127) / 128 * 128 : 0;
  static constexpr int kOffBar = kOffY + kYBytes;
  static constexpr int kSmem = kOffBar + 256 + 1024;        // + alignment slack
  static_assert(kKVBytes % 1024 == 0, "swizzle atoms need 1024-byte aligned tiles");
  static_assert(CA % 32 == 0 && CB % 16 == 0, "column split");
};

// ---- softmax pass 1 over W (32 or 16) score columns: running max, optionally masked to keys < valid
template <int W>
__device__ __forceinline__ float chunk_max(uint32_t taddr, int key0, int valid, float m) {
  uint32_t v[W];
  if (W == 32) tmem_ld_x32(taddr, v); else tmem_ld_x16(taddr, v);
  tmem_ld_wait();
  if (key0 + W <= valid) {
#pragma unroll
    for (int j = 0; j < W; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j) m = (key0 + j < valid) ? fmaxf(m, __uint_as_float(v[j])) : m;
  }
  return m;
}
// ---- softmax pass 2: p = exp2(s*sc - mo), row-sum, bf16 pairs stored at p_addr (W/2 columns)
template <int W>
__device__ __forceinline__ float chunk_exp(uint32_t s_addr, uint32_t p_addr, int key0, int valid, float sc, float mo,
                                           float l) {
  uint32_t v[W];
  if (W == 32) tmem_ld_x32(s_addr, v); else tmem_ld_x16(s_addr, v);
  tmem_ld_wait();
  float l1 = 0.f;
  if (key0 + W <= valid) {
    // packed fp32x2 FMA / ADD (Blackwell FFMA2 / FADD2): 2.75 instead of 3.5 issue slots per exponential
    uint64_t sc2, mo2, acc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sc2) : "f"(sc));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mo2) : "f"(-mo));
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(l), "f"(l1));
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      uint64_t s2, x2, p2;
      asm("mov.b64 %0, {%1, %2};" : "=l"(s2) : "r"(v[j]), "r"(v[j + 1]));
      asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x2) : "l"(s2), "l"(sc2), "l"(mo2));
      float x0, x1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x2));
      const float p0 = ex2_approx(x0);
      const float p1 = ex2_approx(x1);
      asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(p0), "f"(p1));
      asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(acc), "l"(p2));
      v[j >> 1] = pack_bf16x2(p0, p1);
    }
    asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(l1) : "l"(acc));
  } else {
#pragma unroll
    for (int j = 0; j < W; j += 2) {
      float p0 = ex2_approx(fmaf(__uint_as_float(v[j]), sc, -mo));
      float p1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), sc, -mo));
      p0 = (key0 + j < valid) ? p0 : 0.f;
      p1 = (key0 + j + 1 < valid) ? p1 : 0.f;
      l += p0;
      l1 += p1;
      v[j >> 1] = pack_bf16x2(p0, p1);
    }
  }
  if (W == 32) tmem_st_x16(p_addr, v); else tmem_st_x8(p_addr, v);
  return l + l1;
}

template <int NMAIN, bool STRAY>
__global__ void __launch_bounds__(kThreads, 2)
mha_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
               const __grid_constant__ CUtensorMap tmStray, const AParams p) {
  using C = Cfg<NMAIN, STRAY>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smK = smem;
  uint8_t* smV = smem + C::kOffV;
  uint8_t* smQ = smem + C::kOffQ;
  float* xm = reinterpret_cast<float*>(smem + C::kOffX);   // [2 groups][128 rows] partial row max
  float* xl = xm + 256;                                    // [2][128] partial row sum
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* q_full = bars + 4;     // [2]
  uint64_t* q_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;
  uint64_t* sc_full = bars + 9;
  uint64_t* sc_read = bars + 10;   // 8 warps
  uint64_t* pa_ready = bars + 11;  // 4 warps
  uint64_t* pb_ready = bars + 12;  // 4 warps
  uint64_t* o_full = bars + 13;
  uint64_t* o_read = bars + 14;    // 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;   // provably warp-uniform: see agb_ptx.cuh

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    if (STRAY) tma_prefetch_desc(&tmStray);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1); mbar_init(&q_empty[0], 1); mbar_init(&q_empty[1], 1);
    mbar_init(s_full, 1); mbar_init(sc_full, 1); mbar_init(sc_read, 8);
    mbar_init(pa_ready, 4); mbar_init(pb_ready, 4);
    mbar_init(o_full, 1); mbar_init(o_read, 8);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_items = p.nseq * 16;
  const int QT = p.qtiles;
  const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int T = my_items * QT;

  if (warp == 0) {
    // ================================================================ TMA producer
    // warp-wide loop on uniform values, one elected lane issues (no per-issue operand waterfall)
    {
      const bool issuer = elect_one();
      int t = 0;
      for (int j = 0; j < my_items; ++j) {
        const int item = p.reverse ? num_items - 1 - (int)(blockIdx.x + j * gridDim.x) : (int)(blockIdx.x + j * gridDim.x);
        const int seq = item >> 4, head = item & 15;
        const int r0 = 32 * (j & 3);
        mbar_wait(k_empty, (j & 1) ^ 1u, 21);
        if (issuer) {
          mbar_arrive_expect_tx(k_full, C::kKVBytes);
          tma_load_3d(smK, &tmKV, k_full, 1024 + head * 64, 0, seq, kEvictNormal);
          if (STRAY) tma_load_3d(smK + NMAIN * 128, &tmStray, k_full, 1024 + head * 64, NMAIN, seq, kEvictNormal);
        }
        __syncwarp();
        for (int qt = 0; qt < QT; ++qt, ++t) {
          const int qs = t & 1;
          mbar_wait(&q_empty[qs], ((t >> 1) & 1) ^ 1u, 22);
          if (issuer) {
            mbar_arrive_expect_tx(&q_full[qs], C::kQBytes);
            tma_load_3d(smQ + qs * C::kQBytes, &tmQ, &q_full[qs], head * 64, qt < 2 ? qt * 128 : 256 - r0, seq,
                        kEvictFirst);
          }
          __syncwarp();
          if (qt == 0) {
            mbar_wait(v_empty, (j & 1) ^ 1u, 23);
            if (issuer) {
              mbar_arrive_expect_tx(v_full, C::kKVBytes);
              tma_load_3d(smV, &tmKV, v_full, 2048 + head * 64, 0, seq, kEvictNormal);
              if (STRAY) tma_load_3d(smV + NMAIN * 128, &tmStray, v_full, 2048 + head * 64, NMAIN, seq, kEvictNormal);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    // warp-wide loop on uniform values; ONE elected lane (the same throughout) issues every tcgen05.mma / commit
    if (T > 0) {
      const bool issuer = elect_one();
      constexpr uint32_t idS = idesc_bf16(128, NMAIN, 0);
      constexpr uint32_t idSC = idesc_bf16(128, 16, 0);
      constexpr uint32_t idO = idesc_bf16(128, 64, 1);
      const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(smK));
      const uint64_t dKC = umma_desc_kmajor_sw128(smem_u32(smK + NMAIN * 128));
      const uint64_t dV = umma_desc_mnmajor_sw128(smem_u32(smV));
      const uint64_t dQ0 = umma_desc_kmajor_sw128(smem_u32(smQ));
      const uint64_t dQ1 = umma_desc_kmajor_sw128(smem_u32(smQ + C::kQBytes));
      auto issue_qk_stray = [&](int t) {
        const uint64_t da = (t & 1) ? dQ1 : dQ0;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + C::kColSC, da + 2 * k, dKC + 2 * k, idSC, k != 0);
          umma_commit(sc_full);
        }
        __syncwarp();
      };
      if (STRAY) {
        mbar_wait(k_full, 0, 24);
        mbar_wait(&q_full[0], 0, 25);
        tc_fence_after();
        issue_qk_stray(0);
      }
      int t = 0;
      for (int j = 0; j < my_items; ++j) {
        for (int qt = 0; qt < QT; ++qt, ++t) {
          const uint32_t ph = t & 1;
          const bool first = qt == 0, last = qt == QT - 1;
          if (!STRAY) {
            if (first) mbar_wait(k_full, j & 1, 24);
            mbar_wait(&q_full[t & 1], (t >> 1) & 1, 25);
          }
          if (t > 0) mbar_wait(o_read, ph ^ 1u, 26);            // epilogue of tile t-1 has drained TMEM
          if (STRAY) mbar_wait(sc_read, ph, 27);                // S_C(t) is in registers
          tc_fence_after();
          const bool tr = p.trace != nullptr && blockIdx.x == 0 && t < 64 && issuer;
          if (tr) p.trace[t * 16 + 0] = clock64();
          {
            const uint64_t da = (t & 1) ? dQ1 : dQ0;
            if (issuer) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, da + 2 * k, dK + 2 * k, idS, k != 0);
              umma_commit(s_full);
              umma_commit(&q_empty[t & 1]);
              if (last) umma_commit(k_empty);
            }
            __syncwarp();
          }
          if (first) mbar_wait(v_full, j & 1, 28);
          mbar_wait(pa_ready, ph, 29);
          tc_fence_after();
          if (tr) p.trace[t * 16 + 1] = clock64();
          if (issuer) {
#pragma unroll
            for (int k = 0; k < C::CA / 16; ++k)
              umma_bf16_ts(tmem_base + C::kColO, tmem_base + C::kColPA + 8 * k, dV + 128 * k, idO, k != 0);
          }
          __syncwarp();
          mbar_wait(pb_ready, ph, 30);
          tc_fence_after();
          if (tr) p.trace[t * 16 + 2] = clock64();
          if (issuer) {
#pragma unroll
            for (int k = 0; k < C::CB / 16; ++k)
              umma_bf16_ts(tmem_base + C::kColO, tmem_base + C::kColPB + 8 * k, dV + 128 * (C::CA / 16 + k), idO, 1u);
            if (STRAY) umma_bf16_ts(tmem_base + C::kColO, tmem_base + C::kColPC, dV + 128 * (NMAIN / 16), idO, 1u);
            umma_commit(o_full);
            if (last) umma_commit(v_empty);
          }
          __syncwarp();
          if (tr) p.trace[t * 16 + 3] = clock64();
          if (STRAY && t + 1 < T) {                               // stray-key scores of the next tile, one tile ahead
            if (last) mbar_wait(k_full, (j + 1) & 1, 24);
            mbar_wait(&q_full[(t + 1) & 1], ((t + 1) >> 1) & 1, 25);
            tc_fence_after();
            issue_qk_stray(t + 1);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue (8 warps)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;                      // 0 = A (low key columns), 1 = B
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const int valid = p.seq_valid;
    const float sc = p.scale_log2e;
    const bool stray_key = STRAY && valid > NMAIN;
    float sc_cur = 0.f, sc_next = 0.f;
    if (STRAY && T > 0) {
      mbar_wait(sc_full, 0, 31);
      tc_fence_after();
      sc_cur = __uint_as_float(tmem_ld_x1(lane_base + C::kColSC));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sc_read);
    }
    const bool stray_q = STRAY && p.stray_q != 0;
    float* sp = reinterpret_cast<float*>(smem + C::kOffY);   // unnormalised probabilities of the odd query row
    float* sred = sp + C::NKROWS;                             // [0,8) warp maxima, [8,16) warp sums
    float* sq = sred + 16;                                    // [64] the odd query row, fp32, pre-scaled (log2 domain)
    float* so = sq + 64;                                      // [8 key eighths][64 channels]
    const int tid2 = (warp - 4) * 32 + lane;                  // 0..255 over the softmax warps
    // the odd query row's element for thread tid2 < 64, fetched one item ahead (a global-load latency per item otherwise)
    auto load_q = [&](int jj) -> float {
      if (!stray_q || tid2 >= 64 || jj >= my_items) return 0.f;
      const int it2 = p.reverse ? num_items - 1 - (int)(blockIdx.x + jj * gridDim.x) : (int)(blockIdx.x + jj * gridDim.x);
      return __bfloat162float(p.qkv[((long long)(it2 >> 4) * p.seq_stride + NMAIN) * 3072 + (it2 & 15) * 64 + tid2]);
    };
    float q_next = load_q(0);
    int t = 0;
    for (int j = 0; j < my_items; ++j) {
      const int item = p.reverse ? num_items - 1 - (int)(blockIdx.x + j * gridDim.x) : (int)(blockIdx.x + j * gridDim.x);
      const int seq = item >> 4, head = item & 15;
      const int r0 = 32 * (j & 3);
      if (stray_q) {
        // ---- odd query row, part 1 (while Q.K^T of the first tile runs): thread = key, scores in the log2 domain,
        // block max / sum, probabilities to shared memory.  K rows as TMA wrote them: chunk c of row r at c ^ (r & 7).
        if (tid2 < 64) sq[tid2] = q_next * sc;
        q_next = load_q(j + 1);
        mbar_wait(k_full, j & 1, 35);
        softmax_bar_sync();
        float s0, s1;
        {
          const uint8_t* row = smK + tid2 * 128;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 u = *reinterpret_cast<const uint4*>(row + ((c ^ (tid2 & 7)) << 4));
            const float4 qa = *reinterpret_cast<const float4*>(sq + c * 8), qb = *reinterpret_cast<const float4*>(sq + c * 8 + 4);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]),
                         f3 = __bfloat1622float2(h[3]);
            a0 = fmaf(qa.x, f0.x, a0); a1 = fmaf(qa.y, f0.y, a1); a2 = fmaf(qa.z, f1.x, a2); a3 = fmaf(qa.w, f1.y, a3);
            a0 = fmaf(qb.x, f2.x, a0); a1 = fmaf(qb.y, f2.y, a1); a2 = fmaf(qb.z, f3.x, a2); a3 = fmaf(qb.w, f3.y, a3);
          }
          s0 = (a0 + a1) + (a2 + a3);
          // key 256 = row 0 of the stray block (chunk order unswizzled): every warp forms the dot product with a shuffle sum
          const float2 kf = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(smK + NMAIN * 128)[lane]);
          s1 = warp_sum(fmaf(sq[2 * lane], kf.x, sq[2 * lane + 1] * kf.y));
          if (tid2 != 0) s1 = -INFINITY;                                   // counted once, by thread 0
        }
        float mx = warp_max(fmaxf(s0, s1));
        if (lane == 0) sred[warp - 4] = mx;
        softmax_bar_sync();
        mx = fmaxf(fmaxf(fmaxf(sred[0], sred[1]), fmaxf(sred[2], sred[3])), fmaxf(fmaxf(sred[4], sred[5]), fmaxf(sred[6], sred[7])));
        const float p0 = ex2_approx(s0 - mx), p1 = tid2 == 0 ? ex2_approx(s1 - mx) : 0.f;
        sp[tid2] = p0;
        if (tid2 == 0) sp[NMAIN] = p1;
        const float ls = warp_sum(p0 + p1);
        if (lane == 0) sred[8 + warp - 4] = ls;      // sp / sred[8..] are read in part 2, after the barrier of tile 0's pass 1
      }
      for (int qt = 0; qt < QT; ++qt, ++t) {
        const uint32_t ph = t & 1;
        const int grow = (qt < 2 ? qt * 128 : 256 - r0) + row_in_tile;
        const bool mine = grow < valid && (qt < 2 || row_in_tile >= r0);
        const bool active = __any_sync(0xffffffffu, mine);
        mbar_wait(s_full, ph, 32);
        tc_fence_after();
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && t < 64 && warp == 4 && lane == 0;
        if (tr) p.trace[t * 16 + 8] = clock64();
        // ---------------- pass 1: row max over this group's key columns
        float m = -INFINITY;
        if (active) {
          if (grp == 0) {
#pragma unroll
            for (int c = 0; c < C::CA / 32; ++c) m = chunk_max<32>(lane_base + c * 32, c * 32, valid, m);
            if (stray_key) m = fmaxf(m, sc_cur);
          } else {
#pragma unroll
            for (int c = 0; c < C::CB / 32; ++c) m = chunk_max<32>(lane_base + C::CA + c * 32, C::CA + c * 32, valid, m);
            if (C::CB % 32) m = chunk_max<16>(lane_base + C::CA + (C::CB / 32) * 32, C::CA + (C::CB / 32) * 32, valid, m);
          }
        }
        xm[grp * 128 + row_in_tile] = m;
        if (tr) p.trace[t * 16 + 9] = clock64();
        softmax_bar_sync();
        if (tr) p.trace[t * 16 + 10] = clock64();
        m = fmaxf(m, xm[(grp ^ 1) * 128 + row_in_tile]);
        const float mo = m * sc;
        // ---------------- pass 2: exponentials, row sum, P (bf16) written over the consumed S columns
        float l = 0.f;
        if (active) {
          if (grp == 0) {
#pragma unroll
            for (int c = 0; c < C::CA / 32; ++c)
              l = chunk_exp<32>(lane_base + c * 32, lane_base + C::kColPA + c * 16, c * 32, valid, sc, mo, l);
          } else {
#pragma unroll
            for (int c = 0; c < C::CB / 32; ++c)
              l = chunk_exp<32>(lane_base + C::CA + c * 32, lane_base + C::kColPB + c * 16, C::CA + c * 32, valid, sc, mo, l);
            if (C::CB % 32)
              l = chunk_exp<16>(lane_base + C::CA + (C::CB / 32) * 32, lane_base + C::kColPB + (C::CB / 32) * 16,
                                C::CA + (C::CB / 32) * 32, valid, sc, mo, l);
            if (STRAY) {
              uint32_t pc[8];
              const float pz = stray_key ? ex2_approx(fmaf(sc_cur, sc, -mo)) : 0.f;
              l += pz;
              pc[0] = pack_bf16x2(pz, 0.f);
#pragma unroll
              for (int i = 1; i < 8; ++i) pc[i] = 0u;
              tmem_st_x8(lane_base + C::kColPC, pc);
            }
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (tr) p.trace[t * 16 + 11] = clock64();
        if (lane == 0) mbar_arrive(grp == 0 ? pa_ready : pb_ready);
        xl[grp * 128 + row_in_tile] = l;
        if (STRAY && t + 1 < T) {                     // pick up the next tile's stray-key score while P.V runs
          mbar_wait(sc_full, ph ^ 1u, 33);
          tc_fence_after();
          sc_next = __uint_as_float(tmem_ld_x1(lane_base + C::kColSC));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sc_read);
        }
        if (stray_q && qt == 0) {
          // ---- odd query row, part 2 (while P.V of this tile runs): thread = (channel, key quarter)
          mbar_wait(v_full, j & 1, 36);
          const int part = warp - 4;                                       // 32 keys per warp, lane = channel pair
          const uint32_t cofs = (uint32_t)(lane & 3) * 4u, chunk = (uint32_t)lane >> 2;
          float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll 8
          for (int key = part * 32; key < part * 32 + 32; key += 2) {
            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(smV + key * 128 + (((chunk ^ (key & 7)) << 4) + cofs)));
            const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(smV + (key + 1) * 128 + (((chunk ^ ((key + 1) & 7)) << 4) + cofs)));
            const float pa = sp[key], pb = sp[key + 1];
            o0 = fmaf(pa, fa.x, o0); o1 = fmaf(pa, fa.y, o1);
            o2 = fmaf(pb, fb.x, o2); o3 = fmaf(pb, fb.y, o3);
          }
          if (part == 7) {
            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(smV + NMAIN * 128 + ((chunk << 4) + cofs)));
            o0 = fmaf(sp[NMAIN], fa.x, o0); o1 = fmaf(sp[NMAIN], fa.y, o1);
          }
          so[part * 64 + 2 * lane] = o0 + o2;
          so[part * 64 + 2 * lane + 1] = o1 + o3;
          softmax_bar_sync();
          if (tid2 < 64) {
            const float l_stray = ((sred[8] + sred[9]) + (sred[10] + sred[11])) + ((sred[12] + sred[13]) + (sred[14] + sred[15]));
            float o = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) o += so[g * 64 + tid2];
            p.out[((long long)seq * p.seq_stride + NMAIN) * 1024 + head * 64 + tid2] = __float2bfloat16_rn(o / l_stray);
          }
        }
        // ---------------- epilogue: O / l -> bf16 -> global; each group writes 32 of the 64 head channels
        mbar_wait(o_full, ph, 34);
        tc_fence_after();
        if (tr) p.trace[t * 16 + 12] = clock64();
        softmax_bar_sync();
        l += xl[(grp ^ 1) * 128 + row_in_tile];
        if (active) {
          uint32_t o[32];
          tmem_ld_x32(lane_base + C::kColO + grp * 32, o);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_read);
          if (mine) {
            const float inv = 1.0f / l;
            bf16* dst = p.out + ((long long)seq * p.seq_stride + grow) * 1024 + head * 64 + grp * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + i) = w;
            }
          }
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_read);
        }
        if (tr) p.trace[t * 16 + 13] = clock64();
        sc_cur = sc_next;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int NMAIN, bool STRAY>
int launch_tc2(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s, bool reverse) {
  using C = Cfg<NMAIN, STRAY>;
  auto kern = mha_tc2_kernel<NMAIN, STRAY>;
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  // one 3-D view of the packed QKV activation: {3072 channels, seq_valid rows, nseq}; rows >= seq_valid read as zero
  uint64_t dims[3] = {3072, (uint64_t)seq_valid, (uint64_t)nseq};
  uint64_t strides[2] = {3072 * 2, (uint64_t)seq_stride * 3072 * 2};
  CUtensorMap tq, tkv, tst;
  if (int r = make_tmap_generic(&tq, 0, qkv, 3, dims, strides, 64, 128)) return r;
  if (int r = make_tmap_generic(&tkv, 0, qkv, 3, dims, strides, 64, NMAIN)) return r;
  if (int r = make_tmap_generic(&tst, 0, qkv, 3, dims, strides, 64, 16)) return r;
  AParams p{};
  p.nseq = nseq; p.seq_stride = seq_stride; p.seq_valid = seq_valid;
  p.reverse = reverse ? 1 : 0;
  // 257 tokens (ViT): two full query tiles on the tensor cores, row 256 on the CUDA cores inside the softmax warps' idle windows
  static int stray_mode = -1;
  if (stray_mode < 0) { const char* e = getenv("AGB_ATTN_STRAY_TILE"); stray_mode = (e && e[0] == '1') ? 0 : 1; }   // A/B: 1 = third tile
  p.stray_q = (STRAY && stray_mode == 1 && seq_valid == NMAIN + 1) ? 1 : 0;
  p.qtiles = p.stray_q ? NMAIN / 128 : (seq_valid + 127) / 128;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.qkv = qkv;
  p.out = out;
  static long long* trace_buf = nullptr;
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("AGB_ATTN_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
  if (trace_on && trace_buf == nullptr) AGB_CHECK_CUDA(cudaMalloc(&trace_buf, 64 * 16 * sizeof(long long)));
  p.trace = trace_on ? trace_buf : nullptr;
  const int items = nseq * 16;
  const int resident = sm_count() * 2;
  const int grid = items < resident ? items : resident;
  kern<<<grid, kThreads, C::kSmem, s>>>(tq, tkv, tst, p);
  AGB_CHECK_LAUNCH();
  if (trace_on) {                                       // diagnostic mode only: synchronous dump of the stamps
    static long long host[64 * 16];
    AGB_CHECK_CUDA(cudaStreamSynchronize(s));
    AGB_CHECK_CUDA(cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost));
    printf("attn-trace NMAIN=%d qtiles=%d: tile | MMA: qk_issue pa_ready pb_ready pv_issued | softmax: s_full pass1 bar pass2 o_full done (clk from tile 8's qk_issue)\n", NMAIN, p.qtiles);
    const long long t00 = host[8 * 16];
    for (int t = 8; t < 20; ++t) {
      printf("attn-trace %2d |", t);
      for (int k : {0, 1, 2, 3}) printf(" %7lld", host[t * 16 + k] - t00);
      printf(" |");
      for (int k : {8, 9, 10, 11, 12, 13}) printf(" %7lld", host[t * 16 + k] - t00);
      printf("\n");
    }
  }
  return 0;
}

}  // namespace

// returns -1 when the shape is not covered by an instantiation (caller falls back)
int attention_tc2(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s, bool reverse) {
  if (seq_valid <= 112) return launch_tc2<112, false>(qkv, out, nseq, seq_stride, seq_valid, scale, s, reverse);
  if (seq_valid > 128 && seq_valid <= 257) return launch_tc2<256, true>(qkv, out, nseq, seq_stride, seq_valid, scale, s, reverse);
  return -1;
}

}  // namespace agb
