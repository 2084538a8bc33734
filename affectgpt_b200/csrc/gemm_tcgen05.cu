// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+= into fp32 C)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one thread: tcgen05.mma.kind::f16, fp32 accumulators in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..11 : epilogue      (tcgen05.ld -> bias / GELU -> swizzled smem staging -> TMA store, or TMA
//                                reduce-add for the residual stream: the "+= residual" happens in L2, the
//                                epilogue never reads the residual and issues no LSU global traffic)
//
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static tile
// schedule (N fastest so concurrently running CTAs share the A row block in L2).
//
// CG = 2 (the default for the big GEMMs): two CTAs of a cluster form a pair (tcgen05 cta_group::2).  The pair
// computes a 256 x BN tile; each CTA stages only its own 128 A rows and HALF of the W tile, the tensor cores
// read the other half from the peer's shared memory.  Per-SM operand traffic from L2 drops from 48 KB to
// 32 KB per 64-deep K block ("GEMM roofline").  Protocol (both CTAs run all roles except MMA issue):
//   full[s]   (leader's copy) : 2 producer arrivals + complete_tx bytes of BOTH CTAs' TMA loads
//   empty[s]  (every CTA)     : multicast tcgen05.commit from the leader's MMA thread
//   tfull[a]  (every CTA)     : multicast tcgen05.commit when an accumulator stage is complete
//   tempty[a] (leader's copy) : epilogue warps of both CTAs arrive remotely
//
// A-operand modes (AMode): plain row-major; k=3/s=2 conv as a GEMM over the "pair-of-rows" view of a
// channels-last activation (no im2col buffer, the third tap is a +1-row TMA coordinate); grouped k=128
// positional conv (one tap per K block, TMA row coordinate shifted by tap-64, out-of-range rows
// zero-filled by TMA = the conv's zero padding).
//
// LNF (fused LayerNorm tail, residual GEMMs with N = 1024 only): the LayerNorm that follows "resid += A W^T + b" in a
// pre-LN block needs complete rows, and a tile owns only 256 of a row's 1024 columns.  The n_tiles * CG CTAs that share a
// 256-row block therefore meet at a counter in global memory once their reduce-adds are performed; each then normalises
// 32 of the block's rows - read back from L2, where the reduce-adds just happened, not from HBM - and writes the bf16
// operand of the next GEMM.  The epilogue warps do this while the MMA warp is already in the next tile's main loop.  It
// replaces a stand-alone LayerNorm pass (0.8 GB of HBM traffic per launch at C2) by a 0.27 GB write.  The arithmetic is
// ln_row_apply (agb_ptx.cuh), the same code as the stand-alone kernel: bit-identical output.  The meeting needs all CTAs
// of the grid co-resident: the grid is sized by cudaOccupancyMaxActiveClusters (see launch()); two such kernels must not
// share the GPU (AGB_OVERLAP=1 and this switch exclude each other).  Opt-in (AGB_FUSED_LN=1): see fused_ln_on().
//
// Replaces (reference): torch.nn.Linear / Conv1d calls inside HF CLIPVisionTransformer and HubertModel
// (transformers/models/clip/modeling_clip.py:310-350, models/hubert/modeling_hubert.py:45-231, 262-370).
#include <cuda.h>

#include <cstring>
#include <mutex>

#include "agb_internal.h"
#include "agb_ptx.cuh"

#ifdef AGB_GEMM_GELU_EXACT
#define AGB_GEMM_GELU(x) gelu_erf_fast(x)       // A&S 7.1.26: 2 MUFU + 15 FMA, |err| < 1.5e-7
#else
#define AGB_GEMM_GELU(x) gelu_erf_tanhform(x)   // 1 MUFU + 7 FMA, |formula err| < 2.5e-5
#endif

namespace agb {

namespace {

constexpr int BM = 128;               // rows per CTA
constexpr int BK = 64;
constexpr int kNumThreads = 384;      // 4 control warps + 8 epilogue warps
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kStageBufBytes = 4096;  // per epilogue warp: 32 rows x 128 B, 128B-swizzled (TMA store source)

template <int BN, int CG>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / CG) * BK * 2;            // per CTA: the pair splits W's rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = kNumEpiWarps * kStageBufBytes;
  static constexpr int kBarBytes = 512;                         // 2 * kStages + 8 mbarriers + the TMEM slot
  static constexpr int kBudget = 232448 - 1024 - kBarBytes - kEpiBytes;
  static constexpr int kMaxStages = 8;
  static constexpr int kStages = kBudget / kStageBytes > kMaxStages ? kMaxStages : kBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;                      // two accumulator stages (power of two >= 32)
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align*/ + kBarBytes;
  static_assert((kStages * kStageBytes) % 1024 == 0, "staging buffers need 1024-byte alignment");
  static_assert((2 * kStages + 8) * 8 + 16 <= kBarBytes, "barrier region");
};

struct KParams {
  int M, N, K;
  int num_kb;
  int m_tiles, n_tiles;   // m_tiles counts (128 * CG)-row tiles
  const float* bias;
  int has_f32;            // fp32 output through tmF
  int reduce_f32;         // ... as an L2 reduce-add into the existing fp32 tensor (residual stream)
  int has_bf16;           // bf16 output through tmH
  int reverse;            // tile schedule walks from the last tile to the first
  // LNF: fused LayerNorm of the rows this launch completes (resid -> bf16), see the header comment
  const float* ln_src;    // == the fp32 output tensor (row stride 1024)
  const float* ln_g;
  const float* ln_b;
  float ln_eps;
  bf16* ln_out;           // [M, 1024]
  int* ln_cnt;            // [m_tiles] arrival counters, zero before and after every launch
};

template <int BN, int MODE, int ACT, int CG, bool LNF = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmH, const KParams p) {
  constexpr bool kPos = MODE == A_POSCONV;
  using C = Cfg<BN, CG>;
  constexpr int S = C::kStages;
  static_assert(CG == 1 || !kPos, "the positional-conv modes run single-CTA");
  extern __shared__ uint8_t smem_raw[];
  // identical carve-up in both CTAs of a pair: the peer's operands / barriers live at the same offsets
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + S * C::kABytes;
  uint8_t* smE = smem + S * C::kStageBytes;                       // epilogue staging, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smE + C::kEpiBytes);
  uint64_t* full = bars;            // [S]  TMA -> MMA
  uint64_t* empty = bars + S;       // [S]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S;   // [2]  MMA -> epilogue
  uint64_t* tempty = tfull + 2;     // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = uniform_warp_id();
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.has_f32) tma_prefetch_desc(&tmF);
    if (p.has_bf16) tma_prefetch_desc(&tmH);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], CG);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kNumEpiWarps * CG);
    }
    fence_mbar_init();
  }
  if constexpr (CG == 2) cluster_sync_all();   // peer is resident before the paired TMEM allocation
  if (warp == 2) {
    if constexpr (CG == 2) tmem_alloc_pair(tmem_slot, C::kTmemCols); else tmem_alloc(tmem_slot, C::kTmemCols);
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int tile0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ================================================================ TMA producer
    // the whole warp runs the loop on warp-uniform values; one elected lane issues (see uniform_warp_id)
    {
      const bool issuer = elect_one();
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int tl = p.reverse ? num_tiles - 1 - tile : tile;
        const int mt = tl / p.n_tiles, nt = tl % p.n_tiles;
        const int row0 = mt * (BM * CG) + (int)cta_rank * BM;          // this CTA's 128 A rows
        const int wrow0 = nt * BN + (int)cta_rank * (BN / CG);          // this CTA's share of the W tile
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 1);
          void* dA = smA + stage * C::kABytes;
          void* dB = smB + stage * C::kBBytes;
          if (issuer) {
            if constexpr (CG == 2) {
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
              if constexpr (MODE == A_PLAIN) {
                tma_load_2d_pair(dA, &tmA, &full[stage], kb * BK, row0, kEvictNormal);
              } else {
                tma_load_2d_pair(dA, &tmA, &full[stage], (kb & 15) * BK, row0 + (kb >> 4), kEvictNormal);
              }
              tma_load_2d_pair(dB, &tmB, &full[stage], kb * BK, wrow0, kEvictLast);
              if (!leader) mbar_arrive_remote(&full[stage], 0);
            } else {
              mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
              if constexpr (MODE == A_PLAIN) {
                tma_load_2d(dA, &tmA, &full[stage], kb * BK, row0, kEvictNormal);
              } else if constexpr (MODE == A_CONV3) {
                // K blocks 0..15: pair row m, cols kb*64; K blocks 16..23: pair row m+1, cols (kb-16)*64
                tma_load_2d(dA, &tmA, &full[stage], (kb & 15) * BK, row0 + (kb >> 4), kEvictNormal);
              } else if constexpr (MODE == A_POSCONV) {
                // mt = sequence, nt = group; tap kb reads rows t + kb - 64 (TMA zero-fills t outside [0, valid))
                tma_load_3d(dA, &tmA, &full[stage], nt * 64, kb - 64, mt, kEvictNormal);
              }
              tma_load_2d(dB, &tmB, &full[stage], kb * BK, wrow0, kEvictLast);
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA only)
    // warp-wide loop on uniform values, one elected lane issues tcgen05.mma / tcgen05.commit (the SAME lane throughout:
    // a commit tracks the MMAs of the thread that issues it)
    if (leader) {
      const bool issuer = elect_one();
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1u, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smA + stage * C::kABytes));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smB + stage * C::kBBytes));
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128B swizzle atom: +2 in the (addr >> 4) field
              if constexpr (CG == 2)
                umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else
                umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            if constexpr (CG == 2) {
              umma_commit_pair(&empty[stage], 0x3);                        // smem slot free in both CTAs
              if (kb == p.num_kb - 1) umma_commit_pair(&tfull[as], 0x3);   // accumulator complete in both CTAs
            } else {
              umma_commit(&empty[stage]);
              if (kb == p.num_kb - 1) umma_commit(&tfull[as]);
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
      if constexpr (CG == 2) {
        // the peer's remote tempty arrivals must have landed before this CTA's shared memory goes away
        for (int d = 0; d < 2 && d < it; ++d) {
          const int last = it - 1 - d;
          mbar_wait(&tempty[last & 1], (last >> 1) & 1, 5);
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================================ epilogue (8 warps per CTA)
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    const int half = ew >> 2;          // column half handled by this warp
    constexpr int HC = BN / 2;         // columns per warp
    constexpr int NCH = HC / 32;       // 32-column chunks per warp (1, 2 or 4)
    uint8_t* stg = smE + ew * kStageBufBytes;            // this warp's staging tile: 32 rows x 128 B
    uint8_t* my_row = stg + lane * 128;
    const int sw = lane & 7;                             // 128B swizzle: 16-byte chunk index ^= row % 8
    bool store_pending = false;
    int it = 0;
    // LNF: the row block (mt, nt) whose LayerNorm share this CTA still owes (see the header comment)
    int pend_mt = -1, pend_nt = 0;
    auto ln_share = [&](int bmt, int bnt) {
      // (a) meet the other CTAs of the 256-row block: n_tiles * CG first arrivals = all of its reduce-adds are in L2
      const int total = p.n_tiles * CG;
      int* cnt = p.ln_cnt + bmt;
      if (ew == 0 && lane == 0) {
        while (ld_acquire_gpu(cnt) < total) __nanosleep(64);
        __threadfence();
      }
      named_bar_sync(2, kNumEpiWarps * 32);
      // (b) normalise this CTA's share of the block: 256 / total rows, two rows per warp in flight, read from L2
      const int rows_cta = (BM * CG) / total;                    // 32 for N = 1024
      const int rows_warp = rows_cta / kNumEpiWarps;             // 4
      const long long r0 = (long long)bmt * (BM * CG) + (long long)(bnt * CG + (int)cta_rank) * rows_cta + ew * rows_warp;
      for (int rr = 0; rr < rows_warp; rr += 2) {
        float4 va[8], vb[8];
        const long long ra = r0 + rr, rb = r0 + rr + 1;
        const bool oka = ra < p.M, okb = rb < p.M && rr + 1 < rows_warp;
        if (oka) {
          const float4* src = reinterpret_cast<const float4*>(p.ln_src + ra * 1024);
#pragma unroll
          for (int i = 0; i < 8; ++i) va[i] = __ldcg(src + lane + 32 * i);
        }
        if (okb) {
          const float4* src = reinterpret_cast<const float4*>(p.ln_src + rb * 1024);
#pragma unroll
          for (int i = 0; i < 8; ++i) vb[i] = __ldcg(src + lane + 32 * i);
        }
        if (oka) ln_row_apply<8>(va, lane, p.ln_g, p.ln_b, p.ln_eps, p.ln_out + ra * 1024, nullptr);
        if (okb) ln_row_apply<8>(vb, lane, p.ln_g, p.ln_b, p.ln_eps, p.ln_out + rb * 1024, nullptr);
      }
      // (c) second arrival: whoever completes the count puts the counter back to zero for the next launch
      __syncwarp();
      named_bar_sync(2, kNumEpiWarps * 32);
      if (ew == 0 && lane == 0) {
        if (atom_add_release_gpu(cnt, 1) == 2 * total - 1) *reinterpret_cast<volatile int*>(cnt) = 0;
      }
    };
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int tl = p.reverse ? num_tiles - 1 - tile : tile;
      const int mt = tl / p.n_tiles, nt = tl % p.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row_base = kPos ? q * 32 : mt * (BM * CG) + (int)cta_rank * BM + q * 32;
      const int col0 = nt * BN + half * HC;

      mbar_wait(&tfull[as], aphase, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + half * HC);

      uint32_t v[2][32];
      tmem_ld_32x32b_x32(taddr, v[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        tmem_ld_wait();
        if (c + 1 < NCH) {
          tmem_ld_32x32b_x32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
        } else {
          // all of this warp's accumulator columns are in registers: hand the TMEM stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]);
          }
        }
        const int cbase = col0 + c * 32;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[c & 1][j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (cbase + j < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j));
              x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
            }
          }
        }
        if constexpr (ACT == ACT_QUICK_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = quick_gelu(x[j]);
        } else if constexpr (ACT == ACT_GELU_ERF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = AGB_GEMM_GELU(x[j]);
        }
        if (p.has_f32) {
          // ---- fp32 tile [32 rows x 32 cols] -> staging -> TMA store / reduce-add
          if (store_pending) {
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4)) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && cbase < p.N) {
            if constexpr (kPos) {
              tma_reduce_add_3d(&tmF, stg, cbase, row_base, mt);
            } else {
              if (p.reduce_f32) tma_reduce_add_2d(&tmF, stg, cbase, row_base); else tma_store_2d(&tmF, stg, cbase, row_base);
            }
            tma_store_commit();
          }
          store_pending = true;
        }
        if constexpr (HC >= 64) {
          if (p.has_bf16) {
            // ---- bf16 tile [32 rows x 64 cols]: two chunks fill the left / right 64 B of each 128 B row
            if ((c & 1) == 0 && store_pending) {
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
              store_pending = false;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(x[8 * j], x[8 * j + 1]);
              o.y = pack_bf16x2(x[8 * j + 2], x[8 * j + 3]);
              o.z = pack_bf16x2(x[8 * j + 4], x[8 * j + 5]);
              o.w = pack_bf16x2(x[8 * j + 6], x[8 * j + 7]);
              *reinterpret_cast<uint4*>(my_row + ((((c & 1) * 4 + j) ^ sw) << 4)) = o;
            }
            if (c & 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0 && cbase - 32 < p.N) {
                tma_store_2d(&tmH, stg, cbase - 32, row_base);
                tma_store_commit();
              }
              store_pending = true;
            }
          }
        }
      }
      if constexpr (LNF) {
        // ---- fused LayerNorm tail, one tile behind: first this CTA's share of the PREVIOUS tile's row block (whose
        // other CTAs arrived a tile ago, so the meeting costs no waiting; meanwhile this tile's reduce-adds complete) ...
        if (pend_mt >= 0) ln_share(pend_mt, pend_nt);
        // ... then the arrival for this tile: every reduce-add of this CTA's tile has been performed in L2
        if (lane == 0) {
          tma_store_wait_all();
          fence_proxy_async_all();
        }
        store_pending = false;
        __syncwarp();
        named_bar_sync(2, kNumEpiWarps * 32);
        if (ew == 0 && lane == 0) {
          __threadfence();
          atom_add_release_gpu(p.ln_cnt + mt, 1);
        }
        pend_mt = mt;
        pend_nt = nt;
      }
    }
    if constexpr (LNF) {
      if (pend_mt >= 0) ln_share(pend_mt, pend_nt);
    }
    if (store_pending && lane == 0) tma_store_wait_read();   // staging must outlive the last bulk store's read
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// rank-2 or rank-3 tensor, dim0 contiguous; box = {box0, box1, 1} with box0 * elem = 128 B; 128B swizzle; OOB -> zeros
int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, uint32_t box0, uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  AGB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  cuuint64_t gdim[3] = {dims[0], dims[1], rank == 3 ? dims[2] : 1};
  cuuint64_t gstr[2] = {strides_bytes[0], rank == 3 ? strides_bytes[1] : 0};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 3;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct Maps {
  CUtensorMap a, b, f, h;
};

template <int BN, int MODE, int ACT, int CG, bool LNF = false>
int launch(const Maps& m, const KParams& kp, cudaStream_t stream) {
  auto kern = gemm_kernel<BN, MODE, ACT, CG, LNF>;
  using C = Cfg<BN, CG>;
  static bool configured = false;
  static int max_units = 0;                               // LNF: clusters that are guaranteed to be co-resident
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  const int tiles = kp.m_tiles * kp.n_tiles;
  int units = num_sms() / CG;                             // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
  if constexpr (LNF) {
    // the CTAs of a row block wait for each other inside the kernel: every CTA of the grid must be resident at once.
    // Size the grid by what the hardware can co-schedule (pairs need two free SMs of one GPC) and make the launch
    // cooperative, so that an impossible placement is a launch error instead of a hang.
    static_assert(CG == 2, "LNF runs on CTA pairs");
    if (max_units == 0) {
      cfg.gridDim = dim3((unsigned)(units * CG));
      int n = 0;
      AGB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      AGB_REQUIRE(n > 0, "gemm: no co-resident cluster placement for the fused-LayerNorm kernel");
      max_units = n;
    }
    if (units > max_units) units = max_units;
    static int coop = -1;                 // AGB_LNF_COOP=1: also mark the launch cooperative (diagnostics)
    if (coop < 0) { const char* e = getenv("AGB_LNF_COOP"); coop = (e && e[0] == '1') ? 1 : 0; }
    if (coop) {
      attr[1].id = cudaLaunchAttributeCooperative;
      attr[1].val.cooperative = 1;
      cfg.numAttrs = 2;
    }
  }
  const int grid = (tiles < units ? tiles : units) * CG;
  cfg.gridDim = dim3((unsigned)grid);
  AGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, m.a, m.b, m.f, m.h, kp));
  AGB_CHECK_LAUNCH();
  return 0;
}

// AGB_FUSED_LN=1: LayerNorm in the residual GEMM's tail instead of its own pass (bit-identical results)
// Off by default.  Measured on B200 in one process, graphs replayed alternately (tools/ab_graph.py, 30 rounds, C2 step):
// fused 123.7 ms vs separate LayerNorm passes 120.6 ms, although the LayerNorm category drops from 10.2 to 1.9 ms and 51 GB
// of HBM reads per step disappear.  The step runs against the 1 kW power cap, where time follows joules: the stand-alone
// pass costs 0.118 J (profiles/r02_energy_probe.jsonl), of which only the DRAM read (~0.02 J) is saved, and the tail costs
// the GEMM more than that (168 registers with spills, epilogue warps busy while the tensor pipe waits).  A variant with
// four dedicated LayerNorm warps (512 threads, 128 registers, arrivals posted one tile late with cp.async.bulk.wait_group
// NCH so that neither the epilogue nor the MMA warp ever waits) was bit-identical too and worse: 123.2 vs 113.4 ms.  Kept opt-in and
// tested for bit-identity (tests/test_gpu_kernels.py); read per call so one process can capture a graph per setting.
bool fused_ln_on() {
  const char* e = getenv("AGB_FUSED_LN");
  return e != nullptr && e[0] == '1';
}

}  // namespace

int make_tmap_generic(CUtensorMap* out, int dtype_is_f32, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, uint32_t box0, uint32_t box1) {
  return make_tmap(out, dtype_is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims,
                   strides_bytes, box0, box1);
}

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  AGB_REQUIRE(a.A != nullptr && a.W != nullptr, "gemm: null operand");
  AGB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  AGB_REQUIRE(a.N % 8 == 0, "gemm: N must be a multiple of 8");
  AGB_REQUIRE(a.out_f32 != nullptr || a.out_bf16 != nullptr, "gemm: no output");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  AGB_REQUIRE(a.res == nullptr || a.res == a.out_f32, "gemm: the residual must alias the fp32 output (in-place +=)");

  const double rows_eff = a.mode == A_POSCONV ? (double)a.nseq * a.seq_valid : (double)a.M;
  LaunchTimer timer(PROF_GEMM, 2.0 * rows_eff * a.N * a.K,
                    2.0 * (rows_eff * a.K + (double)a.N * a.K) + rows_eff * a.N * (a.out_f32 ? 4.0 : 0.0) +
                        rows_eff * a.N * (a.out_bf16 ? 2.0 : 0.0) + rows_eff * a.N * (a.res ? 4.0 : 0.0),
                    stream);
  KParams kp{};
  kp.M = a.M; kp.N = a.N; kp.K = a.K;
  kp.bias = a.bias;
  kp.has_f32 = a.out_f32 != nullptr;
  kp.reduce_f32 = a.res != nullptr;
  kp.has_bf16 = a.out_bf16 != nullptr;
  kp.reverse = a.reverse ? 1 : 0;
  if (a.out_f32) AGB_REQUIRE(a.ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0, "gemm: f32 out alignment");
  if (a.out_bf16) AGB_REQUIRE(a.ld_bf16 % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out_bf16) & 15) == 0, "gemm: bf16 out alignment");
  if (a.bias) AGB_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias alignment");
  AGB_REQUIRE(!(a.out_f32 && a.out_bf16), "gemm: exactly one output (fp32 or bf16)");

  // CTA pairs pay off once there are enough 256-row tiles to fill the machine
  const bool pairs = a.mode != A_POSCONV && a.M >= 2048;
  const int cg = pairs ? 2 : 1;
  // fused LayerNorm tail (see the header comment): residual GEMMs with complete 1024-wide rows on CTA pairs; any other
  // shape runs the plain kernel followed by the stand-alone LayerNorm (same arithmetic, same bits)
  const bool want_ln = a.ln_out != nullptr;
  if (want_ln)
    AGB_REQUIRE(a.ln_gamma && a.ln_beta && a.res != nullptr && a.out_f32 != nullptr && a.N == 1024,
                "gemm: the fused LayerNorm needs gamma / beta and a 1024-wide fp32 residual output");
  const bool fuse_ln = want_ln && fused_ln_on() && pairs && a.mode == A_PLAIN && a.act == ACT_NONE && a.ld_f32 == 1024 &&
                       a.ln_counters != nullptr;
  kp.ln_src = a.out_f32; kp.ln_g = a.ln_gamma; kp.ln_b = a.ln_beta; kp.ln_eps = a.ln_eps; kp.ln_out = a.ln_out;
  kp.ln_cnt = a.ln_counters;
  auto ln_after = [&]() -> int {            // fallback: separate pass, walking the rows in the opposite direction
    if (!want_ln || fuse_ln) return 0;
    return layernorm_f32_in(a.out_f32, a.ld_f32, a.M, 1024, a.ln_gamma, a.ln_beta, a.ln_eps, a.ln_out, 1024, nullptr, 0, stream,
                            !a.reverse);
  };

  Maps m{};
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  if (a.mode == A_PLAIN) {
    AGB_REQUIRE(a.K % BK == 0, "gemm: K must be a multiple of 64");
    AGB_REQUIRE(a.lda >= a.K && (a.lda * 2) % 16 == 0, "gemm: bad lda");
    uint64_t dA[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t sA[1] = {(uint64_t)a.lda * 2};
    if (int r = make_tmap(&m.a, BF, a.A, 2, dA, sA, 64, BM)) return r;
    kp.num_kb = a.K / BK;
    kp.m_tiles = (a.M + BM * cg - 1) / (BM * cg);
  } else if (a.mode == A_CONV3) {
    AGB_REQUIRE(a.K == 1536, "gemm: conv3 mode expects K = 3 * 512");
    uint64_t dA[2] = {1024, (uint64_t)a.M};  // pair rows; row M (third tap of the last row) is OOB -> zeros
    uint64_t sA[1] = {1024 * 2};
    if (int r = make_tmap(&m.a, BF, a.A, 2, dA, sA, 64, BM)) return r;
    kp.num_kb = 24;
    kp.m_tiles = (a.M + BM * cg - 1) / (BM * cg);
  } else if (a.mode == A_POSCONV) {
    AGB_REQUIRE(a.N == 1024 && a.K == 8192, "gemm: posconv mode expects N = 1024, K = 128 * 64");
    AGB_REQUIRE(a.nseq > 0 && a.seq_valid > 0 && a.seq_valid <= BM && a.seq_stride >= a.seq_valid, "gemm: posconv seq shape");
    AGB_REQUIRE(a.out_f32 != nullptr && a.res != nullptr && a.out_bf16 == nullptr && a.ld_f32 == 1024,
                "gemm: posconv writes resid += gelu(conv) (fp32, row stride 1024)");
    uint64_t dA[3] = {1024, (uint64_t)a.seq_valid, (uint64_t)a.nseq};
    uint64_t sA[2] = {1024 * 2, (uint64_t)a.seq_stride * 1024 * 2};
    if (int r = make_tmap(&m.a, BF, a.A, 3, dA, sA, 64, BM)) return r;
    uint64_t sF[2] = {1024 * 4, (uint64_t)a.seq_stride * 1024 * 4};
    if (int r = make_tmap(&m.f, F32, a.out_f32, 3, dA, sF, 32, 32)) return r;   // rows >= seq_valid are clipped
    kp.num_kb = 128;
    kp.m_tiles = a.nseq;
  } else {
    AGB_REQUIRE(false, "gemm: unknown A mode");
  }

  uint64_t dB[2] = {(uint64_t)a.K, (uint64_t)a.N};
  uint64_t sB[1] = {(uint64_t)a.K * 2};
  uint64_t dC[2] = {(uint64_t)a.N, (uint64_t)a.M};
  if (a.mode != A_POSCONV && a.out_f32) {
    uint64_t sF[1] = {(uint64_t)a.ld_f32 * 4};
    if (int r = make_tmap(&m.f, F32, a.out_f32, 2, dC, sF, 32, 32)) return r;
  }
  if (a.out_bf16) {
    uint64_t sH[1] = {(uint64_t)a.ld_bf16 * 2};
    if (int r = make_tmap(&m.h, BF, a.out_bf16, 2, dC, sH, 64, 32)) return r;   // 64-column boxes: full 128 B rows
  }

  if (a.mode == A_POSCONV) {
    if (int r = make_tmap(&m.b, BF, a.W, 2, dB, sB, 64, 64)) return r;
    kp.n_tiles = 16;
    AGB_REQUIRE(a.act == ACT_GELU_ERF, "gemm: posconv epilogue is bias + GELU (+ residual)");
    return launch<64, A_POSCONV, ACT_GELU_ERF, 1>(m, kp, stream);
  }
  if (int r = make_tmap(&m.b, BF, a.W, 2, dB, sB, 64, 256 / cg)) return r;
  kp.n_tiles = (a.N + 255) / 256;
  if (a.mode == A_CONV3) {
    if (a.act == ACT_NONE) return pairs ? launch<256, A_CONV3, ACT_NONE, 2>(m, kp, stream) : launch<256, A_CONV3, ACT_NONE, 1>(m, kp, stream);
    if (a.act == ACT_GELU_ERF) return pairs ? launch<256, A_CONV3, ACT_GELU_ERF, 2>(m, kp, stream) : launch<256, A_CONV3, ACT_GELU_ERF, 1>(m, kp, stream);
    AGB_REQUIRE(false, "gemm: unsupported activation for conv3 mode");
  }
  if (fuse_ln) return launch<256, A_PLAIN, ACT_NONE, 2, true>(m, kp, stream);
  if (want_ln) {
    AGB_REQUIRE(a.act == ACT_NONE, "gemm: LayerNorm output is only offered for the residual GEMMs (no activation)");
    if (int r = pairs ? launch<256, A_PLAIN, ACT_NONE, 2>(m, kp, stream) : launch<256, A_PLAIN, ACT_NONE, 1>(m, kp, stream)) return r;
    return ln_after();
  }
  switch (a.act) {
    case ACT_NONE: return pairs ? launch<256, A_PLAIN, ACT_NONE, 2>(m, kp, stream) : launch<256, A_PLAIN, ACT_NONE, 1>(m, kp, stream);
    case ACT_QUICK_GELU: return pairs ? launch<256, A_PLAIN, ACT_QUICK_GELU, 2>(m, kp, stream) : launch<256, A_PLAIN, ACT_QUICK_GELU, 1>(m, kp, stream);
    case ACT_GELU_ERF: return pairs ? launch<256, A_PLAIN, ACT_GELU_ERF, 2>(m, kp, stream) : launch<256, A_PLAIN, ACT_GELU_ERF, 1>(m, kp, stream);
    default: break;
  }
  AGB_REQUIRE(false, "gemm: unknown activation");
}

}  // namespace agb
