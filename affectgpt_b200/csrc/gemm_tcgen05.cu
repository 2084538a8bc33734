// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual)
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one thread: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16, fp32 in TMEM)
//   warp 2      : TMEM allocator (2 accumulator stages x BN columns)
//   warps 4..11 : epilogue      (tcgen05.ld -> bias / GELU / residual -> 16-byte global stores)
//
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static tile
// schedule (tile = blockIdx.x + i * gridDim.x, N fastest so concurrent CTAs share the A row block in L2).
//
// A-operand modes (AMode): plain row-major; k=3/s=2 conv as a GEMM over the "pair" view of a
// channels-last activation (no im2col buffer, the third tap is a +1-row TMA coordinate); grouped k=128
// positional conv (one tap per K block, TMA row coordinate shifted by tap-64, out-of-range rows
// zero-filled by TMA = the conv's zero padding).
//
// Replaces (reference): torch.nn.Linear / Conv1d calls inside HF CLIPVisionTransformer and HubertModel
// (transformers/models/clip/modeling_clip.py:310-350, models/hubert/modeling_hubert.py:45-231, 262-370).
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>

#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 384;      // 4 control warps + 8 epilogue warps
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;

template <int BN>
struct Cfg {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32)
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 256 /*barriers*/;
};

struct KParams {
  int M, N, K;
  int num_kb;
  int m_tiles, n_tiles;
  const float* bias;
  const float* res;
  long long ld_res;
  float* out_f32;
  long long ld_f32;
  bf16* out_bf16;
  long long ld_bf16;
  int seq_stride, seq_valid;
};

template <int BN, int MODE, int ACT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  using C = Cfg<BN>;
  constexpr int S = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + S * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * (C::kABytes + C::kBBytes));
  uint64_t* full = bars;            // [S]  TMA -> MMA
  uint64_t* empty = bars + S;       // [S]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S;   // [2]  MMA -> epilogue
  uint64_t* tempty = tfull + 2;     // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ================================================================ TMA producer (one thread)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u, 1);
          mbar_arrive_expect_tx(&full[stage], C::kABytes + C::kBBytes);
          void* dA = smA + stage * C::kABytes;
          void* dB = smB + stage * C::kBBytes;
          if constexpr (MODE == A_PLAIN) {
            tma_load_2d(dA, &tmA, &full[stage], kb * BK, mt * BM, kEvictNormal);
            tma_load_2d(dB, &tmB, &full[stage], kb * BK, nt * BN, kEvictLast);
          } else if constexpr (MODE == A_CONV3) {
            // K blocks 0..15: pair row m, cols kb*64; K blocks 16..23: pair row m+1, cols (kb-16)*64
            tma_load_2d(dA, &tmA, &full[stage], (kb & 15) * BK, mt * BM + (kb >> 4), kEvictNormal);
            tma_load_2d(dB, &tmB, &full[stage], kb * BK, nt * BN, kEvictLast);
          } else {
            // mt = sequence, nt = group; tap kb reads rows t + kb - 64 (TMA zero-fills t outside [0, valid))
            tma_load_3d(dA, &tmA, &full[stage], nt * 64, kb - 64, mt, kEvictNormal);
            tma_load_2d(dB, &tmB, &full[stage], kb * BK, nt * BN, kEvictLast);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                         // smem slot free once these MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tfull[as]);    // accumulator complete
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ================================================================ epilogue (8 warps)
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access (warp id % 4)
    const int half = ew >> 2;          // column half handled by this warp
    constexpr int HC = BN / 2;         // columns per warp
    constexpr int NCH = HC / 32;       // 32-column chunks per warp (1, 2 or 4)
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int r_in_tile = q * 32 + lane;
      long long grow;
      bool row_ok;
      if constexpr (MODE == A_POSCONV) {
        grow = (long long)mt * p.seq_stride + r_in_tile;
        row_ok = r_in_tile < p.seq_valid;
      } else {
        grow = (long long)mt * BM + r_in_tile;
        row_ok = grow < p.M;
      }
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
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
        if (row_ok) {
          const uint32_t* vv = v[c & 1];
          const int cbase = col0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = cbase + j;
            if (col < p.N) {
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(vv[j + i]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
              }
              if constexpr (ACT == ACT_QUICK_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = quick_gelu(x[i]);
              } else if constexpr (ACT == ACT_GELU_ERF) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = gelu_erf(x[i]);
              }
              if (p.res != nullptr) {
                const float4* rp = reinterpret_cast<const float4*>(p.res + grow * p.ld_res + col);
                const float4 r0 = rp[0], r1 = rp[1];
                x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
                x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
              }
              if (p.out_f32 != nullptr) {
                float4* op = reinterpret_cast<float4*>(p.out_f32 + grow * p.ld_f32 + col);
                op[0] = make_float4(x[0], x[1], x[2], x[3]);
                op[1] = make_float4(x[4], x[5], x[6], x[7]);
              }
              if (p.out_bf16 != nullptr) {
                uint4 o;
                o.x = pack_bf16x2(x[0], x[1]);
                o.y = pack_bf16x2(x[2], x[3]);
                o.z = pack_bf16x2(x[4], x[5]);
                o.w = pack_bf16x2(x[6], x[7]);
                *reinterpret_cast<uint4*>(p.out_bf16 + grow * p.ld_bf16 + col) = o;
              }
            }
          }
        }
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

// rank-2 or rank-3 bf16 tensor, dim0 contiguous; box = {64, box1, 1}; 128B swizzle; OOB -> zeros
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              uint32_t box1) {
  EncodeTiledFn fn = get_encode_fn();
  AGB_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  cuuint64_t gdim[3] = {dims[0], dims[1], rank == 3 ? dims[2] : 1};
  cuuint64_t gstr[2] = {strides_bytes[0], rank == 3 ? strides_bytes[1] : 0};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

template <int BN, int MODE, int ACT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const KParams& kp, cudaStream_t stream) {
  auto kern = gemm_kernel<BN, MODE, ACT>;
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes));
    configured = true;
  }
  const int tiles = kp.m_tiles * kp.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kNumThreads, Cfg<BN>::kSmemBytes, stream>>>(tmA, tmB, kp);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int gemm_bf16(const GemmArgs& a, cudaStream_t stream) {
  AGB_REQUIRE(a.A != nullptr && a.W != nullptr, "gemm: null operand");
  AGB_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  AGB_REQUIRE(a.N % 8 == 0, "gemm: N must be a multiple of 8");
  AGB_REQUIRE(a.out_f32 != nullptr || a.out_bf16 != nullptr, "gemm: no output");
  AGB_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
              "gemm: operands must be 16-byte aligned");

  const double rows_eff = a.mode == A_POSCONV ? (double)a.nseq * a.seq_valid : (double)a.M;
  LaunchTimer timer(PROF_GEMM, 2.0 * rows_eff * a.N * a.K,
                    2.0 * (rows_eff * a.K + (double)a.N * a.K) + rows_eff * a.N * (a.out_f32 ? 4.0 : 0.0) +
                        rows_eff * a.N * (a.out_bf16 ? 2.0 : 0.0) + rows_eff * a.N * (a.res ? 4.0 : 0.0),
                    stream);
  KParams kp{};
  kp.M = a.M; kp.N = a.N; kp.K = a.K;
  kp.bias = a.bias; kp.res = a.res; kp.ld_res = a.ld_res;
  kp.out_f32 = a.out_f32; kp.ld_f32 = a.ld_f32;
  kp.out_bf16 = a.out_bf16; kp.ld_bf16 = a.ld_bf16;
  kp.seq_stride = a.seq_stride; kp.seq_valid = a.seq_valid;
  if (a.out_f32) AGB_REQUIRE(a.ld_f32 % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0, "gemm: f32 out alignment");
  if (a.out_bf16) AGB_REQUIRE(a.ld_bf16 % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out_bf16) & 15) == 0, "gemm: bf16 out alignment");
  if (a.res) AGB_REQUIRE(a.ld_res % 4 == 0 && (reinterpret_cast<uintptr_t>(a.res) & 15) == 0, "gemm: residual alignment");
  if (a.bias) AGB_REQUIRE((reinterpret_cast<uintptr_t>(a.bias) & 15) == 0, "gemm: bias alignment");

  CUtensorMap tmA, tmB;
  if (a.mode == A_PLAIN) {
    AGB_REQUIRE(a.K % BK == 0, "gemm: K must be a multiple of 64");
    AGB_REQUIRE(a.lda >= a.K && (a.lda * 2) % 16 == 0, "gemm: bad lda");
    uint64_t dA[2] = {(uint64_t)a.K, (uint64_t)a.M};
    uint64_t sA[1] = {(uint64_t)a.lda * 2};
    if (int r = make_tmap(&tmA, a.A, 2, dA, sA, BM)) return r;
    kp.num_kb = a.K / BK;
    kp.m_tiles = (a.M + BM - 1) / BM;
  } else if (a.mode == A_CONV3) {
    AGB_REQUIRE(a.K == 1536, "gemm: conv3 mode expects K = 3 * 512");
    uint64_t dA[2] = {1024, (uint64_t)a.M};  // pair rows; row M (third tap of the last row) is OOB -> zeros
    uint64_t sA[1] = {1024 * 2};
    if (int r = make_tmap(&tmA, a.A, 2, dA, sA, BM)) return r;
    kp.num_kb = 24;
    kp.m_tiles = (a.M + BM - 1) / BM;
  } else if (a.mode == A_POSCONV) {
    AGB_REQUIRE(a.N == 1024 && a.K == 8192, "gemm: posconv mode expects N = 1024, K = 128 * 64");
    AGB_REQUIRE(a.nseq > 0 && a.seq_valid > 0 && a.seq_valid <= BM && a.seq_stride >= a.seq_valid, "gemm: posconv seq shape");
    uint64_t dA[3] = {1024, (uint64_t)a.seq_valid, (uint64_t)a.nseq};
    uint64_t sA[2] = {1024 * 2, (uint64_t)a.seq_stride * 1024 * 2};
    if (int r = make_tmap(&tmA, a.A, 3, dA, sA, BM)) return r;
    kp.num_kb = 128;
    kp.m_tiles = a.nseq;
  } else {
    AGB_REQUIRE(false, "gemm: unknown A mode");
  }

  uint64_t dB[2] = {(uint64_t)a.K, (uint64_t)a.N};
  uint64_t sB[1] = {(uint64_t)a.K * 2};

  if (a.mode == A_POSCONV) {
    if (int r = make_tmap(&tmB, a.W, 2, dB, sB, 64)) return r;
    kp.n_tiles = 16;
    AGB_REQUIRE(a.act == ACT_GELU_ERF, "gemm: posconv epilogue is bias + GELU (+ residual)");
    return launch<64, A_POSCONV, ACT_GELU_ERF>(tmA, tmB, kp, stream);
  }
  if (int r = make_tmap(&tmB, a.W, 2, dB, sB, 256)) return r;
  kp.n_tiles = (a.N + 255) / 256;
  if (a.mode == A_CONV3) {
    if (a.act == ACT_NONE) return launch<256, A_CONV3, ACT_NONE>(tmA, tmB, kp, stream);
    if (a.act == ACT_GELU_ERF) return launch<256, A_CONV3, ACT_GELU_ERF>(tmA, tmB, kp, stream);
    AGB_REQUIRE(false, "gemm: unsupported activation for conv3 mode");
  }
  switch (a.act) {
    case ACT_NONE: return launch<256, A_PLAIN, ACT_NONE>(tmA, tmB, kp, stream);
    case ACT_QUICK_GELU: return launch<256, A_PLAIN, ACT_QUICK_GELU>(tmA, tmB, kp, stream);
    case ACT_GELU_ERF: return launch<256, A_PLAIN, ACT_GELU_ERF>(tmA, tmB, kp, stream);
    default: break;
  }
  AGB_REQUIRE(false, "gemm: unknown activation");
}

}  // namespace agb
