// tcgen05 multi-head attention for the path's short non-causal sequences (ViT: 257 tokens, HuBERT: 99 frames),
// 16 heads x 64.  Persistent CTAs (one per SM), work item = (sequence, head), query tiles of 128 rows.
//
//   warp 0     : TMA producer   K/V of the head (double-buffered per work item), Q tiles (ring of 2)
//   warp 1     : MMA issuer     S = Q K^T  (SS, fp32 in TMEM, keys split in two column halves A | B)
//                               O = P V    (TS: P is read from TMEM as the A operand, V from smem MN-major)
//   warp 2     : TMEM allocator
//   warps 4..7 : softmax + epilogue warpgroup, one thread per query row (TMEM lane):
//                pass 1 row max over S; pass 2 exp2 / row sum / bf16 P written back over S (tcgen05.st);
//                epilogue O * (1 / row sum) -> bf16 -> global
//
// Exact softmax (global row max, fp32), no online rescaling: both score halves of a tile are in TMEM before the
// first exp.  The MMA issue order is software-pipelined (PV_A(t), QK_A(t+1), PV_B(t), QK_B(t+1)) so the tensor pipe
// works on the next tile's scores while the warpgroup is still exponentiating the current tile's second half.
//
// TMEM columns: S_A [0,NKA) | S_B [NKA,NKA+NKB) | O0 | O1 (64 each);  P_A / P_B alias the first half of S_A / S_B
// (bf16 pairs per 32-bit column).  ViT: NKA=128, NKB=144 (keys 257 -> 272, last 15 masked); HuBERT: NKA=112, NKB=0.
//
// Replaces: CLIPAttention.forward (HF modeling_clip.py:262-332, scale 1/8, fp32 softmax) and
// HubertAttention.forward (HF modeling_hubert.py:262-370).
#include <cuda.h>

#include <mutex>

#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

int make_tmap_generic(CUtensorMap* out, int dtype_is_f32, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, uint32_t box0, uint32_t box1);

namespace {

constexpr int kThreads = 256;   // warps 0-3 control, 4-7 softmax/epilogue

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// V tile [keys][64 d] as the B operand: N (= d) contiguous -> MN-major, 128B swizzle, 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t m, uint32_t n, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct AParams {
  int nseq, seq_stride, seq_valid, qtiles;
  float scale_log2e;
  bf16* out;
};

template <int NKA, int NKB>
struct ACfg {
  static constexpr int NK = NKA + NKB;
  static constexpr int kKVBytes = NK * 128;                 // one of K or V for one head
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kSmem = 2 * 2 * kKVBytes + 2 * kQBytes + 1024 + 256;
  static constexpr int kColO = NK;                          // O0 at NK, O1 at NK + 64
  static constexpr int kTmemCols = (NK + 128) <= 256 ? 256 : 512;
};

// masked running max over one chunk of W (32 or 16) score columns starting at key index key0
template <int W>
__device__ __forceinline__ float chunk_max(const uint32_t* v, int key0, int valid, float m) {
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const float s = __uint_as_float(v[j]);
    m = (key0 + j < valid) ? fmaxf(m, s) : m;
  }
  return m;
}

template <int NKA, int NKB>
__global__ void __launch_bounds__(kThreads, 1)
mha_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKA,
              const __grid_constant__ CUtensorMap tmKB, const AParams p) {
  using C = ACfg<NKA, NKB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smK = smem;                              // [2 stages][NK rows][128 B]
  uint8_t* smV = smem + 2 * C::kKVBytes;
  uint8_t* smQ = smem + 4 * C::kKVBytes;            // [2][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smQ + 2 * C::kQBytes);
  uint64_t* kv_full = bars;         // [2]
  uint64_t* kv_empty = bars + 2;    // [2]
  uint64_t* q_full = bars + 4;      // [2]
  uint64_t* q_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [1]
  uint64_t* pa_ready = bars + 9;    // [1]
  uint64_t* pb_ready = bars + 10;   // [1]
  uint64_t* o_full = bars + 11;     // [2]
  uint64_t* o_empty = bars + 13;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKA);
    if (NKB) tma_prefetch_desc(&tmKB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4);
    }
    mbar_init(s_full, 1);
    mbar_init(pa_ready, 4);
    mbar_init(pb_ready, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_items = p.nseq * 16;
  const int QT = p.qtiles;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int j = 0, t = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++j) {
        const int seq = item >> 4, head = item & 15;
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1u, 11);
        mbar_arrive_expect_tx(&kv_full[st], 2 * C::kKVBytes);
        uint8_t* k = smK + st * C::kKVBytes;
        uint8_t* v = smV + st * C::kKVBytes;
        tma_load_3d(k, &tmKA, &kv_full[st], 1024 + head * 64, 0, seq, kEvictNormal);
        tma_load_3d(v, &tmKA, &kv_full[st], 2048 + head * 64, 0, seq, kEvictNormal);
        if (NKB) {
          tma_load_3d(k + NKA * 128, &tmKB, &kv_full[st], 1024 + head * 64, NKA, seq, kEvictNormal);
          tma_load_3d(v + NKA * 128, &tmKB, &kv_full[st], 2048 + head * 64, NKA, seq, kEvictNormal);
        }
        for (int qt = 0; qt < QT; ++qt, ++t) {
          const int qs = t & 1;
          mbar_wait(&q_empty[qs], ((t >> 1) & 1) ^ 1u, 12);
          mbar_arrive_expect_tx(&q_full[qs], C::kQBytes);
          tma_load_3d(smQ + qs * C::kQBytes, &tmQ, &q_full[qs], head * 64, qt * 128, seq, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idS_A = idesc_bf16(128, NKA, 0);
      constexpr uint32_t idS_B = idesc_bf16(128, NKB ? NKB : 16, 0);
      constexpr uint32_t idO = idesc_bf16(128, 64, 1);
      const int my_items = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      const int T = my_items * QT;
      auto issue_qk = [&](int t, int half) {
        const int j = t / QT;
        const int st = j & 1, qs = t & 1;
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smQ + qs * C::kQBytes));
        const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smK + st * C::kKVBytes + (half ? NKA * 128 : 0)));
        const uint32_t d = tmem_base + (half ? NKA : 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(d, da + 2 * k, db + 2 * k, half ? idS_B : idS_A, k != 0);
      };
      auto issue_pv = [&](int t, int half) {
        const int j = t / QT;
        const int st = j & 1;
        const uint32_t d = tmem_base + C::kColO + (t & 1) * 64;
        const uint32_t a = tmem_base + (half ? NKA : 0);
        const uint64_t db = umma_desc_mnmajor_sw128(smem_u32(smV + st * C::kKVBytes + (half ? NKA * 128 : 0)));
        const int steps = (half ? NKB : NKA) / 16;
        for (int k = 0; k < steps; ++k) umma_bf16_ts(d, a + 8 * k, db + 128 * k, idO, (half | k) != 0 ? 1u : 0u);
      };
      auto wait_inputs = [&](int t) {   // operands of tile t's QK: its Q tile and (first tile of an item) its K/V
        const int j = t / QT;
        if (t % QT == 0) mbar_wait(&kv_full[j & 1], (j >> 1) & 1, 13);
        mbar_wait(&q_full[t & 1], (t >> 1) & 1, 14);
        tc_fence_after();
      };
      if (T > 0) {
        wait_inputs(0);
        issue_qk(0, 0);
        if (NKB) issue_qk(0, 1);
        umma_commit(&q_empty[0]);
        umma_commit(s_full);
      }
      for (int t = 0; t < T; ++t) {
        const uint32_t ph = t & 1;
        const bool has_next = t + 1 < T;
        mbar_wait(pa_ready, ph, 15);
        mbar_wait(&o_empty[t & 1], ((t >> 1) & 1) ^ 1u, 16);
        tc_fence_after();
        issue_pv(t, 0);
        if (has_next) {
          wait_inputs(t + 1);
          issue_qk(t + 1, 0);
        }
        if (NKB) {
          mbar_wait(pb_ready, ph, 17);
          tc_fence_after();
          issue_pv(t, 1);
        }
        umma_commit(&o_full[t & 1]);
        if ((t % QT) == QT - 1) umma_commit(&kv_empty[(t / QT) & 1]);   // last tile of the item: K/V stage free
        if (has_next) {
          if (NKB) issue_qk(t + 1, 1);
          umma_commit(&q_empty[(t + 1) & 1]);
          umma_commit(s_full);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue warpgroup
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const int valid = p.seq_valid;
    const float sc = p.scale_log2e;
    constexpr int NK = C::NK;
    constexpr int NFULL = NK / 32;        // full 32-column chunks
    constexpr bool TAIL = (NK % 32) != 0; // one 16-column tail chunk
    int t = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int seq = item >> 4, head = item & 15;
      for (int qt = 0; qt < QT; ++qt, ++t) {
        mbar_wait(s_full, t & 1, 18);
        tc_fence_after();
        // ---------------- pass 1: row max (masked)
        float m = -INFINITY;
        {
          uint32_t v[2][32];
          tmem_ld_32x32b_x32(lane_base, v[0]);
#pragma unroll
          for (int c = 0; c < NFULL; ++c) {
            tmem_ld_wait();
            if (c + 1 < NFULL) tmem_ld_32x32b_x32(lane_base + (c + 1) * 32, v[(c + 1) & 1]);
            else if (TAIL) tmem_ld_32x32b_x16(lane_base + NFULL * 32, v[(c + 1) & 1]);
            if (c * 32 < valid) m = chunk_max<32>(v[c & 1], c * 32, valid, m);
          }
          if (TAIL) {
            tmem_ld_wait();
            if (NFULL * 32 < valid) m = chunk_max<16>(v[NFULL & 1], NFULL * 32, valid, m);
          }
        }
        const float mo = m * sc;
        float l = 0.f;
        // ---------------- pass 2: p = exp2(s*sc - m*sc), row sum, bf16 P written over S (half as wide)
        {
          uint32_t v[2][32];
          tmem_ld_32x32b_x32(lane_base, v[0]);
#pragma unroll
          for (int c = 0; c < NFULL; ++c) {
            tmem_ld_wait();
            if (c + 1 < NFULL) tmem_ld_32x32b_x32(lane_base + (c + 1) * 32, v[(c + 1) & 1]);
            else if (TAIL) tmem_ld_32x32b_x16(lane_base + NFULL * 32, v[(c + 1) & 1]);
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int key = c * 32 + j;
              float p0 = exp2f(fmaf(__uint_as_float(v[c & 1][j]), sc, -mo));
              float p1 = exp2f(fmaf(__uint_as_float(v[c & 1][j + 1]), sc, -mo));
              p0 = key < valid ? p0 : 0.f;
              p1 = key + 1 < valid ? p1 : 0.f;
              l += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
            // S columns [32c, 32c+32) of half h -> P columns half_base + (32c - half_base)/2
            const int s_col = c * 32;
            const int hb = (NKB && s_col >= NKA) ? NKA : 0;
            tmem_st_32x32b_x16(lane_base + hb + ((s_col - hb) >> 1), pk);
            if (NKB && (c + 1) * 32 == NKA) {        // half A complete (NKA is a multiple of 32 when NKB > 0)
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(pa_ready);
            }
          }
          if (TAIL) {
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const int key = NFULL * 32 + j;
              float p0 = exp2f(fmaf(__uint_as_float(v[NFULL & 1][j]), sc, -mo));
              float p1 = exp2f(fmaf(__uint_as_float(v[NFULL & 1][j + 1]), sc, -mo));
              p0 = key < valid ? p0 : 0.f;
              p1 = key + 1 < valid ? p1 : 0.f;
              l += p0 + p1;
              pk[j >> 1] = pack_bf16x2(p0, p1);
            }
            const int s_col = NFULL * 32;
            const int hb = (NKB && s_col >= NKA) ? NKA : 0;
            tmem_st_32x32b_x8(lane_base + hb + ((s_col - hb) >> 1), pk);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(NKB ? pb_ready : pa_ready);
        }
        // ---------------- epilogue: O / l -> bf16 -> global
        const int ob = t & 1;
        mbar_wait(&o_full[ob], (t >> 1) & 1, 19);
        tc_fence_after();
        {
          uint32_t o[2][32];
          tmem_ld_32x32b_x32(lane_base + C::kColO + ob * 64, o[0]);
          tmem_ld_32x32b_x32(lane_base + C::kColO + ob * 64 + 32, o[1]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_empty[ob]);
          const int row = qt * 128 + row_in_tile;
          if (row < valid) {
            const float inv = 1.0f / l;
            bf16* dst = p.out + ((long long)seq * p.seq_stride + row) * 1024 + head * 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(o[h][j]) * inv, __uint_as_float(o[h][j + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(o[h][j + 2]) * inv, __uint_as_float(o[h][j + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(o[h][j + 4]) * inv, __uint_as_float(o[h][j + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(o[h][j + 6]) * inv, __uint_as_float(o[h][j + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + h * 32 + j) = w;
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

template <int NKA, int NKB>
int launch_tc(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s) {
  using C = ACfg<NKA, NKB>;
  auto kern = mha_tc_kernel<NKA, NKB>;
  static bool configured = false;
  if (!configured) {
    AGB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  // one 3-D view of the packed QKV activation: {3072 channels, seq_valid rows, nseq}; rows >= seq_valid read as zero
  uint64_t dims[3] = {3072, (uint64_t)seq_valid, (uint64_t)nseq};
  uint64_t strides[2] = {3072 * 2, (uint64_t)seq_stride * 3072 * 2};
  CUtensorMap tq, tka, tkb;
  if (int r = make_tmap_generic(&tq, 0, qkv, 3, dims, strides, 64, 128)) return r;
  if (int r = make_tmap_generic(&tka, 0, qkv, 3, dims, strides, 64, NKA)) return r;
  if (int r = make_tmap_generic(&tkb, 0, qkv, 3, dims, strides, 64, NKB ? NKB : 16)) return r;
  AParams p{};
  p.nseq = nseq; p.seq_stride = seq_stride; p.seq_valid = seq_valid;
  p.qtiles = (seq_valid + 127) / 128;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = out;
  const int items = nseq * 16;
  const int resident = sm_count() * (C::kTmemCols <= 256 ? 2 : 1);   // two CTAs per SM when TMEM and smem allow
  const int grid = items < resident ? items : resident;
  kern<<<grid, kThreads, C::kSmem, s>>>(tq, tka, tkb, p);
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// returns -1 when the shape is not covered by an instantiation (caller falls back to the mma.sync kernel)
int attention_tc(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s) {
  if (seq_valid <= 112) return launch_tc<112, 0>(qkv, out, nseq, seq_stride, seq_valid, scale, s);
  if (seq_valid > 128 && seq_valid <= 272) return launch_tc<128, 144>(qkv, out, nseq, seq_stride, seq_valid, scale, s);
  return -1;
}

}  // namespace agb
