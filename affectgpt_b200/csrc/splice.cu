// Token splice: overwrite the placeholder runs of the prompt embedding with the fused modality tokens.
//
// Replaces the per-sample torch.where / torch.cat chain of AffectGPT.forward
// (A/my_affectgpt/models/affectgpt.py:956-1009; same rule in Chat.answer_sample, conversation_video.py:330-360):
// for every sample and every modality whose placeholder id occurs in that sample's input_ids, the id must occur
// exactly n_query times and consecutively; the run [start, start + n_query) is replaced by the modality's tokens.
// One block per (sample, modality): scan the ids, validate, copy.  The verdict (count, first position, consecutive)
// goes to a status array so that the host wrapper can raise exactly the reference's exceptions.
#include "agb_internal.h"
#include "agb_ptx.cuh"

namespace agb {

namespace {

struct SpliceMods {
  long long token_id[8];
  int n_query[8];
  const float* embeds[8];   // [B, n_query, D] fp32 or null
};

template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 cvt_out<bf16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) splice_kernel(T* __restrict__ out, const long long* __restrict__ ids, int L, int D,
                                                    SpliceMods mods, int n_mod, int* __restrict__ status) {
  __shared__ int s_cnt[8], s_min[8], s_max[8];
  const int b = blockIdx.x, m = blockIdx.y;
  const long long tid = mods.token_id[m];
  const long long* row = ids + (long long)b * L;
  int cnt = 0, lo = L, hi = -1;
  for (int p = threadIdx.x; p < L; p += 256)
    if (row[p] == tid) { ++cnt; lo = min(lo, p); hi = max(hi, p); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_cnt[warp] = cnt; s_min[warp] = lo; s_max[warp] = hi; }
  __syncthreads();
  cnt = 0; lo = L; hi = -1;
  for (int i = 0; i < 8; ++i) { cnt += s_cnt[i]; lo = min(lo, s_min[i]); hi = max(hi, s_max[i]); }
  const bool consecutive = cnt > 0 && (hi - lo + 1 == cnt);
  if (threadIdx.x == 0) {
    int* st = status + ((long long)b * n_mod + m) * 3;
    st[0] = cnt; st[1] = cnt > 0 ? lo : -1; st[2] = consecutive ? 1 : 0;
  }
  const int nq = mods.n_query[m];
  const float* src = mods.embeds[m];
  if (cnt != nq || !consecutive || src == nullptr) return;       // nothing to do, or an error the host will raise
  src += (long long)b * nq * D;
  T* dst = out + ((long long)b * L + lo) * D;
  for (long long i = threadIdx.x; i < (long long)nq * D; i += 256) dst[i] = cvt_out<T>(src[i]);
}

}  // namespace

int splice_tokens(void* embeds_inout, int dtype, const long long* ids, int B, int L, int D, int n_mod,
                  const long long* token_ids, const int* n_query, const float* const* mod_embeds, int* status_dev,
                  cudaStream_t s) {
  AGB_REQUIRE(embeds_inout && ids && status_dev && B > 0 && L > 0 && D > 0, "splice: bad arguments");
  AGB_REQUIRE(n_mod > 0 && n_mod <= 8, "splice: 1..8 modalities");
  SpliceMods mods{};
  for (int i = 0; i < n_mod; ++i) {
    mods.token_id[i] = token_ids[i];
    mods.n_query[i] = n_query[i];
    mods.embeds[i] = mod_embeds[i];
  }
  dim3 grid(B, n_mod);
  if (dtype == 0) splice_kernel<float><<<grid, 256, 0, s>>>(static_cast<float*>(embeds_inout), ids, L, D, mods, n_mod, status_dev);
  else if (dtype == 2) splice_kernel<bf16><<<grid, 256, 0, s>>>(static_cast<bf16*>(embeds_inout), ids, L, D, mods, n_mod, status_dev);
  else AGB_REQUIRE(false, "splice: embeddings must be fp32 (0) or bf16 (2)");
  AGB_CHECK_LAUNCH();
  return 0;
}

}  // namespace agb
