// Internal declarations shared by the .cu translation units of libagb.so (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>

namespace agb {

using bf16 = __nv_bfloat16;

// thread-local error text; the C ABI copies it into the context
void set_error(const std::string& msg);
const std::string& get_error();

#define AGB_CHECK_CUDA(expr)                                                                         \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::agb::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                       ":" + std::to_string(__LINE__) + ")");                                        \
      return 1;                                                                                      \
    }                                                                                                \
  } while (0)

// after every kernel launch: count it (bench.py reports gpu_launches) and surface launch errors
extern std::atomic<long long> g_launches;
#define AGB_CHECK_LAUNCH()                  \
  do {                                      \
    ++::agb::g_launches;                    \
    AGB_CHECK_CUDA(cudaGetLastError());     \
  } while (0)

#define AGB_REQUIRE(cond, msg)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::agb::set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +             \
                       std::to_string(__LINE__) + ")");                                   \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

// profiling categories (agb_profile_*): wrapper-level CUDA-event timing, off unless enabled
enum ProfCat : int { PROF_GEMM = 0, PROF_ATTENTION = 1, PROF_LAYERNORM = 2, PROF_FRONTEND = 3, PROF_OTHER = 4 };
bool profiling_on();
struct LaunchTimer {
  LaunchTimer(int cat, double flops, double bytes, cudaStream_t s);
  ~LaunchTimer();
  int idx_;
  cudaStream_t s_;
};

enum Act : int { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_GELU_ERF = 2 };
enum AMode : int {
  A_PLAIN = 0,    // A is [M, K] row-major (row stride lda elements)
  A_CONV3 = 1,    // k=3,s=2 conv over channels-last rows of 512: A is the "pair" view [M(+1), 1024];
                  // logical K = 1536 = pair row m (1024) ++ first half of pair row m+1 (512)
  A_POSCONV = 2,  // grouped k=128 conv over [nseq, seq_stride, 1024]; one M tile per sequence,
                  // one N tile (64) per group, K block kb = tap kb of that group's 64 channels
};

struct GemmArgs {
  int mode = A_PLAIN;
  int act = ACT_NONE;
  const bf16* A = nullptr;
  int64_t lda = 0;      // elements (PLAIN); ignored otherwise
  const bf16* W = nullptr;  // [N, K] row-major (K contiguous); POSCONV: [1024, 8192]
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;  // [N] or null
  const float* res = nullptr;   // fp32 [M, ld_res] added after activation, or null (may alias out_f32)
  int64_t ld_res = 0;
  float* out_f32 = nullptr;
  int64_t ld_f32 = 0;
  bf16* out_bf16 = nullptr;
  int64_t ld_bf16 = 0;
  // POSCONV only
  int nseq = 0, seq_stride = 0, seq_valid = 0;
  // walk the tiles from the last row block to the first (see encoder_block: consumers start where the producer finished)
  bool reverse = false;
  // optional: ln_out[M, 1024] (bf16) = LayerNorm(out_f32 rows after the += ) * ln_gamma + ln_beta, fused into the kernel's
  // tail where the shape allows (N = 1024 residual GEMM on CTA pairs), otherwise a separate pass; ln_counters: >= ceil(M/256)
  // zero-initialised ints owned by the caller (left zero by every launch)
  bf16* ln_out = nullptr;
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float ln_eps = 1e-5f;
  int* ln_counters = nullptr;
};

// tcgen05/TMEM/TMA GEMM (gemm_tcgen05.cu).  Returns 0 on success.
int gemm_bf16(const GemmArgs& a, cudaStream_t stream);

// layernorm.cu ------------------------------------------------------------------------------
// y = LN(x) * gamma + beta over the last dim (D = 1024 or 512), eps.  x fp32 or bf16 rows with
// row stride ldx; writes bf16 and/or fp32.
int layernorm_f32_in(const float* x, int64_t ldx, int64_t rows, int D, const float* gamma, const float* beta,
                     float eps, bf16* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32,
                     cudaStream_t s, bool reverse = false);
// in-place over bf16 rows of 512 channels: y = act(LN(x)) (HuBERT conv stack, "layer" variant)
int layernorm512_bf16_gelu(bf16* x, int64_t rows, const float* gamma, const float* beta, float eps,
                           bool apply_gelu, bf16* out, cudaStream_t s);

// attention.cu -------------------------------------------------------------------------------
// qkv: [nseq * seq_stride, 3*1024] bf16 (q | k | v, 16 heads x 64); out: [nseq * seq_stride, 1024] bf16
int attention_mha16x64(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale,
                       cudaStream_t s, bool reverse = false);

// attention_tc2.cu: tcgen05 path (S / P / O in TMEM, two CTAs per SM); returns -1 when no instantiation covers the
// shape (the caller then falls back to the mma.sync kernel)
int attention_tc2(const bf16* qkv, bf16* out, int nseq, int seq_stride, int seq_valid, float scale, cudaStream_t s,
                  bool reverse = false);

// vit.cu -------------------------------------------------------------------------------------
int vit_patchify(const void* frames, int in_dtype /*0=f32 0..255, 1=u8*/, int64_t nframes, bf16* a_patch /*[n*256,640]*/,
                 cudaStream_t s);
int vit_embed_preln(const float* patch_out /*[n*256,1024]*/, const float* cls, const float* pos,
                    const float* g, const float* b, float eps, int64_t nframes, float* resid /*[n*257,1024]*/,
                    cudaStream_t s);

// last block only: softmax(q_cls K^T * scale) V for the CLS query of each frame; q_cls [n,1024], qkv [n*257,3072] (K, V
// columns filled), out [n,1024]
int vit_cls_attention(const bf16* q_cls, const bf16* qkv, int64_t nframes, float scale, bf16* out, cudaStream_t s);

// hubert.cu ----------------------------------------------------------------------------------
int wave_stats(const float* wav, int64_t nseg, int64_t samples, int64_t seg_per_group, float* stats /*[ngroups,2] mean,rstd*/,
               int do_normalize, double* acc_scratch /*[ngroups,2]*/, cudaStream_t s);
int hubert_conv0(const float* wav, const float* stats, int64_t seg_per_group, int64_t seg_offset, int64_t nseg, const float* w /*[512,10]*/,
                 const float* bias, const float* g, const float* b, int norm_mode /*0 layer,1 group*/,
                 float* gn_scratch /*nseg*1024 doubles, group mode only*/, bf16* out /*[nseg*6400,512]*/, cudaStream_t s);
// conv0_tc.cu: the "layer" variant on tcgen05 (im2col tile x filter matrix in TMEM, LayerNorm + GELU epilogue)
int hubert_conv0_layer_tc(const float* wav, const float* stats, int64_t seg_per_group, int64_t seg_offset, int64_t nseg,
                          const float* w, const float* bias, const float* g, const float* b, bf16* wp_scratch /*[512*32]*/,
                          bf16* out, cudaStream_t s);
int gelu_bf16_inplace(bf16* x, int64_t n, cudaStream_t s);
// pooled[seq, :] (+)= scale * mean_{t<valid} f(x[seq, t, :]); f = identity or LayerNorm(gamma,beta)
int seq_mean_pool(const float* x, int nseq, int seq_stride, int seq_valid, const float* gamma, const float* beta,
                  float eps, float scale, int accumulate, float* pooled /*[nseq,1024]*/, cudaStream_t s);

// fusion.cu ----------------------------------------------------------------------------------
// out[M,N] = in[M,K] @ W[N,K]^T + bias  (fp32 SIMT; the heads are ~0.03 GF per clip)
int linear_f32(const float* in, int64_t ld_in, const float* W, const float* bias, float* out, int64_t ld_out, int M,
               int N, int K, cudaStream_t s);
// fused[b,:] = sum_t H[b,t,:] * (H[b,t,:].w + bias)   (no softmax: affectgpt.py:526-529)
int attn_pool_nosoftmax(const float* H, int B, int T, int D, const float* w, const float* bias, int bypass_if_t1,
                        float* fused, cudaStream_t s);
int mean_over_t(const float* H, int B, int T, int D, float* out, cudaStream_t s);
// multi[b,:] = att[b,0] * v[b,:] + att[b,1] * a[b,:]
int weighted_pair_sum(const float* v, const float* a, const float* att, int B, int D, float* out, cudaStream_t s);
int concat2_f32(const float* a, const float* b, int B, int Da, int Db, float* out, cudaStream_t s);

// qformer.cu ---------------------------------------------------------------------------------
struct QFLayerW {   // reference keys under <x>_Qformer.bert.encoder.layer.<i>.
  float *sq_w, *sq_b, *sk_w, *sk_b, *sv_w, *sv_b, *so_w, *so_b, *sln_g, *sln_b;   // attention.{self.*,output.*}
  float *cq_w, *cq_b, *ck_w, *ck_b, *cv_w, *cv_b, *co_w, *co_b, *cln_g, *cln_b;   // crossattention.{...}
  float *fi_w, *fi_b, *fo_w, *fo_b, *fln_g, *fln_b;                               // intermediate_query / output_query
};
struct QFormerW {
  int nq = 0, enc_width = 0, pos_rows = 0;
  float *emb_g = nullptr, *emb_b = nullptr;     // bert.embeddings.LayerNorm
  float *query = nullptr;                       // <x>_query_tokens [1, nq, 768]
  float *pos = nullptr;                         // position embedding [pos_rows, enc_width]
  float *proj_w = nullptr, *proj_b = nullptr;   // projector 768 -> llm_dim
  QFLayerW layer[2] = {};
};
struct QFormerWs {
  float *x = nullptr, *t0 = nullptr, *t1 = nullptr, *t2 = nullptr, *kx = nullptr, *vx = nullptr, *ff = nullptr,
        *enc = nullptr, *ev = nullptr, *ea = nullptr;
  int64_t cap_q = 0, cap_kv = 0, cap_enc = 0;
};
int ln_residual_f32(const float* x, int x_rows, const float* res, int64_t rows, int D, const float* g, const float* b,
                    float eps, float* y, cudaStream_t s);
int add_position_f32(const float* in, const float* pos, int B, int T, int D, int t0, int T_total, float* out,
                     cudaStream_t s);
int gelu_erf_f32_inplace(float* x, int64_t n, cudaStream_t s);
int mha12x64_f32(const float* q, const float* k, const float* v, int B, int Tq, int Tk, float* out, cudaStream_t s);
int qformer_forward_f32(const QFormerW& W, const float* enc, int B, int T, QFormerWs& ws, cudaStream_t s);

// staging.cu ---------------------------------------------------------------------------------
// faces [n_src,h,w,3] u8 -> out [T,3,224,224] u8: cv2.resize(INTER_LINEAR) of the crops selected by idx_dev, bit-exact
int stage_faces_u8(const uint8_t* faces, int64_t n_src, int h, int w, const int* idx_dev, int T, uint8_t* out,
                   cudaStream_t s);
// wav [ch,n] f32 -> out [clips,clip_len]: mono mix (2 channels: mean), zero padding, slices at starts_dev
int stage_audio_f32(const float* wav, int ch, int64_t n, const long long* starts_dev, int clips, int clip_len, float* out,
                    cudaStream_t s);

// wav [ch,n] f32 -> out [ch,n_out]: polyphase sinc FIR, bank [nw, taps = 2*width + orig] (torchaudio.functional.resample)
int resample_f32(const float* wav, int ch, int64_t n, const float* bank, int orig, int nw, int width, int taps, float* out,
                 int64_t n_out, cudaStream_t s);

// splice.cu ----------------------------------------------------------------------------------
// in-place on embeds [B,L,D] (fp32 or bf16); status [B,n_mod,3] = {count, first position or -1, consecutive}
int splice_tokens(void* embeds_inout, int dtype, const long long* ids, int B, int L, int D, int n_mod,
                  const long long* token_ids, const int* n_query, const float* const* mod_embeds, int* status_dev,
                  cudaStream_t s);

// pack.cu ------------------------------------------------------------------------------------
int convert_f32_to_bf16(const float* src, bf16* dst, int64_t n, cudaStream_t s);
// dst[r, c] (row stride ld_dst) = src[r, c] for r<rows, c<cols; zero-fills c in [cols, cols_pad)
int pack_rows_f32_to_bf16(const float* src, int64_t rows, int64_t cols, bf16* dst, int64_t ld_dst, int64_t cols_pad,
                          cudaStream_t s);
// conv weight [512, 512, k] -> [512, k*512] with K index = tap*512 + c_in
int pack_conv_weight(const float* src, int k, bf16* dst, cudaStream_t s);
// pos-conv: w = g * v / ||v||_(dims 0,1) per tap; v [1024,64,128], g [128]; dst [1024, 128*64], K index = tap*64 + c_in
int pack_posconv_weight(const float* v, const float* g, bf16* dst, float* norm_scratch /*[128]*/, cudaStream_t s);

}  // namespace agb
