/* libagb -- B200 (sm_100a) encode-and-fuse path of AffectGPT behind a plain C ABI.
 *
 * Everything a reference-side binding needs: plain pointers and sizes, no torch types.  All data
 * pointers are DEVICE pointers on the context's device unless stated otherwise; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Every function returns 0 on success and a
 * non-zero code on failure; agb_last_error() then holds the text (the Python shim raises RuntimeError
 * with it, which is the reference's error convention: its extractor catches Exception per sample,
 * A/extract_multimodal_features_precompute.py:227-229, 587-589, 610-612).
 * One context per (process, device); a context is not thread-safe.  There is no CPU fallback.
 *
 * Reference interfaces replaced (file:line under /root/reference/AffectGPT):
 *   agb_clip_forward    <- CLIP_VIT_LARGE.forward        my_affectgpt/models/encoder.py:197-208
 *   agb_hubert_forward  <- HUBERT_LARGE.forward          my_affectgpt/models/encoder.py:418-432
 *   agb_video_fuse      <- AffectGPT.encode_video_attention / _mean (+ pre-extracted twins)
 *                                                        my_affectgpt/models/affectgpt.py:505-535, 543-610
 *   agb_audio_fuse      <- AffectGPT.encode_audio_attention / _mean   affectgpt.py:690-716, 719-773
 *   agb_multi_fuse      <- AffectGPT.encode_multi_attention           affectgpt.py:883-908
 *   agb_load_tensor     <- AutoModel.from_pretrained / ckpt['model'] load_state_dict
 *                                                        encoder.py:184, 403; affectgpt.py:1099-1120
 */
#ifndef AGB_H_
#define AGB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGB_VERSION 100

typedef struct agb_ctx agb_ctx;

typedef struct agb_cfg {
  int32_t vit_layers;             /* CLIP ViT-L/14 encoder layers (24) */
  int32_t hubert_layers;          /* HuBERT-large encoder layers (24), >= 3 */
  int32_t hubert_conv_norm;       /* 0 = "layer" (hubert-large family), 1 = "group" (base family) */
  int32_t hubert_do_normalize;    /* preprocessor_config.json do_normalize (1) */
  int32_t llm_dim;                /* projector output width (3584 = Qwen2.5-7B) */
  int32_t max_frames_per_chunk;   /* workspace sizing; 0 = default (512) */
  int32_t max_segments_per_chunk; /* workspace sizing; 0 = default (512) */
  int32_t qf_video_queries;       /* num_video_query_token when video_fusion_type == "qformer", else 0 (not built) */
  int32_t qf_audio_queries;       /* num_audio_query_token when audio_fusion_type == "qformer", else 0 */
  int32_t qf_multi_queries;       /* num_multi_query_token when multi_fusion_type == "qformer", else 0 */
  int32_t vit_dense_last_block;   /* 0 (default): the last ViT block computes only what the pooled CLS row needs (K/V for all
                                     tokens, everything else for the CLS row); 1: all 257 rows, as the reference does */
  int32_t hubert_post_layer_norm; /* 0 (default): do_stable_layer_norm = true, pre-LN blocks (hubert-large family,
                                     HF HubertEncoderStableLayerNorm, modeling_hubert.py); 1: do_stable_layer_norm = false,
                                     post-LN blocks with encoder.layer_norm in front of the stack (HF HubertEncoder /
                                     HubertEncoderLayer) */
  int32_t hubert_no_feat_proj_ln; /* 0 (default): feat_proj_layer_norm = true; 1: the projection reads the raw conv features */
  int32_t reserved[3];
} agb_cfg;

enum { AGB_DTYPE_F32 = 0, AGB_DTYPE_U8 = 1, AGB_DTYPE_BF16 = 2 };
enum { AGB_ACT_NONE = 0, AGB_ACT_QUICK_GELU = 1, AGB_ACT_GELU_ERF = 2 };
enum { AGB_FUSE_ATTENTION = 0, AGB_FUSE_MEAN = 1 };

int agb_version(void);
void agb_default_cfg(agb_cfg* cfg);
int agb_create(agb_ctx** out, int device, const agb_cfg* cfg);
void agb_destroy(agb_ctx* ctx);
/* text of the last error on this context (or the last global error when ctx is NULL) */
const char* agb_last_error(agb_ctx* ctx);

/* Weights.  `name` is "clip.<HF CLIPModel key>", "hubert.<HF HubertModel key>", "fusion.<AffectGPT ckpt['model'] key>"
 * (attention / mean heads) or "qf_video." / "qf_audio." / "qf_multi." + <ckpt key> (Q-Former variants: *_Qformer.bert.*,
 * *_query_tokens, *position_embedding.weight and the 768-wide projectors); fp32 device tensor; the context converts/packs into its
 * own bf16 buffers (QKV concatenated, conv taps interleaved, weight-norm folded). */
int agb_load_tensor(agb_ctx* ctx, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim,
                    void* stream);
/* 1 if `name` is a tensor this build consumes (checkpoints carry extra keys, e.g. masked_spec_embed) */
int agb_has_tensor(agb_ctx* ctx, const char* name);
/* number of still-missing tensors of a family ("clip", "hubert", "fusion"); name of the i-th one */
int agb_missing_count(agb_ctx* ctx, const char* family);
const char* agb_missing_name(agb_ctx* ctx, const char* family, int index);

/* raw_frames: [n,3,224,224], f32 holding 0..255 (AGB_DTYPE_F32) or u8; out: [n,768] f32 */
int agb_clip_forward(agb_ctx* ctx, const void* raw_frames, int in_dtype, int64_t n, float* out, void* stream);
/* wav: [n,32000] f32; zero-mean/unit-var statistics are shared by each run of `seg_per_norm_group`
 * consecutive segments (<= 0: the whole call, the reference's behaviour for one call); out: [n,1024] f32 */
int agb_hubert_forward(agb_ctx* ctx, const float* wav, int64_t n, int64_t seg_per_norm_group, float* out, void* stream);

/* H: [B,T,768] f32 -> llm [B,llm_dim] f32.  bypass_t1: pre-extracted single vector skips the attention */
int agb_video_fuse(agb_ctx* ctx, const float* H, int B, int T, int fusion_type, int bypass_t1, float* llm, void* stream);
/* H: [B,T,1024] f32 -> llm [B,llm_dim] f32 */
int agb_audio_fuse(agb_ctx* ctx, const float* H, int B, int T, int fusion_type, float* llm, void* stream);
/* Hv: [B,Tv,768], Ha: [B,Ta,1024] -> multi_hidden [B,1024], llm [B,llm_dim] */
int agb_multi_fuse(agb_ctx* ctx, const float* Hv, int Tv, const float* Ha, int Ta, int B, float* multi_hidden, float* llm,
                   void* stream);

/* Q-Former variants (`*_fusion_type: qformer`): AffectGPT.encode_video_qformer affectgpt.py:420-468 (+ pre-extracted
 * twin :563-580), encode_audio_qformer :628-664, encode_multi_qformer :848-881 over Qformer.py's 2-layer query
 * transformer.  which: 0 video (Hv [B,Tv,768]), 1 audio (Ha [B,Ta,1024]), 2 multi (both).  hidden: [B,n_q,768] or NULL;
 * llm: [B,n_q,llm_dim]. */
enum { AGB_QF_VIDEO = 0, AGB_QF_AUDIO = 1, AGB_QF_MULTI = 2 };
int agb_qformer_fuse(agb_ctx* ctx, int which, const float* Hv, int Tv, const float* Ha, int Ta, int B, float* hidden,
                     float* llm, void* stream);

/* Input staging (device pointers throughout).
 * agb_stage_faces: load_face, video_processor.py:262-296 - faces [n_src,h,w,3] u8 (OpenFace crop stack), idx [T] int32
 *   (uniform index rule, computed by the host) -> out [T,3,224,224] u8 = cv2.resize(face, (224,224)) bit for bit.
 * agb_stage_audio: load_audio, ImageBind/data.py:188-212 after decode/resample - wav [ch,n] f32, starts [clips] int64
 *   (host, exact rational arithmetic) -> out [clips,clip_len] f32: two channels are averaged, samples past n are 0. */
int agb_stage_faces(const void* faces_u8, int64_t n_src, int h, int w, const int32_t* idx, int T, void* out_u8, void* stream);
int agb_stage_audio(const float* wav, int ch, int64_t n, const int64_t* starts, int clips, int clip_len, float* out,
                    void* stream);
/* agb_resample_audio: the torchaudio.functional.resample call of load_audio (ImageBind/data.py:186-187) - wav [ch,n] f32 at
 *   the file's rate -> out [ch,n_out] f32 at 16 kHz, n_out = ceil(new * n / orig) with orig / new the rates divided by
 *   their gcd; bank [new, taps] f32, taps = 2 * width + orig, is torchaudio's windowed-sinc filter bank (built by the host:
 *   affectgpt_b200.staging.resample_bank). */
int agb_resample_audio(const float* wav, int ch, int64_t n, const float* bank, int orig, int nw, int width, int taps,
                       float* out, int64_t n_out, void* stream);

/* Token splice (AffectGPT.forward, affectgpt.py:956-1009): in place on embeds [B,L,D] (AGB_DTYPE_F32 / AGB_DTYPE_BF16).
 * input_ids [B,L] int64 (device); token_ids / n_query / mod_embeds are HOST arrays of n_mod (<= 8) entries, mod_embeds[i]
 * a device pointer to [B,n_query[i],D] fp32 or NULL.  A modality is spliced into a sample when its placeholder occurs
 * exactly n_query times, consecutively; status (device, [B,n_mod,3] int32 = count, first position, consecutive flag)
 * lets the caller raise the reference's ValueErrors for the other cases. */
int agb_splice_tokens(void* embeds, int dtype, const int64_t* input_ids, int B, int L, int D, int n_mod,
                      const int64_t* token_ids, const int32_t* n_query, const float* const* mod_embeds, int32_t* status,
                      void* stream);

/* Debug taps: after the next forward, copy the named intermediate (fp32 residual stream) to dst.
 * Names: "vit.embed", "vit.block<i>", "hubert.conv" (bf16 [n*100,512]), "hubert.featproj",
 * "hubert.posconv", "hubert.block<i>".  dst == NULL removes the tap. */
int agb_set_tap(agb_ctx* ctx, const char* name, void* dst, int64_t capacity_bytes);

/* kernels launched by this library since process start (for bench.py's gpu_launches) */
int64_t agb_launch_count(void);

/* Per-category device timing with CUDA events around every kernel wrapper (roofline legs of bench.py).
 * Categories: 0 GEMM (tcgen05), 1 attention, 2 LayerNorm, 3 front-end (patchify / conv0 / pooling), 4 other.
 * `flops` / `bytes` are the ALGORITHMIC counts of the wrapped launches. */
#define AGB_PROFILE_CATEGORIES 5
int agb_profile_begin(void);
int agb_profile_end(double* ms, int64_t* count, double* flops, double* bytes);

/* ---- kernel-level entry points (unit tests, micro-benchmarks) ---------------------------------- */
/* C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ res); bf16 operands, fp32 accumulate in TMEM.
 * a_mode: 0 plain (lda elements), 1 conv k3/s2 pair view (K = 1536), 2 grouped positional conv. */
int agb_op_gemm(int a_mode, int act, const void* A, int64_t lda, const void* W, int M, int N, int K, const float* bias,
                const float* res, int64_t ld_res, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                int nseq, int seq_stride, int seq_valid, void* stream);
int agb_op_layernorm(const float* x, int64_t ldx, int64_t rows, int D, const float* gamma, const float* beta, float eps,
                     void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32, void* stream);
int agb_op_layernorm512_bf16(const void* x, int64_t rows, const float* gamma, const float* beta, float eps, int gelu,
                             void* out, void* stream);
int agb_op_attention(const void* qkv_bf16, void* out_bf16, int nseq, int seq_stride, int seq_valid, float scale,
                     void* stream);
int agb_op_patchify(const void* frames, int in_dtype, int64_t n, void* a_patch_bf16, void* stream);
int agb_op_conv0(const float* wav, int64_t nseg, int64_t seg_per_norm_group, int do_normalize, const float* w,
                 const float* bias, const float* gamma, const float* beta, int norm_mode, void* out_bf16, void* stream);
int agb_op_seq_mean_pool(const float* x, int nseq, int seq_stride, int seq_valid, const float* gamma, const float* beta,
                         float eps, float scale, int accumulate, float* pooled, void* stream);
int agb_op_linear_f32(const float* in, int64_t ld_in, const float* W, const float* bias, float* out, int64_t ld_out, int M,
                      int N, int K, void* stream);
int agb_op_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AGB_H_ */
