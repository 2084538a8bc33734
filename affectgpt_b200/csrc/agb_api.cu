// C ABI of libagb.so: context, weight packing by checkpoint key, forward orchestration.
// Mirrors (reference): CLIP_VIT_LARGE / HUBERT_LARGE wrappers (A/my_affectgpt/models/encoder.py:177-208,
// 395-432) over HF CLIPVisionTransformer / HubertModel, and AffectGPT's fusion heads + projectors
// (A/my_affectgpt/models/affectgpt.py:505-535, 690-716, 883-908).
#include <atomic>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/agb.h"
#include "agb_internal.h"

namespace agb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }
std::atomic<long long> g_launches{0};

}  // namespace agb

using namespace agb;

namespace {

struct EncLayerW {
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *lnA_g = nullptr, *lnA_b = nullptr, *lnB_g = nullptr, *lnB_b = nullptr;
};

struct Loader {
  std::vector<int64_t> shape;
  std::function<int(const float*, cudaStream_t)> fn;
  bool loaded = false;
};

struct Tap {
  void* dst;
  int64_t cap;
};

}  // namespace

struct agb_ctx {
  int device = 0;
  agb_cfg cfg{};
  std::string err;
  std::vector<void*> allocs;
  std::map<std::string, Loader> loaders;
  std::map<std::string, Tap> taps;

  // ---- CLIP ViT-L/14
  bf16* v_patch_w = nullptr;  // [1024, 640]
  float *v_cls = nullptr, *v_pos = nullptr, *v_pre_g = nullptr, *v_pre_b = nullptr, *v_post_g = nullptr, *v_post_b = nullptr;
  bf16* v_proj = nullptr;  // [768, 1024]
  std::vector<EncLayerW> v_layers;
  // ---- HuBERT-large
  float* h_conv0_w = nullptr;  // [512,10]
  float *h_conv_b[7] = {}, *h_conv_g[7] = {}, *h_conv_be[7] = {};
  bf16* h_conv_w[7] = {};  // 1..6 packed [512, k*512]
  float *h_fp_g = nullptr, *h_fp_b = nullptr, *h_fp_bias = nullptr;
  bf16* h_fp_w = nullptr;  // [1024,512]
  float *h_pos_v = nullptr, *h_pos_g = nullptr, *h_pos_bias = nullptr, *h_pos_norm = nullptr;
  bf16* h_pos_w = nullptr;  // [1024, 8192]
  bool h_pos_v_loaded = false, h_pos_g_loaded = false, h_pos_packed = false;
  float *h_enc_g = nullptr, *h_enc_b = nullptr;
  std::vector<EncLayerW> h_layers;
  // ---- fusion heads / projectors (fp32)
  std::map<std::string, float*> f;

  // ---- workspaces
  int64_t v_cap = 0;  // frames
  bf16 *v_apatch = nullptr, *v_xln = nullptr, *v_qkv = nullptr, *v_attn = nullptr, *v_h1 = nullptr, *v_clsln = nullptr;
  bf16 *v_qcls = nullptr, *v_acls = nullptr;   // pruned last block: CLS queries / CLS attention output [frames,1024]
  float *v_patch_out = nullptr, *v_resid = nullptr;
  int *v_lncnt = nullptr, *h_lncnt = nullptr;   // per-256-row-block arrival counters of the fused LayerNorm (zero at rest)
  int64_t h_cap = 0;  // segments
  bf16 *h_c0 = nullptr, *h_c1 = nullptr, *h_xln = nullptr, *h_qkv = nullptr, *h_attn = nullptr, *h_h1 = nullptr;
  float *h_resid = nullptr, *h_stats = nullptr;
  double *h_stat_acc = nullptr, *h_gn_acc = nullptr;
  int64_t h_stat_cap = 0;
  int64_t f_cap = 0;
  float *f_a = nullptr, *f_b = nullptr, *f_c = nullptr, *f_d = nullptr, *f_e = nullptr;
  // ---- Q-Former variants (fp32): [0] video, [1] audio, [2] multi; multi also owns its 768/1024 -> 1024 embeddings
  QFormerW qf[3];
  float *qf_mv_w = nullptr, *qf_mv_b = nullptr, *qf_ma_w = nullptr, *qf_ma_b = nullptr;
  QFormerWs qf_ws;
};

namespace {

int fail(agb_ctx* c, int code) {
  if (c) c->err = get_error();
  return code;
}

template <typename T>
int dev_alloc(agb_ctx* c, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  AGB_CHECK_CUDA(cudaMalloc(&q, count * sizeof(T)));
  if (zero) {
    // the memset runs on the legacy default stream, which non-blocking streams (torch side streams) do not wait for:
    // finish it before any caller stream can touch the buffer (allocation happens only when a workspace grows)
    AGB_CHECK_CUDA(cudaMemset(q, 0, count * sizeof(T)));
    AGB_CHECK_CUDA(cudaDeviceSynchronize());
  }
  c->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

int dev_free(agb_ctx* c, void* p) {
  if (!p) return 0;
  for (auto it = c->allocs.begin(); it != c->allocs.end(); ++it)
    if (*it == p) {
      c->allocs.erase(it);
      break;
    }
  cudaFree(p);
  return 0;
}

// free + null: a failed re-allocation must never leave a dangling pointer behind a still-valid capacity
template <typename T>
void dev_release(agb_ctx* c, T*& p) {
  dev_free(c, p);
  p = nullptr;
}

void add_loader(agb_ctx* c, const std::string& name, std::vector<int64_t> shape,
                std::function<int(const float*, cudaStream_t)> fn) {
  Loader l;
  l.shape = std::move(shape);
  l.fn = std::move(fn);
  c->loaders[name] = std::move(l);
}

int64_t numel(const std::vector<int64_t>& s) {
  int64_t n = 1;
  for (auto d : s) n *= d;
  return n;
}

// plain fp32 copy into a freshly allocated buffer
int reg_f32(agb_ctx* c, const std::string& name, std::vector<int64_t> shape, float** slot) {
  const int64_t n = numel(shape);
  if (int r = dev_alloc(c, slot, (size_t)n)) return r;
  float* dst = *slot;
  add_loader(c, name, shape, [dst, n](const float* src, cudaStream_t s) -> int {
    AGB_CHECK_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
  });
  return 0;
}

// fp32 [rows, cols] -> bf16 rows of an existing matrix (row offset, row stride ld, zero pad to cols_pad)
void reg_bf16_rows(agb_ctx* c, const std::string& name, std::vector<int64_t> shape, bf16* dst, int64_t rows, int64_t cols,
                   int64_t ld, int64_t cols_pad) {
  add_loader(c, name, std::move(shape), [=](const float* src, cudaStream_t s) -> int {
    return pack_rows_f32_to_bf16(src, rows, cols, dst, ld, cols_pad, s);
  });
}

int reg_encoder_layer(agb_ctx* c, const std::string& prefix, EncLayerW& L, const char* attn, const char* lnA,
                      const char* lnB, const char* fc1, const char* fc2) {
  if (int r = dev_alloc(c, &L.wqkv, 3072ull * 1024)) return r;
  if (int r = dev_alloc(c, &L.wo, 1024ull * 1024)) return r;
  if (int r = dev_alloc(c, &L.w1, 4096ull * 1024)) return r;
  if (int r = dev_alloc(c, &L.w2, 1024ull * 4096)) return r;
  if (int r = dev_alloc(c, &L.bqkv, 3072)) return r;
  const char* qkv[3] = {"q_proj", "k_proj", "v_proj"};
  for (int i = 0; i < 3; ++i) {
    reg_bf16_rows(c, prefix + attn + "." + qkv[i] + ".weight", {1024, 1024}, L.wqkv + (size_t)i * 1024 * 1024, 1024, 1024,
                  1024, 1024);
    float* bdst = L.bqkv + i * 1024;
    add_loader(c, prefix + attn + "." + qkv[i] + ".bias", {1024}, [bdst](const float* src, cudaStream_t s) -> int {
      AGB_CHECK_CUDA(cudaMemcpyAsync(bdst, src, 1024 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      return 0;
    });
  }
  reg_bf16_rows(c, prefix + attn + ".out_proj.weight", {1024, 1024}, L.wo, 1024, 1024, 1024, 1024);
  if (int r = reg_f32(c, prefix + attn + ".out_proj.bias", {1024}, &L.bo)) return r;
  reg_bf16_rows(c, prefix + fc1 + ".weight", {4096, 1024}, L.w1, 4096, 1024, 1024, 1024);
  if (int r = reg_f32(c, prefix + fc1 + ".bias", {4096}, &L.b1)) return r;
  reg_bf16_rows(c, prefix + fc2 + ".weight", {1024, 4096}, L.w2, 1024, 4096, 4096, 4096);
  if (int r = reg_f32(c, prefix + fc2 + ".bias", {1024}, &L.b2)) return r;
  if (int r = reg_f32(c, prefix + lnA + ".weight", {1024}, &L.lnA_g)) return r;
  if (int r = reg_f32(c, prefix + lnA + ".bias", {1024}, &L.lnA_b)) return r;
  if (int r = reg_f32(c, prefix + lnB + ".weight", {1024}, &L.lnB_g)) return r;
  if (int r = reg_f32(c, prefix + lnB + ".bias", {1024}, &L.lnB_b)) return r;
  return 0;
}

int maybe_pack_posconv(agb_ctx* c, cudaStream_t s) {
  if (c->h_pos_v_loaded && c->h_pos_g_loaded) {
    if (int r = pack_posconv_weight(c->h_pos_v, c->h_pos_g, c->h_pos_w, c->h_pos_norm, s)) return r;
    c->h_pos_packed = true;
  }
  return 0;
}

int register_weights(agb_ctx* c) {
  // ------------------------------------------------------------------ CLIP
  const std::string V = "clip.vision_model.";
  if (int r = dev_alloc(c, &c->v_patch_w, 1024ull * 640)) return r;
  reg_bf16_rows(c, V + "embeddings.patch_embedding.weight", {1024, 3, 14, 14}, c->v_patch_w, 1024, 588, 640, 640);
  if (int r = reg_f32(c, V + "embeddings.class_embedding", {1024}, &c->v_cls)) return r;
  if (int r = reg_f32(c, V + "embeddings.position_embedding.weight", {257, 1024}, &c->v_pos)) return r;
  if (int r = reg_f32(c, V + "pre_layrnorm.weight", {1024}, &c->v_pre_g)) return r;
  if (int r = reg_f32(c, V + "pre_layrnorm.bias", {1024}, &c->v_pre_b)) return r;
  if (int r = reg_f32(c, V + "post_layernorm.weight", {1024}, &c->v_post_g)) return r;
  if (int r = reg_f32(c, V + "post_layernorm.bias", {1024}, &c->v_post_b)) return r;
  if (int r = dev_alloc(c, &c->v_proj, 768ull * 1024)) return r;
  reg_bf16_rows(c, "clip.visual_projection.weight", {768, 1024}, c->v_proj, 768, 1024, 1024, 1024);
  c->v_layers.resize(c->cfg.vit_layers);
  for (int i = 0; i < c->cfg.vit_layers; ++i)
    if (int r = reg_encoder_layer(c, V + "encoder.layers." + std::to_string(i) + ".", c->v_layers[i], "self_attn",
                                  "layer_norm1", "layer_norm2", "mlp.fc1", "mlp.fc2"))
      return r;

  // ------------------------------------------------------------------ HuBERT
  const std::string H = "hubert.";
  const int kk[7] = {10, 3, 3, 3, 3, 2, 2};
  for (int i = 0; i < 7; ++i) {
    const std::string p = H + "feature_extractor.conv_layers." + std::to_string(i) + ".";
    if (i == 0) {
      if (int r = reg_f32(c, p + "conv.weight", {512, 1, 10}, &c->h_conv0_w)) return r;
    } else {
      if (int r = dev_alloc(c, &c->h_conv_w[i], 512ull * 512 * kk[i])) return r;
      bf16* dst = c->h_conv_w[i];
      const int k = kk[i];
      add_loader(c, p + "conv.weight", {512, 512, k},
                 [dst, k](const float* src, cudaStream_t s) -> int { return pack_conv_weight(src, k, dst, s); });
    }
    if (int r = reg_f32(c, p + "conv.bias", {512}, &c->h_conv_b[i])) return r;
    const bool has_norm = c->cfg.hubert_conv_norm == 0 || i == 0;
    if (has_norm) {
      if (int r = reg_f32(c, p + "layer_norm.weight", {512}, &c->h_conv_g[i])) return r;
      if (int r = reg_f32(c, p + "layer_norm.bias", {512}, &c->h_conv_be[i])) return r;
    }
  }
  if (!c->cfg.hubert_no_feat_proj_ln) {
    if (int r = reg_f32(c, H + "feature_projection.layer_norm.weight", {512}, &c->h_fp_g)) return r;
    if (int r = reg_f32(c, H + "feature_projection.layer_norm.bias", {512}, &c->h_fp_b)) return r;
  }
  if (int r = dev_alloc(c, &c->h_fp_w, 1024ull * 512)) return r;
  reg_bf16_rows(c, H + "feature_projection.projection.weight", {1024, 512}, c->h_fp_w, 1024, 512, 512, 512);
  if (int r = reg_f32(c, H + "feature_projection.projection.bias", {1024}, &c->h_fp_bias)) return r;
  if (int r = reg_f32(c, H + "encoder.pos_conv_embed.conv.bias", {1024}, &c->h_pos_bias)) return r;
  if (int r = dev_alloc(c, &c->h_pos_v, 1024ull * 64 * 128)) return r;
  if (int r = dev_alloc(c, &c->h_pos_g, 128)) return r;
  if (int r = dev_alloc(c, &c->h_pos_norm, 128)) return r;
  if (int r = dev_alloc(c, &c->h_pos_w, 1024ull * 8192)) return r;
  {
    // weight-norm parametrisation: original0 = g [1,1,128], original1 = v [1024,64,128]
    // (transformers 4.x checkpoints name them weight_g / weight_v; both spellings fill the same slot)
    agb_ctx* cc = c;
    auto load_g = [cc](const float* src, cudaStream_t s) -> int {
      AGB_CHECK_CUDA(cudaMemcpyAsync(cc->h_pos_g, src, 128 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      cc->h_pos_g_loaded = true;
      return maybe_pack_posconv(cc, s);
    };
    auto load_v = [cc](const float* src, cudaStream_t s) -> int {
      AGB_CHECK_CUDA(cudaMemcpyAsync(cc->h_pos_v, src, 1024ull * 64 * 128 * sizeof(float), cudaMemcpyDeviceToDevice, s));
      cc->h_pos_v_loaded = true;
      return maybe_pack_posconv(cc, s);
    };
    add_loader(c, H + "encoder.pos_conv_embed.conv.parametrizations.weight.original0", {1, 1, 128}, load_g);
    add_loader(c, H + "encoder.pos_conv_embed.conv.parametrizations.weight.original1", {1024, 64, 128}, load_v);
  }
  if (int r = reg_f32(c, H + "encoder.layer_norm.weight", {1024}, &c->h_enc_g)) return r;
  if (int r = reg_f32(c, H + "encoder.layer_norm.bias", {1024}, &c->h_enc_b)) return r;
  c->h_layers.resize(c->cfg.hubert_layers);
  for (int i = 0; i < c->cfg.hubert_layers; ++i)
    if (int r = reg_encoder_layer(c, H + "encoder.layers." + std::to_string(i) + ".", c->h_layers[i], "attention",
                                  "layer_norm", "final_layer_norm", "feed_forward.intermediate_dense",
                                  "feed_forward.output_dense"))
      return r;

  // ------------------------------------------------------------------ fusion heads / projectors
  const int D = c->cfg.llm_dim;
  struct FS { const char* name; int out, in; };
  const FS fs[] = {{"video_attention_mlp", 1, 768},   {"affectgpt_proj", D, 768},     {"audio_attention_mlp", 1, 1024},
                   {"audio_llama_proj", D, 1024},     {"multi_video_embs", 1024, 768}, {"multi_audio_embs", 1024, 1024},
                   {"attention_mlp", 1024, 2048},     {"fc_att", 2, 1024},             {"multi_llama_proj", D, 1024}};
  for (const FS& e : fs) {
    if (int r = reg_f32(c, std::string("fusion.") + e.name + ".weight", {e.out, e.in}, &c->f[std::string(e.name) + ".weight"])) return r;
    if (int r = reg_f32(c, std::string("fusion.") + e.name + ".bias", {e.out}, &c->f[std::string(e.name) + ".bias"])) return r;
  }
  // ------------------------------------------------------------------ Q-Former variants (built only when requested)
  struct QS { const char* fam; const char* stem; const char* pos_key; int pos_rows; int enc_width; int nq; const char* proj; };
  const QS qs[3] = {
      {"qf_video.", "video", "video_frame_position_embedding.weight", 32, 768, c->cfg.qf_video_queries, "affectgpt_proj"},
      {"qf_audio.", "audio", "audio_position_embedding.weight", 8, 1024, c->cfg.qf_audio_queries, "audio_llama_proj"},
      {"qf_multi.", "multi", "multi_position_embedding.weight", 264, 1024, c->cfg.qf_multi_queries, "multi_llama_proj"}};
  for (int i = 0; i < 3; ++i) {
    const QS& q = qs[i];
    if (q.nq <= 0) continue;
    QFormerW& W = c->qf[i];
    W.nq = q.nq; W.enc_width = q.enc_width; W.pos_rows = q.pos_rows;
    const std::string F = q.fam, S = q.stem;
    if (int r = reg_f32(c, F + q.pos_key, {q.pos_rows, q.enc_width}, &W.pos)) return r;
    if (int r = reg_f32(c, F + S + "_query_tokens", {1, q.nq, 768}, &W.query)) return r;
    if (int r = reg_f32(c, F + q.proj + ".weight", {D, 768}, &W.proj_w)) return r;
    if (int r = reg_f32(c, F + q.proj + ".bias", {D}, &W.proj_b)) return r;
    const std::string Bk = F + S + "_Qformer.bert.";
    if (int r = reg_f32(c, Bk + "embeddings.LayerNorm.weight", {768}, &W.emb_g)) return r;
    if (int r = reg_f32(c, Bk + "embeddings.LayerNorm.bias", {768}, &W.emb_b)) return r;
    for (int l = 0; l < 2; ++l) {
      QFLayerW& L = W.layer[l];
      const std::string P = Bk + "encoder.layer." + std::to_string(l) + ".";
      struct LW { const char* key; int out, in; float** w; float** b; };
      const LW lw[] = {{"attention.self.query", 768, 768, &L.sq_w, &L.sq_b},
                       {"attention.self.key", 768, 768, &L.sk_w, &L.sk_b},
                       {"attention.self.value", 768, 768, &L.sv_w, &L.sv_b},
                       {"attention.output.dense", 768, 768, &L.so_w, &L.so_b},
                       {"crossattention.self.query", 768, 768, &L.cq_w, &L.cq_b},
                       {"crossattention.self.key", 768, q.enc_width, &L.ck_w, &L.ck_b},
                       {"crossattention.self.value", 768, q.enc_width, &L.cv_w, &L.cv_b},
                       {"crossattention.output.dense", 768, 768, &L.co_w, &L.co_b},
                       {"intermediate_query.dense", 3072, 768, &L.fi_w, &L.fi_b},
                       {"output_query.dense", 768, 3072, &L.fo_w, &L.fo_b}};
      for (const LW& e : lw) {
        if (int r = reg_f32(c, P + e.key + ".weight", {e.out, e.in}, e.w)) return r;
        if (int r = reg_f32(c, P + e.key + ".bias", {e.out}, e.b)) return r;
      }
      struct LN { const char* key; float** g; float** b; };
      const LN ln[] = {{"attention.output.LayerNorm", &L.sln_g, &L.sln_b},
                       {"crossattention.output.LayerNorm", &L.cln_g, &L.cln_b},
                       {"output_query.LayerNorm", &L.fln_g, &L.fln_b}};
      for (const LN& e : ln) {
        if (int r = reg_f32(c, P + e.key + ".weight", {768}, e.g)) return r;
        if (int r = reg_f32(c, P + e.key + ".bias", {768}, e.b)) return r;
      }
    }
    if (i == 2) {
      if (int r = reg_f32(c, F + "multi_video_embs.weight", {1024, 768}, &c->qf_mv_w)) return r;
      if (int r = reg_f32(c, F + "multi_video_embs.bias", {1024}, &c->qf_mv_b)) return r;
      if (int r = reg_f32(c, F + "multi_audio_embs.weight", {1024, 1024}, &c->qf_ma_w)) return r;
      if (int r = reg_f32(c, F + "multi_audio_embs.bias", {1024}, &c->qf_ma_b)) return r;
    }
  }
  return 0;
}

int check_family_loaded(agb_ctx* c, const std::string& family) {
  for (auto& kv : c->loaders)
    if (kv.first.compare(0, family.size() + 1, family + ".") == 0 && !kv.second.loaded) {
      set_error("weights not loaded: " + kv.first + " (and possibly more; see agb_missing_count)");
      return 4;
    }
  return 0;
}

// the attention / mean heads are loaded per variant in use: check exactly the tensors a call reads
int check_heads_loaded(agb_ctx* c, std::initializer_list<const char*> heads) {
  for (const char* h : heads)
    for (const char* suffix : {".weight", ".bias"}) {
      auto it = c->loaders.find(std::string("fusion.") + h + suffix);
      if (it == c->loaders.end() || !it->second.loaded) {
        set_error(std::string("weights not loaded: fusion.") + h + suffix);
        return 4;
      }
    }
  return 0;
}

void tap(agb_ctx* c, const std::string& name, const void* src, int64_t bytes, cudaStream_t s) {
  auto it = c->taps.find(name);
  if (it == c->taps.end() || it->second.dst == nullptr) return;
  const int64_t n = bytes < it->second.cap ? bytes : it->second.cap;
  cudaMemcpyAsync(it->second.dst, src, (size_t)n, cudaMemcpyDeviceToDevice, s);
}

// ------------------------------------------------------------------------------------ workspaces
int ensure_vit_ws(agb_ctx* c, int64_t frames) {
  if (frames <= c->v_cap) return 0;
  c->v_cap = 0;
  dev_release(c, c->v_apatch); dev_release(c, c->v_xln); dev_release(c, c->v_qkv); dev_release(c, c->v_attn);
  dev_release(c, c->v_h1); dev_release(c, c->v_clsln); dev_release(c, c->v_patch_out); dev_release(c, c->v_resid);
  dev_release(c, c->v_qcls); dev_release(c, c->v_acls); dev_release(c, c->v_lncnt);
  const size_t M = (size_t)frames * 257;
  if (int r = dev_alloc(c, &c->v_apatch, (size_t)frames * 256 * 640)) return r;
  if (int r = dev_alloc(c, &c->v_patch_out, (size_t)frames * 256 * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_resid, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_xln, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_qkv, M * 3072)) return r;
  if (int r = dev_alloc(c, &c->v_attn, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_h1, M * 4096)) return r;
  if (int r = dev_alloc(c, &c->v_clsln, (size_t)frames * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_qcls, (size_t)frames * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_acls, (size_t)frames * 1024)) return r;
  if (int r = dev_alloc(c, &c->v_lncnt, M / 256 + 2)) return r;
  c->v_cap = frames;
  return 0;
}

int ensure_hub_ws(agb_ctx* c, int64_t segs) {
  if (segs <= c->h_cap) return 0;
  c->h_cap = 0;
  dev_release(c, c->h_c0); dev_release(c, c->h_c1); dev_release(c, c->h_xln); dev_release(c, c->h_qkv);
  dev_release(c, c->h_attn); dev_release(c, c->h_h1); dev_release(c, c->h_resid); dev_release(c, c->h_gn_acc);
  dev_release(c, c->h_lncnt);
  const size_t M = (size_t)segs * 100;
  if (int r = dev_alloc(c, &c->h_c0, (size_t)segs * 6400 * 512 + 4096)) return r;
  if (int r = dev_alloc(c, &c->h_c1, (size_t)segs * 3200 * 512 + 4096)) return r;
  if (int r = dev_alloc(c, &c->h_resid, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->h_xln, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->h_qkv, M * 3072)) return r;
  if (int r = dev_alloc(c, &c->h_attn, M * 1024)) return r;
  if (int r = dev_alloc(c, &c->h_h1, M * 4096)) return r;
  // group variant: per-(segment, channel) fp64 sum / sumsq; layer variant: the 32 KB packed conv0 filter matrix
  if (int r = dev_alloc(c, &c->h_gn_acc, c->cfg.hubert_conv_norm == 1 ? (size_t)segs * 1024 : (size_t)4096)) return r;
  if (int r = dev_alloc(c, &c->h_lncnt, M / 256 + 2)) return r;
  c->h_cap = segs;
  return 0;
}

int ensure_stats(agb_ctx* c, int64_t groups) {
  if (groups <= c->h_stat_cap) return 0;
  c->h_stat_cap = 0;                      // capacities are set only after every allocation succeeded
  dev_release(c, c->h_stats);
  dev_release(c, c->h_stat_acc);
  if (int r = dev_alloc(c, &c->h_stats, (size_t)groups * 2)) return r;
  if (int r = dev_alloc(c, &c->h_stat_acc, (size_t)groups * 2)) return r;
  c->h_stat_cap = groups;
  return 0;
}

int ensure_fuse_ws(agb_ctx* c, int64_t B) {
  if (B <= c->f_cap) return 0;
  c->f_cap = 0;
  dev_release(c, c->f_a); dev_release(c, c->f_b); dev_release(c, c->f_c); dev_release(c, c->f_d); dev_release(c, c->f_e);
  if (int r = dev_alloc(c, &c->f_a, (size_t)B * 1024)) return r;
  if (int r = dev_alloc(c, &c->f_b, (size_t)B * 1024)) return r;
  if (int r = dev_alloc(c, &c->f_c, (size_t)B * 2048)) return r;
  if (int r = dev_alloc(c, &c->f_d, (size_t)B * 1024)) return r;
  if (int r = dev_alloc(c, &c->f_e, (size_t)B * 2)) return r;
  c->f_cap = B;
  return 0;
}

// ------------------------------------------------------------------------------------ encoder block
// pre-LN transformer block shared by CLIP (quick_gelu) and HuBERT stable-LN (erf GELU)
// Every tensor between two kernels of a block (0.3-1.1 GB) is larger than the 126 MB L2, but the LAST ~100 MB a producer
// wrote are still there when its consumer starts.  The kernels of a chain therefore walk their rows in alternating
// directions ("zigzag", `flip` toggles per launch): a consumer begins with the rows its producer finished with and reads them
// from L2 instead of HBM.  Results do not depend on the order (AGB_ZIGZAG=0 switches it off for A/B runs).
bool zigzag_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AGB_ZIGZAG"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// On entry xln holds LayerNorm_A(resid) (written by the previous block's FC2, or by the caller for the first block).  The
// two residual GEMMs emit the LayerNorm that follows them (out-proj -> this block's LN_B; FC2 -> the NEXT block's LN_A when
// `next` is given), fused into the GEMM's tail where the shape allows (gemm_tcgen05.cu, "LNF"): per block the chain is
// QKV -> MHA -> out-proj(+LN) -> FC1 -> FC2(+LN), five kernels, no stand-alone LayerNorm pass.
int encoder_block(const EncLayerW& L, const EncLayerW* next, float* resid, bf16* xln, bf16* qkv, bf16* attn, bf16* h1, int* lncnt,
                  int64_t M, int nseq, int seq_stride, int seq_valid, int act, cudaStream_t s, bool& flip) {
  const bool zz = zigzag_on();
  auto dir = [&]() { const bool r = zz && flip; flip = !flip; return r; };
  GemmArgs g;
  g.A = xln; g.lda = 1024; g.W = L.wqkv; g.M = (int)M; g.N = 3072; g.K = 1024; g.bias = L.bqkv;
  g.out_bf16 = qkv; g.ld_bf16 = 3072; g.reverse = dir();
  if (int r = gemm_bf16(g, s)) return r;
  if (int r = attention_mha16x64(qkv, attn, nseq, seq_stride, seq_valid, 0.125f, s, dir())) return r;
  g = GemmArgs();
  g.A = attn; g.lda = 1024; g.W = L.wo; g.M = (int)M; g.N = 1024; g.K = 1024; g.bias = L.bo;
  g.res = resid; g.ld_res = 1024; g.out_f32 = resid; g.ld_f32 = 1024; g.reverse = dir();
  g.ln_out = xln; g.ln_gamma = L.lnB_g; g.ln_beta = L.lnB_b; g.ln_eps = 1e-5f; g.ln_counters = lncnt;
  if (int r = gemm_bf16(g, s)) return r;
  g = GemmArgs();
  g.A = xln; g.lda = 1024; g.W = L.w1; g.M = (int)M; g.N = 4096; g.K = 1024; g.bias = L.b1; g.act = act;
  g.out_bf16 = h1; g.ld_bf16 = 4096; g.reverse = dir();
  if (int r = gemm_bf16(g, s)) return r;
  g = GemmArgs();
  g.A = h1; g.lda = 4096; g.W = L.w2; g.M = (int)M; g.N = 1024; g.K = 4096; g.bias = L.b2;
  g.res = resid; g.ld_res = 1024; g.out_f32 = resid; g.ld_f32 = 1024; g.reverse = dir();
  if (next != nullptr) {
    g.ln_out = xln; g.ln_gamma = next->lnA_g; g.ln_beta = next->lnA_b; g.ln_eps = 1e-5f; g.ln_counters = lncnt;
  }
  return gemm_bf16(g, s);
}

// Last ViT block when only the pooled CLS row is consumed (CLIPVisionTransformer: pooled_output = last_hidden_state[:, 0],
// modeling_clip.py:685-686): LN1 and the K/V projection over all 257 rows, then Q, attention, out-proj, LN2 and the MLP for
// the CLS row of each frame only (strided views of the residual stream: row pitch 257 * 1024).  Same arithmetic for that
// row as the dense block; 5.66 of the block's 6.75 GFLOP per frame are not executed.
int vit_last_block_cls(agb_ctx* c, const EncLayerW& L, int64_t F, cudaStream_t s) {
  const int64_t M = F * 257, pitch = 257 * 1024;
  // v_xln already holds LayerNorm_1(resid): written by the previous block's FC2 (or by clip_chunk for a 1-layer tower)
  GemmArgs g;
  g.A = c->v_xln; g.lda = 1024; g.W = L.wqkv + 1024ull * 1024; g.M = (int)M; g.N = 2048; g.K = 1024; g.bias = L.bqkv + 1024;
  g.out_bf16 = c->v_qkv + 1024; g.ld_bf16 = 3072;
  if (int r = gemm_bf16(g, s)) return r;
  g = GemmArgs();
  g.A = c->v_xln; g.lda = pitch; g.W = L.wqkv; g.M = (int)F; g.N = 1024; g.K = 1024; g.bias = L.bqkv;
  g.out_bf16 = c->v_qcls; g.ld_bf16 = 1024;
  if (int r = gemm_bf16(g, s)) return r;
  if (int r = vit_cls_attention(c->v_qcls, c->v_qkv, F, 0.125f, c->v_acls, s)) return r;
  g = GemmArgs();
  g.A = c->v_acls; g.lda = 1024; g.W = L.wo; g.M = (int)F; g.N = 1024; g.K = 1024; g.bias = L.bo;
  g.res = c->v_resid; g.ld_res = pitch; g.out_f32 = c->v_resid; g.ld_f32 = pitch;
  if (int r = gemm_bf16(g, s)) return r;
  if (int r = layernorm_f32_in(c->v_resid, pitch, F, 1024, L.lnB_g, L.lnB_b, 1e-5f, c->v_clsln, 1024, nullptr, 0, s)) return r;
  g = GemmArgs();
  g.A = c->v_clsln; g.lda = 1024; g.W = L.w1; g.M = (int)F; g.N = 4096; g.K = 1024; g.bias = L.b1; g.act = ACT_QUICK_GELU;
  g.out_bf16 = c->v_h1; g.ld_bf16 = 4096;
  if (int r = gemm_bf16(g, s)) return r;
  g = GemmArgs();
  g.A = c->v_h1; g.lda = 4096; g.W = L.w2; g.M = (int)F; g.N = 1024; g.K = 4096; g.bias = L.b2;
  g.res = c->v_resid; g.ld_res = pitch; g.out_f32 = c->v_resid; g.ld_f32 = pitch;
  return gemm_bf16(g, s);
}

int clip_chunk(agb_ctx* c, const void* frames, int in_dtype, int64_t F, float* out, cudaStream_t s) {
  const int64_t M = F * 257;
  if (int r = vit_patchify(frames, in_dtype, F, c->v_apatch, s)) return r;
  GemmArgs g;
  g.A = c->v_apatch; g.lda = 640; g.W = c->v_patch_w; g.M = (int)(F * 256); g.N = 1024; g.K = 640;
  g.out_f32 = c->v_patch_out; g.ld_f32 = 1024;
  if (int r = gemm_bf16(g, s)) return r;
  if (int r = vit_embed_preln(c->v_patch_out, c->v_cls, c->v_pos, c->v_pre_g, c->v_pre_b, 1e-5f, F, c->v_resid, s)) return r;
  tap(c, "vit.embed", c->v_resid, M * 1024 * 4, s);
  // the last block is pruned to the CLS row unless the caller wants it dense (cfg) or taps its full output
  const int last = c->cfg.vit_layers - 1;
  const bool dense_last = c->cfg.vit_dense_last_block != 0 || c->taps.count("vit.block" + std::to_string(last)) != 0;
  bool flip = true;        // embed_preln wrote the rows in ascending order: the first LayerNorm starts from the end
  AGB_CHECK_CUDA(cudaMemsetAsync(c->v_lncnt, 0, (size_t)(M / 256 + 2) * sizeof(int), s));
  {                        // LayerNorm_1 of the first block; every later one is emitted by the preceding FC2
    const EncLayerW& L0 = c->v_layers[0];
    const bool r0 = zigzag_on() && flip;
    flip = !flip;
    if (int r = layernorm_f32_in(c->v_resid, 1024, M, 1024, L0.lnA_g, L0.lnA_b, 1e-5f, c->v_xln, 1024, nullptr, 0, s, r0)) return r;
  }
  for (int l = 0; l < c->cfg.vit_layers; ++l) {
    if (l == last && !dense_last) {
      if (int r = vit_last_block_cls(c, c->v_layers[l], F, s)) return r;
      break;
    }
    if (int r = encoder_block(c->v_layers[l], l + 1 < c->cfg.vit_layers ? &c->v_layers[l + 1] : nullptr, c->v_resid, c->v_xln,
                              c->v_qkv, c->v_attn, c->v_h1, c->v_lncnt, M, (int)F, 257, 257, ACT_QUICK_GELU, s, flip))
      return r;
    tap(c, "vit.block" + std::to_string(l), c->v_resid, M * 1024 * 4, s);
  }
  // CLS rows -> post_layernorm -> visual_projection (no bias)   (modeling_clip.py:685-686, 860-861)
  if (int r = layernorm_f32_in(c->v_resid, 257 * 1024, F, 1024, c->v_post_g, c->v_post_b, 1e-5f, c->v_clsln, 1024, nullptr,
                               0, s))
    return r;
  g = GemmArgs();
  g.A = c->v_clsln; g.lda = 1024; g.W = c->v_proj; g.M = (int)F; g.N = 768; g.K = 1024; g.out_f32 = out; g.ld_f32 = 768;
  return gemm_bf16(g, s);
}

// do_stable_layer_norm = false: HF HubertEncoder + HubertEncoderLayer (modeling_hubert.py).  The residual stream is
// normalised AFTER every sub-layer:  h0 = LN_enc(x + posconv(x));  a = h + out(MHA(h));  h' = LN(a);  f = h' + FFN(h');
// h = LN_final(f).  hidden_states[l] = the stream entering layer l (already normalised), hidden_states[L] = the last layer's
// output - no extra LayerNorm at the end.  Same kernels as the pre-LN block: the residual GEMMs += into the fp32 stream,
// the LayerNorm then rewrites that stream in place (fp32) and emits the bf16 operand of the next GEMM.
int hubert_post_ln_encoder(agb_ctx* c, int64_t G, float* out, cudaStream_t s) {
  const int64_t M = G * 100;
  const int L = c->cfg.hubert_layers;
  if (int r = layernorm_f32_in(c->h_resid, 1024, M, 1024, c->h_enc_g, c->h_enc_b, 1e-5f, c->h_xln, 1024, c->h_resid, 1024, s)) return r;
  bool first_pool = true;
  for (int l = 0; l < L; ++l) {
    const EncLayerW& W = c->h_layers[l];
    if (l >= L - 3) {  // hidden_states[-4], [-3], [-2]
      if (int r = seq_mean_pool(c->h_resid, (int)G, 100, 99, nullptr, nullptr, 0.f, 0.25f, first_pool ? 0 : 1, out, s)) return r;
      first_pool = false;
    }
    GemmArgs g;
    g.A = c->h_xln; g.lda = 1024; g.W = W.wqkv; g.M = (int)M; g.N = 3072; g.K = 1024; g.bias = W.bqkv;
    g.out_bf16 = c->h_qkv; g.ld_bf16 = 3072;
    if (int r = gemm_bf16(g, s)) return r;
    if (int r = attention_mha16x64(c->h_qkv, c->h_attn, (int)G, 100, 99, 0.125f, s)) return r;
    g = GemmArgs();
    g.A = c->h_attn; g.lda = 1024; g.W = W.wo; g.M = (int)M; g.N = 1024; g.K = 1024; g.bias = W.bo;
    g.res = c->h_resid; g.ld_res = 1024; g.out_f32 = c->h_resid; g.ld_f32 = 1024;
    if (int r = gemm_bf16(g, s)) return r;
    if (int r = layernorm_f32_in(c->h_resid, 1024, M, 1024, W.lnA_g, W.lnA_b, 1e-5f, c->h_xln, 1024, c->h_resid, 1024, s)) return r;
    g = GemmArgs();
    g.A = c->h_xln; g.lda = 1024; g.W = W.w1; g.M = (int)M; g.N = 4096; g.K = 1024; g.bias = W.b1; g.act = ACT_GELU_ERF;
    g.out_bf16 = c->h_h1; g.ld_bf16 = 4096;
    if (int r = gemm_bf16(g, s)) return r;
    g = GemmArgs();
    g.A = c->h_h1; g.lda = 4096; g.W = W.w2; g.M = (int)M; g.N = 1024; g.K = 4096; g.bias = W.b2;
    g.res = c->h_resid; g.ld_res = 1024; g.out_f32 = c->h_resid; g.ld_f32 = 1024;
    if (int r = gemm_bf16(g, s)) return r;
    if (int r = layernorm_f32_in(c->h_resid, 1024, M, 1024, W.lnB_g, W.lnB_b, 1e-5f, c->h_xln, 1024, c->h_resid, 1024, s)) return r;
    tap(c, "hubert.block" + std::to_string(l), c->h_resid, M * 1024 * 4, s);
  }
  return seq_mean_pool(c->h_resid, (int)G, 100, 99, nullptr, nullptr, 0.f, 0.25f, 1, out, s);   // hidden_states[-1]
}

int hubert_chunk(agb_ctx* c, const float* wav, int64_t seg0, int64_t G, int64_t spg, float* out, cudaStream_t s) {
  const bool layer_norm = c->cfg.hubert_conv_norm == 0;
  // conv0 (+norm +GELU) -> c0 [G*6400, 512]
  if (int r = hubert_conv0(wav, c->h_stats, spg, seg0, G, c->h_conv0_w, c->h_conv_b[0], c->h_conv_g[0], c->h_conv_be[0],
                           layer_norm ? 0 : 1, reinterpret_cast<float*>(c->h_gn_acc), c->h_c0, s))
    return r;
  // conv1..6 as GEMMs over the pair view (rows per segment: 3200, 1600, 800, 400, 200, 100)
  bf16* src = c->h_c0;
  bf16* dst = c->h_c1;
  int64_t rows = 3200;
  for (int i = 1; i <= 6; ++i) {
    GemmArgs g;
    g.A = src; g.W = c->h_conv_w[i]; g.M = (int)(G * rows); g.N = 512; g.bias = c->h_conv_b[i];
    g.out_bf16 = dst; g.ld_bf16 = 512;
    if (i <= 4) { g.mode = A_CONV3; g.K = 1536; } else { g.mode = A_PLAIN; g.K = 1024; g.lda = 1024; }
    g.act = layer_norm ? ACT_NONE : ACT_GELU_ERF;
    if (int r = gemm_bf16(g, s)) return r;
    if (layer_norm)
      if (int r = layernorm512_bf16_gelu(dst, G * rows, c->h_conv_g[i], c->h_conv_be[i], 1e-5f, true, dst, s)) return r;
    bf16* t = src; src = dst; dst = t;
    rows /= 2;
  }
  const int64_t M = G * 100;
  tap(c, "hubert.conv", src, M * 512 * 2, s);
  // feature projection: [LN(512) ->] Linear 512->1024 (fp32 residual stream + bf16 copy for the pos-conv)
  if (!c->cfg.hubert_no_feat_proj_ln) {
    if (int r = layernorm512_bf16_gelu(src, M, c->h_fp_g, c->h_fp_b, 1e-5f, false, dst, s)) return r;
  } else {
    dst = src;                                                  // feat_proj_layer_norm = false (modeling_hubert.py, HubertFeatureProjection)
  }
  GemmArgs g;
  g.A = dst; g.lda = 512; g.W = c->h_fp_w; g.M = (int)M; g.N = 1024; g.K = 512; g.bias = c->h_fp_bias;
  g.out_f32 = c->h_resid; g.ld_f32 = 1024;
  if (int r = gemm_bf16(g, s)) return r;
  if (int r = convert_f32_to_bf16(c->h_resid, c->h_xln, M * 1024, s)) return r;   // pos-conv A operand
  tap(c, "hubert.featproj", c->h_resid, M * 1024 * 4, s);
  // positional conv embedding: resid += GELU(grouped conv(resid))
  g = GemmArgs();
  g.mode = A_POSCONV; g.act = ACT_GELU_ERF; g.A = c->h_xln; g.W = c->h_pos_w; g.M = (int)M; g.N = 1024; g.K = 8192;
  g.bias = c->h_pos_bias; g.res = c->h_resid; g.ld_res = 1024; g.out_f32 = c->h_resid; g.ld_f32 = 1024;
  g.nseq = (int)G; g.seq_stride = 100; g.seq_valid = 99;
  if (int r = gemm_bf16(g, s)) return r;
  tap(c, "hubert.posconv", c->h_resid, M * 1024 * 4, s);
  const int L = c->cfg.hubert_layers;
  bool first_pool = true;
  bool flip = true;
  if (c->cfg.hubert_post_layer_norm) return hubert_post_ln_encoder(c, G, out, s);
  AGB_CHECK_CUDA(cudaMemsetAsync(c->h_lncnt, 0, (size_t)(M / 256 + 2) * sizeof(int), s));
  {                        // LayerNorm of the first block; every later one is emitted by the preceding FC2
    const EncLayerW& L0 = c->h_layers[0];
    const bool r0 = zigzag_on() && flip;
    flip = !flip;
    if (int r = layernorm_f32_in(c->h_resid, 1024, M, 1024, L0.lnA_g, L0.lnA_b, 1e-5f, c->h_xln, 1024, nullptr, 0, s, r0)) return r;
  }
  for (int l = 0; l < L; ++l) {
    if (l >= L - 3) {  // hidden_states[-4], [-3], [-2] = residual stream entering layers L-3, L-2, L-1
      if (int r = seq_mean_pool(c->h_resid, (int)G, 100, 99, nullptr, nullptr, 0.f, 0.25f, first_pool ? 0 : 1, out, s)) return r;
      first_pool = false;
    }
    if (int r = encoder_block(c->h_layers[l], l + 1 < L ? &c->h_layers[l + 1] : nullptr, c->h_resid, c->h_xln, c->h_qkv,
                              c->h_attn, c->h_h1, c->h_lncnt, M, (int)G, 100, 99, ACT_GELU_ERF, s, flip))
      return r;
    tap(c, "hubert.block" + std::to_string(l), c->h_resid, M * 1024 * 4, s);
  }
  // hidden_states[-1] = encoder.layer_norm(final residual)
  return seq_mean_pool(c->h_resid, (int)G, 100, 99, c->h_enc_g, c->h_enc_b, 1e-5f, 0.25f, 1, out, s);
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int agb_version(void) { return AGB_VERSION; }

void agb_default_cfg(agb_cfg* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->vit_layers = 24;
  cfg->hubert_layers = 24;
  cfg->hubert_conv_norm = 0;
  cfg->hubert_do_normalize = 1;
  cfg->llm_dim = 3584;
  cfg->max_frames_per_chunk = 512;
  cfg->max_segments_per_chunk = 512;
}

const char* agb_last_error(agb_ctx* ctx) {
  (void)ctx;  // errors are recorded per calling thread; a context is single-threaded by contract
  return get_error().c_str();
}

int agb_create(agb_ctx** out, int device, const agb_cfg* cfg) {
  if (!out) { set_error("agb_create: null out"); return 2; }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    set_error("agb_create: no CUDA device visible (this library has no CPU fallback)");
    return 5;
  }
  if (device < 0 || device >= ndev) { set_error("agb_create: bad device index"); return 2; }
  cudaDeviceProp prop;
  AGB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("agb_create: device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
              ", this library is built for sm_100a only");
    return 5;
  }
  AGB_CHECK_CUDA(cudaSetDevice(device));
  agb_ctx* c = new agb_ctx();
  c->device = device;
  if (cfg) c->cfg = *cfg; else agb_default_cfg(&c->cfg);
  if (c->cfg.vit_layers <= 0) c->cfg.vit_layers = 24;
  if (c->cfg.hubert_layers <= 0) c->cfg.hubert_layers = 24;
  if (c->cfg.llm_dim <= 0) c->cfg.llm_dim = 3584;
  if (c->cfg.max_frames_per_chunk <= 0) c->cfg.max_frames_per_chunk = 512;
  if (c->cfg.max_segments_per_chunk <= 0) c->cfg.max_segments_per_chunk = 512;
  if (const char* e = getenv("AGB_VIT_DENSE_LAST")) c->cfg.vit_dense_last_block = e[0] == '1';   // A/B switch (bench, tools)
  if (c->cfg.hubert_layers < 3) { set_error("agb_create: hubert_layers must be >= 3 (last-4 hidden states)"); delete c; return 2; }
  if (int r = register_weights(c)) { agb_destroy(c); return r; }
  *out = c;
  return 0;
}

void agb_destroy(agb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (void* p : ctx->allocs) cudaFree(p);
  delete ctx;
}

int agb_load_tensor(agb_ctx* ctx, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim,
                    void* stream) {
  if (!ctx || !name || !dev_ptr) { set_error("agb_load_tensor: null argument"); return fail(ctx, 2); }
  std::string key(name);
  // transformers 4.x spelling of the weight-norm parametrisation
  const std::string g4 = "encoder.pos_conv_embed.conv.weight_g", v4 = "encoder.pos_conv_embed.conv.weight_v";
  if (key == "hubert." + g4) key = "hubert.encoder.pos_conv_embed.conv.parametrizations.weight.original0";
  if (key == "hubert." + v4) key = "hubert.encoder.pos_conv_embed.conv.parametrizations.weight.original1";
  auto it = ctx->loaders.find(key);
  if (it == ctx->loaders.end()) { set_error("agb_load_tensor: unknown tensor name '" + key + "'"); return fail(ctx, 6); }
  if (dtype != AGB_DTYPE_F32) { set_error("agb_load_tensor: only fp32 source tensors are accepted"); return fail(ctx, 2); }
  // the full shape must match (leading / trailing singleton dimensions aside): an [in, out] tensor passed for an [out, in]
  // weight has the right element count and would otherwise load silently
  auto squeeze = [](const int64_t* d, int n) {
    std::vector<int64_t> v;
    for (int i = 0; i < n; ++i)
      if (d[i] != 1) v.push_back(d[i]);
    return v;
  };
  if (!shape || ndim <= 0 || squeeze(shape, ndim) != squeeze(it->second.shape.data(), (int)it->second.shape.size())) {
    std::string want, got;
    for (auto d : it->second.shape) want += (want.empty() ? "" : ",") + std::to_string(d);
    for (int i = 0; shape && i < ndim; ++i) got += (got.empty() ? "" : ",") + std::to_string(shape[i]);
    set_error("agb_load_tensor: shape mismatch for '" + key + "': expected [" + want + "], got [" + got + "]");
    return fail(ctx, 7);
  }
  if (int r = it->second.fn(static_cast<const float*>(dev_ptr), static_cast<cudaStream_t>(stream))) return fail(ctx, r);
  it->second.loaded = true;
  return 0;
}

int agb_missing_count(agb_ctx* ctx, const char* family) {
  if (!ctx || !family) return -1;
  const std::string f = std::string(family) + ".";
  int n = 0;
  for (auto& kv : ctx->loaders)
    if (kv.first.compare(0, f.size(), f) == 0 && !kv.second.loaded) ++n;
  return n;
}

const char* agb_missing_name(agb_ctx* ctx, const char* family, int index) {
  if (!ctx || !family) return nullptr;
  const std::string f = std::string(family) + ".";
  int n = 0;
  for (auto& kv : ctx->loaders)
    if (kv.first.compare(0, f.size(), f) == 0 && !kv.second.loaded) {
      if (n == index) return kv.first.c_str();
      ++n;
    }
  return nullptr;
}

int agb_has_tensor(agb_ctx* ctx, const char* name) {
  if (!ctx || !name) return 0;
  return ctx->loaders.count(name) ? 1 : 0;
}

int agb_set_tap(agb_ctx* ctx, const char* name, void* dst, int64_t capacity_bytes) {
  if (!ctx || !name) return 2;
  if (dst == nullptr) ctx->taps.erase(name); else ctx->taps[name] = Tap{dst, capacity_bytes};
  return 0;
}

int64_t agb_launch_count(void) { return g_launches.load(); }

int agb_clip_forward(agb_ctx* ctx, const void* raw_frames, int in_dtype, int64_t n, float* out, void* stream) {
  if (!ctx) { set_error("agb_clip_forward: null context"); return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) return 0;
  if (n < 0 || !raw_frames || !out) { set_error("agb_clip_forward: bad arguments"); return fail(ctx, 2); }
  if (int r = check_family_loaded(ctx, "clip")) return fail(ctx, r);
  const int64_t cap = ctx->cfg.max_frames_per_chunk;
  if (int r = ensure_vit_ws(ctx, n < cap ? n : cap)) return fail(ctx, r);
  const size_t esz = in_dtype == AGB_DTYPE_U8 ? 1 : 4;
  for (int64_t off = 0; off < n; off += cap) {
    const int64_t F = (n - off) < cap ? (n - off) : cap;
    const char* fp = static_cast<const char*>(raw_frames) + (size_t)off * 3 * 224 * 224 * esz;
    if (int r = clip_chunk(ctx, fp, in_dtype, F, out + off * 768, s)) return fail(ctx, r);
  }
  return 0;
}

int agb_hubert_forward(agb_ctx* ctx, const float* wav, int64_t n, int64_t seg_per_norm_group, float* out, void* stream) {
  if (!ctx) { set_error("agb_hubert_forward: null context"); return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) return 0;
  if (n < 0 || !wav || !out) { set_error("agb_hubert_forward: bad arguments"); return fail(ctx, 2); }
  if (int r = check_family_loaded(ctx, "hubert")) return fail(ctx, r);
  if (!ctx->h_pos_packed) { set_error("hubert: positional conv weight-norm tensors incomplete"); return fail(ctx, 4); }
  int64_t spg = seg_per_norm_group <= 0 ? n : seg_per_norm_group;
  if (n % spg != 0) { set_error("agb_hubert_forward: n must be a multiple of seg_per_norm_group"); return fail(ctx, 2); }
  const int64_t groups = n / spg;
  if (int r = ensure_stats(ctx, groups)) return fail(ctx, r);
  if (int r = wave_stats(wav, n, 32000, spg, ctx->h_stats, ctx->cfg.hubert_do_normalize, ctx->h_stat_acc, s)) return fail(ctx, r);
  const int64_t cap = ctx->cfg.max_segments_per_chunk;
  if (int r = ensure_hub_ws(ctx, n < cap ? n : cap)) return fail(ctx, r);
  for (int64_t off = 0; off < n; off += cap) {
    const int64_t G = (n - off) < cap ? (n - off) : cap;
    if (int r = hubert_chunk(ctx, wav + off * 32000, off, G, spg, out + off * 1024, s)) return fail(ctx, r);
  }
  return 0;
}

int agb_video_fuse(agb_ctx* ctx, const float* H, int B, int T, int fusion_type, int bypass_t1, float* llm, void* stream) {
  if (!ctx) { set_error("agb_video_fuse: null context"); return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  if (int r = fusion_type == AGB_FUSE_ATTENTION ? check_heads_loaded(ctx, {"video_attention_mlp", "affectgpt_proj"})
                                                : check_heads_loaded(ctx, {"affectgpt_proj"})) return fail(ctx, r);
  if (int r = ensure_fuse_ws(ctx, B)) return fail(ctx, r);
  int r;
  if (fusion_type == AGB_FUSE_ATTENTION)
    r = attn_pool_nosoftmax(H, B, T, 768, ctx->f["video_attention_mlp.weight"], ctx->f["video_attention_mlp.bias"], bypass_t1,
                            ctx->f_a, s);
  else
    r = mean_over_t(H, B, T, 768, ctx->f_a, s);
  if (r) return fail(ctx, r);
  r = linear_f32(ctx->f_a, 768, ctx->f["affectgpt_proj.weight"], ctx->f["affectgpt_proj.bias"], llm, ctx->cfg.llm_dim, B,
                 ctx->cfg.llm_dim, 768, s);
  return r ? fail(ctx, r) : 0;
}

int agb_audio_fuse(agb_ctx* ctx, const float* H, int B, int T, int fusion_type, float* llm, void* stream) {
  if (!ctx) { set_error("agb_audio_fuse: null context"); return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  if (int r = fusion_type == AGB_FUSE_ATTENTION ? check_heads_loaded(ctx, {"audio_attention_mlp", "audio_llama_proj"})
                                                : check_heads_loaded(ctx, {"audio_llama_proj"})) return fail(ctx, r);
  if (int r = ensure_fuse_ws(ctx, B)) return fail(ctx, r);
  int r;
  if (fusion_type == AGB_FUSE_ATTENTION)
    r = attn_pool_nosoftmax(H, B, T, 1024, ctx->f["audio_attention_mlp.weight"], ctx->f["audio_attention_mlp.bias"], 0,
                            ctx->f_a, s);
  else
    r = mean_over_t(H, B, T, 1024, ctx->f_a, s);
  if (r) return fail(ctx, r);
  r = linear_f32(ctx->f_a, 1024, ctx->f["audio_llama_proj.weight"], ctx->f["audio_llama_proj.bias"], llm, ctx->cfg.llm_dim,
                 B, ctx->cfg.llm_dim, 1024, s);
  return r ? fail(ctx, r) : 0;
}

int agb_multi_fuse(agb_ctx* ctx, const float* Hv, int Tv, const float* Ha, int Ta, int B, float* multi_hidden, float* llm,
                   void* stream) {
  if (!ctx) { set_error("agb_multi_fuse: null context"); return 2; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  if (int r = check_heads_loaded(ctx, {"multi_video_embs", "multi_audio_embs", "attention_mlp", "fc_att", "multi_llama_proj"}))
    return fail(ctx, r);
  if (int r = ensure_fuse_ws(ctx, B)) return fail(ctx, r);
  auto& f = ctx->f;
  int r;
  // v = multi_video_embs(mean_t Hv), a = multi_audio_embs(mean_t Ha)          (affectgpt.py:890-893)
  if ((r = mean_over_t(Hv, B, Tv, 768, ctx->f_a, s))) return fail(ctx, r);
  if ((r = mean_over_t(Ha, B, Ta, 1024, ctx->f_b, s))) return fail(ctx, r);
  float* v = ctx->f_c;            // [B,1024] (first half of the concat buffer, row stride 2048)
  float* a = ctx->f_c + 1024;     // second half
  if ((r = linear_f32(ctx->f_a, 768, f["multi_video_embs.weight"], f["multi_video_embs.bias"], v, 2048, B, 1024, 768, s))) return fail(ctx, r);
  if ((r = linear_f32(ctx->f_b, 1024, f["multi_audio_embs.weight"], f["multi_audio_embs.bias"], a, 2048, B, 1024, 1024, s))) return fail(ctx, r);
  // att = fc_att(attention_mlp([v | a]))                                      (:895-897)
  if ((r = linear_f32(ctx->f_c, 2048, f["attention_mlp.weight"], f["attention_mlp.bias"], ctx->f_d, 1024, B, 1024, 2048, s))) return fail(ctx, r);
  if ((r = linear_f32(ctx->f_d, 1024, f["fc_att.weight"], f["fc_att.bias"], ctx->f_e, 2, B, 2, 1024, s))) return fail(ctx, r);
  // multi = [v, a] . att                                                      (:900-902)
  // (v and a live interleaved in the concat buffer with row stride 2048: copy out densely first)
  AGB_CHECK_CUDA(cudaMemcpy2DAsync(ctx->f_a, 1024 * 4, v, 2048 * 4, 1024 * 4, B, cudaMemcpyDeviceToDevice, s));
  AGB_CHECK_CUDA(cudaMemcpy2DAsync(ctx->f_b, 1024 * 4, a, 2048 * 4, 1024 * 4, B, cudaMemcpyDeviceToDevice, s));
  if ((r = weighted_pair_sum(ctx->f_a, ctx->f_b, ctx->f_e, B, 1024, multi_hidden, s))) return fail(ctx, r);
  r = linear_f32(multi_hidden, 1024, f["multi_llama_proj.weight"], f["multi_llama_proj.bias"], llm, ctx->cfg.llm_dim, B,
                 ctx->cfg.llm_dim, 1024, s);
  return r ? fail(ctx, r) : 0;
}

namespace {
int ensure_qf_ws(agb_ctx* c, int64_t q_rows, int64_t kv_rows, int64_t enc_elems) {
  QFormerWs& w = c->qf_ws;
  if (q_rows > w.cap_q) {
    w.cap_q = 0;
    dev_release(c, w.x); dev_release(c, w.t0); dev_release(c, w.t1); dev_release(c, w.t2); dev_release(c, w.ff);
    if (int r = dev_alloc(c, &w.x, (size_t)q_rows * 768)) return r;
    if (int r = dev_alloc(c, &w.t0, (size_t)q_rows * 768)) return r;
    if (int r = dev_alloc(c, &w.t1, (size_t)q_rows * 768)) return r;
    if (int r = dev_alloc(c, &w.t2, (size_t)q_rows * 768)) return r;
    if (int r = dev_alloc(c, &w.ff, (size_t)q_rows * 3072)) return r;
    w.cap_q = q_rows;
  }
  if (kv_rows > w.cap_kv) {
    w.cap_kv = 0;
    dev_release(c, w.kx); dev_release(c, w.vx);
    if (int r = dev_alloc(c, &w.kx, (size_t)kv_rows * 768)) return r;
    if (int r = dev_alloc(c, &w.vx, (size_t)kv_rows * 768)) return r;
    w.cap_kv = kv_rows;
  }
  if (enc_elems > w.cap_enc) {
    w.cap_enc = 0;
    dev_release(c, w.enc); dev_release(c, w.ev); dev_release(c, w.ea);
    if (int r = dev_alloc(c, &w.enc, (size_t)enc_elems)) return r;
    if (int r = dev_alloc(c, &w.ev, (size_t)enc_elems)) return r;
    if (int r = dev_alloc(c, &w.ea, (size_t)enc_elems)) return r;
    w.cap_enc = enc_elems;
  }
  return 0;
}
}  // namespace

int agb_qformer_fuse(agb_ctx* ctx, int which, const float* Hv, int Tv, const float* Ha, int Ta, int B, float* hidden,
                     float* llm, void* stream) {
  if (!ctx) { set_error("agb_qformer_fuse: null context"); return 2; }
  if (which < 0 || which > 2) { set_error("agb_qformer_fuse: which must be 0 (video), 1 (audio) or 2 (multi)"); return fail(ctx, 2); }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B == 0) return 0;
  const QFormerW& W = ctx->qf[which];
  static const char* fam[3] = {"qf_video", "qf_audio", "qf_multi"};
  if (W.nq <= 0) {
    set_error(std::string("agb_qformer_fuse: the ") + fam[which] + " Q-Former was not configured (agb_cfg.qf_*_queries == 0)");
    return fail(ctx, 2);
  }
  if (int r = check_family_loaded(ctx, fam[which])) return fail(ctx, r);
  const bool need_v = which != AGB_QF_AUDIO, need_a = which != AGB_QF_VIDEO;
  if ((need_v && (!Hv || Tv <= 0)) || (need_a && (!Ha || Ta <= 0)) || !llm) {
    set_error("agb_qformer_fuse: missing input / output for this variant");
    return fail(ctx, 2);
  }
  const int T = (need_v ? Tv : 0) + (need_a ? Ta : 0);
  if (T > W.pos_rows) {
    set_error("agb_qformer_fuse: " + std::to_string(T) + " encoder tokens exceed the position embedding table (" +
              std::to_string(W.pos_rows) + " rows)");
    return fail(ctx, 2);
  }
  const int64_t q_rows = (int64_t)B * W.nq, kv_rows = (int64_t)B * (T > W.nq ? T : W.nq);
  if (int r = ensure_qf_ws(ctx, q_rows, kv_rows, (int64_t)B * T * 1024)) return fail(ctx, r);
  QFormerWs& ws = ctx->qf_ws;
  int r;
  if (which == AGB_QF_VIDEO) {                       // frames + video_frame_position_embedding (:439-451)
    if ((r = add_position_f32(Hv, W.pos, B, Tv, 768, 0, Tv, ws.enc, s))) return fail(ctx, r);
  } else if (which == AGB_QF_AUDIO) {                // segments + audio_position_embedding (:645-648)
    if ((r = add_position_f32(Ha, W.pos, B, Ta, 1024, 0, Ta, ws.enc, s))) return fail(ctx, r);
  } else {                                           // both streams to 1024, concatenated along t, + position (:855-865)
    if ((r = linear_f32(Hv, 768, ctx->qf_mv_w, ctx->qf_mv_b, ws.ev, 1024, B * Tv, 1024, 768, s))) return fail(ctx, r);
    if ((r = linear_f32(Ha, 1024, ctx->qf_ma_w, ctx->qf_ma_b, ws.ea, 1024, B * Ta, 1024, 1024, s))) return fail(ctx, r);
    if ((r = add_position_f32(ws.ev, W.pos, B, Tv, 1024, 0, T, ws.enc, s))) return fail(ctx, r);
    if ((r = add_position_f32(ws.ea, W.pos, B, Ta, 1024, Tv, T, ws.enc, s))) return fail(ctx, r);
  }
  if ((r = qformer_forward_f32(W, ws.enc, B, T, ws, s))) return fail(ctx, r);
  if (hidden) AGB_CHECK_CUDA(cudaMemcpyAsync(hidden, ws.x, (size_t)q_rows * 768 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  r = linear_f32(ws.x, 768, W.proj_w, W.proj_b, llm, ctx->cfg.llm_dim, (int)q_rows, ctx->cfg.llm_dim, 768, s);
  return r ? fail(ctx, r) : 0;
}

int agb_splice_tokens(void* embeds, int dtype, const int64_t* input_ids, int B, int L, int D, int n_mod,
                      const int64_t* token_ids, const int32_t* n_query, const float* const* mod_embeds, int32_t* status,
                      void* stream) {
  return splice_tokens(embeds, dtype, reinterpret_cast<const long long*>(input_ids), B, L, D, n_mod,
                       reinterpret_cast<const long long*>(token_ids), n_query, mod_embeds, status,
                       static_cast<cudaStream_t>(stream));
}

int agb_stage_faces(const void* faces_u8, int64_t n_src, int h, int w, const int32_t* idx, int T, void* out_u8, void* stream) {
  return stage_faces_u8(static_cast<const uint8_t*>(faces_u8), n_src, h, w, idx, T, static_cast<uint8_t*>(out_u8),
                        static_cast<cudaStream_t>(stream));
}

int agb_stage_audio(const float* wav, int ch, int64_t n, const int64_t* starts, int clips, int clip_len, float* out,
                    void* stream) {
  return stage_audio_f32(wav, ch, n, reinterpret_cast<const long long*>(starts), clips, clip_len, out,
                         static_cast<cudaStream_t>(stream));
}

int agb_resample_audio(const float* wav, int ch, int64_t n, const float* bank, int orig, int nw, int width, int taps,
                       float* out, int64_t n_out, void* stream) {
  return resample_f32(wav, ch, n, bank, orig, nw, width, taps, out, n_out, static_cast<cudaStream_t>(stream));
}

// ---- kernel-level entry points ---------------------------------------------------------------
int agb_op_gemm(int a_mode, int act, const void* A, int64_t lda, const void* W, int M, int N, int K, const float* bias,
                const float* res, int64_t ld_res, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, int nseq,
                int seq_stride, int seq_valid, void* stream) {
  GemmArgs g;
  g.mode = a_mode; g.act = act; g.A = static_cast<const bf16*>(A); g.lda = lda; g.W = static_cast<const bf16*>(W);
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.res = res; g.ld_res = ld_res; g.out_f32 = out_f32; g.ld_f32 = ld_f32;
  g.out_bf16 = static_cast<bf16*>(out_bf16); g.ld_bf16 = ld_bf16; g.nseq = nseq; g.seq_stride = seq_stride; g.seq_valid = seq_valid;
  return gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

int agb_op_layernorm(const float* x, int64_t ldx, int64_t rows, int D, const float* gamma, const float* beta, float eps,
                     void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32, void* stream) {
  return layernorm_f32_in(x, ldx, rows, D, gamma, beta, eps, static_cast<bf16*>(out_bf16), ld_bf16, out_f32, ld_f32,
                          static_cast<cudaStream_t>(stream));
}

int agb_op_layernorm512_bf16(const void* x, int64_t rows, const float* gamma, const float* beta, float eps, int gelu,
                             void* out, void* stream) {
  return layernorm512_bf16_gelu(const_cast<bf16*>(static_cast<const bf16*>(x)), rows, gamma, beta, eps, gelu != 0,
                                static_cast<bf16*>(out), static_cast<cudaStream_t>(stream));
}

int agb_op_attention(const void* qkv, void* out, int nseq, int seq_stride, int seq_valid, float scale, void* stream) {
  return attention_mha16x64(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), nseq, seq_stride, seq_valid, scale,
                            static_cast<cudaStream_t>(stream));
}

int agb_op_patchify(const void* frames, int in_dtype, int64_t n, void* a_patch, void* stream) {
  return vit_patchify(frames, in_dtype, n, static_cast<bf16*>(a_patch), static_cast<cudaStream_t>(stream));
}

int agb_op_conv0(const float* wav, int64_t nseg, int64_t seg_per_norm_group, int do_normalize, const float* w,
                 const float* bias, const float* gamma, const float* beta, int norm_mode, void* out, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t spg = seg_per_norm_group <= 0 ? nseg : seg_per_norm_group;
  AGB_REQUIRE(nseg > 0 && nseg % spg == 0, "conv0: bad normalisation grouping");
  const int64_t groups = nseg / spg;
  float* stats = nullptr;
  double* acc = nullptr;
  double* gn = nullptr;
  AGB_CHECK_CUDA(cudaMalloc(&stats, sizeof(float) * 2 * groups));
  AGB_CHECK_CUDA(cudaMalloc(&acc, sizeof(double) * 2 * groups));
  AGB_CHECK_CUDA(cudaMalloc(&gn, norm_mode == 1 ? sizeof(double) * 1024 * nseg : sizeof(double) * 4096));
  int r = wave_stats(wav, nseg, 32000, spg, stats, do_normalize, acc, s);
  if (!r) r = hubert_conv0(wav, stats, spg, 0, nseg, w, bias, gamma, beta, norm_mode, reinterpret_cast<float*>(gn),
                           static_cast<bf16*>(out), s);
  cudaStreamSynchronize(s);
  cudaFree(stats);
  cudaFree(acc);
  if (gn) cudaFree(gn);
  return r;
}

int agb_op_seq_mean_pool(const float* x, int nseq, int seq_stride, int seq_valid, const float* gamma, const float* beta,
                         float eps, float scale, int accumulate, float* pooled, void* stream) {
  return seq_mean_pool(x, nseq, seq_stride, seq_valid, gamma, beta, eps, scale, accumulate, pooled,
                       static_cast<cudaStream_t>(stream));
}

int agb_op_linear_f32(const float* in, int64_t ld_in, const float* W, const float* bias, float* out, int64_t ld_out, int M,
                      int N, int K, void* stream) {
  return linear_f32(in, ld_in, W, bias, out, ld_out, M, N, K, static_cast<cudaStream_t>(stream));
}

int agb_op_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  return convert_f32_to_bf16(src, static_cast<bf16*>(dst), n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
