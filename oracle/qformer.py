"""fp32 CPU restatement of the reference's Q-Former fusion variants (TEST INFRASTRUCTURE ONLY).

Follows, line by line, the eval-mode forward of
  A/my_affectgpt/models/Qformer.py   BertEmbeddings.forward :81-109 (query_embeds only -> LayerNorm, eps 1e-12)
                                     BertSelfAttention.forward :173-276 (12 heads x 64, 1/sqrt(64), additive mask = 0)
                                     BertSelfOutput :279-290, BertAttention :293-347
                                     BertLayer.forward :411-485 (self-attn -> cross-attn every layer -> query FFN)
                                     BertIntermediate/BertOutput :350-376 (gelu, LN(x + residual))
  A/my_affectgpt/models/affectgpt.py init_video_Qformer :23-37 (2 layers, cross_attention_freq=1, BertConfig defaults)
                                     encode_video_qformer :420-468, encode_audio_qformer :628-664,
                                     encode_multi_qformer :848-881
Pinned against the reference's own code run in the build container: tests/golden/make_reference_golden.py
(fixtures tests/golden/reference_fusion.npz, checked in tests/test_reference_golden.py).

``qformer_state`` builds deterministic weights under the reference's state-dict key names, so the reference module,
this restatement and the CUDA path can all be fed identical parameters without shipping an 80 MB fixture.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

HIDDEN = 768
HEADS = 12
FFN = 3072
LN_EPS = 1e-12


def qformer_state(seed: int, enc_width: int, n_query: int, n_layers: int = 2, prefix: str = "",
                  pos_rows: int = 0, llm_dim: int = 3584, proj_name: str = "") -> Dict[str, torch.Tensor]:
    """Seeded weights, reference key names.  ``prefix`` e.g. 'multi_Qformer.'; also returns ``<x>_query_tokens`` etc.
    when the caller asks for them through pos_rows / proj_name (those keys are added by the callers below)."""
    g = torch.Generator().manual_seed(seed)

    def w(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def ln(n):
        return 1.0 + 0.1 * torch.randn(n, generator=g), 0.05 * torch.randn(n, generator=g)

    st: Dict[str, torch.Tensor] = {}
    st[prefix + "bert.embeddings.LayerNorm.weight"], st[prefix + "bert.embeddings.LayerNorm.bias"] = ln(HIDDEN)
    for i in range(n_layers):
        L = f"{prefix}bert.encoder.layer.{i}."
        for att, kv_in in (("attention", HIDDEN), ("crossattention", enc_width)):
            st[L + att + ".self.query.weight"] = w(HIDDEN, HIDDEN)
            st[L + att + ".self.query.bias"] = w(HIDDEN)
            st[L + att + ".self.key.weight"] = w(HIDDEN, kv_in)
            st[L + att + ".self.key.bias"] = w(HIDDEN)
            st[L + att + ".self.value.weight"] = w(HIDDEN, kv_in)
            st[L + att + ".self.value.bias"] = w(HIDDEN)
            st[L + att + ".output.dense.weight"] = w(HIDDEN, HIDDEN)
            st[L + att + ".output.dense.bias"] = w(HIDDEN)
            st[L + att + ".output.LayerNorm.weight"], st[L + att + ".output.LayerNorm.bias"] = ln(HIDDEN)
        st[L + "intermediate_query.dense.weight"] = w(FFN, HIDDEN)
        st[L + "intermediate_query.dense.bias"] = w(FFN)
        st[L + "output_query.dense.weight"] = w(HIDDEN, FFN)
        st[L + "output_query.dense.bias"] = w(HIDDEN)
        st[L + "output_query.LayerNorm.weight"], st[L + "output_query.LayerNorm.bias"] = ln(HIDDEN)
    return st


def fusion_qformer_state(seed: int = 4321, n_query: int = 4, llm_dim: int = 3584) -> Dict[str, torch.Tensor]:
    """Every parameter the three *_fusion_type == 'qformer' branches read (affectgpt.py:150-349 shapes)."""
    g = torch.Generator().manual_seed(seed)

    def w(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    st: Dict[str, torch.Tensor] = {}
    # video (:157-166, :195): position embedding [32, 768], Q-Former over 768-wide frames, proj 768 -> llm
    st["video_frame_position_embedding.weight"] = w(32, 768, std=1.0)
    st["video_query_tokens"] = w(1, n_query, HIDDEN)
    st.update(qformer_state(seed + 1, 768, n_query, prefix="video_Qformer."))
    st["affectgpt_proj.weight"], st["affectgpt_proj.bias"] = w(llm_dim, HIDDEN), w(llm_dim)
    # audio (:223-231, :255): position embedding [8, 1024], Q-Former over 1024-wide segments, proj 768 -> llm
    st["audio_position_embedding.weight"] = w(8, 1024, std=1.0)
    st["audio_query_tokens"] = w(1, n_query, HIDDEN)
    st.update(qformer_state(seed + 2, 1024, n_query, prefix="audio_Qformer."))
    st["audio_llama_proj.weight"], st["audio_llama_proj.bias"] = w(llm_dim, HIDDEN), w(llm_dim)
    # multi (:312-327, :349): both streams to 1024, position embedding [264, 1024], proj 768 -> llm
    st["multi_video_embs.weight"], st["multi_video_embs.bias"] = w(1024, 768), w(1024)
    st["multi_audio_embs.weight"], st["multi_audio_embs.bias"] = w(1024, 1024), w(1024)
    st["multi_position_embedding.weight"] = w(264, 1024, std=1.0)
    st["multi_query_tokens"] = w(1, n_query, HIDDEN)
    st.update(qformer_state(seed + 3, 1024, n_query, prefix="multi_Qformer."))
    st["multi_llama_proj.weight"], st["multi_llama_proj.bias"] = w(llm_dim, HIDDEN), w(llm_dim)
    return st


def _attention(st, pre, x_q, x_kv):
    """BertSelfAttention + BertSelfOutput (Qformer.py:173-290): returns LN(dense(ctx) + x_q)."""
    b, nq, _ = x_q.shape
    d = HIDDEN // HEADS

    def heads(t):
        return t.view(b, -1, HEADS, d).permute(0, 2, 1, 3)

    q = heads(F.linear(x_q, st[pre + ".self.query.weight"], st[pre + ".self.query.bias"]))
    k = heads(F.linear(x_kv, st[pre + ".self.key.weight"], st[pre + ".self.key.bias"]))
    v = heads(F.linear(x_kv, st[pre + ".self.value.weight"], st[pre + ".self.value.bias"]))
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)          # all-ones masks add zeros (:714-803)
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).reshape(b, nq, HIDDEN)
    out = F.linear(ctx, st[pre + ".output.dense.weight"], st[pre + ".output.dense.bias"])
    return F.layer_norm(out + x_q, (HIDDEN,), st[pre + ".output.LayerNorm.weight"], st[pre + ".output.LayerNorm.bias"],
                        LN_EPS)


def qformer_forward(st: Dict[str, torch.Tensor], prefix: str, query_tokens: torch.Tensor, enc: torch.Tensor,
                    n_layers: int = 2) -> torch.Tensor:
    """``<prefix>bert(query_embeds=query_tokens, encoder_hidden_states=enc)`` -> last_hidden_state [b, n_q, 768]."""
    b = enc.shape[0]
    x = query_tokens.expand(b, -1, -1)
    x = F.layer_norm(x, (HIDDEN,), st[prefix + "bert.embeddings.LayerNorm.weight"],
                     st[prefix + "bert.embeddings.LayerNorm.bias"], LN_EPS)
    for i in range(n_layers):
        L = f"{prefix}bert.encoder.layer.{i}."
        x = _attention(st, L + "attention", x, x)
        x = _attention(st, L + "crossattention", x, enc)
        h = F.gelu(F.linear(x, st[L + "intermediate_query.dense.weight"], st[L + "intermediate_query.dense.bias"]))
        h = F.linear(h, st[L + "output_query.dense.weight"], st[L + "output_query.dense.bias"])
        x = F.layer_norm(h + x, (HIDDEN,), st[L + "output_query.LayerNorm.weight"], st[L + "output_query.LayerNorm.bias"],
                         LN_EPS)
    return x


def encode_video_qformer(st, hidden: torch.Tensor):
    """affectgpt.py:420-468 for 3-D input [b, t, 768] -> (store_hidden [b,t,768], llm [b, n_q, llm])."""
    b, t, _ = hidden.shape
    x = hidden + st["video_frame_position_embedding.weight"][:t].unsqueeze(0)
    out = qformer_forward(st, "video_Qformer.", st["video_query_tokens"], x)
    return hidden, F.linear(out, st["affectgpt_proj.weight"], st["affectgpt_proj.bias"])


def encode_audio_qformer(st, hidden: torch.Tensor):
    """affectgpt.py:628-664: [b, t, 1024] -> (store_hidden, llm [b, n_q, llm])."""
    b, t, _ = hidden.shape
    x = hidden + st["audio_position_embedding.weight"][:t].unsqueeze(0)
    out = qformer_forward(st, "audio_Qformer.", st["audio_query_tokens"], x)
    return hidden, F.linear(out, st["audio_llama_proj.weight"], st["audio_llama_proj.bias"])


def encode_multi_qformer(st, video_hidden: torch.Tensor, audio_hidden: torch.Tensor):
    """affectgpt.py:848-881: ([b,tv,768],[b,ta,1024]) -> (multi_hidden [b,n_q,768], llm [b,n_q,llm])."""
    v = F.linear(video_hidden, st["multi_video_embs.weight"], st["multi_video_embs.bias"])
    a = F.linear(audio_hidden, st["multi_audio_embs.weight"], st["multi_audio_embs.bias"])
    x = torch.cat([v, a], dim=1)
    x = x + st["multi_position_embedding.weight"][: x.shape[1]].unsqueeze(0)
    out = qformer_forward(st, "multi_Qformer.", st["multi_query_tokens"], x)
    return out, F.linear(out, st["multi_llama_proj.weight"], st["multi_llama_proj.bias"])
