"""Pin the token layout against the REFERENCE's own loop.

    python tests/golden/make_reference_splice.py          # build container only (/root/reference must exist)

What runs is the unmodified ``AffectGPT.forward`` of /root/reference/AffectGPT/my_affectgpt/models/affectgpt.py
(:924-1025; the splice loop is :956-1009) on an ``AffectGPT`` instance built without ``__init__`` (ref_loader):

* ``llama_model`` is a stub whose ``model.model.embed_tokens`` is a seeded ``nn.Embedding`` and whose ``__call__``
  records the ``inputs_embeds`` the reference hands to the LLM - the call is cut there.
* "stub" cases replace ``encode_{video,audio,multi}_merge`` on the instance by functions that return seeded token
  tensors (width 16), so that only the reference's layout rule is under test; the "heads" case runs the reference's
  real ``encode_*_merge`` (attention heads + projectors to 3584, pre-extracted inputs) in front of the same loop.
* error cases record the exception type and message the reference raises (both ``ValueError``s and the
  ``AssertionError``).

Output: ``reference_splice.npz``.  ``oracle.sampling.splice`` (CPU suite) and ``agb_splice_tokens`` (GPU suite) are
held to it bit for bit.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as L  # noqa: E402
from oracle import reference_path as R  # noqa: E402

TOK = {"frame": 1001, "face": 1002, "audio": 1003, "multi": 1004, "image": 1005}
VOCAB = 1100


class _Captured(Exception):
    pass


class _StubLLM(nn.Module):
    """llama_model stand-in: embed_tokens under .model.model (affectgpt.py:966), forward records inputs_embeds."""

    def __init__(self, dim: int, seed: int):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        emb = nn.Embedding(VOCAB, dim)
        emb.weight.data.copy_(torch.randn(VOCAB, dim, generator=g))
        self.model = types.SimpleNamespace(model=types.SimpleNamespace(embed_tokens=emb))
        self.seen = None

    def forward(self, inputs_embeds=None, attention_mask=None, return_dict=True, labels=None):
        self.seen = inputs_embeds.detach().clone()
        return types.SimpleNamespace(loss=torch.zeros(()))


def _model(A, nq_video, nq_audio, nq_multi, dim, seed):
    m = L.bare_affectgpt(A, num_video_query_token=nq_video, num_audio_query_token=nq_audio,
                         num_multi_query_token=nq_multi, num_image_query_token=1,
                         video_fusion_type="attention", audio_fusion_type="attention", multi_fusion_type="attention")
    m.llama_model = _StubLLM(dim, seed)
    m.FRAME_PATCH_TOKEN_ID, m.FACE_PATCH_TOKEN_ID = TOK["frame"], TOK["face"]
    m.AUDIO_PATCH_TOKEN_ID, m.MULTI_PATCH_TOKEN_ID, m.IMAGE_PATCH_TOKEN_ID = TOK["audio"], TOK["multi"], TOK["image"]
    return m


def _prompt(b, Lp, layout, nq, gen, skip=()):
    """ids in [5, 900) with one consecutive run per modality, prompt order ``layout``, shifted per sample."""
    ids = torch.randint(5, 900, (b, Lp), generator=gen)
    for i in range(b):
        if i in skip:
            continue
        pos = 2 + i
        for name in layout:
            ids[i, pos:pos + nq[name]] = TOK[name]
            pos += nq[name] + 1 + i
    return ids


def _run(m, samples):
    with torch.no_grad():
        m.forward(samples)
    return m.llama_model.seen


def _samples(ids, face_or_frame, **extra):
    s = {"face_or_frame": face_or_frame, "input_ids": ids, "labels": ids.clone(), "attention_masks": torch.ones_like(ids)}
    s.update(extra)
    return s


def stub_case(A, name, nq, layout, b, Lp, dim, seed, skip=(), mutate=None, drop=()):
    """One run of the reference loop with seeded tokens per modality.  Returns the fixture entries."""
    g = torch.Generator().manual_seed(seed)
    m = _model(A, nq["frame"], nq["audio"], nq["multi"], dim, seed + 1)
    ids = _prompt(b, Lp, layout, nq, g, skip)
    if mutate is not None:
        mutate(ids)
    toks = {k: torch.randn(b, nq[k], dim, generator=g) for k in ("frame", "face", "audio", "multi")}
    m.encode_video_merge = lambda v, raw, is_preextracted=False: (None, toks[v])       # samples['frames'] = 'frame' / 'face'
    m.encode_audio_merge = lambda a, raw, is_preextracted=False: (None, toks["audio"])
    m.encode_multi_merge = lambda vh, ah: (None, toks["multi"])
    extra = {}
    if "frame" not in drop:
        extra.update(frames="frame", raw_frames=None)
    if "face" not in drop:
        extra.update(faces="face", raw_faces=None)
    if "audio" not in drop:
        extra.update(audios="audio", raw_audios=None)
    out = {f"{name}.ids": ids.numpy(), f"{name}.embed_seed": np.int64(seed + 1), f"{name}.dim": np.int64(dim),
           f"{name}.nq": np.array([nq["frame"], nq["face"], nq["audio"], nq["multi"]], dtype=np.int64),
           f"{name}.present": np.array([k not in drop for k in ("frame", "face", "audio", "multi")])}
    for k, v in toks.items():
        out[f"{name}.tok_{k}"] = v.numpy()
    try:
        seen = _run(m, _samples(ids, "multiface_audio_face_frame_text", **extra))
        out[f"{name}.inputs_embeds"] = seen.numpy()
        out[f"{name}.error"] = np.array("")
    except (ValueError, AssertionError) as e:                   # what the reference raises (:984-1002)
        out[f"{name}.error"] = np.array(json.dumps({"type": type(e).__name__, "message": str(e)}))
    return out


def heads_case(A, name="heads"):
    """The reference's real encode_*_merge (attention heads, projectors to 3584, pre-extracted features) in front of the
    same loop: fused tokens -> inputs_embeds, end to end."""
    st = R.build_fusion_state(seed=1234)
    g = torch.Generator().manual_seed(11)
    b, Lp, D = 2, 14, 3584
    m = _model(A, 1, 1, 1, D, 99)
    from make_reference_golden import _linear
    for n in ("video_attention_mlp", "affectgpt_proj", "audio_attention_mlp", "audio_llama_proj", "multi_video_embs",
              "multi_audio_embs", "attention_mlp", "fc_att", "multi_llama_proj"):
        setattr(m, n, _linear(st, n))
    m.eval()
    nq = {"frame": 1, "face": 1, "audio": 1, "multi": 1}
    ids = _prompt(b, Lp, ["multi", "audio", "face", "frame"], nq, g)                  # base_dataset.py:882-889 order
    frame_h, face_h = torch.randn(b, 8, 768, generator=g), torch.randn(b, 8, 768, generator=g)
    audio_h = torch.randn(b, 8, 1024, generator=g)
    seen = _run(m, _samples(ids, "multiface_audio_face_frame_text", frames=frame_h, raw_frames=None, frame_preextracted=True,
                            faces=face_h, raw_faces=None, face_preextracted=True, audios=audio_h, raw_audios=None,
                            audio_preextracted=True))
    return {f"{name}.ids": ids.numpy(), f"{name}.embed_seed": np.int64(99), f"{name}.frame_h": frame_h.numpy(),
            f"{name}.face_h": face_h.numpy(), f"{name}.audio_h": audio_h.numpy(), f"{name}.inputs_embeds": seen.numpy()}


def embedding_table(dim: int, seed: int) -> torch.Tensor:
    """The stub LLM's embed_tokens weight, rebuilt from its seed (tests use this instead of storing the table)."""
    return torch.randn(VOCAB, dim, generator=torch.Generator().manual_seed(seed))


def main():
    A = L.load_affectgpt()
    fix = {}
    order = ["multi", "audio", "face", "frame"]
    one = {"frame": 1, "face": 1, "audio": 1, "multi": 1}
    many = {"frame": 4, "face": 4, "audio": 2, "multi": 3}
    fix.update(stub_case(A, "nq1", one, order, 3, 24, 16, 100))
    fix.update(stub_case(A, "nq_many", many, order, 3, 48, 16, 200))
    fix.update(stub_case(A, "other_order", many, ["frame", "multi", "face", "audio"], 2, 40, 16, 300))
    fix.update(stub_case(A, "sample_without_placeholder", one, order, 3, 24, 16, 400, skip=(1,)))
    fix.update(stub_case(A, "frame_audio_only", many, ["audio", "frame"], 2, 32, 16, 500))

    def extra_frame(ids):                                       # count != n_query
        ids[1, 30] = TOK["frame"]
    fix.update(stub_case(A, "err_count", many, order, 2, 48, 16, 600, mutate=extra_frame))

    def split_run(ids):                                         # right count, not consecutive
        pos = int((ids[0] == TOK["audio"]).nonzero()[0])
        ids[0, pos + 1] = 7
        ids[0, 40] = TOK["audio"]
    fix.update(stub_case(A, "err_consecutive", many, order, 2, 48, 16, 700, mutate=split_run))
    fix.update(stub_case(A, "err_missing_face", one, order, 2, 24, 16, 800, drop=("face",)))
    fix.update(heads_case(A))
    np.savez_compressed(os.path.join(HERE, "reference_splice.npz"), **fix)
    for k in sorted(fix):
        if k.endswith(".error"):
            print(k, str(fix[k]))
    print("wrote reference_splice.npz with", len(fix), "entries")


if __name__ == "__main__":
    main()
