"""Generate fixtures by running the REFERENCE's own code (only possible in the build container).

    python tests/golden/make_reference_golden.py

What runs is /root/reference/AffectGPT's unmodified source, loaded by ``ref_loader`` with stub modules for the
third-party imports that are not installed (decord, omegaconf, pytorchvideo, peft ...):

* ``processors/video_processor.py``: ``load_video`` (uniform / headtail / emotion_peak), ``load_face`` (index rule and
  the cv2 bilinear resize), ``_calculate_smart_frame_indices``, ``_load_au_info``
* ``models/ImageBind/data.py``: ``load_audio`` (mono mix-down, 2 s zero padding, clip slicing; the pytorchvideo clip
  sampler itself is restated from its published source - the one piece that is NOT the reference's own code)
* ``models/affectgpt.py``: ``encode_{video,audio}_{attention,mean,qformer}``, ``encode_{video,audio}_merge``
  (pre-extracted twins), ``encode_multi_{attention,qformer}``, ``init_video_Qformer``
* ``models/Qformer.py``: the 2-layer Q-Former those methods call
* ``models/encoder.py``: ``CLIP_VIT_LARGE`` / ``HUBERT_LARGE`` (constructor + forward) on small random-init checkpoints

Outputs: ``reference_sampling.json`` (integer contracts, bit-exact) and ``reference_fusion.npz`` (fp32 outputs for
seeded inputs and seeded weights; weights are rebuilt from seeds by oracle/, not stored).
"""
import json
import os
import random
import sys
import tempfile

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_loader as L  # noqa: E402
from oracle import qformer as OQ  # noqa: E402
from oracle import reference_path as R  # noqa: E402


def random_au_info(rng: random.Random):
    total = rng.choice([3, 5, 7, 8, 9, 12, 20, 37, 88, 150])
    peak = rng.randrange(total)
    return {"total_frames": total,
            "peak_frames": [{"peak_index": peak, "frames_before_peak": rng.choice([0, 1, 2, 3, peak]),
                             "frames_after_peak": rng.choice([0, 1, 2, 3, total - 1 - peak])}]}


def sampling_fixtures():
    vp = L.load_video_processor()
    out = {"uniform8": {}, "uniform16": {}, "emotion_peak_noinfo": {}, "headtail": [], "smart": [], "emotion_peak_json": [],
           "face_uniform8": {}, "audio": []}
    vlens = list(range(1, 131)) + [199, 250, 256, 300, 1000, 4321]
    for vlen in vlens:
        L.FakeVideoReader.vlen = vlen
        out["uniform8"][str(vlen)] = [int(x) for x in vp.load_video("v", n_frms=8, sampling="uniform")[0, :, 0, 0]]
        out["uniform16"][str(vlen)] = [int(x) for x in vp.load_video("v", n_frms=16, sampling="uniform")[0, :, 0, 0]]
        out["emotion_peak_noinfo"][str(vlen)] = [int(x) for x in
                                                 vp.load_video("v", n_frms=8, sampling="emotion_peak")[0, :, 0, 0]]
    for seed in range(24):
        vlen = [16, 17, 30, 88, 101, 300][seed % 6]
        L.FakeVideoReader.vlen = vlen
        vp.rnd.seed(seed)
        idx = [int(x) for x in vp.load_video("v", n_frms=8, sampling="headtail")[0, :, 0, 0]]
        out["headtail"].append({"seed": seed, "vlen": vlen, "indices": idx})
    rng = random.Random(2024)
    cases = [None, {}, {"peak_frames": []}] + [random_au_info(rng) for _ in range(120)]
    for au in cases:
        tv = au["total_frames"] if au and "total_frames" in au else rng.choice([8, 20, 88])
        out["smart"].append({"au_info": au, "total_video_frames": tv,
                             "indices": [int(i) for i in vp._calculate_smart_frame_indices(au, tv)]})
    with tempfile.TemporaryDirectory() as d:               # full emotion_peak path incl. _load_au_info (JSON on disk)
        for k in range(12):
            au = random_au_info(rng)
            name = f"clip{k}"
            os.makedirs(os.path.join(d, name))
            with open(os.path.join(d, name, f"{name}_au_analysis.json"), "w") as f:
                json.dump({"au_info": au}, f)
            L.FakeVideoReader.vlen = au["total_frames"]
            idx = [int(x) for x in vp.load_video("v", n_frms=8, sampling="emotion_peak", video_name=name,
                                                 mer_factory_output=d)[0, :, 0, 0]]
            out["emotion_peak_json"].append({"au_info": au, "vlen": au["total_frames"], "indices": idx})
        for vlen in (1, 3, 8, 9, 37, 88):                   # load_face: same index rule on a face stack
            faces = np.stack([np.full((5, 7, 3), i, np.uint8) for i in range(vlen)])
            path = os.path.join(d, f"f{vlen}.npy")
            np.save(path, faces)
            fr = vp.load_face(path, n_frms=8, height=224, width=224, sampling="uniform")
            assert tuple(fr.shape) == (3, 8, 224, 224)
            out["face_uniform8"][str(vlen)] = [int(x) for x in fr[0, :, 0, 0]]
        # cv2 bilinear resize golden for the GPU staging kernel: 3 random crops -> 224x224
        g = np.random.default_rng(7)
        faces = g.integers(0, 256, size=(3, 61, 47, 3), dtype=np.uint8)
        path = os.path.join(d, "faces.npy")
        np.save(path, faces)
        fr = vp.load_face(path, n_frms=3, height=224, width=224, sampling="uniform")       # [3, 3, 224, 224] float
        np.savez_compressed(os.path.join(HERE, "reference_face_resize.npz"), faces=faces,
                            resized=fr.numpy().astype(np.uint8))
    ib = L.load_imagebind_data()
    for n, ch in [(55357, 1), (64000, 1), (128000, 1), (32000, 1), (20000, 1), (32001, 1), (47999, 2), (160000, 2),
                  (16000 * 7 + 123, 1), (1, 1), (31999, 2), (480000, 1)]:
        base = torch.arange(n, dtype=torch.float32)
        wav = torch.stack([base + 1.0, base + 1.0]) if ch == 2 else (base + 1.0)[None]    # sample k holds k+1 (0 = padding)
        L.FakeAudio.table["a"] = (wav, 16000)
        clips = ib.load_audio(["a"], "cpu")
        assert tuple(clips.shape) == (1, 8, 1, 32000)
        first = clips[0, :, 0, 0]
        last = clips[0, :, 0, -1]
        out["audio"].append({"n": n, "channels": ch, "starts": [int(x) - 1 for x in first],
                             "last_sample": [int(x) - 1 for x in last]})
    return out


class _Enc:
    """registry encoder stand-in: returns the hidden states it was given (the encoders are tested on their own)."""

    def __init__(self, h):
        self.h = h

    def __call__(self, a, b):
        return self.h


def _linear(state, name):
    w, b = state[name + ".weight"], state[name + ".bias"]
    m = nn.Linear(w.shape[1], w.shape[0])
    m.weight.data.copy_(w)
    m.bias.data.copy_(b)
    return m


def fusion_fixtures():
    A = L.load_affectgpt()
    out = {}
    g = torch.Generator().manual_seed(7)
    frame_h = torch.randn(3, 8, 768, generator=g)
    face_h = torch.randn(3, 8, 768, generator=g)
    audio_h = torch.randn(3, 8, 1024, generator=g)
    frame_vec = torch.randn(3, 768, generator=g)
    dummy = torch.zeros(1)
    # ---- attention / mean variants: the weights of oracle.reference_path.build_fusion_state(1234)
    st = R.build_fusion_state(seed=1234)
    m = L.bare_affectgpt(A, num_video_query_token=1, num_audio_query_token=1, num_multi_query_token=1,
                         video_fusion_type="attention", audio_fusion_type="attention", multi_fusion_type="attention")
    for name in ("video_attention_mlp", "affectgpt_proj", "audio_attention_mlp", "audio_llama_proj", "multi_video_embs",
                 "multi_audio_embs", "attention_mlp", "fc_att", "multi_llama_proj"):
        setattr(m, name, _linear(st, name))
    m.eval()
    with torch.no_grad():
        m.visual_encoder = _Enc(frame_h)
        out["att_video_realtime_llm"] = m.encode_video_attention(dummy, dummy)[1].contiguous().numpy()
        out["att_video_merge_realtime_llm"] = m.encode_video_merge(dummy, dummy)[1].contiguous().numpy()
        out["att_video_pre3d_llm"] = m.encode_video_merge(face_h, None, is_preextracted=True)[1].contiguous().numpy()
        out["att_video_pre2d_llm"] = m.encode_video_merge(frame_vec, None, is_preextracted=True)[1].contiguous().numpy()
        out["att_video_pre2d_hidden"] = m.encode_video_merge(frame_vec, None, is_preextracted=True)[0].contiguous().numpy()
        m.acoustic_encoder = _Enc(audio_h)
        out["att_audio_realtime_llm"] = m.encode_audio_attention(dummy, dummy)[1].contiguous().numpy()
        out["att_audio_pre_llm"] = m.encode_audio_merge(audio_h, None, is_preextracted=True)[1].contiguous().numpy()
        mh, ml = m.encode_multi_merge(face_h, audio_h)
        out["att_multi_hidden"], out["att_multi_llm"] = mh.contiguous().numpy(), ml.contiguous().numpy()
        m.video_fusion_type = m.audio_fusion_type = "mean"
        out["mean_video_realtime_llm"] = m.encode_video_merge(dummy, dummy)[1].contiguous().numpy()
        out["mean_video_pre3d_llm"] = m.encode_video_merge(face_h, None, is_preextracted=True)[1].contiguous().numpy()
        out["mean_audio_realtime_llm"] = m.encode_audio_merge(dummy, dummy)[1].contiguous().numpy()
        out["mean_audio_pre_llm"] = m.encode_audio_merge(audio_h, None, is_preextracted=True)[1].contiguous().numpy()
    # ---- qformer variants: reference init_video_Qformer structure, weights from oracle.qformer.fusion_qformer_state
    nq = 4
    qs = OQ.fusion_qformer_state(seed=4321, n_query=nq)
    q = L.bare_affectgpt(A, num_video_query_token=nq, num_audio_query_token=nq, num_multi_query_token=nq,
                         video_fusion_type="qformer", audio_fusion_type="qformer", multi_fusion_type="qformer")
    for which, width in (("video", 768), ("audio", 1024), ("multi", 1024)):
        Q, tok = A.init_video_Qformer(num_query_token=nq, vision_width=width, num_hidden_layers=2)
        Q.cls = None                                                       # affectgpt.py:161-166 / :227-232 / :322-327
        Q.bert.embeddings.word_embeddings = None
        Q.bert.embeddings.position_embeddings = None
        for layer in Q.bert.encoder.layer:
            layer.output = None
            layer.intermediate = None
        sub = {k[len(which) + len("_Qformer."):]: v for k, v in qs.items() if k.startswith(which + "_Qformer.")}
        missing, unexpected = Q.load_state_dict(sub, strict=False)
        assert not unexpected and all(k.endswith("position_ids") for k in missing), (missing, unexpected)
        setattr(q, which + "_Qformer", Q)
        setattr(q, which + "_query_tokens", nn.Parameter(qs[which + "_query_tokens"].clone()))
    q.video_frame_position_embedding = nn.Embedding.from_pretrained(qs["video_frame_position_embedding.weight"])
    q.audio_position_embedding = nn.Embedding.from_pretrained(qs["audio_position_embedding.weight"])
    q.multi_position_embedding = nn.Embedding.from_pretrained(qs["multi_position_embedding.weight"])
    for name in ("affectgpt_proj", "audio_llama_proj", "multi_video_embs", "multi_audio_embs", "multi_llama_proj"):
        setattr(q, name, _linear(qs, name))
    q.eval()
    with torch.no_grad():
        q.visual_encoder = _Enc(frame_h)
        out["qf_video_realtime_llm"] = q.encode_video_qformer(dummy, dummy)[1].contiguous().numpy()
        out["qf_video_pre3d_llm"] = q.encode_video_merge(face_h, None, is_preextracted=True)[1].contiguous().numpy()
        out["qf_video_pre2d_llm"] = q.encode_video_merge(frame_vec, None, is_preextracted=True)[1].contiguous().numpy()
        q.acoustic_encoder = _Enc(audio_h)
        out["qf_audio_realtime_llm"] = q.encode_audio_qformer(dummy, dummy)[1].contiguous().numpy()
        mh, ml = q.encode_multi_merge(face_h, audio_h)
        out["qf_multi_hidden"], out["qf_multi_llm"] = mh.contiguous().numpy(), ml.contiguous().numpy()
        # the restatement in oracle/qformer.py must agree with the reference's module on the spot
        for key, got in (("qf_video_realtime_llm", OQ.encode_video_qformer(qs, frame_h)[1]),
                         ("qf_audio_realtime_llm", OQ.encode_audio_qformer(qs, audio_h)[1]),
                         ("qf_multi_llm", OQ.encode_multi_qformer(qs, face_h, audio_h)[1])):
            err = float((got - torch.from_numpy(out[key])).abs().max())
            print(f"oracle vs reference {key}: max abs diff {err:.3e}")
            assert err < 2e-5, (key, err)
    return out


def encoder_fixtures():
    """The reference's CLIP_VIT_LARGE / HUBERT_LARGE wrapper classes run on small random-init checkpoints (the oracle's
    seeded builders, written to disk with save_pretrained): pins the preprocessing quirks (uint8 truncation, joint
    waveform normalisation over the call, last-4-layer / time mean) against the reference's own forward code."""
    from transformers import CLIPConfig, CLIPImageProcessor, CLIPModel, Wav2Vec2FeatureExtractor
    out = {}
    with tempfile.TemporaryDirectory() as d:
        vis, aud = os.path.join(d, "clip"), os.path.join(d, "hubert")
        tower = R.build_clip_vision(num_layers=2, seed=1234)
        cfg = CLIPConfig(vision_config=tower.config.to_dict(),
                         text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                                          vocab_size=100, max_position_embeddings=16, projection_dim=768),
                         projection_dim=768)
        full = CLIPModel(cfg).eval()
        missing, unexpected = full.load_state_dict(tower.state_dict(), strict=False)
        assert not unexpected, unexpected
        full.save_pretrained(vis)
        CLIPImageProcessor(do_resize=True, size={"shortest_edge": 224}, do_center_crop=True,
                           crop_size={"height": 224, "width": 224}, do_normalize=True, image_mean=list(R.CLIP_MEAN),
                           image_std=list(R.CLIP_STD), do_convert_rgb=True).save_pretrained(vis)
        hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
        hub.save_pretrained(aud)
        Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                 return_attention_mask=True).save_pretrained(aud)
        enc = L.load_encoders({"CLIP_VIT_LARGE": vis}, {"HUBERT_LARGE": aud})
        clip = enc.CLIP_VIT_LARGE()
        assert clip.hidden_size == 768
        hubert = enc.HUBERT_LARGE()
        assert hubert.hidden_size == 1024
        with torch.no_grad():
            frames = R.synth_frames(1, 2, seed=1) + 0.6                     # fractional part: exercises the uint8 truncation
            out["clip2_frames_seed1_plus06_b1_t2"] = clip(None, frames).numpy()
            audio = R.synth_audio(1, 2, seed=3)
            out["hubert3_layer_audio_seed3_b1_t2"] = hubert(None, audio).numpy()
            audio2 = R.synth_audio(2, 2, seed=5)                             # b = 2: joint normalisation over 4 segments
            out["hubert3_layer_audio_seed5_b2_t2"] = hubert(None, audio2).numpy()
            for key, got in (("clip2_frames_seed1_plus06_b1_t2", R.clip_vit_large_forward(tower, frames)),
                             ("hubert3_layer_audio_seed3_b1_t2", R.hubert_large_forward(hub, audio)),
                             ("hubert3_layer_audio_seed5_b2_t2", R.hubert_large_forward(hub, audio2))):
                err = float((got - torch.from_numpy(out[key])).abs().max())
                print(f"oracle vs reference {key}: max abs diff {err:.3e}")
                assert err < 1e-4, (key, err)
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    samp = sampling_fixtures()
    with open(os.path.join(HERE, "reference_sampling.json"), "w") as f:
        json.dump(samp, f, separators=(",", ":"))
    fus = fusion_fixtures()
    np.savez_compressed(os.path.join(HERE, "reference_fusion.npz"), **fus)
    enc = encoder_fixtures()
    np.savez_compressed(os.path.join(HERE, "reference_encoders.npz"), **enc)
    print("wrote", len(samp), "sampling tables,", len(fus), "fusion tensors and", len(enc), "encoder outputs")
