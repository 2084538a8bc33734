"""Run by tests/test_gpu_switches.py in a subprocess with one AGB_* switch set: the small path (2-layer ViT, 3-layer
HuBERT, real heads) against the oracle under the BASELINE gate, plus the attention op on the path's two shapes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import affectgpt_b200 as agb  # noqa: E402
from affectgpt_b200 import ops  # noqa: E402
from oracle import reference_path as R  # noqa: E402
from tests.gpu_util import gate  # noqa: E402

DEV = "cuda:0"
clip = R.build_clip_vision(num_layers=2, seed=1234)
hub = R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234)
fus = R.build_fusion_state(seed=1234)
model = agb.build_path(clip.state_dict(), hub.state_dict(), fus, device=DEV, vit_layers=2, hubert_layers=3)
frames, audio = R.synth_frames(1, 9, seed=1), R.synth_audio(1, 24, seed=3)        # 2313 / 2400 rows: the CTA-pair GEMMs
for rep in range(2):                                                              # twice: state left behind by a launch
    out = model.encode_clip_batch({"raw_frames": frames.to(DEV), "raw_audios": audio.to(DEV),
                                   "face_or_frame": "multiframe_audio_frame_text"})
    fh, ah = R.clip_vit_large_forward(clip, frames), R.hubert_large_forward(hub, audio)
    gate(out["frame_hiddens"], fh, "frame hiddens")
    gate(out["audio_hiddens"], ah, "audio hiddens")
    want = R.fused_tokens(fus, fh, None, ah, "multiframe_audio_frame_text")
    for k in ("frame", "audio", "multi"):
        gate(out[k], want[k], f"token[{k}]")
g = torch.Generator().manual_seed(0)
for nseq, stride, valid in ((6, 257, 257), (7, 100, 99)):
    qkv = (torch.randn(nseq * stride, 3072, generator=g) * 0.7).bfloat16().to(DEV)
    got = ops.attention(qkv, nseq, stride, valid).float().reshape(nseq, stride, 16, 64)[:, :valid]
    x = qkv.float().reshape(nseq, stride, 3, 16, 64)[:, :valid]
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=0.125).permute(0, 2, 1, 3)
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-2, (nseq, stride, valid, err)
print("switch-check ok:", {k: v for k, v in os.environ.items() if k.startswith("AGB_")}, "launches", agb.launch_count())
