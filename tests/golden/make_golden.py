"""Generate tests/golden/*.npz from the oracle (oracle/reference_path.py = HF modules + restated wrappers).

Run here (CPU container): ``python tests/golden/make_golden.py``.  The fixtures pin the ORACLE's behaviour
(seeded random-init weights, seeded synthetic inputs); the reference ships no golden vectors for this path
(SURVEY F10), so these are "reference-shaped outputs of the restatement", regenerated and compared on every CPU
test run (tests/test_oracle.py) and compared against the CUDA path on the GPU (tests/test_gpu_path.py).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_path as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def small():
    out = {}
    clip = R.build_clip_vision(num_layers=2, seed=1234)
    out["clip2_frames_seed1_b1_t2"] = R.clip_vit_large_forward(clip, R.synth_frames(1, 2, seed=1)).numpy()
    for norm in ("layer", "group"):
        hub = R.build_hubert(num_layers=3, feat_extract_norm=norm, seed=1234)
        out[f"hubert3_{norm}_audio_seed3_b1_t2"] = R.hubert_large_forward(hub, R.synth_audio(1, 2, seed=3)).numpy()
    st = R.build_fusion_state(seed=1234)
    g = torch.Generator().manual_seed(7)
    fh, ch, ah = torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 768, generator=g), torch.randn(2, 8, 1024, generator=g)
    toks = R.fused_tokens(st, fh, ch, ah, "multiface_audio_face_frame_text")
    for k, v in toks.items():
        out[f"fusion_seed7_{k}"] = v.contiguous().numpy()
    return out


def full():
    """BASELINE config C1: one clip, 8 frames + 8 x 2 s audio, 24-layer encoders, fp32 CPU."""
    out = {}
    t0 = time.time()
    clip = R.build_clip_vision(num_layers=24, seed=1234)
    fh = R.clip_vit_large_forward(clip, R.synth_frames(1, 8, seed=1))
    ch = R.clip_vit_large_forward(clip, R.synth_frames(1, 8, seed=2))
    del clip
    hub = R.build_hubert(num_layers=24, feat_extract_norm="layer", seed=1234)
    ah = R.hubert_large_forward(hub, R.synth_audio(1, 8, seed=3))
    del hub
    st = R.build_fusion_state(seed=1234)
    toks = R.fused_tokens(st, fh, ch, ah, "multiface_audio_face_frame_text")
    out["c1_frame_hidden"], out["c1_face_hidden"], out["c1_audio_hidden"] = fh.numpy(), ch.numpy(), ah.numpy()
    for k, v in toks.items():
        out[f"c1_token_{k}"] = v.contiguous().numpy()
    print(f"full config took {time.time() - t0:.1f}s")
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small())
    if "--small-only" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "golden_c1.npz"), **full())
    print("written", os.listdir(HERE))
