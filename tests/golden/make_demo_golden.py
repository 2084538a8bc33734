"""Structured-input golden: the reference's own demo clip instead of noise.

Run here (CPU container, /root/reference present): ``python tests/golden/make_demo_golden.py``.
Inputs: ``A/demo/sample_00000000.mp4`` (88 frames -> the reference's uniform indices [0, 11, ..., 77], decoded with
OpenCV, BGR -> RGB, resized to 224x224 the way the reference asks its decoder to) and ``A/demo/sample_00000000.wav``
(55 357 samples -> the eight 2-second slices of ImageBind/data.py:170-214).  Outputs: the oracle's features and fused
tokens for that clip on the seeded 24-layer random-init weights (seed 1234) - the inputs travel in the fixture, the
weights are regenerated from the seed on the GPU box.
"""
import os
import sys

import cv2
import numpy as np
import torch
from scipy.io import wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_path as R  # noqa: E402
from oracle import sampling as OS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DEMO = "/root/reference/AffectGPT/demo/sample_00000000"


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    cap = cv2.VideoCapture(DEMO + ".mp4")
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    idx = OS.uniform_indices(len(frames), 8)
    assert len(frames) == 88 and idx == [0, 11, 22, 33, 44, 55, 66, 77], (len(frames), idx)
    u8 = np.stack([cv2.resize(cv2.cvtColor(frames[i], cv2.COLOR_BGR2RGB), (224, 224), interpolation=cv2.INTER_LINEAR) for i in idx])
    sr, pcm = wavfile.read(DEMO + ".wav")
    assert sr == 16000 and pcm.shape == (55357,) and pcm.dtype == np.int16
    wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]                 # [1, n]
    segs, bounds = OS.audio_segments(wav)                                          # [8,1,32000]
    assert [a for a, _ in bounds] == [0, 3336, 6673, 10010, 13346, 16683, 20020, 23356]
    raw = torch.from_numpy(u8).permute(3, 0, 1, 2).float()[None]                   # [1,3,8,224,224], 0..255
    clip = R.build_clip_vision(num_layers=24, seed=1234)
    fh = R.clip_vit_large_forward(clip, raw)
    del clip
    hub = R.build_hubert(num_layers=24, feat_extract_norm="layer", seed=1234)
    ah = R.hubert_large_forward(hub, segs[None])
    del hub
    toks = R.fused_tokens(R.build_fusion_state(seed=1234), fh, None, ah, "multiframe_audio_frame_text")
    out = {"frames_u8": u8, "pcm_i16": pcm, "frame_hidden": fh.numpy(), "audio_hidden": ah.numpy()}
    for k, v in toks.items():
        out[f"token_{k}"] = v.contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "golden_demo_clip.npz"), **out)
    print("written golden_demo_clip.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
