"""Probe: does one C2 step replay faster from a CUDA graph than from 376 separate launches?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affectgpt_b200 as agb  # noqa: E402
from affectgpt_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
model = agb.build_path(S.clip_state(24, dev), S.hubert_state(24, "layer", dev), S.fusion_state(3584, dev), device=dev)
B = 64
frames, audio = S.frames(B, 8, seed=1).to(dev), S.audio(B, 8, seed=2).to(dev)


def step():
    out = model.encode_clip_batch({"raw_frames": frames, "raw_audios": audio, "face_or_frame": "multiframe_audio_frame_text"})
    return torch.cat([out["multi"], out["audio"], out["frame"]], dim=1)


def timeit(fn, k=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


for _ in range(3):
    ref = step()
print(f"eager : {timeit(step):.2f} ms/step")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = step()
g.replay()
torch.cuda.synchronize()
print("graph output equals eager:", torch.equal(out, ref))
print(f"graph : {timeit(g.replay):.2f} ms/step")
