import torch, sys
sys.path.insert(0, '.')
import affectgpt_b200 as agb
from affectgpt_b200 import synthetic as S
DEV='cuda:0'
m = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV)
B=4
def toks(fr, au):
    o = m.encode_clip_batch({"raw_frames": fr, "raw_audios": au, "face_or_frame": "multiframe_audio_frame_text"})
    return {k: o[k].clone() for k in ("multi","audio","frame","frame_hiddens","audio_hiddens")}
fr, au = S.frames(B, 8, seed=61).to(DEV), S.audio(B, 8, seed=71).to(DEV)
a = toks(fr, au); b = toks(fr, au)
print("eager vs eager:", {k: bool(torch.equal(a[k], b[k])) for k in a})
g = agb.GraphedClipBatch(m, B, 8, 8, device=DEV)
o = g(fr, au); c = {k: o[k].clone() for k in a}
print("graph vs eager:", {k: (bool(torch.equal(a[k], c[k])), float((a[k]-c[k]).abs().max())) for k in a})
o = g(fr, au); d = {k: o[k].clone() for k in a}
print("graph vs graph:", {k: bool(torch.equal(d[k], c[k])) for k in a})
