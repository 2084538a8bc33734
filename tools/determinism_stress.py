"""Run the same C2-shaped batch repeatedly (eager and graph replay) and compare every output bit for bit with the first
run: catches races in the fused-LayerNorm GEMM tail (inter-CTA meeting) that parity tolerances would hide."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affectgpt_b200 as agb
from affectgpt_b200 import synthetic as S
DEV = 'cuda:0'
B = int(os.environ.get("CLIPS", "64")); N = int(os.environ.get("REPS", "20"))
m = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV)
fr, au = S.frames(B, 8, seed=61).to(DEV), S.audio(B, 8, seed=71).to(DEV)
keys = ("multi", "audio", "frame", "frame_hiddens", "audio_hiddens")
def run():
    o = m.encode_clip_batch({"raw_frames": fr, "raw_audios": au, "face_or_frame": "multiframe_audio_frame_text"})
    return {k: o[k].clone() for k in keys}
ref = run()
bad = 0
for i in range(N):
    o = run()
    for k in keys:
        if not torch.equal(o[k], ref[k]):
            bad += 1
            d = (o[k] - ref[k]).abs()
            print(f"eager rep {i} {k}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}, rows {torch.nonzero(d.flatten(1).amax(1) > 0).flatten()[:8].tolist()}")
g = agb.GraphedClipBatch(m, B, 8, 8, device=DEV)
for i in range(N):
    o = g(fr, au)
    torch.cuda.synchronize()
    for k in keys:
        if not torch.equal(o[k], ref[k]):
            bad += 1
            d = (o[k] - ref[k]).abs()
            print(f"graph rep {i} {k}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e}, rows {torch.nonzero(d.flatten(1).amax(1) > 0).flatten()[:8].tolist()}")
print("mismatching outputs:", bad, "of", 2 * N * len(keys))
