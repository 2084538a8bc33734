"""A/B of environment switches inside ONE process: one CUDA graph of the C2 step per setting, replayed alternately
(A B A B ...), so that box-to-box and minute-to-minute clock / power differences cancel.

    python tools/ab_graph.py AGB_FUSED_LN=1 AGB_FUSED_LN=0 [--clips 64] [--rounds 12]

Only switches that libagb reads per call (not cached in a static) can be compared this way."""
import argparse, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affectgpt_b200 as agb
from affectgpt_b200 import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("settings", nargs="+", help="KEY=VALUE[,KEY=VALUE...] per arm")
ap.add_argument("--clips", type=int, default=64)
ap.add_argument("--rounds", type=int, default=12)
args = ap.parse_args()
DEV = "cuda:0"
_models = {}


def model_for(chunk):
    """One engine per workspace chunking (pseudo-setting CHUNK=<frames>x<segments>; default 512x512)."""
    if chunk not in _models:
        f, sgm = (int(x) for x in chunk.split("x"))
        _models[chunk] = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV),
                                        device=DEV, max_frames_per_chunk=f, max_segments_per_chunk=sgm)
    return _models[chunk]


fr, au = S.frames(args.clips, 8, seed=1).to(DEV), S.audio(args.clips, 8, seed=2).to(DEV)
arms = []
for setting in args.settings:
    kv = dict(x.split("=", 1) for x in setting.split(",") if x)
    model = model_for(kv.pop("CHUNK", "512x512"))
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    g = agb.GraphedClipBatch(model, args.clips, 8, 8, device=DEV)
    g.raw_frames.copy_(fr); g.raw_audios.copy_(au)
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    arms.append((setting, g, []))
ref = None
for _, g, _ in arms:                      # all arms must agree bit for bit (these switches do not change the arithmetic)
    o = g()
    t = torch.cat([o["multi"], o["audio"], o["frame"]], dim=1).clone()
    if ref is None: ref = t
    else: print("bit-identical to arm 0:", bool(torch.equal(ref, t)))
for _ in range(3):
    for _, g, _ in arms: g()
torch.cuda.synchronize()
for r in range(args.rounds):
    for _, g, ms in arms:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g(); g(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / 2)
for name, g, ms in arms:
    print(f"{name:40s} kernels {g.kernels_per_replay:4d}  median {statistics.median(ms):8.3f} ms  min {min(ms):8.3f}  "
          f"mean {statistics.mean(ms):8.3f}  -> {args.clips / statistics.median(ms) * 1e3:7.1f} clips/s")
