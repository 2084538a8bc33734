"""End-to-end check of the sweep CLI under torchrun (dynamic shards): small random-init HF checkpoints written with
save_pretrained, a synthetic dataset of pre-decoded sources, then

    python -m affectgpt_b200.extract ... --batch_size 2                      (one process)
    torchrun --nproc-per-node W -m affectgpt_b200.extract ... --batch_size 2 (W ranks claiming batches from the shared counter)

and the two output trees must hold the same files with the same bytes.  Needs W GPUs (default 2)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_path as R  # noqa: E402  (checkpoint builders only; the CLI itself never imports oracle)

W = int(os.environ.get("RANKS", "2"))
tmp = tempfile.mkdtemp(prefix="agb_cli_")
vis, aud = os.path.join(tmp, "clip"), os.path.join(tmp, "hubert")
from transformers import CLIPConfig, CLIPModel  # noqa: E402
tower = R.build_clip_vision(num_layers=2, seed=1234)
cfg = CLIPConfig(vision_config=tower.config.to_dict(),
                 text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                                  vocab_size=100, max_position_embeddings=16, projection_dim=768), projection_dim=768)
full = CLIPModel(cfg).eval()
full.load_state_dict(tower.state_dict(), strict=False)
full.save_pretrained(vis)
R.build_hubert(num_layers=3, feat_extract_norm="layer", seed=1234).save_pretrained(aud)
torch.save({"model": R.build_fusion_state(seed=1234)}, os.path.join(tmp, "fusion.pth"))

from scipy.io import wavfile  # noqa: E402
rng = np.random.default_rng(3)
names = [f"clip{i:02d}" for i in range(9)]
for d in ("video", "face", "audio"):
    os.makedirs(os.path.join(tmp, d))
for i, n in enumerate(names):
    np.save(os.path.join(tmp, "video", n + ".npy"), rng.integers(0, 256, size=(12 + i, 224, 224, 3), dtype=np.uint8))
    os.makedirs(os.path.join(tmp, "face", n))
    np.save(os.path.join(tmp, "face", n, n + ".npy"), rng.integers(0, 256, size=(9 + i, 64, 48, 3), dtype=np.uint8))
    wavfile.write(os.path.join(tmp, "audio", n + ".wav"), 16000, (rng.standard_normal(40000 + 3000 * i) * 3000).astype(np.int16))
with open(os.path.join(tmp, "names.txt"), "w") as f:
    f.write("\n".join(names + ["missing_everything"]) + "\n")

env = dict(os.environ, AGB_CLIP_VIT_LARGE=vis, AGB_HUBERT_LARGE=aud, PYTHONPATH=ROOT)
common = ["--dataset", "mer2024", "--modality", "all", "--video_root", f"{tmp}/video", "--face_root", f"{tmp}/face",
          "--audio_root", f"{tmp}/audio", "--sample_list", f"{tmp}/names.txt", "--fusion_ckpt", f"{tmp}/fusion.pth",
          "--batch_size", "2", "--quiet"]
r1 = subprocess.run([sys.executable, "-m", "affectgpt_b200.extract", *common, "--save_root", f"{tmp}/one"], env=env, cwd=ROOT,
                    stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
print(r1.stdout[-600:])
assert r1.returncode == 0
r2 = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={W}", "--master-addr", "127.0.0.1",
                     "--master-port", "29711", "-m", "affectgpt_b200.extract", *common, "--save_root", f"{tmp}/many"], env=env, cwd=ROOT,
                    stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
print("\n".join(ln for ln in r2.stdout.splitlines() if "rank" in ln or "Error" in ln or "error" in ln)[-1200:])
assert r2.returncode == 0, r2.stdout[-2000:]


def tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            out[os.path.relpath(os.path.join(d, f), root)] = np.load(os.path.join(d, f))
    return out


a, b = tree(f"{tmp}/one"), tree(f"{tmp}/many")
assert sorted(a) == sorted(b), (sorted(set(a) ^ set(b)))
assert len(a) == 4 * len(names) + 1, len(a)                       # + the zero-filled audio of the missing sample
bad = [k for k in a if not np.array_equal(a[k], b[k])]
assert not bad, bad
print(f"cli-torchrun-check ok: {len(a)} files, {W} ranks, bit-identical to the single-process sweep")
