"""Throughput of the pre-extraction sweep CLI path (C5 shape: frame + face + audio + multi files per sample) on a
bounded synthetic dataset: the reference-style per-sample loop (--batch_size 1) vs device batches (--batch_size N).
Full 24-layer random-init encoders; sources are pre-decoded uint8 stacks / wav files on local disk, so decoding cost is
np.load + scipy wav read (mp4 decoding is outside the path)."""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affectgpt_b200 as agb  # noqa: E402
from affectgpt_b200 import extract as X, synthetic as S  # noqa: E402

N = int(os.environ.get("SWEEP_CLIPS", "384"))
N_SINGLE = int(os.environ.get("SWEEP_CLIPS_SINGLE", "48"))
BATCH = int(os.environ.get("SWEEP_BATCH", "64"))
RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"


def write_dataset(root, n):
    from scipy.io import wavfile
    rng = np.random.default_rng(5)
    for d in ("video", "face", "audio"):
        os.makedirs(os.path.join(root, d))
    names = [f"clip_{i:05d}" for i in range(n)]
    for i, nm in enumerate(names):
        np.save(os.path.join(root, "video", nm + ".npy"), rng.integers(0, 256, size=(16, 224, 224, 3), dtype=np.uint8))
        os.makedirs(os.path.join(root, "face", nm))
        np.save(os.path.join(root, "face", nm, nm + ".npy"), rng.integers(0, 256, size=(12, 112, 96, 3), dtype=np.uint8))
        wavfile.write(os.path.join(root, "audio", nm + ".wav"), 16000,
                      (rng.standard_normal(64000) * 4000).astype(np.int16))          # 4 s
    with open(os.path.join(root, "names.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    return names


def main():
    """Under torchrun every rank sweeps its own N samples (what `python -m affectgpt_b200.extract` does with samples
    rank::world): no collective on the data path, a barrier on both sides of the timed region, time = max over ranks."""
    root = tempfile.mkdtemp(prefix=f"agb_sweep_r{RANK}_")
    if WORLD > 1:
        import torch.distributed as dist
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    try:
        names = write_dataset(root, N)
        model = agb.build_path(S.clip_state(24, dev), S.hubert_state(24, "layer", dev), S.fusion_state(3584, dev), device=dev)
        ex = X.MultimodalFeatureExtractor(model.visual_encoder, model.acoustic_encoder, model, device=dev)
        res = {"clips": N, "batch": BATCH, "files_per_clip": 4}

        def args_for(save, bs):
            a = X.build_parser().parse_args([
                "--dataset", "mer2024", "--video_root", f"{root}/video", "--face_root", f"{root}/face", "--audio_root",
                f"{root}/audio", "--sample_list", f"{root}/names.txt", "--save_root", f"{root}/{save}", "--quiet",
                "--batch_size", str(bs)])
            X.validate_args(a)
            return a
        X.run_sweep_batched(args_for("warm", BATCH), ex, names[:BATCH], batch_size=BATCH, log=lambda *_: None)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = X.run_sweep(args_for("single", 1), ex, names[:N_SINGLE], log=lambda *_: None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["per_sample_loop"] = {"clips": N_SINGLE, "seconds": dt, "clips_per_s": N_SINGLE / dt, "stats": st}
        if RANK == 0:
            print(f"per-sample loop : {N_SINGLE / dt:7.1f} clips/s per GPU ({N_SINGLE} clips, {dt:.1f} s)", flush=True)
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = X.run_sweep_batched(args_for("batched", BATCH), ex, names, batch_size=BATCH, workers=16, log=lambda *_: None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if WORLD > 1:
            tmax = torch.tensor([dt], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax)
        res["batched"] = {"clips": N * WORLD, "gpus": WORLD, "seconds": dt, "clips_per_s": N * WORLD / dt, "stats_rank0": st}
        if RANK == 0:
            print(f"batched ({BATCH:3d})   : {N * WORLD / dt:7.1f} clips/s on {WORLD} GPU(s) ({N * WORLD} clips, {dt:.1f} s)", flush=True)
        same = all(np.array_equal(np.load(f"{root}/single/mer2024/{sub}/{nm}.npy"), np.load(f"{root}/batched/mer2024/{sub}/{nm}.npy"))
                   for sub in os.listdir(f"{root}/single/mer2024") for nm in names[:N_SINGLE])
        res["files_identical"] = bool(same)
        if RANK == 0:
            print("files identical:", same)
            out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
            os.makedirs(out, exist_ok=True)
            json.dump(res, open(os.path.join(out, f"sweep_bench_n{WORLD}.json"), "w"), indent=1)
            print(json.dumps(res))
        if WORLD > 1:
            dist.barrier()
            dist.destroy_process_group()
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
