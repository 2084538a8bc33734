"""BASELINE config C5 at its stated size: the pre-extraction sweep of extract_multimodal_features_precompute.py
(A/extract_multimodal_features_precompute.py:892-1024) over 132 171 synthetic samples on all GPUs of the node.

    torchrun --nproc-per-node 8 tools/c5_sweep.py [--samples 132171] [--batch 64] [--save-root /tmp/agb_c5]

* Inputs are generated ON THE DEVICE per sample from ``seed = clip_index``: 8 frames + 8 face crops (uint8
  224 x 224) and a 4 s waveform cut into the reference's 8 x 2 s segments (ImageBind/data.py:70-78 rule, computed by
  affectgpt_b200.sampling).  A sample therefore does not depend on the rank or the batch it lands in.
* Sharding as in the sweep CLI (`python -m affectgpt_b200.extract` under torchrun): batches of 64 consecutive samples are
  claimed from a counter all ranks share (extract.shared_batch_counter), so a GPU that runs slower under the power cap
  claims fewer batches; `--static-shards` gives rank r the batches r::W instead.  No collective on the data path.
* Every sample writes the reference's four files - frame [8,768], face [8,768], audio [8,1024], multi [1024], fp32 .npy -
  into ``save_root/<dataset>/{frame_*, face_*, audio_*, multi_*}/<name>.npy`` (extract.save_dirs) from a writer pool,
  while the next batch is on the GPU (two pinned host slots per rank).
* Reported: steady-state samples/s of the whole node (first batch of every rank excluded), the device-bound rate of the
  same loop without D2H / file writes, files and bytes written, per-rank times, a determinism probe (samples of another
  rank's shard recomputed here and compared with the files on disk bit for bit), and the reference-style CPU loop (oracle,
  fp32, per-sample) on a bounded sample beside it.
"""
import argparse
import json
import os
import shutil
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import affectgpt_b200 as agb  # noqa: E402
from affectgpt_b200 import extract as X, sampling, synthetic as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=132171)                 # /root/reference/README.md:33,47
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--save-root", default="/tmp/agb_c5")
ap.add_argument("--writers", type=int, default=8)
ap.add_argument("--cpu-samples", type=int, default=2)
ap.add_argument("--keep", action="store_true", help="leave the written files on disk")
ap.add_argument("--static-shards", action="store_true", help="batches rank::world instead of the shared counter")
args = ap.parse_args()

RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
DEV = torch.device("cuda", LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=DEV)

B = args.batch
STARTS = [a for a, _ in sampling.audio_segment_bounds(64000, 16000, 2, 8)]     # [0, 4571, ..., 32000]
_gen = torch.Generator(device=DEV)


def make_sample(idx: int, frames: torch.Tensor, faces: torch.Tensor, audio: torch.Tensor, slot: int) -> None:
    """Sample ``idx`` into row ``slot`` of the batch buffers: seed = clip index, nothing else."""
    _gen.manual_seed(idx)
    frames[slot] = torch.randint(0, 256, (3, 8, 224, 224), generator=_gen, device=DEV, dtype=torch.uint8)
    faces[slot] = torch.randint(0, 256, (3, 8, 224, 224), generator=_gen, device=DEV, dtype=torch.uint8)
    wav = torch.randn(64000, generator=_gen, device=DEV) * (0.02 + 0.48 * (idx % 97) / 96.0)
    for t, s0 in enumerate(STARTS):
        audio[slot, t, 0] = wav[s0:s0 + 32000]


def main():
    model = agb.build_path(S.clip_state(24, DEV), S.hubert_state(24, "layer", DEV), S.fusion_state(3584, DEV), device=DEV)
    graph = agb.GraphedClipBatch(model, B, 8, 8, with_faces=True, device=DEV, frame_dtype=torch.uint8)
    xargs = X.build_parser().parse_args(["--dataset", "mercaptionplus", "--save_root", args.save_root, "--quiet"])
    dirs = X.save_dirs(xargs)                 # the sources are synthetic: no --video_root / --face_root / --audio_root to validate
    n_batches = (args.samples + B - 1) // B
    batch_of = lambda b: list(range(b * B, min((b + 1) * B, args.samples)))          # noqa: E731
    if WORLD > 1 and not args.static_shards:
        claim = X.shared_batch_counter(RANK, WORLD)
    else:
        _it = iter(range(RANK, n_batches, WORLD))
        claim = lambda: next(_it, n_batches)                                          # noqa: E731
    own_batches = []
    widths = {"frame": (8, 768), "face": (8, 768), "audio": (8, 1024), "multi": (1024,)}
    keys = {"frame": "frame_hiddens", "face": "face_hiddens", "audio": "audio_hiddens", "multi": "multi_hiddens"}
    host = [{k: torch.empty((B,) + w, dtype=torch.float32).pin_memory() for k, w in widths.items()} for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    cur = torch.cuda.current_stream(DEV)
    written = {"files": 0, "bytes": 0}
    lock = threading.Lock()

    def write_batch(idxs, slot, ev):
        ev.synchronize()
        n_files = n_bytes = 0
        for j, idx in enumerate(idxs):
            name = f"sample_{idx:08d}"
            for kind in widths:
                arr = host[slot][kind][j].numpy()
                np.save(os.path.join(dirs[kind], name + ".npy"), arr)
                n_files += 1
                n_bytes += arr.nbytes + 128
        with lock:
            written["files"] += n_files
            written["bytes"] += n_bytes

    def run_batch(idxs):
        for j, idx in enumerate(idxs):
            make_sample(idx, graph.raw_frames, graph.raw_faces, graph.raw_audios, j)
        return graph()

    def barrier():
        if WORLD > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up + device-bound rate of the same loop (generation + encoders, no D2H, no files)
    run_batch(batch_of(RANK % n_batches))
    barrier()
    t0 = time.perf_counter()
    n_dev = 0
    for k in range(1, 9):
        b = batch_of((RANK + k * WORLD) % n_batches)
        if len(b) == B:
            run_batch(b)
            n_dev += B
    torch.cuda.synchronize()
    dev_rate = n_dev / (time.perf_counter() - t0) if n_dev else 0.0

    # ---- the sweep
    barrier()
    t_start = time.perf_counter()
    t_first = None
    futs = []
    with ThreadPoolExecutor(max_workers=args.writers) as pool:
        slot_busy = [None, None]
        bi = -1
        while True:
            b = claim()
            if b >= n_batches:
                break
            own_batches.append(b)
            idxs = batch_of(b)
            bi += 1
            slot = bi % 2
            if slot_busy[slot] is not None:
                slot_busy[slot].result()                               # the writer has drained this host slot
            out = run_batch(idxs)
            for kind, k in keys.items():
                host[slot][kind][:len(idxs)].copy_(out[k][:len(idxs)].reshape((len(idxs),) + widths[kind]), non_blocking=True)
            copied[slot].record(cur)
            slot_busy[slot] = pool.submit(write_batch, list(idxs), slot, copied[slot])
            futs.append(slot_busy[slot])
            if bi == 0:
                torch.cuda.synchronize()
                t_first = time.perf_counter()
        for f in futs:
            f.result()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    n_own = sum(len(batch_of(b)) for b in own_batches)
    own = {"seconds": t_end - t_start, "steady_seconds": t_end - (t_first or t_start), "samples": n_own, "device_bound_rate": dev_rate}

    # ---- determinism probe: 4 samples of the NEXT rank's shard recomputed here vs that rank's files
    barrier()
    mine_set = set(own_batches)
    foreign = [b for b in range(n_batches) if b not in mine_set] or list(range(n_batches))
    probe = [batch_of(b)[0] for b in foreign[:: max(1, len(foreign) // 4)][:4]]        # samples some OTHER rank wrote
    run_batch((probe + batch_of(0))[:B])
    out = {k: graph.outputs[v][:len(probe)].float().cpu().numpy() for k, v in keys.items()}
    same = True
    for j, idx in enumerate(probe):
        for kind in widths:
            disk = np.load(os.path.join(dirs[kind], f"sample_{idx:08d}.npy"))
            same = same and np.array_equal(disk, out[kind][j].reshape(widths[kind]))
    rows = torch.tensor([own["seconds"], own["steady_seconds"], float(own["samples"]), dev_rate, float(written["files"]),
                         float(written["bytes"]), 1.0 if same else 0.0], dtype=torch.float64, device=DEV)
    allr = [rows.clone() for _ in range(WORLD)]
    if WORLD > 1:
        dist.all_gather(allr, rows)
    if RANK == 0:
        sec = max(float(r[0]) for r in allr)
        steady = max(float(r[1]) for r in allr)
        total = int(sum(float(r[2]) for r in allr))
        res = {"config": "C5: pre-extraction sweep, three streams per sample (8 frames + 8 face crops + 8 x 2 s audio), "
                         "4 files per sample in the reference layout",
               "samples": total, "gpus": WORLD, "batch": B, "seconds": sec,
               "samples_per_s": total / sec,
               "steady_state_samples_per_s": (total - WORLD * B) / steady,
               "device_bound_samples_per_s": sum(float(r[3]) for r in allr),
               "host_pipeline_loss": 1.0 - (total - WORLD * B) / steady / max(sum(float(r[3]) for r in allr), 1e-9),
               "files_written": int(sum(float(r[4]) for r in allr)), "bytes_written": int(sum(float(r[5]) for r in allr)),
               "sharding": "static (batches rank::world)" if args.static_shards or WORLD == 1 else "dynamic (batches claimed from a shared counter)",
               "per_rank_samples": [int(float(r[2])) for r in allr],
               "per_rank_seconds": [round(float(r[0]), 2) for r in allr],
               "per_rank_device_bound": [round(float(r[3]), 1) for r in allr],
               "cross_rank_files_bit_identical": all(float(r[6]) == 1.0 for r in allr),
               "inputs": "generated on the device per sample from seed = clip_index", "save_root": args.save_root}
        if args.cpu_samples > 0:
            from oracle import reference_path as R
            torch.set_num_threads(os.cpu_count() or 1)
            clip, hub = R.build_clip_vision(24, seed=1234), R.build_hubert(24, "layer", seed=1234)
            fus = R.build_fusion_state(seed=1234)

            def one(i):
                fh = R.clip_vit_large_forward(clip, R.synth_frames(1, 8, seed=100 + i))
                ch = R.clip_vit_large_forward(clip, R.synth_frames(1, 8, seed=300 + i))
                ah = R.hubert_large_forward(hub, R.synth_audio(1, 8, seed=200 + i))
                return R.fused_tokens(fus, fh, ch, ah, "multiface_audio_face_frame_text")
            one(0)
            t0 = time.perf_counter()
            for i in range(args.cpu_samples):
                one(1 + i)
            dt = time.perf_counter() - t0
            res["cpu_reference_loop"] = {"samples_per_s": args.cpu_samples / dt, "cores": os.cpu_count(),
                                         "sample": f"{args.cpu_samples} samples, per-sample loop, oracle fp32 (HF modules + restated wrappers)",
                                         "hours_for_all_samples": total / (args.cpu_samples / dt) / 3600.0}
        print(json.dumps(res), flush=True)
        out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        json.dump(res, open(os.path.join(out_dir, f"c5_sweep_n{WORLD}.json"), "w"), indent=1)
    barrier()
    if RANK == 0 and not args.keep:
        shutil.rmtree(args.save_root, ignore_errors=True)
    if WORLD > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
