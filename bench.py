#!/usr/bin/env python
"""Headline benchmark: fused clips/sec of the encode-and-fuse path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--impl ours|reference]

* step      = one pass of the hot path over one batch of B synthetic clips per GPU (config C2: B = 64 clips,
              each 8 frames 224x224 + 8 x 2 s of 16 kHz audio -> frame / audio / multi tokens at width 3584).
* value     = clips/s summed over all ranks, inputs already resident in HBM, CUDA-event timed, max over ranks.
* e2e       = same metric through the reference-facing wrappers with PINNED HOST buffers: H2D of frames and
              audio, forward, D2H of the fused tokens, all inside the timed region of every step (the next step's
              upload overlaps this step's compute on a copy stream, as a prefetching loader would).
* roofline  = tensor bound: algorithmic FLOPs of the tcgen05 GEMM launches / their summed CUDA-event duration
              (agb_profile_* event pairs on the launching stream, separate instrumented pass) vs the measured
              sustained cuBLAS bf16 peak in MEASURED_PEAKS.json.
* cpu_baseline = the oracle (HF fp32 modules + restated reference wrappers) timed on this box's host cores on a
              bounded sample (rank 0, N = 1 only).  `--impl reference` times that same CPU path as the arm.
* library_gpu_baseline = the reference's own arithmetic on this GPU: HuggingFace CLIPVisionModelWithProjection
              + HubertModel (what A/my_affectgpt/models/encoder.py:184-205, 403-430 call) under torch.autocast(bf16) on
              the same batch shape, cuBLASLt / cuDNN / SDPA kernels of torch 2.11 - "the kernels to beat" (N = 1 only).
Multi-GPU: one process per GPU (torchrun), clips sharded across ranks, one NCCL all_gather_into_tensor of the fused tokens
per step on a side stream (parallel.ClipGather: step i's gather overlaps step i+1's encoders; the timed region ends after
the last gather).  Every rank then recomputes two clips of ANOTHER rank's shard locally and compares them with the gathered
rows bit for bit (`multi_gpu_parity`); `per_rank` lists every rank's own step time and SM clock.  `--balance` (opt-in)
splits the global batch in proportion to each GPU's measured speed instead of equally (`balance` in the JSON line).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_VIT_FRAME = 162.026          # SURVEY 8d / BASELINE.md section 3 (2*MAC, dense)
GF_HUBERT_SEG = 72.35
GF_FUSE_CLIP = 0.034
# executed by default: the last ViT block keeps K/V for all tokens but runs Q / attention / out-proj / MLP for the pooled
# CLS row only (256 of 257 rows of 0.537 + 0.2705 + 0.537 + 4.295 GF skipped); every reported fraction uses the DENSE figure
GF_VIT_FRAME_PRUNED = 5.638
METRIC = "fused clips/sec (8 frames+4 s audio)"


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def gemm_traffic_per_launch():
    """Mean DRAM bytes (read + write) per GEMM launch of one C2 step, from the committed ncu launch list
    (profiles/r02_traffic.json, produced by tools/summarize_launches.py from an ncu capture of this very command);
    None when no capture has been committed."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return float(json.load(open(p))["gemm_kernel"]["dram_bytes_per_launch"])
    except Exception:
        return None


def gpu_energy_mj(index: int):
    """NVML total-energy counter of one GPU in mJ (None when NVML is unavailable).  Read on both sides of the timed
    region: the step runs against the power cap, so joules per step is the quantity the step time follows."""
    try:
        import pynvml
        pynvml.nvmlInit()
        return int(pynvml.nvmlDeviceGetTotalEnergyConsumption(pynvml.nvmlDeviceGetHandleByIndex(index)))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons, watts = [], 0, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                watts.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": statistics.median(watts) if watts else None}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_clips_per_sec(n_clips: int, warmup: int = 1):
    """Oracle (the reference's CPU path restated over HF modules), fp32, all host threads, b = 1 per call as the
    reference's extractor does (A/extract_multimodal_features_precompute.py:892)."""
    from oracle import reference_path as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clip = R.build_clip_vision(24, seed=1234)
    hub = R.build_hubert(24, "layer", seed=1234)
    fus = R.build_fusion_state(seed=1234)

    def one(i):
        fh = R.clip_vit_large_forward(clip, R.synth_frames(1, 8, seed=100 + i))
        ah = R.hubert_large_forward(hub, R.synth_audio(1, 8, seed=200 + i))
        return R.fused_tokens(fus, fh, None, ah, "multiframe_audio_frame_text")

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(n_clips):
        one(warmup + i)
    dt = time.perf_counter() - t0
    return n_clips / dt, cores, dt


def clip_tflop(n_frames: int = 8, n_segments: int = 8, faces: bool = False) -> float:
    """Dense algorithmic TFLOP per clip: 1.875 for C2, 5.764 for C4 (16 frames + 16 faces + 8 s audio)."""
    return (n_frames * (2 if faces else 1) * GF_VIT_FRAME + n_segments * GF_HUBERT_SEG + GF_FUSE_CLIP) / 1e3


def workload_config(clips_per_gpu: int, world: int, n_frames: int = 8, n_segments: int = 8, faces: bool = False) -> dict:
    """`config` of the JSON line; shared by both arms (the reference arm measures the same workload on the host)."""
    tf_per_clip = clip_tflop(n_frames, n_segments, faces)
    vit_frames = n_frames * (2 if faces else 1)
    if (n_frames, n_segments, faces) != (8, 8, False):
        return {"workload": f"{clips_per_gpu} clips/GPU x ({n_frames} frames{' + ' + str(n_frames) + ' face crops' if faces else ''} 224^2 + "
                            f"{n_segments}x2s 16kHz audio) -> {'frame/face' if faces else 'frame'}/audio/multi tokens @3584 (C4 shape when 16/16/faces)",
                "clips_per_gpu": clips_per_gpu, "weights": "random-init CLIP ViT-L/14 + HuBERT-large (layer-norm conv) + heads",
                "cache": "inputs and activations exceed the 126 MB L2", "parallelism": f"dp{world}", "tflop_per_clip": tf_per_clip,
                "tflop_per_clip_executed": tf_per_clip - (0.0 if os.environ.get("AGB_VIT_DENSE_LAST") == "1" else vit_frames * GF_VIT_FRAME_PRUNED / 1e3)}
    return {"workload": f"C2: {clips_per_gpu} clips/GPU x (8 frames 224^2 + 8x2s 16kHz audio) -> frame/audio/multi tokens @3584",
            "clips_per_gpu": clips_per_gpu, "weights": "random-init CLIP ViT-L/14 + HuBERT-large (layer-norm conv) + heads",
            "cache": "inputs (373 MB/step) and activations exceed the 126 MB L2", "parallelism": f"dp{world}",
            "tflop_per_clip": tf_per_clip,
            "tflop_per_clip_executed": tf_per_clip - (0.0 if os.environ.get("AGB_VIT_DENSE_LAST") == "1" else 8 * GF_VIT_FRAME_PRUNED / 1e3),
            "last_vit_block": "dense" if os.environ.get("AGB_VIT_DENSE_LAST") == "1" else "CLS row only after K/V (identical pooled output)"}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (oracle port: HF fp32 modules + the restated wrappers, one clip
    per call as the reference's extractor runs it), all host threads, on rank 0 only.  Each "step" is a bounded sample
    of the C2 workload: ONE clip of the 64 (the per-clip cost is batch-independent on the CPU), so that
    `--steps K --warmup W` stays within minutes; value = clips / second over the K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    per_step = 1
    cps, cores, dt = cpu_oracle_clips_per_sec(per_step * max(args.steps, 1), warmup=max(min(args.warmup, 2), 1))
    line = {"impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": max(args.gpus, world),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * per_step / cps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.clips, max(args.gpus, world)),
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} clip of the batch per step, {max(args.steps, 1)} timed steps at batch 1 "
                                       f"(oracle: HF fp32 + restated wrappers), {dt:.1f} s"},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def library_gpu_baseline(B: int, n_frames: int, n_segments: int, dev) -> dict:
    """HF transformers modules (the library the reference calls) on this GPU under bf16 autocast, random-init full-size
    configs, same batch shape; encoders only (> 99.9 % of the work; the reference's per-frame PIL preprocessing is left
    out, which favours this arm).  Uses transformers directly - nothing of this repository and nothing under oracle/."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection, HubertConfig, HubertModel
    torch.manual_seed(1234)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, patch_size=14,
        image_size=224, hidden_act="quick_gelu", projection_dim=768, layer_norm_eps=1e-5)).eval().to(dev)
    hub = HubertModel(HubertConfig(
        hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096, feat_extract_norm="layer",
        do_stable_layer_norm=True, conv_bias=True, num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16,
        conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2), conv_kernel=(10, 3, 3, 3, 3, 2, 2),
        feat_proj_layer_norm=True, layer_norm_eps=1e-5, hidden_act="gelu", feat_extract_activation="gelu")).eval().to(dev)
    px = torch.randn(B * n_frames, 3, 224, 224, device=dev)
    wav = torch.randn(B * n_segments, 32000, device=dev)

    @torch.no_grad()
    def hf_step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f = clip(pixel_values=px).image_embeds
            hs = hub(input_values=wav, output_hidden_states=True).hidden_states
            return f, torch.stack(hs)[-4:].sum(dim=0).mean(dim=1)          # encoder.py:428-430

    for _ in range(2):
        hf_step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 3
    e0.record()
    for _ in range(k):
        hf_step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / k
    del clip, hub, px, wav
    torch.cuda.empty_cache()
    return {"value": B / ms * 1e3, "unit": "clips/s", "ms_per_step": ms, "kind": "HF transformers modules on this GPU, "
            "torch.autocast(bfloat16), torch " + torch.__version__ + " library kernels (cuBLASLt / cuDNN / SDPA)",
            "sample": f"{B} clips per step, 2 warm-up + {k} timed steps, encoders only"}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=64, help="clips per GPU per step (C2: 64)")
    ap.add_argument("--n-frames", type=int, default=8, help="frames per clip (C2: 8; C4: 16)")
    ap.add_argument("--n-segments", type=int, default=8, help="2-second audio segments per clip (C2: 8; C4 with 8 s audio: 16)")
    ap.add_argument("--with-faces", action="store_true", help="three-stream clips: as many face crops as frames (C4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the HF-modules-on-this-GPU comparator (N = 1)")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: shares of the global batch proportional to each GPU's measured speed instead of equal shards "
                         "(parallel.balanced_counts; opt-in, see the comment at the calibration)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel separately instead of replaying a CUDA graph")
    ap.add_argument("--only-resident", action="store_true",
                    help="profiling aid (tools/launch_list.sh): run the warm-up and the device-resident timed steps only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    import affectgpt_b200 as agb
    from affectgpt_b200 import engine as E, parallel as P, synthetic as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    B_nominal = args.clips                       # clips per GPU of the workload; the global batch is B_nominal * world
    B = B_nominal

    NF, NS, FACES = args.n_frames, args.n_segments, args.with_faces
    B_max = B_nominal + (max(B_nominal // 8, 1) if world > 1 and args.balance else 0)            # largest balanced share
    model = agb.build_path(S.clip_state(24, dev), S.hubert_state(24, "layer", dev), S.fusion_state(3584, dev), device=dev,
                           max_frames_per_chunk=max(512, B_max * NF), max_segments_per_chunk=max(512, B_max * NS))
    fof = "multiface_audio_face_frame_text" if FACES else "multiframe_audio_frame_text"
    total_clips = B_nominal * world

    # ---- N > 1: shares of the global batch in proportion to each GPU's measured speed.  The GPUs of a node sit at different
    # points of the 1 kW power cap (SM clocks 1.26-1.42 GHz on one box) and equal shards run at the pace of the slowest.
    # Calibration = the equal-shard step itself, timed per rank on its own stream; the tokens do not depend on which rank
    # computes a clip (checked below: multi_gpu_parity), so only the split changes.  OPT-IN (--balance), because a static
    # calibration is only as good as the speeds are stable: on one 8 x B200 box it gained 2.9 % (equal shards 4017 clips/s,
    # own-kernel times 123.5-127.4 ms; balanced 4134, 121.9-123.8 ms, shares 62-66), on a 4-GPU run the GPU it had measured
    # fastest was 3 % slower a few seconds later and the balanced split lost 1 % (2040 vs ~2060) - the power-cap operating
    # point of a GPU drifts on the scale of seconds.  Calibration = 12 + 16 equal-shard steps per rank.
    counts, calib = None, None
    if world > 1 and args.balance and not args.no_graph:
        g0 = agb.GraphedClipBatch(model, B_nominal, NF, NS, with_faces=FACES, device=dev)
        g0.raw_frames.copy_(S.frames(B_nominal, NF, seed=1000 + rank).to(dev))
        g0.raw_audios.copy_(S.audio(B_nominal, NS, seed=2000 + rank).to(dev))
        if FACES:
            g0.raw_faces.copy_(S.frames(B_nominal, NF, seed=3000 + rank).to(dev))
        for _ in range(12):                      # ~1.5 s: past the boost budget a cold GPU starts with
            g0()
        dist.barrier()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(16):
            g0()
        c1.record()
        torch.cuda.synchronize()
        t_own = torch.tensor([c0.elapsed_time(c1) / 16], device=dev)
        t_all = [torch.empty_like(t_own) for _ in range(world)]
        dist.all_gather(t_all, t_own)
        calib = [float(t) for t in t_all]
        counts = P.balanced_counts(total_clips, [1.0 / t for t in calib], max_shift=max(B_nominal // 8, 1))
        B = counts[rank]
        del g0
        torch.cuda.empty_cache()
    offsets = [sum(counts[:r]) for r in range(world)] if counts is not None else None
    # host (pinned) and device copies of this rank's batch; frames fp32 0..255 as the reference passes them
    h_frames = S.frames(B, NF, seed=1000 + rank).pin_memory()
    h_audio = S.audio(B, NS, seed=2000 + rank).pin_memory()
    h_faces = S.frames(B, NF, seed=3000 + rank).pin_memory() if FACES else None
    d_frames, d_audio = h_frames.to(dev), h_audio.to(dev)
    d_faces = h_faces.to(dev) if FACES else None
    h_tokens = torch.empty((B, 4 if FACES else 3, 3584), dtype=torch.float32).pin_memory()

    n_tok = 4 if FACES else 3
    gather = P.ClipGather(total_clips, (n_tok, 3584), torch.float32, dev, counts=counts) if world > 1 else None

    def finish(out):
        """This rank's fused tokens in prompt order [B, n_tok, 3584]; at N > 1 their all_gather is queued on the gather's
        side stream (it overlaps the next step; `drain` makes the calling stream wait for the last one)."""
        toks = torch.cat([out["multi"], out["audio"]] + ([out["face"]] if FACES else []) + [out["frame"]], dim=1)   # prompt order
        if gather is not None:
            gather.start(toks)
        return toks

    def drain():
        return gather.wait() if gather is not None else None

    def step(frames, audio, faces=None):                                             # eager: one launch per kernel
        smp = {"raw_frames": frames, "raw_audios": audio, "face_or_frame": fof}
        if FACES:
            smp["raw_faces"] = faces
        return finish(model.encode_clip_batch(smp))

    # the public fast path for a fixed batch shape: the whole step replayed from a CUDA graph (GraphedClipBatch);
    # three instances = the device-resident one + the two upload slots of the end-to-end arm
    use_graph = not args.no_graph
    graphed = [agb.GraphedClipBatch(model, B, NF, NS, with_faces=FACES, device=dev) for _ in range(3)] if use_graph else None
    if use_graph:
        graphed[2].raw_frames.copy_(d_frames)
        graphed[2].raw_audios.copy_(d_audio)
        if FACES:
            graphed[2].raw_faces.copy_(d_faces)

    def step_resident():
        return finish(graphed[2]()) if use_graph else step(d_frames, d_audio, d_faces)

    # end-to-end: every step's inputs come from pinned host memory and its tokens go back to the host inside the timed
    # region; the NEXT step's H2D runs on a copy stream while this step computes (two device input slots), which is
    # what a sweep with a prefetching loader does.  K timed steps = K uploads + K downloads.
    copy_stream = torch.cuda.Stream(device=dev)
    slots = ([(g.raw_frames, g.raw_audios, g.raw_faces) for g in graphed[:2]] if use_graph
             else [(torch.empty_like(d_frames), torch.empty_like(d_audio), torch.empty_like(d_faces) if FACES else None)
                   for _ in range(2)])
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        sl = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[sl])                 # the step that last read this slot has finished
            slots[sl][0].copy_(h_frames, non_blocking=True)
            slots[sl][1].copy_(h_audio, non_blocking=True)
            if FACES:
                slots[sl][2].copy_(h_faces, non_blocking=True)
            ev_copied[sl].record(copy_stream)

    def run_e2e(k):
        cur = torch.cuda.current_stream(dev)
        upload(0)
        for i in range(k):
            if i + 1 < k:
                upload(i + 1)
            sl = i % 2
            cur.wait_event(ev_copied[sl])
            toks = finish(graphed[sl]()) if use_graph else step(slots[sl][0], slots[sl][1], slots[sl][2])
            ev_consumed[sl].record(cur)
            h_tokens.copy_(toks, non_blocking=True)
        drain()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    own_ms = {}

    def timed(fn, k, tag):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        ec = torch.cuda.Event(enable_timing=True)
        ec.record()                              # this rank's own kernels are done here (the gathers run on a side stream)
        drain()                                  # the last step's gather is part of the timed region
        e1.record()
        barrier()
        own_ms[tag] = e0.elapsed_time(e1)
        own_ms[tag + "_compute"] = e0.elapsed_time(ec)
        ms = torch.tensor([own_ms[tag]], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step_resident()
    drain()
    clocks = ClockSampler(local)
    clocks.start()                               # every rank samples its own GPU (per_rank in the JSON line)
    l0 = E.launch_count()
    j0, tj0 = gpu_energy_mj(local), time.perf_counter()
    ms_dev = timed(step_resident, args.steps, "resident")
    j1, tj1 = gpu_energy_mj(local), time.perf_counter()
    joules_per_step = (j1 - j0) / 1e3 / args.steps if (j0 is not None and j1 is not None and j1 > j0) else None
    launches = (graphed[2].kernels_per_replay * args.steps) if use_graph else (E.launch_count() - l0)
    if args.only_resident:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": total_clips * args.steps / (ms_dev / 1e3), "unit": "clips/s",
                              "ms_per_step": ms_dev / args.steps, "note": "--only-resident: no e2e / roofline legs"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    run_e2e(2)
    ms_e2e = timed(lambda: run_e2e(args.steps), 1, "e2e")
    clk = clocks.stop()

    # ---- W GPUs == 1 GPU, bit for bit: recompute two clips of the NEXT rank's shard here, eagerly, at batch 2,
    # and compare them with the rows the gather delivered (clip j of rank o is global row j * W + o)
    parity, per_rank = None, None
    if world > 1:
        last = step_resident()
        gathered = drain().clone()
        other = (rank + 1) % world
        B_other = counts[other] if counts is not None else B
        pick = [0, B_other - 1]
        o_frames = S.frames(B_other, NF, seed=1000 + other)[pick].to(dev)
        o_audio = S.audio(B_other, NS, seed=2000 + other)[pick].to(dev)
        o_faces = S.frames(B_other, NF, seed=3000 + other)[pick].to(dev) if FACES else None
        smp = {"raw_frames": o_frames, "raw_audios": o_audio, "face_or_frame": fof}
        if FACES:
            smp["raw_faces"] = o_faces
        o = model.encode_clip_batch(smp)
        redo = torch.cat([o["multi"], o["audio"]] + ([o["face"]] if FACES else []) + [o["frame"]], dim=1)
        rows_of = (lambda r, js: [offsets[r] + j for j in js]) if counts is not None else (lambda r, js: [j * world + r for j in js])
        same = torch.equal(redo, gathered[rows_of(other, pick)])
        mine = torch.equal(gathered[rows_of(rank, range(B))], last)
        flag = torch.tensor([1 if (same and mine) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = bool(int(flag))
        rows = torch.tensor([own_ms["resident"] / args.steps, own_ms["e2e"] / args.steps, float((clk or {}).get("sm_mhz") or 0.0),
                             own_ms["resident_compute"] / args.steps], device=dev)
        allr = [torch.empty_like(rows) for _ in range(world)]
        dist.all_gather(allr, rows)
        per_rank = {"ms_per_step": [round(float(r[0]), 3) for r in allr],
                    "own_kernels_ms_per_step": [round(float(r[3]), 3) for r in allr],      # before waiting for the last gather
                    "e2e_ms_per_step": [round(float(r[1]), 3) for r in allr],
                    "sm_mhz": [float(r[2]) for r in allr], "slowest_rank": int(max(range(world), key=lambda i: float(allr[i][3])))}

    # instrumented pass: CUDA-event pairs around every kernel wrapper (separate from the timed value)
    barrier()
    overlap = model.overlap_encoders
    model.overlap_encoders = False          # one stream: every kernel's event pair then brackets that kernel alone
    E.profile_begin()
    for _ in range(args.steps):
        step(d_frames, d_audio, d_faces)
    prof = E.profile_end()
    drain()
    model.overlap_encoders = overlap

    value = total_clips * args.steps / (ms_dev / 1e3)
    e2e_value = total_clips * args.steps / (ms_e2e / 1e3)
    peaks = read_peaks()
    g = prof["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    tf_per_clip = clip_tflop(NF, NS, FACES)
    cfg = workload_config(B_nominal, world, NF, NS, FACES)
    step_tflops = value / world * cfg["tflop_per_clip_executed"]       # FLOPs that actually ran (pruned last ViT block)
    step_tflops_dense = value / world * tf_per_clip                     # SURVEY 8d dense figure, for comparison
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": "clips/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int((h_frames.numel() + h_audio.numel() + (h_faces.numel() if FACES else 0)) * 4),
                "d2h_bytes_per_step": int(h_tokens.numel() * 4)},
        "gpu_launches": int(launches), "launch_mode": "cuda-graph replay (GraphedClipBatch)" if use_graph else "eager",
        "encoder_streams": 2 if overlap else 1,
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["bf16_sustained"], "traffic": gemm_traffic_per_launch(),
                     "traffic_unit": "bytes per launch (dram read+write, mean over the GEMM launches of one step; ncu, profiles/)",
                     "flops_per_launch": g["flops"] / max(g["launches"], 1), "ms_per_launch": g["ms"] / max(g["launches"], 1),
                     "kernel": "gemm_kernel (tcgen05, all GEMM launches)",
                     "peak_source": peaks["source"] + ", sustained", "gemm_launches": g["launches"], "gemm_ms_per_step": g["ms"] / args.steps,
                     "timing_pass": "eager instrumented pass (one CUDA-event pair per kernel, same stream): categories sum to "
                                    f"{tot_ms / args.steps:.2f} ms per step vs {ms_dev / args.steps:.2f} ms graph-replayed",
                     "whole_step_tflops_per_gpu": step_tflops, "whole_step_frac": step_tflops / peaks["bf16_sustained"],
                     "whole_step_frac_basis": "executed FLOPs (tflop_per_clip_executed) / graph-replayed step time / sustained peak",
                     "whole_step_frac_dense": step_tflops_dense / peaks["bf16_sustained"],
                     "time_share": {k: v["ms"] / tot_ms for k, v in prof.items()},
                     "category_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
                     "attention_tflops": (prof["attention"]["flops"] / (prof["attention"]["ms"] / 1e3) / 1e12) if prof["attention"]["ms"] > 0 else 0.0,
                     "layernorm_gbs": (prof["layernorm"]["bytes"] / (prof["layernorm"]["ms"] / 1e3) / 1e9) if prof["layernorm"]["ms"] > 0 else 0.0,
                     "frontend_gbs": (prof["frontend"]["bytes"] / (prof["frontend"]["ms"] / 1e3) / 1e9) if prof["frontend"]["ms"] > 0 else 0.0},
        "clocks": clk,
        "energy": {"joules_per_step": joules_per_step, "joules_per_clip": (joules_per_step / B if joules_per_step else None),
                   "avg_w": (joules_per_step * args.steps / (tj1 - tj0) if joules_per_step else None),
                   "wall_s_between_reads": tj1 - tj0,
                   "source": "NVML total-energy counter of this rank's GPU read on both sides of the timed device-resident "
                             "steps (wall clock between the reads includes the two barriers; the counter is coarse for few steps)"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline and (NF, NS, FACES) == (8, 8, False):   # the CPU sample is C2-shaped
        cps, cores, dt = cpu_oracle_clips_per_sec(args.cpu_clips, warmup=1)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_clips} clips at batch 1 (oracle: HF fp32 + restated wrappers), 1 warm-up, {dt:.1f} s"}
    if world > 1:
        line["multi_gpu_parity"] = parity
        line["multi_gpu_parity_check"] = "each rank recomputed 2 clips of the next rank's shard at batch 2 (eager) and compared them, and its own rows, with the all_gathered tokens: torch.equal"
        line["per_rank"] = per_rank
        line["gather"] = "all_gather_into_tensor on a side stream (parallel.ClipGather), one per step, overlapped with the next step"
        line["balance"] = ({"clips_per_rank": counts, "equal_shard_ms_per_step": [round(t, 3) for t in calib],
                            "rule": "global batch of clips_per_gpu x N split in proportion to each GPU's measured equal-shard speed "
                                    "(parallel.balanced_counts, +-1/8 of a shard at most; --balance)"}
                           if counts is not None else {"clips_per_rank": [B_nominal] * world, "rule": "equal shards (default; --balance splits by measured speed)"})
    if rank == 0 and world == 1 and not args.no_library_baseline:
        try:
            line["library_gpu_baseline"] = library_gpu_baseline(B, NF, NS, dev)
            line["library_gpu_baseline"]["speedup_of_this_repo"] = value / line["library_gpu_baseline"]["value"]
        except Exception as e:                                  # a comparator must never take the bench line down
            line["library_gpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
