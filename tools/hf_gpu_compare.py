"""On-box GPU comparison: "the reference's kernels on this box".

The reference's arithmetic is HuggingFace ``transformers`` (CLIPVisionModelWithProjection / HubertModel) called from
A/my_affectgpt/models/encoder.py:184-205, 403-430.  This script runs those same library modules on the GPU, random-init
weights of the full configuration, on one C2 batch (64 clips = 512 frames + 512 two-second segments), in three settings:

  fp32            what the reference's extractor does (torch.no_grad, fp32, no TF32; extract_...precompute.py:221, 581)
  autocast-fp16   what AffectGPT.forward does (maybe_autocast, blip2.py:36-49)
  autocast-bf16   the dtype this repository computes in

and prints clips/s for each, next to this repository's path on the same batch.  Only the encoders are timed for the
library arm (they are > 99.9 % of the work; the reference's per-frame PIL/CPU preprocessing is left out, which favours
the library arm).  Stand-alone on purpose: it uses transformers directly and nothing under oracle/.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

dev = torch.device("cuda:0")
B = int(os.environ.get("CLIPS", "64"))


def build_hf():
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
    return clip, hub


def timeit(fn, warm=2, k=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    res = {"clips": B, "gpu": torch.cuda.get_device_name(0)}
    clip, hub = build_hf()
    px = torch.randn(B * 8, 3, 224, 224, device=dev)
    wav = torch.randn(B * 8, 32000, device=dev)

    @torch.no_grad()
    def hf_step():
        f = clip(pixel_values=px).image_embeds
        hs = hub(input_values=wav, output_hidden_states=True).hidden_states
        a = torch.stack(hs)[-4:].sum(dim=0).mean(dim=1)          # encoder.py:428-430 (last-4 mean, time mean)
        return f, a

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ms = timeit(hf_step, warm=1, k=2)
    res["hf_fp32"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
    print(f"HF modules fp32 (no TF32)  : {ms:9.1f} ms / {B} clips = {B / ms * 1e3:7.1f} clips/s", flush=True)
    for name, dt in (("hf_autocast_fp16", torch.float16), ("hf_autocast_bf16", torch.bfloat16)):
        def ac():
            with torch.autocast("cuda", dtype=dt):
                return hf_step()
        ms = timeit(ac)
        res[name] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
        print(f"HF modules {name[3:]:17s}: {ms:9.1f} ms / {B} clips = {B / ms * 1e3:7.1f} clips/s", flush=True)
    del clip, hub
    torch.cuda.empty_cache()

    import affectgpt_b200 as agb
    from affectgpt_b200 import synthetic as S
    model = agb.build_path(S.clip_state(24, dev), S.hubert_state(24, "layer", dev), S.fusion_state(3584, dev), device=dev)
    g = agb.GraphedClipBatch(model, B, 8, 8, device=dev)
    g.raw_frames.copy_(S.frames(B, 8, seed=1).to(dev))
    g.raw_audios.copy_(S.audio(B, 8, seed=2).to(dev))
    ms = timeit(lambda: g(), warm=3, k=5)
    res["affectgpt_b200"] = {"ms_per_step": ms, "clips_per_s": B / ms * 1e3}
    print(f"affectgpt_b200 (bf16, graph): {ms:9.1f} ms / {B} clips = {B / ms * 1e3:7.1f} clips/s", flush=True)
    for k in ("hf_fp32", "hf_autocast_fp16", "hf_autocast_bf16"):
        res[k]["speedup_of_this_repo"] = res[k]["ms_per_step"] / ms
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(res, open(os.path.join(out, "hf_gpu_compare.json"), "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
