#!/usr/bin/env python
"""The UNMODIFIED reference TIM through eager PyTorch on this GPU - the "library call" path the hand-written kernels have to beat
(BASELINE.md §4 secondary baseline). Imports the reference from baseline/_ref (staged by tools/stage_reference.py)
or /root/reference; measurement tool only, nothing under tim_b200/ uses it.

    python tools/eager_reference.py [--workload cfg2] [--clips 512] [--steps 5]

Modes: fp32 with TF32 off (the reference's eval precision), fp32 with TF32 on, bf16 autocast and fp16 autocast (its training
precision, recognition/scripts/train.py:197). Device-resident inputs, CUDA events, model.eval() under no_grad. Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200.config import named_config   # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402
from tools.refload import build_reference   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--clips", type=int, default=512)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg, Qv, Qa = named_config(args.workload)
    model = build_reference(cfg)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in synth_state_dict(cfg, 0, "trained").items()}, strict=True)
    model = model.to(dev).eval()
    B = args.clips
    g = torch.Generator(device=dev).manual_seed(1234)
    F = cfg.num_feats
    vis = torch.randn((B, F, cfg.visual_input_dim), generator=g, device=dev) if cfg.has_visual_input else None
    aud = torch.randn((B, F, cfg.audio_input_dim), generator=g, device=dev) if cfg.has_audio_input else None
    t1 = torch.from_numpy(synth_inputs(cfg, 1, Qv, Qa, 1234, shared_queries=cfg.variant == "detection")["times"]).to(dev)
    times = t1.repeat(B, 1, 1).contiguous()
    if cfg.variant == "detection":
        model.inference_queries = times[0:1, cfg.F_tot:cfg.F_tot + max(Qv, Qa)].clone()
        model.num_queries = max(Qv, Qa)

    def fwd():
        if cfg.variant == "recognition":
            te = model(times, "time_mlp")
            return model([vis, aud], "encoder", te, Qv, Qa)[0][2 if Qv else 3]
        return model([vis, aud], "encoder", times[:, :cfg.F_tot], None, False)[0][0][2 if Qv else 3]

    res = {"workload": args.workload, "clips_per_step": B, "queries_per_clip": Qv + Qa, "torch": torch.__version__,
           "what": "unmodified reference TIM, eager PyTorch, model.eval() under no_grad, device-resident inputs, CUDA events", "modes": {}}
    base = None
    for mode in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast", "fp16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "fp32_tf32_on"
        torch.backends.cudnn.allow_tf32 = mode == "fp32_tf32_on"
        ac = {"bf16_autocast": torch.bfloat16, "fp16_autocast": torch.float16}.get(mode)

        def step():
            with torch.no_grad(), torch.autocast("cuda", dtype=ac or torch.bfloat16, enabled=ac is not None):
                return fwd()
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                y = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            y = y.float()
            base = y if base is None else base
            res["modes"][mode] = {"ms_per_step": ms, "clips_x_queries_per_sec": B * (Qv + Qa) / (ms * 1e-3),
                                  "rel_l2_vs_fp32": float((y - base).norm() / base.norm())}
        except RuntimeError as e:          # e.g. out of memory for the dense [B*H, S, S] mask at this batch
            res["modes"][mode] = {"error": str(e).splitlines()[0][:200]}
            torch.cuda.empty_cache()
    res["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
