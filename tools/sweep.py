#!/usr/bin/env python
"""BASELINE.json configs[4]: token-length sweep S = 128 .. 4096 (d=512, 6 layers, detection-style as cfg4) on one B200.

    python tools/sweep.py [--tokens 200000] [--steps 10] [--out gpurun_out/sweep.json]
For every S: F_tot = 100 feature tokens (50 for S = 128), Q = S - F_tot interval queries (visual data modality, as cfg4),
clips per step B = tokens // S (the same number of token rows at every S, so the points are comparable; --fill raises B to
the largest power of two whose workspace stays under --hbm-frac of the GPU memory instead). Reports ms/step,
clips x queries / s, tokens/s, algorithmic TFLOP/s (SURVEY.md §8d formula, mask-aware attention) and its fraction of the
measured sustained bf16 peak, plus the GEMM-class TFLOP/s from the library's live per-class timing.
Under torchrun every rank runs its own shard (weak scaling) and rank 0 reports the whole-job numbers (max time over ranks).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200.config import DETECTION, TIMConfig   # noqa: E402
from tim_b200.plugin import TIMEngine   # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=200000, help="token rows per GPU per step (B = tokens // S)")
    ap.add_argument("--fill", action="store_true", help="size B to fill --hbm-frac of the GPU memory instead")
    ap.add_argument("--hbm-frac", type=float, default=0.5)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--seqs", default="128,256,512,1024,2048,4096")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak = 1400.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["bf16_tflops_sustained"])
    rows = []
    for S in [int(x) for x in args.seqs.split(",")]:
        F = 25 if S == 128 else 50
        Q = S - 2 * F
        cfg = TIMConfig(num_class=[97, 44], visual_input_dim=2048, num_layers=6, num_feats=F, data_modality="visual",
                        include_verb_noun=False, variant=DETECTION)
        bytes_per_token = 20.6e3 + 4.0 * (2048 + 2304) * (2 * F) / S * 1.5       # workspace + inputs (fp32 + 16-bit copy)
        if args.fill:
            free, total = torch.cuda.mem_get_info()
            B = 1
            while 2 * B * S * bytes_per_token <= args.hbm_frac * total:
                B *= 2
        else:
            B = max(1, args.tokens // S)
        eng = TIMEngine(cfg, local, args.dtype)
        eng.load_state_dict(synth_state_dict(cfg, 0, "trained"))
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        vis = torch.randn((B, F, cfg.visual_input_dim), generator=g, device=dev)
        aud = torch.randn((B, F, cfg.audio_input_dim), generator=g, device=dev)
        t1 = torch.from_numpy(synth_inputs(cfg, 1, Q, 0, 1234, shared_queries=True)["times"]).to(dev)
        times = t1.repeat(B, 1, 1).contiguous()

        def step():
            return eng.encoder(vis, aud, eng.time_mlp(times), Q, 0)

        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        eng.profile_begin()
        for _ in range(3):
            step()
        prof = eng.profile_end()
        gemm_tf = prof["gemm"]["flops"] / (prof["gemm"]["ms"] * 1e-3) / 1e12
        flops = cfg.flops_fwd_per_clip(Q, 0) * B * world
        row = {"S": S, "F_tot": 2 * F, "queries": Q, "clips_per_gpu": B, "n_gpus": world, "ms_per_step": ms,
               "clips_x_queries_per_sec": world * B * Q / (ms * 1e-3), "tokens_per_sec": world * B * S / (ms * 1e-3),
               "algorithmic_tflops": flops / (ms * 1e-3) / 1e12, "frac_of_sustained_peak": flops / (ms * 1e-3) / 1e12 / (peak * world),
               "gemm_class_tflops": gemm_tf, "class_ms": {k: v["ms"] / 3 for k, v in prof.items() if not k.startswith("gemm_")},
               "workspace_gb": eng.workspace_bytes / 1e9}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        eng.close()
        del vis, aud, times
        torch.cuda.empty_cache()
    if rank == 0 and args.out:
        json.dump({"peak_tflops_sustained": peak, "dtype": args.dtype, "rows": rows}, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
