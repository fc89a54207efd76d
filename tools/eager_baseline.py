#!/usr/bin/env python
"""Secondary baseline of BASELINE.md §4: the reference forward through eager PyTorch ON THE B200 - the "library call" GPU path the
hand-written kernels have to beat. The reference's package cannot travel to the GPU box, so this runs oracle/tim_oracle_torch.py
(the reference's own PyTorch calls restated op for op: dense S x S masked attention through F.multi_head_attention_forward, the
materialised [B*H, S, S] mask, F.linear / F.layer_norm / F.gelu) on CUDA tensors, device-resident inputs, CUDA-event timing:

    python tools/eager_baseline.py [--workload cfg2] [--clips 256] [--steps 10] [--out gpurun_out/eager_baseline.json]

Modes: fp32 with TF32 off (the reference's eval precision), fp32 with TF32 on, bf16 autocast (its training precision). Reports
ms/step and clips x queries / s next to the library's own number for the same clips per step. Measurement tool only; written at the
end of round 1 when no GPU minutes were left - its first run is item 2 of DESIGN.md §10.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.tim_oracle_torch import TIMOracleTorch   # noqa: E402
from tim_b200.config import named_config   # noqa: E402
from tim_b200.plugin import TIMEngine   # noqa: E402
from tim_b200.synth import rel_l2, synth_inputs, synth_state_dict   # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--clips", type=int, default=256, help="clips per step (the dense [B*H, S, S] mask and scores bound it)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg, Qv, Qa = named_config(args.workload)
    sd = synth_state_dict(cfg, 0, "trained")
    inp = synth_inputs(cfg, args.clips, Qv, Qa, 1234, shared_queries=cfg.variant == "detection")
    t = {k: torch.from_numpy(v).to(dev) for k, v in inp.items()}
    ref = TIMOracleTorch(cfg, sd, device=dev)
    res = {"workload": args.workload, "clips_per_step": args.clips, "queries_per_clip": Qv + Qa, "torch": torch.__version__, "modes": {}}
    base = None
    for mode in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "fp32_tf32_on"
        torch.backends.cudnn.allow_tf32 = mode == "fp32_tf32_on"

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16_autocast"):
                return ref.forward_tensors(t.get("vis"), t.get("aud"), t["times"], Qv, Qa)
        ms = timed(step, args.steps)
        out = step()
        key = "action" if out.get("action") is not None else "audio"
        y = out[key].float().cpu().numpy()
        base = y if base is None else base
        res["modes"][mode] = {"ms_per_step": ms, "clips_x_queries_per_sec": args.clips * (Qv + Qa) / (ms * 1e-3),
                              "rel_l2_vs_fp32": rel_l2(y, base)}
    eng = TIMEngine(cfg, 0, "fp16")
    eng.load_state_dict(sd)

    def lib_step():
        return eng.encoder(t.get("vis"), t.get("aud"), eng.time_mlp(t["times"]), Qv, Qa)
    ms = timed(lib_step, args.steps)
    out = lib_step()
    key = "action" if out.get("action") is not None else "audio"
    res["tim_b200_fp16"] = {"ms_per_step": ms, "clips_x_queries_per_sec": args.clips * (Qv + Qa) / (ms * 1e-3),
                            "rel_l2_vs_fp32": rel_l2(out[key].float().cpu().numpy(), base)}
    eng.close()
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
