#!/usr/bin/env python
"""Role profile of the decoupled attention forward kernel IN THE STEP (the clock of a power-capped step, not ncu's isolated launch):
tim_debug_role_prof makes every launch add the cycles one warp of each role spent in its waits to a device buffer; this runs the
bench workload's forward a few times with the hook on and prints the per-tile averages.

    python tools/attn_roles.py [--workload cfg2] [--clips 1024] [--steps 5]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200 import _lib                                   # noqa: E402
from tim_b200.config import named_config                    # noqa: E402
from tim_b200.plugin import TIMEngine                       # noqa: E402
from tim_b200.synth import synth_inputs, synth_state_dict   # noqa: E402

NAMES = ["cta_total(producer)", "producer wait K_f free", "producer wait stage free (o_full)", "producer wait V_f free", "mma idle (nothing ready)",
         "mma total", "softmax wait s_full", "softmax total", "epilogue wait slab read", "epilogue wait p_full", "epilogue wait o_full",
         "epilogue total", "tiles", "epilogue wait tmem ld"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    cfg, Qv, Qa = named_config(args.workload)
    eng = TIMEngine(cfg, 0, "fp16")
    eng.load_state_dict(synth_state_dict(cfg, 0, "trained"))
    inp = synth_inputs(cfg, 1, Qv, Qa, 1234, shared_queries=cfg.variant == "detection")
    B = args.clips
    g = torch.Generator(device=dev).manual_seed(1)
    vis = torch.randn((B, cfg.num_feats, cfg.visual_input_dim), generator=g, device=dev)
    aud = torch.randn((B, cfg.num_feats, cfg.audio_input_dim), generator=g, device=dev)
    times = torch.from_numpy(inp["times"]).to(dev).repeat(B, 1, 1).contiguous()
    for _ in range(10):                                   # bring the step to its power-capped clock
        eng.encoder(vis, aud, eng.time_mlp(times), Qv, Qa)
    buf = torch.zeros((256, 16), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    lib.tim_debug_role_prof(C.c_void_p(buf.data_ptr()))
    for _ in range(args.steps):
        eng.encoder(vis, aud, eng.time_mlp(times), Qv, Qa)
    torch.cuda.synchronize()
    lib.tim_debug_role_prof(None)
    a = buf.cpu().numpy().astype(np.float64)
    a = a[a[:, 12] > 0]
    launches = args.steps * cfg.num_layers
    tiles = a[:, 12].sum() / launches                     # per launch, all CTAs
    print(f"{args.workload} B={B}: {len(a)} CTAs, {tiles / len(a):.1f} tiles per CTA and launch, {launches} launches")
    per_tile = a.sum(0) / a[:, 12].sum()
    for i, n in enumerate(NAMES):
        if i != 12:
            print(f"  {n:36s} {per_tile[i]:9.0f} cycles per tile")
    print(f"  softmax busy  {per_tile[7] - per_tile[6]:9.0f}   epilogue busy {per_tile[11] - per_tile[8] - per_tile[9] - per_tile[10]:9.0f}  (cycles per tile, total minus waits)")


if __name__ == "__main__":
    main()
