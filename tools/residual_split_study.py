#!/usr/bin/env python
"""How much precision does a two-plane 16-bit residual stream cost? (DESIGN.md §8, next-round idea: store the pre-LayerNorm rows z as
hi = fp16(z), lo = fp16(z - hi) instead of fp32 + a separate fp16 copy.) CPU study on the numpy oracle, no GPU:

    python tools/residual_split_study.py [--clips 2]

Runs the cfg2 forward in fp32 with z1 / z2 of every layer passed through the given storage format and reports the rel-L2 error of the
logits against the plain fp32 forward: fp16 hi + fp16 lo, bf16 hi + bf16 lo, and (for scale) a single fp16 / bf16 plane.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.tim_oracle as to   # noqa: E402
from tim_b200.config import named_config   # noqa: E402
from tim_b200.synth import rel_l2, synth_inputs, synth_state_dict   # noqa: E402


def bf16(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7fff)
    return ((u + r) & np.uint32(0xffff0000)).view(np.float32)


FORMATS = {
    "fp32": lambda z: z,
    "fp16 hi + fp16 lo": lambda z: (lambda h: h + (z - h).astype(np.float16).astype(np.float32))(z.astype(np.float16).astype(np.float32)),
    "bf16 hi + bf16 lo": lambda z: (lambda h: h + bf16(z - h))(bf16(z)),
    "fp16 only": lambda z: z.astype(np.float16).astype(np.float32),
    "bf16 only": bf16,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2)
    args = ap.parse_args()
    cfg, Qv, Qa = named_config("cfg2")
    sd = synth_state_dict(cfg, 0, "trained")
    inp = synth_inputs(cfg, args.clips, Qv, Qa, 1234)
    plain_ln = to._layer_norm
    results = {}
    for name, fmt in FORMATS.items():
        def ln(x, w, b, eps=1e-5, _fmt=fmt):
            # every LayerNorm of the encoder stack normalises a residual-stream row z: store it in the format first
            return plain_ln(_fmt(x.astype(np.float32)) if x.shape[-1] == cfg.d_model * 2 else x, w, b, eps)
        to._layer_norm = ln
        try:
            o = to.TIMOracle(cfg, sd, np.float32).forward(inp["vis"], inp["aud"], inp["times"], Qv, Qa)
        finally:
            to._layer_norm = plain_ln
        results[name] = o
    base = results["fp32"]
    for name, o in results.items():
        if name == "fp32":
            continue
        worst = max(rel_l2(o[k], base[k]) for k in ("verb", "noun", "action", "audio", "feats"))
        print(f"{name:20s} worst rel-L2 of the outputs vs fp32 residual stream: {worst:.2e}")


if __name__ == "__main__":
    main()
