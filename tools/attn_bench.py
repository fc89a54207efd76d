#!/usr/bin/env python
"""Isolated timing of the attention kernels (C-ABI hook tim_bench_attention) on a B200.

    python tools/attn_bench.py [--versions 1,2,3] [--dtype fp16] [--shapes cfg2,cfg3,cfg4]
Reports ms per launch and achieved HBM GB/s on the algorithmic bytes (qkv read once + out written once = 8 * E bytes/row),
and checks version 2 against version 1 on the same input.
"""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200 import _lib   # noqa: E402

DT = {"bf16": 1, "fp16": 2}
TDT = {"bf16": torch.bfloat16, "fp16": torch.float16}
SHAPES = {  # name: (B, Ft, Qt, H, hd)
    "cfg2": (1024, 100, 100, 8, 128), "cfg3": (256, 128, 400, 8, 192), "cfg4": (96, 100, 2048, 8, 128),
    "small": (64, 100, 100, 8, 128), "hd64": (512, 100, 100, 8, 64),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--versions", default="1,2,4", help="1 warp-MMA, 2 tcgen05 (r01 form), 4 tcgen05 decoupled pipeline")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--shapes", default="cfg2,cfg3,cfg4")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    for name in args.shapes.split(","):
        B, Ft, Qt, H, hd = SHAPES[name]
        E, M = H * hd, B * (Ft + Qt)
        qkv = torch.randn(M, 3 * E, generator=g, device=dev)
        qkv[:, :E] *= hd ** -0.5 * 1.4426950408889634 * 2.0
        qkv = qkv.to(TDT[args.dtype]).contiguous()
        outs = {}
        for v in [int(x) for x in args.versions.split(",")]:
            out = torch.zeros(M, E, dtype=TDT[args.dtype], device=dev)
            ms = C.c_float(0)
            r = lib.tim_bench_attention(DT[args.dtype], C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), B, Ft, Qt, H, hd, v,
                                        args.iters, C.byref(ms))
            if r != 0:
                print(f"{name} v{v}: ERROR {lib.tim_last_error(None)}")
                continue
            torch.cuda.synchronize()
            gbs = M * 8 * E / (ms.value * 1e-3) / 1e9
            outs[v] = out
            print(f"{name:6s} v{v} B={B} Ft={Ft} Qt={Qt} H={H} hd={hd}: {ms.value * 1e3:9.1f} us  {gbs:8.1f} GB/s algorithmic", flush=True)
        for v in (2, 4):
            if 1 in outs and v in outs:
                d = (outs[1].float() - outs[v].float()).norm() / outs[1].float().norm()
                print(f"{name:6s} v{v} vs v1 rel-L2 {d.item():.2e}")


if __name__ == "__main__":
    main()
