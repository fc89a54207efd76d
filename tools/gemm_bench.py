#!/usr/bin/env python
"""Isolated timing + correctness of the linear-layer kernels (C-ABI hook tim_bench_linear) on a B200.

    python tools/gemm_bench.py [--m 204800] [--versions 1,2] [--dtype fp16]
Shapes are the dense contractions of one TIM encoder layer at E = 1024 / FF = 2048 (cfg2, cfg4) with their epilogues:
in_proj (16-bit out), out_proj (+fp32 residual, fp32 out), linear1 (GELU, 16-bit out), linear2 (+residual, fp32 out).
Each launch is checked on a sample of rows against torch fp32 matmul of the same 16-bit operands.
"""
import argparse
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200 import _lib   # noqa: E402

DT = {"bf16": 1, "fp16": 2}
TDT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=204800)
    ap.add_argument("--versions", default="1,2")
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--shapes", default="layer")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    M = args.m
    if args.shapes == "layer":
        shapes = [("in_proj", M, 3072, 1024, 0, False, False), ("out_proj", M, 1024, 1024, 0, True, True),
                  ("linear1", M, 2048, 1024, 2, False, False), ("linear2", M, 1024, 2048, 0, True, True),
                  ("plain32", M, 1024, 1024, 0, False, True), ("relu16", M, 512, 512, 1, False, False)]
    else:   # edge shapes: ragged M, N = 64-multiples, K tails
        shapes = [("m1", 1, 256, 64, 0, False, False), ("m129", 129, 320, 72, 2, False, False), ("m300", 300, 1024, 1024, 0, True, True),
                  ("m257", 257, 128, 2304, 1, False, True), ("m1000", 1000, 3072, 1024, 0, False, False),
                  ("m5000", 5000, 1536, 1536, 2, True, True), ("m20000", 20000, 4608, 1536, 0, False, False),
                  ("m777", 777, 192, 200, 0, False, True)]
    tdt = TDT[args.dtype]
    g = torch.Generator(device=dev).manual_seed(0)
    for name, m, N, K, act, use_res, out32 in shapes:
        A = torch.randn(m, K, generator=g, device=dev).to(tdt)
        W = (torch.randn(N, K, generator=g, device=dev) / math.sqrt(K)).to(tdt)
        bias = torch.randn(N, generator=g, device=dev)
        resid = torch.randn(m, N, generator=g, device=dev) if use_res else None
        rows = torch.unique(torch.cat([torch.arange(min(m, 384)), torch.arange(max(0, m - 384), m),
                                       torch.randint(0, m, (384,))])).to(dev)
        ref = A[rows].float() @ W.float().T + bias
        if act == 1:
            ref = torch.relu(ref)
        elif act == 2:
            ref = torch.nn.functional.gelu(ref)
        if use_res:
            ref = ref + resid[rows]
        for v in [int(x) for x in args.versions.split(",")]:
            out = torch.full((m, N), float("nan"), device=dev, dtype=torch.float32 if out32 else tdt)
            ms = C.c_float(0)
            r = lib.tim_bench_linear(DT[args.dtype], ptr(A), ptr(W), ptr(bias), ptr(resid), ptr(out), m, N, K, act,
                                     int(out32), v, args.iters, C.byref(ms))
            if r != 0:
                print(f"{name:9s} v{v}: status {r} {lib.tim_last_error(None)}", flush=True)
                continue
            torch.cuda.synchronize()
            got = out[rows].float()
            err = float((got - ref).norm() / ref.norm())
            nan = int(torch.isnan(out.float()).sum())
            tf = 2.0 * m * N * K / (ms.value * 1e-3) / 1e12
            print(f"{name:9s} v{v} M={m:6d} N={N:4d} K={K:4d} act={act} res={int(use_res)} out32={int(out32)}: "
                  f"{ms.value:8.4f} ms  {tf:7.1f} TFLOP/s  rel_l2(sample)={err:.2e} nan={nan}", flush=True)


if __name__ == "__main__":
    main()
