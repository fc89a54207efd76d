#!/usr/bin/env python
"""Is the pair GEMM at the power-limited ceiling of the part? The same layer shape (in_proj: M x 3072 x 1024; linear1: M x 2048 x 1024)
run back to back for several seconds - long enough for the power cap to settle the SM clock - through (a) cuBLAS (torch.matmul) with
fp16 operands, (b) cuBLAS with bf16 operands (what MEASURED_PEAKS.json's sustained figure is), (c) this library's kernel (C-ABI hook
tim_bench_linear, fp16 and bf16). Prints TFLOP/s over the whole window and the median SM clock / power seen during it.

    python tools/sustained_gemm.py [--m 204800] [--seconds 4] [--out gpurun_out/sustained_gemm.json]
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200 import _lib   # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.rows = []

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.2)


def median(v):
    v = sorted(v)
    return v[len(v) // 2] if v else None


def timed(fn, seconds):
    """fn(n) enqueues n back-to-back launches; returns (ms per launch, clock, power) over a window of about `seconds`."""
    fn(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(20); torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / 20
    n = max(20, int(seconds / per))
    smp = Sampler(); smp.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    smp.stop = True; smp.join()
    rows = smp.rows[1:-1] if len(smp.rows) > 4 else smp.rows
    return e0.elapsed_time(e1) / n, median([r[0] for r in rows]), median([r[1] for r in rows]), n


SHAPES = {"in_proj": (3072, 1024), "linear1": (2048, 1024)}
DTYPES = {"fp16": (torch.float16, 2), "bf16": (torch.bfloat16, 1)}


def run(M, seconds, shapes=("in_proj", "linear1"), dtypes=("fp16", "bf16"), verbose=False):
    """rows of {shape, operands, impl (cublas | tim_b200), tflops, sm_mhz_median, power_w_median, ...}; used by bench.py too"""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(0)
    res = []
    for name in shapes:
        N, K = SHAPES[name]
        flops = 2.0 * M * N * K
        for dt_name in dtypes:
            tdt, code = DTYPES[dt_name]
            A = torch.randn(M, K, generator=g, device=dev).to(tdt)
            W = (torch.randn(N, K, generator=g, device=dev) / math.sqrt(K)).to(tdt)
            bias = torch.zeros(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=tdt)
            Wt = W.t()

            def cublas(n):
                for _ in range(n):
                    torch.matmul(A, Wt, out=out)

            def ours(n):
                ms = C.c_float(0)
                r = lib.tim_bench_linear(code, C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()), C.c_void_p(0),
                                         C.c_void_p(out.data_ptr()), M, N, K, 0, 0, 2, n, C.byref(ms))
                assert r == 0, lib.tim_last_error(None)

            for impl, fn in [("cublas", cublas), ("tim_b200", ours)]:
                ms, clk, pw, n = timed(fn, seconds)
                row = {"shape": name, "M": M, "N": N, "K": K, "operands": dt_name, "impl": impl, "launches": n, "ms_per_launch": ms,
                       "tflops": flops / (ms * 1e-3) / 1e12, "sm_mhz_median": clk, "power_w_median": pw}
                res.append(row)
                if verbose:
                    print(json.dumps(row), flush=True)
            del A, W, out, Wt
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=204800)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--out", default="")
    ap.add_argument("--shapes", default="in_proj,linear1")
    ap.add_argument("--dtypes", default="fp16,bf16")
    args = ap.parse_args()
    res = run(args.m, args.seconds, tuple(args.shapes.split(",")), tuple(args.dtypes.split(",")), verbose=True)
    if args.out:
        json.dump({"what": "sustained back-to-back GEMM launches, one shape at a time", "rows": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
