#!/usr/bin/env python
"""Timing of the device query labelling (tim_label_queries + tim_smooth_labels) at the detection dense-query size, next to the
reference's formulation run (a) with eager PyTorch on the same GPU and (b) on the host cores - restated here from
detection/time_interval_machine/models/tim.py:186-270 (the reference checkout is not on the GPU box).

    python tools/label_bench.py [--clips 96] [--queries 2048] [--gts 24] [--classes 97]
Algorithmic bytes of the device path: reads B*Nq*8 + B*Na*(8 + 8*Nl), writes B*Nq*(8 + 8*Nl + 4) + B*Nq*C*4 (smoothed labels).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200.config import named_config   # noqa: E402
from tim_b200.plugin import TIMEngine   # noqa: E402


def torch_reference(queries, segs, labels, thr, C_, s):
    """The reference's tensor program (repeat_interleave to [B,Nq,Na], stack/max/min, argmax, advanced indexing, one_hot)."""
    Nq, Na = queries.shape[1], segs.shape[1]
    q = queries[:, :, None].repeat_interleave(Na, dim=2)
    t = segs[:, None].repeat_interleave(Nq, dim=1)
    ql = labels[:, None].repeat_interleave(Nq, dim=1)
    qs, qe, gs, ge = q[..., 0], q[..., 1], t[..., 0], t[..., 1]
    off = torch.abs(torch.clamp(gs.min(dim=-1)[0], max=0.0))
    qs += off[:, :, None]; qe += off[:, :, None]; gs += off[:, :, None]; ge += off[:, :, None]
    ist = torch.stack([qs, gs], dim=-1).max(dim=-1)[0]
    ien = torch.stack([qe, ge], dim=-1).min(dim=-1)[0]
    inter = torch.clamp(ien - ist, min=0.0)
    ious = inter / ((ge - gs) + (qe - qs) - inter)
    mi = ious.argmax(-1).flatten()
    bi = torch.arange(ious.shape[0], device=ious.device).repeat_interleave(ious.shape[1])
    fi = torch.arange(ious.shape[1], device=ious.device).repeat(ious.shape[0])
    best = ious[bi, fi, mi].reshape(ious.shape[0], -1)
    tg = t[bi, fi, mi].reshape(ious.shape[0], -1, 2)
    lb = ql[bi, fi, mi].reshape(ious.shape[0], -1, ql.shape[-1])
    neg = best < thr
    tg.masked_fill_(neg[:, :, None], float("inf")); lb.masked_fill_(neg[:, :, None], -1)
    lb = torch.flatten(lb, 0, 1)
    a = lb[:, 2].masked_fill(lb[:, 2] == -1, C_)
    sm = ((torch.nn.functional.one_hot(a, C_ + 1) * s) + ((1 - s) / (C_ + 1)))[:, :-1]
    return torch.flatten(tg, 0, 1), sm, torch.flatten(best)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=96)
    ap.add_argument("--queries", type=int, default=2048)
    ap.add_argument("--gts", type=int, default=24)
    ap.add_argument("--classes", type=int, default=97)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    B, Nq, Na, C_ = a.clips, a.queries, a.gts, a.classes
    st = torch.rand(B, Nq, generator=g) * 0.9
    q = torch.stack([st, st + 0.01 + 0.29 * torch.rand(B, Nq, generator=g)], -1)
    gs = torch.rand(B, Na, generator=g) * 0.8
    gt = torch.stack([gs, gs + 0.02 + 0.3 * torch.rand(B, Na, generator=g)], -1)
    lab = torch.randint(0, C_, (B, Na, 3), generator=g)
    cfg, _, _ = named_config("cfg1")
    eng = TIMEngine(cfg, 0, "fp32")
    qd, gd, ld = q.to(dev), gt.to(dev), lab.to(dev)

    def ours():
        t, ids, iou = eng.label_queries(qd, gd, ld, 0.25)
        return t, eng.smooth_labels(ids, 2, C_, 0.9), iou

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, out

    ms_ours, o = timed(ours, a.iters)
    ms_torch, r = timed(lambda: torch_reference(qd, gd, ld, 0.25, C_, 0.9), a.iters)
    same = all(torch.equal(x, y) for x, y in zip(o, r))
    t0 = time.perf_counter()
    for _ in range(3):
        torch_reference(q, gt, lab, 0.25, C_, 0.9)
    ms_cpu = (time.perf_counter() - t0) / 3 * 1e3
    nbytes = B * Nq * 8 + B * Na * (8 + 24) + B * Nq * (8 + 24 + 4) + B * Nq * C_ * 4
    print(f"B={B} Nq={Nq} Na={Na} C={C_}: device kernels {ms_ours * 1e3:.1f} us ({nbytes / ms_ours / 1e6:.0f} GB/s algorithmic), "
          f"reference tensor program on the same GPU (eager torch) {ms_torch * 1e3:.1f} us, on {torch.get_num_threads()} host threads "
          f"{ms_cpu:.1f} ms; identical results: {same}")
    eng.close()


if __name__ == "__main__":
    main()
