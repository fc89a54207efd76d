#!/usr/bin/env python
"""Detection post-processing throughput: soft-NMS of an evaluation's proposals, all (video, class) groups in
one launch (tim_b200.postprocess.batched_nms_videos -> tim_softnms_1d) against the reference's compiled CPU extension driven
class by class the way detection/eval_detection/nms.py:123-155 does (oracle/_ref/nms_1d_cpu.so, one host thread; the reference
runs 32 such workers in a joblib pool, format_predictions_epic.py:44-49).

    python tools/nms_bench.py [--videos 64] [--per-video 20000] [--classes 97] [--out gpurun_out/nms_bench.json]

Workload: per video `per-video` (proposal, class) entries above the score threshold; proposals are jittered copies of ~40 ground
truth actions per video (so classes hold heavily overlapping segments, as detector output does), classes drawn Zipf-like.
Parameters of the reference evaluation: gaussian soft-NMS, sigma 0.25, min_score 0.001 (format_predictions_epic.py:146-157).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tim_b200.postprocess import batched_nms_videos, grouped_nms   # noqa: E402


def workload(videos, per_video, classes, seed=0):
    rng = np.random.default_rng(seed)
    segs, scores, cls, vid = [], [], [], []
    for v in range(videos):
        n_act = 40
        a_start = rng.uniform(0, 600, n_act)
        a_len = rng.lognormal(0.7, 0.8, n_act) + 0.3
        a_cls = np.minimum((rng.zipf(1.3, n_act) - 1), classes - 1)
        which = rng.integers(0, n_act, per_video)
        st = a_start[which] + rng.normal(0, 0.3, per_video) * a_len[which]
        ln = a_len[which] * np.exp(rng.normal(0, 0.25, per_video))
        c = np.where(rng.random(per_video) < 0.7, a_cls[which], np.minimum(rng.zipf(1.3, per_video) - 1, classes - 1))
        segs.append(np.round(np.stack([st, st + ln], 1), 3))
        scores.append(rng.beta(0.6, 3.0, per_video) * 0.97 + 0.03)
        cls.append(c)
        vid.append(np.full(per_video, v))
    return (np.concatenate(segs).astype(np.float32), np.concatenate(scores).astype(np.float32),
            np.concatenate(cls).astype(np.int64), np.concatenate(vid).astype(np.int64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=64)
    ap.add_argument("--per-video", type=int, default=20000)
    ap.add_argument("--classes", type=int, default=97)
    ap.add_argument("--cpu-videos", type=int, default=64, help="videos of the same workload timed on the host (bounded sample)")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    segs, scores, cls, vid = workload(args.videos, args.per_video, args.classes)
    N = len(scores)
    prm = dict(iou_threshold=0.1, min_score=0.001, sigma=0.25, method=2, nms="soft")
    d_segs, d_scores, d_cls, d_vid = (torch.from_numpy(a).to(dev) for a in (segs, scores, cls, vid))
    keys = d_vid * args.classes + d_cls

    def run_all():
        return batched_nms_videos(d_segs, d_scores, d_cls, d_vid, device=dev, **{k: v for k, v in prm.items()})

    out = run_all()
    torch.cuda.synchronize()
    kept = int(out[1].shape[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # whole call, device-resident inputs (sort into groups + kernel + gather + final sort)
    e0.record()
    for _ in range(args.reps):
        run_all()
    e1.record()
    torch.cuda.synchronize()
    ms_call = e0.elapsed_time(e1) / args.reps
    # host inputs -> host outputs (what a drop-in for format_predictions_epic.py sees)
    t = time.perf_counter()
    for _ in range(args.reps):
        r = batched_nms_videos(segs, scores, cls, vid, device=dev, **prm)
        _ = [x.cpu() for x in r]
    ms_e2e = (time.perf_counter() - t) / args.reps * 1e3
    # the kernel alone: groups prepared once, CUDA events around the library call on the current stream
    from tim_b200 import _lib
    lib = _lib.load()
    sk, perm = torch.sort(keys, stable=True)
    counts = torch.unique_consecutive(sk, return_counts=True)[1]
    G = int(counts.numel())
    offs = torch.zeros((G + 1,), dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(counts, 0)
    g_segs, g_scores = d_segs[perm].contiguous(), d_scores[perm].contiguous()
    dets, inds = torch.empty((N, 3), device=dev), torch.empty((N,), dtype=torch.int64, device=dev)
    kept_g = torch.empty((G,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib.tim_nms_workspace_bytes(N))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)

    def kernel():
        _lib.check(lib.tim_softnms_1d(g_segs.data_ptr(), g_scores.data_ptr(), offs.data_ptr(), G, N, 0.1, 0.25, 0.001, 2, dets.data_ptr(),
                                      inds.data_ptr(), kept_g.data_ptr(), ws.data_ptr(), ws_bytes,
                                      torch.cuda.current_stream(dev).cuda_stream), None)

    kernel()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.reps):
        kernel()
    e1.record()
    torch.cuda.synchronize()
    ms_kernel = e0.elapsed_time(e1) / args.reps
    assert int(kept_g.sum()) == kept
    group_sizes = torch.unique(keys, return_counts=True)[1]
    res = {"metric": "proposals_per_sec", "workload": {"videos": args.videos, "entries": N, "classes": args.classes,
                                                        "groups": int(group_sizes.numel()), "largest_group": int(group_sizes.max()),
                                                        "median_group": float(group_sizes.float().median()), "kept": kept, **prm},
           "kernel_ms": ms_kernel, "kernel_proposals_per_sec": N / (ms_kernel * 1e-3),
           "device_call_ms": ms_call, "device_proposals_per_sec": N / (ms_call * 1e-3),
           "e2e_ms": ms_e2e, "e2e_proposals_per_sec": N / (ms_e2e * 1e-3)}
    # CPU baseline: the reference's compiled extension, per class, one thread, on a bounded sample of the videos
    try:
        from oracle import build_ref
        mod = build_ref.load()
    except Exception:
        mod = None
    if mod is not None:
        torch.set_num_threads(1)
        nv = min(args.cpu_videos, args.videos)
        sel = vid < nv
        ts, tp, tc = torch.from_numpy(segs[sel]), torch.from_numpy(scores[sel]), torch.from_numpy(cls[sel])
        tv = torch.from_numpy(vid[sel])
        t = time.perf_counter()
        total = 0
        for v in range(nv):
            m = tv == v
            vs, vp, vc = ts[m], tp[m], tc[m]
            for c in torch.unique(vc):
                idx = torch.where(vc == c)[0]
                dets = vs.new_empty((len(idx), 3))
                inds = mod.softnms(vs[idx].contiguous(), vp[idx].contiguous(), dets, 0.1, 0.25, 0.001, 2)
                total += len(inds)
        dt = time.perf_counter() - t
        res["cpu_baseline"] = {"value": int(sel.sum()) / dt, "unit": "proposals/s", "cores": 1, "kind": "reference",
                               "sample": f"{nv} of {args.videos} videos ({int(sel.sum())} entries), nms_1d_cpu.softnms per class, {dt:.2f} s"}
        # the same sample on the device must keep the same number of proposals
        r = batched_nms_videos(segs[sel], scores[sel], cls[sel], vid[sel], device=dev, **prm)
        res["cpu_baseline"]["kept_reference"] = total
        res["cpu_baseline"]["kept_device"] = int(r[1].shape[0])
    print(json.dumps(res))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
