#!/usr/bin/env python
"""Mints tests/golden/nms.npz from the reference's OWN compiled NMS extension (oracle/_ref/nms_1d_cpu.so, built by
oracle/build_ref.py from /root/reference/detection/eval_detection/csrc/nms_cpu.cpp) and, for the batched cases, from the
reference's unmodified Python driver detection/eval_detection/nms.py (imported from /root/reference). Run in the build container
(the reference tree does not exist on the GPU box):

    python oracle/build_ref.py && python tools/make_golden_nms.py

Inputs are seeded; every case stores inputs, parameters and the reference outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref   # noqa: E402


def make_segments(rng, n, kind):
    if kind == "uniform":                       # proposals spread over a 60 s video
        start = rng.uniform(0, 60, n)
        length = rng.uniform(0.2, 6.0, n)
    elif kind == "clustered":                   # heavy overlap: a handful of actions, many jittered proposals each
        centres = rng.uniform(5, 55, 6)
        start = centres[rng.integers(0, 6, n)] + rng.normal(0, 0.4, n)
        length = np.abs(rng.normal(3.0, 0.5, n)) + 0.1
    elif kind == "quantised":                   # rounded to 3 decimals as format_predictions_epic.py:125 does -> exact duplicates
        start = np.round(rng.uniform(0, 8, n), 1)
        length = np.round(rng.uniform(0.5, 2.0, n), 1)
    else:
        raise ValueError(kind)
    segs = np.stack([start, start + length], 1).astype(np.float32)
    return segs


def cases():
    rng = np.random.default_rng(20240611)
    out = []
    for name, n, kind, thr, sigma, mins, method, qs in (
            ("soft_gauss_uniform", 200, "uniform", 0.1, 0.25, 0.001, 2, False),
            ("soft_gauss_clustered", 300, "clustered", 0.1, 0.25, 0.05, 2, False),
            ("soft_gauss_sigma04", 257, "clustered", 0.1, 0.4, 0.001, 2, False),
            ("soft_gauss_ties", 96, "quantised", 0.1, 0.25, 0.02, 2, True),
            ("soft_linear", 150, "clustered", 0.3, 0.25, 0.05, 1, False),
            ("soft_vanilla", 150, "clustered", 0.5, 0.25, 0.001, 0, False),
            ("soft_vanilla_ties", 64, "quantised", 0.4, 0.25, 0.001, 0, True),
            ("soft_single", 1, "uniform", 0.1, 0.25, 0.001, 2, False),
            ("soft_all_below_min", 17, "uniform", 0.1, 0.25, 2.0, 2, False),
            ("soft_big", 1500, "clustered", 0.1, 0.25, 0.01, 2, False)):
        segs = make_segments(rng, n, kind)
        scores = rng.uniform(0.03, 1.0, n).astype(np.float32)
        if qs:
            scores = np.round(scores, 1).astype(np.float32)
        out.append((name, "soft", segs, scores, None, dict(iou_threshold=thr, sigma=sigma, min_score=mins, method=method)))
    for name, n, kind, thr, qs in (("nms_uniform", 200, "uniform", 0.1, False), ("nms_clustered", 300, "clustered", 0.5, False),
                                   ("nms_ties", 96, "quantised", 0.3, True)):
        segs = make_segments(rng, n, kind)
        scores = rng.uniform(0.03, 1.0, n).astype(np.float32)
        if qs:
            scores = np.round(scores, 1).astype(np.float32)
        out.append((name, "nms", segs, scores, None, dict(iou_threshold=thr)))
    for name, n, ncls, nms in (("batched_soft", 1500, 12, "soft"), ("batched_vanilla", 800, 7, "vanilla")):
        segs = make_segments(rng, n, "clustered")
        scores = rng.uniform(0.03, 1.0, n).astype(np.float32)
        cls = rng.integers(0, ncls, n).astype(np.int64) * 301 + 5         # sparse ids like verb * 300 + noun
        out.append((name, "batched", segs, scores, cls, dict(iou_threshold=0.1, min_score=0.001, sigma=0.25, method=2, nms=nms)))
    return out


def main():
    path = build_ref.build()
    assert path, "the reference tree is needed to mint the golden file"
    mod = build_ref.load()
    sys.modules["nms_1d_cpu"] = mod                       # detection/eval_detection/nms.py:5 does `import nms_1d_cpu`
    sys.path.insert(0, "/root/reference/detection/eval_detection")
    import nms as ref_nms                                 # the reference's Python driver, unmodified
    blob = {}
    names = []
    for name, kind, segs, scores, cls, prm in cases():
        names.append(name)
        blob[f"{name}/kind"] = np.array(kind)
        blob[f"{name}/segs"], blob[f"{name}/scores"] = segs, scores
        for k, v in prm.items():
            blob[f"{name}/p_{k}"] = np.array(v)
        ts, tp = torch.from_numpy(segs), torch.from_numpy(scores)
        if kind == "soft":
            dets = torch.zeros((segs.shape[0], 3))
            inds = mod.softnms(ts.clone(), tp.clone(), dets, float(prm["iou_threshold"]), float(prm["sigma"]), float(prm["min_score"]),
                               int(prm["method"]))
            blob[f"{name}/inds"] = inds.numpy()
            blob[f"{name}/dets"] = dets[:len(inds)].numpy()
        elif kind == "nms":
            blob[f"{name}/inds"] = mod.nms(ts.clone(), tp.clone(), float(prm["iou_threshold"])).numpy()
        else:
            blob[f"{name}/cls"] = cls
            s, p, c = ref_nms.batched_nms(ts.clone(), tp.clone(), torch.from_numpy(cls), iou_threshold=prm["iou_threshold"],
                                          min_score=prm["min_score"], sigma=prm["sigma"], method=prm["method"], nms=prm["nms"],
                                          multi_class=True)
            blob[f"{name}/out_segs"], blob[f"{name}/out_scores"], blob[f"{name}/out_cls"] = s, p, c
        print(name, kind, segs.shape[0], "->", len(blob.get(f"{name}/inds", blob.get(f"{name}/out_scores"))))
    blob["names"] = np.array(names)
    out = os.path.join(ROOT, "tests", "golden", "nms.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
