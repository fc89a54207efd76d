#!/usr/bin/env python
"""Mints tests/golden/detpost.npz: the detection post-processing chain run by the UNMODIFIED reference on CPU.

  stage 1  FeatureMeter.update + finalize_metrics   (detection/time_interval_machine/utils/meters.py:574-800)
           two batches of synthetic head outputs -> 'action' (sigmoid scores), 'v_proposals' (seconds), 'video_ids'
  stage 2  main() of detection/eval_detection/format_predictions.py:98-184 on that file: thresholding loop, per-video soft-NMS
           through the reference's nms.py and its compiled extension (oracle/_ref/nms_1d_cpu.so), submission file tim.json

    python oracle/build_ref.py && python tools/make_golden_detpost.py        (build container only)

Only two things are stubbed, neither on the arithmetic path: misc.gpu_mem_usage (queries a CUDA device for a log line) and the
reference's logging imports (simplejson / fvcore, as in tools/make_golden.py). main()'s last step - a subprocess running the mAP
script on a ground-truth file - fails harmlessly (no check) after tim.json has been written.
"""
import argparse
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref   # noqa: E402

N_CLASS = 29
QUERIES = 48
WINDOW_SIZE = 29.97            # not representable in fp32: the reference multiplies by its fp32 rounding
SCORE_THRESHOLD = 0.03
SIGMA = 0.25


def shim_reference_imports():
    sj = types.ModuleType("simplejson"); sj.dumps = json.dumps; sys.modules["simplejson"] = sj
    fio = types.ModuleType("fvcore.common.file_io"); fio.PathManager = type("PM", (), {"open": staticmethod(open)})
    tm = types.ModuleType("fvcore.common.timer")

    class Timer:
        def reset(self): pass
        def pause(self): pass
        def seconds(self): return 0.0
    tm.Timer = Timer
    sys.modules.update({"fvcore": types.ModuleType("fvcore"), "fvcore.common": types.ModuleType("fvcore.common"),
                        "fvcore.common.file_io": fio, "fvcore.common.timer": tm})


def synth_batches():
    """Two batches of 6 windows (3 videos), as the detection heads / loader hand them to FeatureMeter.update."""
    rng = np.random.default_rng(20240612)
    batches = []
    vids = ["P01_01", "P01_02", "P22_107"]
    for b in range(2):
        B = 6
        logits = rng.normal(-5.0, 1.6, (B * QUERIES, N_CLASS)).astype(np.float32)
        q_start = rng.uniform(0.0, 1.0, (B, QUERIES)).astype(np.float32)
        q = np.stack([q_start, q_start + rng.uniform(0.01, 0.3, (B, QUERIES)).astype(np.float32)], -1)
        # regression outputs cluster around a few actions per window (so the NMS has work), plus out-of-range and empty ones
        centres = rng.uniform(0.1, 0.8, (B, 5))
        which = rng.integers(0, 5, (B, QUERIES))
        st = np.take_along_axis(centres, which, 1) + rng.normal(0, 0.02, (B, QUERIES))
        reg = np.stack([st, st + np.abs(rng.normal(0.1, 0.03, (B, QUERIES)))], -1).astype(np.float32).reshape(B * QUERIES, 2)
        reg[0] = (-0.2, 0.1)                 # clamped at 0
        reg[1] = (0.9, 1.9)                  # clamped at max_time
        reg[2] = (0.4, 0.4)                  # empty -> dropped by the thresholding loop
        reg[3] = (0.5, 0.3)                  # negative length -> dropped
        reg[4] = (0.20004, 0.20044)          # rounds to an empty segment at 3 decimals once scaled? (kept or dropped by np.round)
        meta = {"video_id": [vids[(b * B + i) // 4] for i in range(B)],
                "window_start": torch.tensor([7.5 * ((b * B + i) % 4) + 0.1 * i for i in range(B)], dtype=torch.float64),
                "window_size": torch.tensor([WINDOW_SIZE] * B, dtype=torch.float64)}      # default_collate of python floats
        batches.append((logits, reg, q, meta))
    return batches


def main():
    assert build_ref.build(), "the reference tree is needed to mint the golden file"
    shim_reference_imports()
    sys.path.insert(0, "/root/reference/detection")
    import time_interval_machine.utils.misc as misc
    misc.gpu_mem_usage = lambda: (0.0, 0.0)              # log-line helper that needs a CUDA device
    from time_interval_machine.utils.meters import FeatureMeter
    args = argparse.Namespace(data_modality="visual", include_verb_noun=False, num_class=[N_CLASS, 4])
    meter = FeatureMeter(args)
    blob = {}
    for i, (logits, reg, q, meta) in enumerate(synth_batches()):
        meter.update((None, None, torch.from_numpy(logits), None), (torch.from_numpy(reg), None), (torch.from_numpy(q), None), meta)
        blob[f"b{i}/logits"], blob[f"b{i}/reg"], blob[f"b{i}/queries"] = logits, reg, q
        blob[f"b{i}/window_start"] = meta["window_start"].numpy()
        blob[f"b{i}/video_id"] = np.array(meta["video_id"])
    data = meter.finalize_metrics()
    blob["action"], blob["v_proposals"] = data["action"], data["v_proposals"]
    blob["video_ids"] = np.array([str(v) for v in data["video_ids"]])
    print("stage 1:", data["action"].shape, data["action"].dtype, data["v_proposals"].shape, data["v_proposals"].dtype)

    # stage 2: the reference's formatting script on that file
    sys.modules["nms_1d_cpu"] = build_ref.load()
    sys.path.insert(0, "/root/reference/detection/eval_detection")
    sys.argv = ["format_predictions.py", "x", "y"]
    import format_predictions as fp                       # unmodified; argparse definitions only at import
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            torch.save(data, "preds.pth.tar", pickle_protocol=5)
            # torch >= 2.6 defaults to weights_only=True, which refuses the numpy object array of video ids the reference saves
            _load = torch.load
            torch.load = lambda *a, **k: _load(*a, **{**k, "weights_only": False})
            try:
                fp.main(argparse.Namespace(path_to_preds="preds.pth.tar", path_to_gt="missing.pkl", score_threshold=SCORE_THRESHOLD,
                                           sigma=SIGMA, is_audio=False, n_jobs=1))
            finally:
                torch.load = _load
            sub = json.load(open("tim.json"))
        finally:
            os.chdir(cwd)
    res = sub["results"]
    vids = sorted(res)
    blob["result_videos"] = np.array(vids)
    for v in vids:
        blob[f"result/{v}/action"] = np.array([e["action"] for e in res[v]], np.int64)
        blob[f"result/{v}/score"] = np.array([e["score"] for e in res[v]], np.float64)
        blob[f"result/{v}/segment"] = np.array([e["segment"] for e in res[v]], np.float64).reshape(-1, 2)
        print("stage 2:", v, len(res[v]), "detections")
    blob["params"] = np.array([WINDOW_SIZE, SCORE_THRESHOLD, SIGMA, QUERIES], np.float64)
    out = os.path.join(ROOT, "tests", "golden", "detpost.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
