#!/usr/bin/env python
"""Golden vectors for the detection query labelling (detection/time_interval_machine/models/tim.py:158-270: get_query_ious,
label_queries, assign_positive_labels), minted by calling the UNMODIFIED reference methods on CPU fp32.

    python tools/make_golden_labels.py        # writes tests/golden/label_queries.npz  (build container only)
Inputs are stored too (they are tiny). Cases: visual with verb/noun/action labels, visual action-only, audio; ground-truth
segments starting before the window (negative starts -> the reference's offset shift), zero-length padded segments, exact ties.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_golden import build_reference   # noqa: E402
from tim_b200.config import TIMConfig   # noqa: E402


def main():
    import torch
    rng = np.random.default_rng(7)
    out = {}
    cases = {
        "vn": dict(num_class=[[5, 7, 11], 4], include_verb_noun=True, modality="visual", B=3, Nq=37, Na=6, neg=True),
        "act": dict(num_class=[9, 4], include_verb_noun=False, modality="visual", B=2, Nq=64, Na=5, neg=False),
        "aud": dict(num_class=(9, 4), include_verb_noun=False, modality="audio", B=2, Nq=20, Na=3, neg=True, data_modality="audio_visual"),
    }
    for name, c in cases.items():
        cfg = TIMConfig(num_class=c["num_class"], visual_input_dim=48, audio_input_dim=40, d_model=64, nhead=4, num_layers=1,
                        num_feats=6, data_modality=c.get("data_modality", "visual"), include_verb_noun=c["include_verb_noun"],
                        variant="detection")
        model = build_reference(cfg)
        B, Nq, Na = c["B"], c["Nq"], c["Na"]
        st = rng.uniform(0.0, 0.9, (B, Nq)).astype(np.float32)
        q = np.stack([st, st + rng.uniform(0.01, 0.3, (B, Nq)).astype(np.float32)], -1)
        gs = rng.uniform(-0.2 if c["neg"] else 0.0, 0.8, (B, Na)).astype(np.float32)
        g = np.stack([gs, gs + rng.uniform(0.02, 0.4, (B, Na)).astype(np.float32)], -1)
        g[:, -1] = 0.0                                        # a zero-length padded segment per clip
        q[0, 0] = g[0, 0]                                     # an exact match (IoU 1)
        q[0, 1] = q[0, 2]                                     # two identical queries
        g[1, 1] = g[1, 0]                                     # two identical ground truths: argmax must take the first
        nl = 3 if c["modality"] == "visual" else 1
        hi = [5, 7, 11] if c["include_verb_noun"] else ([9, 9, 9] if c["modality"] == "visual" else [4])
        lab = np.stack([rng.integers(0, hi[k], (B, Na)) for k in range(nl)], -1).astype(np.int64)
        target = {"v_gt_segments": torch.from_numpy(g.copy()), "a_gt_segments": torch.from_numpy(g.copy()),
                  "verb": torch.from_numpy(lab[..., 0].copy()), "noun": torch.from_numpy(lab[..., min(1, nl - 1)].copy()),
                  "action": torch.from_numpy(lab[..., nl - 1].copy()), "class_id": torch.from_numpy(lab[..., 0].copy())}
        with torch.no_grad():
            tg, lb, iou = model.label_queries(torch.from_numpy(q.copy()), target, c["modality"], model.iou_threshold)
        out[f"{name}_queries"], out[f"{name}_gt"] = q, g
        out[f"{name}_labels"] = lab if c["modality"] == "visual" else lab[..., :1]
        out[f"{name}_targets"], out[f"{name}_ious"] = tg.numpy(), iou.numpy()
        lbs = lb if isinstance(lb, (list, tuple)) else [lb]
        for k, t in enumerate(lbs):
            out[f"{name}_smooth{k}"] = t.numpy().astype(np.float32)
        out[f"{name}_meta"] = np.array([model.iou_threshold, model.label_smoothing], np.float64)
        print(name, tg.shape, [tuple(t.shape) for t in lbs], iou.shape, "positives", int(np.isfinite(tg.numpy()[:, 0]).sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "label_queries.npz"), **out)


if __name__ == "__main__":
    main()
