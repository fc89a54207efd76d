"""CPU restatement (numpy, fp32 in the reference's operation order) of the detection query labelling — TEST INFRASTRUCTURE ONLY
(same rules as oracle/tim_oracle.py: only tests/ may import it).

Reference: detection/time_interval_machine/models/tim.py
  get_query_ious          :186-212   IoU of every query with every ground-truth segment, after shifting both by
                                     |min(min_a gt_start, 0)| (the reference adds the offset IN PLACE, so the returned
                                     regression targets are the shifted segments)
  label_queries           :214-270   argmax over ground truths (first maximum; NaN counts as maximal, as torch.argmax does),
                                     negatives (IoU < threshold) get targets = +inf and labels = -1
  assign_positive_labels  :158-185   smoothed one-hot: one_hot(id, C + 1) * s + (1 - s) / (C + 1), last column dropped; id -1 -> C

Parity pin: tests/golden/label_queries.npz, minted by tools/make_golden_labels.py from the unmodified reference on CPU fp32.
"""
import numpy as np


def label_queries(queries, gt_segs, gt_labels, iou_threshold):
    """queries [B,Nq,2] f32, gt_segs [B,Na,2] f32, gt_labels [B,Na,Nl] i64 ->
    targets [B*Nq,2] f32, label_ids [B*Nq,Nl] i64 (-1 = negative), ious [B*Nq] f32."""
    q = np.asarray(queries, np.float32)
    g = np.asarray(gt_segs, np.float32)
    lab = np.asarray(gt_labels, np.int64)
    B, Nq, _ = q.shape
    Na = g.shape[1]
    off = np.abs(np.minimum(g[:, :, 0].min(axis=1), np.float32(0.0))).astype(np.float32)        # [B]
    qs = (q[:, :, None, 0] + off[:, None, None]).astype(np.float32)
    qe = (q[:, :, None, 1] + off[:, None, None]).astype(np.float32)
    gs = (g[:, None, :, 0] + off[:, None, None]).astype(np.float32)
    ge = (g[:, None, :, 1] + off[:, None, None]).astype(np.float32)
    inter = np.maximum((np.minimum(qe, ge) - np.maximum(qs, gs)).astype(np.float32), np.float32(0.0))
    union = (((ge - gs).astype(np.float32) + (qe - qs).astype(np.float32)).astype(np.float32) - inter).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = (inter / union).astype(np.float32)                                                 # [B,Nq,Na]
    key = np.where(np.isnan(iou), np.float32(np.inf), iou)                                        # NaN is maximal for torch.argmax
    # first index of the maximum (np.argmax does that); a NaN beats +inf only if it comes first among equals: make NaN strictly larger
    key = np.where(np.isnan(iou), np.float64(np.inf) * 2, key.astype(np.float64))
    idx = np.argmax(key, axis=-1)                                                                # [B,Nq]
    bi, qi = np.meshgrid(np.arange(B), np.arange(Nq), indexing="ij")
    best = iou[bi, qi, idx]
    targets = np.stack([np.broadcast_to(gs, iou.shape)[bi, qi, idx], np.broadcast_to(ge, iou.shape)[bi, qi, idx]], -1).astype(np.float32)
    ids = lab[bi, idx]                                                                           # [B,Nq,Nl]
    neg = best < np.float32(iou_threshold)                                                       # NaN -> False
    targets = np.where(neg[..., None], np.float32(np.inf), targets)
    ids = np.where(neg[..., None], np.int64(-1), ids)
    return targets.reshape(B * Nq, 2), ids.reshape(B * Nq, -1), best.reshape(B * Nq)


def smooth_labels(ids, num_classes, smoothing):
    """ids [rows] i64 (-1 = none) -> [rows, num_classes] f32 (tim.py:172-182)."""
    ids = np.asarray(ids, np.int64)
    ids = np.where(ids == -1, num_classes, ids)
    s = np.float32(smoothing)
    base = np.float32((1 - smoothing) / (num_classes + 1))       # python float arithmetic, then one fp32 rounding, as torch does
    onehot = (ids[:, None] == np.arange(num_classes + 1)[None]).astype(np.float32)
    return ((onehot * s).astype(np.float32) + base).astype(np.float32)[:, :-1]
