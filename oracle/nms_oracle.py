"""CPU restatement of the reference's 1-D (soft-)NMS post-processing — TEST INFRASTRUCTURE ONLY (same rules as
oracle/tim_oracle.py: only tests/, tools/*_bench.py's cpu_baseline leg and tools/make_golden_nms.py may import it).

Reference:
  detection/eval_detection/csrc/nms_cpu.cpp
      nms_1d_cpu       :19-58    greedy suppression in descending-score order, IoU >= threshold suppresses
      softnms_1d_cpu   :67-160   selection sort with score decay: every round picks the maximum of the live tail (first maximum
                                 in CURRENT array order), swaps it to the front, decays the rest (0 vanilla / 1 linear /
                                 2 gaussian exp(-iou^2 / sigma)), and deletes entries whose score fell below min_score by moving
                                 the LAST live entry into the hole (that entry is then decayed in the same round)
  detection/eval_detection/nms.py
      SoftNMSop / NMSop :7-61, batched_nms :98-180 (per-class loop in ascending class id, then one descending sort by score)
  detection/eval_detection/format_predictions_epic.py
      filter_nms :51-112, main :114-157 (score threshold, one entry per (proposal, class) above it, per-video NMS)

All arithmetic is IEEE fp32 in the reference's operation order (np.float32 scalars); areas = fl32(fl32(x2 - x1) + fl32(1e-6)).
The gaussian weight is std::exp(float): here exp is taken in double and rounded once to fp32 (the correctly rounded value;
glibc's expf differs from it by one ulp in rare cases, which the parity tests allow for: scores to 1e-6 relative).

Parity pin: tests/golden/nms.npz — outputs of the reference's own compiled extension (oracle/_ref/nms_1d_cpu.so, built from
/root/reference by oracle/build_ref.py) minted by tools/make_golden_nms.py; tests/test_oracle_golden.py checks this file against it.
"""
import math

import numpy as np

F = np.float32


def _areas(x1, x2):
    return ((x2 - x1).astype(F) + F(1e-6)).astype(F)


def nms_1d(segs, scores, iou_threshold):
    """nms_cpu.cpp:19-58 -> indices kept, in descending-score order (stable for equal scores, as torch's CPU sort is)."""
    segs = np.asarray(segs, F).reshape(-1, 2)
    scores = np.asarray(scores, F)
    n = segs.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    x1, x2 = segs[:, 0].copy(), segs[:, 1].copy()
    areas = _areas(x1, x2)
    order = np.argsort(-scores.astype(np.float64), kind="stable")
    thr = F(iou_threshold)
    select = np.ones(n, bool)
    for _i in range(n):
        if not select[_i]:
            continue
        i = order[_i]
        for _j in range(_i + 1, n):
            if not select[_j]:
                continue
            j = order[_j]
            xx1 = max(x1[i], x1[j])
            xx2 = min(x2[i], x2[j])
            inter = max(F(0.0), F(xx2 - xx1))
            ovr = F(inter / F(F(areas[i] + areas[j]) - inter))
            if ovr >= thr:
                select[_j] = False
    return order[select].astype(np.int64)


def softnms_1d(segs, scores, iou_threshold, sigma, min_score, method):
    """nms_cpu.cpp:67-160 -> (inds [k] i64 into the input, dets [k,3] f32 = (start, end, decayed score)) in pick order."""
    segs = np.asarray(segs, F).reshape(-1, 2)
    n = segs.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64), np.zeros((0, 3), F)
    x1, x2 = segs[:, 0].copy(), segs[:, 1].copy()
    sc = np.asarray(scores, F).copy()
    areas = _areas(x1, x2)
    inds = np.arange(n, dtype=np.int64)
    dets = np.zeros((n, 3), F)
    thr, sig, mins = F(iou_threshold), F(sigma), F(min_score)
    i = 0
    while i < n:
        m = i + int(np.argmax(sc[i:n]))                  # first maximum of the live tail (strict '<' in the reference)
        for arr in (x1, x2, sc, areas, inds):
            arr[i], arr[m] = arr[m], arr[i]
        ix1, ix2, iarea = x1[i], x2[i], areas[i]
        dets[i] = (ix1, ix2, sc[i])
        pos = i + 1
        while pos < n:
            xx1 = max(ix1, x1[pos])
            xx2 = min(ix2, x2[pos])
            inter = max(F(0.0), F(xx2 - xx1))
            ovr = F(inter / F(F(iarea + areas[pos]) - inter))
            w = F(1.0)
            if method == 0:
                if ovr >= thr:
                    w = F(0.0)
            elif method == 1:
                if ovr >= thr:
                    w = F(F(1.0) - ovr)
            elif method == 2:
                w = F(math.exp(float(F(-F(ovr * ovr)) / sig)))
            sc[pos] = F(sc[pos] * w)
            if sc[pos] < mins:
                last = n - 1
                x1[pos], x2[pos], sc[pos], areas[pos], inds[pos] = x1[last], x2[last], sc[last], areas[last], inds[last]
                n -= 1
                pos -= 1
            pos += 1
        i += 1
    return inds[:n].copy(), dets[:n].copy()


def batched_nms(segs, scores, cls_idxs, iou_threshold, min_score, sigma=0.5, method=2, nms="soft", max_seg_num=2000000):
    """nms.py:98-180 with multi_class=True (what format_predictions_epic.py:146-157 uses): per class in ascending id, then one
    descending sort by score. Returns (segs [k,2] f32, scores [k] f32, cls [k] i64). The order of EQUAL final scores is the
    stable one here; the reference's unstable torch.sort leaves it unspecified."""
    segs = np.asarray(segs, F).reshape(-1, 2)
    scores = np.asarray(scores, F)
    cls_idxs = np.asarray(cls_idxs, np.int64)
    if segs.shape[0] == 0:
        return np.zeros((0, 2), F), np.zeros((0,), F), np.zeros((0,), np.int64)
    out_s, out_p, out_c = [], [], []
    for c in np.unique(cls_idxs):
        cur = np.nonzero(cls_idxs == c)[0]
        if nms == "soft":
            inds, dets = softnms_1d(segs[cur], scores[cur], iou_threshold, sigma, min_score, method)
            out_s.append(dets[:, :2]); out_p.append(dets[:, 2]); out_c.append(cls_idxs[cur][inds])
        else:
            s, p = segs[cur], scores[cur]
            if min_score > 0:                                        # nms.py:15-19
                ok = p > F(min_score)
                s, p = s[ok], p[ok]
            inds = nms_1d(s, p, iou_threshold)
            if max_seg_num > 0:
                inds = inds[:max_seg_num]
            out_s.append(s[inds]); out_p.append(p[inds]); out_c.append(np.full(len(inds), c, np.int64))
    s, p, c = np.concatenate(out_s), np.concatenate(out_p), np.concatenate(out_c)
    order = np.argsort(-p.astype(np.float64), kind="stable")
    return s[order], p[order], c[order]


# ---------------------------------------------------------------------------------------------------------------------
# in front of the NMS: FeatureMeter.update and the thresholding loop of format_predictions.py
# ---------------------------------------------------------------------------------------------------------------------
def decode_predictions(logits, reg, window_start, window_size, max_time):
    """detection/time_interval_machine/utils/meters.py:652-700 (visual branch; the audio branch :702-722 is the same arithmetic):
    preds = sigmoid(logits) fp32; proposals = clamp(reg, 0, max_time) * window_size + window_start of the row's window.
    dtypes as the reference's tensors have them: the product is fp32 (a 0-dim float64 window_size does not promote a fp32
    tensor), the sum is float64 (window_start [R,1] is a float64 tensor with a dimension)."""
    logits = np.asarray(logits, F)
    reg = np.asarray(reg, F).reshape(-1, 2)
    ws = np.asarray(window_start, np.float64).reshape(-1)
    R = reg.shape[0]
    with np.errstate(over="ignore"):
        preds = (F(1.0) / (F(1.0) + np.exp(-logits).astype(F))).astype(F)
    v = np.minimum(np.maximum(reg, F(0.0)), F(max_time)).astype(F)
    scaled = (v * F(window_size)).astype(F)
    return preds, scaled.astype(np.float64) + np.repeat(ws, R // ws.shape[0])[:, None]


def threshold_detections(preds, proposals, score_threshold):
    """detection/eval_detection/format_predictions.py:103-125: proposals rounded to 3 decimals (np.round on float64), rows with
    end - start <= 0 dropped, one detection per class whose score exceeds the threshold, in (row, class) order.
    -> (row [n] i64, cls [n] i64, score [n] f32, segs [n,2] f32 = torch.FloatTensor of the rounded proposals)."""
    preds = np.asarray(preds, F)
    p = np.round(np.asarray(proposals, np.float64).reshape(-1, 2), 3)
    live = (p[:, 1] - p[:, 0]) > 0.0
    row, cls = np.nonzero((preds > F(score_threshold)) & live[:, None])
    return row.astype(np.int64), cls.astype(np.int64), preds[row, cls], p[row].astype(F)


def format_predictions(preds, proposals, video_ids, score_threshold=0.03, sigma=0.25, iou_threshold=0.1, min_score=0.001, method=2):
    """main() of detection/eval_detection/format_predictions.py:98-141: {video: (segs [k,2] f32, scores [k] f32, cls [k] i64)} with
    every video's detections sorted by descending score (stable), after per-class gaussian soft-NMS."""
    row, cls, score, segs = threshold_detections(preds, proposals, score_threshold)
    vids = np.asarray(video_ids)[row]
    out = {}
    for v in np.unique(np.asarray(video_ids)):
        m = vids == v
        out[str(v)] = batched_nms(segs[m], score[m], cls[m], iou_threshold, min_score, sigma=sigma, method=method, nms="soft")
    return out
