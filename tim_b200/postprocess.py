"""Detection post-processing on the device, mirroring the reference's evaluation chain:

  decode_predictions     <- FeatureMeter.update (detection/time_interval_machine/utils/meters.py:652-724): sigmoid scores,
                            regression outputs clamped and mapped back to seconds            (tim_det_decode)
  threshold_detections   <- the thresholding loop of detection/eval_detection/format_predictions.py:103-125
                                                                                              (tim_det_count / tim_det_emit)
  batched_nms            <- detection/eval_detection/nms.py: batched_nms :98-180 (SoftNMSop :35-61, NMSop :7-33), same signature,
                            instead of the scalar CPU extension nms_1d_cpu (csrc/nms_cpu.cpp)  (tim_softnms_1d / tim_nms_1d)
  batched_nms_videos, format_predictions
                         <- main() of format_predictions.py:98-141 / format_predictions_epic.py:114-157: all videos and classes
                            of an evaluation in one call instead of a Python loop over proposals and a joblib pool over videos

PyTorch is used for device memory and the index plumbing (sorting proposals into (video, class) groups, prefix sums, gathering
the kept rows); the arithmetic runs in the library's kernels (tim_b200/csrc/detpost.cu, nms.cu).

No CPU fallback: without a CUDA device or the built library every entry point raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _ptr(t: torch.Tensor):
    return t.data_ptr()


def grouped_nms(segs: torch.Tensor, scores: torch.Tensor, keys: torch.Tensor, *, iou_threshold: float, min_score: float,
                sigma: float = 0.5, method: int = 2, nms: str = "soft", max_seg_num: int = 0,
                device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(Soft-)NMS of every group of proposals sharing a key (e.g. video * n_classes + class), all groups in one launch.

    segs [N,2], scores [N], keys [N] int64. Returns device tensors (segs [k,2], scores [k], keys [k], source [k]) of the kept
    proposals: groups in ascending key order, inside a group in pick order (descending score after decay); `source` is the row of
    the input each kept proposal came from. Inside a group the proposals keep their input order, as `segs[curr_indices]` does in
    nms.py:125-135 — the reference's result depends on it."""
    if nms not in ("soft", "vanilla"):
        raise ValueError(f"nms must be 'soft' or 'vanilla', got {nms!r}")
    lib = _lib.load()
    if device is None:
        device = segs.device if segs.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("tim_b200.postprocess runs on a CUDA device only (no CPU fallback)")
    segs = segs.to(device=device, dtype=torch.float32).reshape(-1, 2).contiguous()
    scores = scores.to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    keys = keys.to(device=device, dtype=torch.int64).reshape(-1).contiguous()
    N = int(segs.shape[0])
    if scores.shape[0] != N or keys.shape[0] != N:
        raise ValueError("segs [N,2], scores [N] and keys [N] must agree on N")
    if N == 0:
        z = torch.zeros
        return (z((0, 2), device=device), z((0,), device=device), z((0,), dtype=torch.int64, device=device),
                z((0,), dtype=torch.int64, device=device))
    with torch.cuda.device(device):
        sorted_keys, perm = torch.sort(keys, stable=True)
        uniq, counts = torch.unique_consecutive(sorted_keys, return_counts=True)
        G = int(uniq.shape[0])
        offs = torch.zeros((G + 1,), dtype=torch.int64, device=device)
        offs[1:] = torch.cumsum(counts, 0)
        g_segs, g_scores = segs[perm].contiguous(), scores[perm].contiguous()
        dets = torch.empty((N, 3), dtype=torch.float32, device=device)
        inds = torch.empty((N,), dtype=torch.int64, device=device)
        kept = torch.empty((G,), dtype=torch.int32, device=device)
        ws_bytes = int(lib.tim_nms_workspace_bytes(N))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        if nms == "soft":
            _lib.check(lib.tim_softnms_1d(_ptr(g_segs), _ptr(g_scores), _ptr(offs), G, N, float(iou_threshold), float(sigma),
                                          float(min_score), int(method), _ptr(dets), _ptr(inds), _ptr(kept), _ptr(ws), ws_bytes, stream),
                       None)
        else:
            _lib.check(lib.tim_nms_1d(_ptr(g_segs), _ptr(g_scores), _ptr(offs), G, N, float(iou_threshold), float(min_score),
                                      int(max_seg_num), _ptr(dets), _ptr(inds), _ptr(kept), _ptr(ws), ws_bytes, stream), None)
        group_of_row = torch.repeat_interleave(torch.arange(G, device=device), counts, output_size=N)
        first = offs[group_of_row]
        live = (torch.arange(N, device=device) - first) < kept[group_of_row].to(torch.int64)
        out = dets[live]
        src = perm[first[live] + inds[live]]
        return out[:, :2].contiguous(), out[:, 2].contiguous(), uniq[group_of_row[live]], src


def batched_nms(segs, scores, cls_idxs, iou_threshold, min_score, sigma=0.5, method=2, nms="soft", multi_class=True,
                voting_thresh=0.75, max_seg_num=2000000, device=None):
    """Drop-in for detection/eval_detection/nms.py: batched_nms (:98-180), multi-class mode (what format_predictions_epic.py
    calls): NMS per class, then one descending sort by score; returns numpy arrays (segs [k,2], scores [k], cls [k]) as the
    reference does. Equal final scores come out in a stable order (class ascending, then pick order); the reference's
    unstable torch.sort leaves their order unspecified. The class-agnostic mode (multi_class=False) is not provided: the
    reference's own soft branch of it cannot run (nms.py:158-161 passes 8 arguments to a 7-argument op)."""
    if not multi_class:
        raise NotImplementedError("class-agnostic NMS (multi_class=False, segment voting) is outside this path")
    s, p, c, _ = grouped_nms(torch.as_tensor(segs), torch.as_tensor(scores), torch.as_tensor(cls_idxs), iou_threshold=iou_threshold,
                             min_score=min_score, sigma=sigma, method=method, nms=nms, max_seg_num=max_seg_num, device=device)
    if s.shape[0]:
        order = torch.sort(p, descending=True, stable=True)[1]
        s, p, c = s[order], p[order], c[order]
    return s.cpu().numpy(), p.cpu().numpy(), c.cpu().numpy()


def batched_nms_videos(segs, scores, cls_idxs, video_idxs, iou_threshold, min_score, sigma=0.5, method=2, nms="soft",
                       max_seg_num=2000000, device=None):
    """batched_nms for all videos of an evaluation at once (the reference runs one filter_nms job per video,
    format_predictions_epic.py:51-112,146-157): groups are (video, class). Returns device tensors (segs, scores, cls, video),
    videos in ascending index, inside a video by descending score."""
    cls_idxs = torch.as_tensor(cls_idxs).to(torch.int64)
    video_idxs = torch.as_tensor(video_idxs).to(torch.int64)
    if cls_idxs.numel() == 0:
        s, p, k, _ = grouped_nms(torch.as_tensor(segs), torch.as_tensor(scores), cls_idxs, iou_threshold=iou_threshold,
                                 min_score=min_score, device=device)
        return s, p, k, k.clone()
    if int(cls_idxs.min()) < 0 or int(video_idxs.min()) < 0:
        raise ValueError("class and video indices must be non-negative")
    span = int(cls_idxs.max()) + 1
    if (int(video_idxs.max()) + 1) * span >= 2 ** 62:
        raise ValueError("video * class key does not fit in int64")
    keys = video_idxs.to(cls_idxs.device) * span + cls_idxs
    s, p, k, _ = grouped_nms(torch.as_tensor(segs), torch.as_tensor(scores), keys, iou_threshold=iou_threshold, min_score=min_score,
                             sigma=sigma, method=method, nms=nms, max_seg_num=max_seg_num, device=device)
    vid, cls = torch.div(k, span, rounding_mode="floor"), k % span
    if s.shape[0]:
        order = torch.sort(p, descending=True, stable=True)[1]
        order = order[torch.sort(vid[order], stable=True)[1]]
        s, p, cls, vid = s[order], p[order], cls[order], vid[order]
    return s, p, cls, vid


# ---------------------------------------------------------------------------------------------------------------------
# in front of the NMS: FeatureMeter.update and the thresholding loop of format_predictions.py, on the device
# ---------------------------------------------------------------------------------------------------------------------
def decode_predictions(logits: torch.Tensor, regressions: torch.Tensor, window_start: torch.Tensor, window_size: float, max_time: float,
                       device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """FeatureMeter.update for one modality (detection/time_interval_machine/utils/meters.py:652-724) without the host round
    trip: logits [R, C] and regression outputs [R, 2] of a batch of B windows (R = B * queries, window-major, as the heads return
    them), window_start [B] seconds -> (preds [R, C] fp32 = sigmoid(logits), proposals [R, 2] float64 in seconds of the video =
    clamp(reg, 0, max_time) * window_size + window_start of the row's window). max_time is the largest query time of the batch
    (meters.py:673), window_size the dataset's window length in seconds (:669)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else (logits.device if logits.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if device.type != "cuda":
        raise RuntimeError("tim_b200.postprocess runs on a CUDA device only (no CPU fallback)")
    logits = logits.to(device=device, dtype=torch.float32).contiguous()
    reg = regressions.to(device=device, dtype=torch.float32).reshape(-1, 2).contiguous()
    ws = torch.as_tensor(window_start).to(device=device, dtype=torch.float64).reshape(-1).contiguous()
    R, C_ = int(logits.shape[0]), int(logits.shape[1])
    if reg.shape[0] != R or ws.numel() == 0 or R % ws.numel():
        raise ValueError("logits [R,C], regressions [R,2] and window_start [B] with R = B * queries expected")
    with torch.cuda.device(device):
        preds = torch.empty((R, C_), dtype=torch.float32, device=device)
        props = torch.empty((R, 2), dtype=torch.float64, device=device)
        _lib.check(lib.tim_det_decode(_ptr(logits), _ptr(reg), _ptr(ws), R // int(ws.numel()), R, C_, float(window_size), float(max_time),
                                      _ptr(preds), _ptr(props), torch.cuda.current_stream(device).cuda_stream), None)
    return preds, props


def threshold_detections(preds: torch.Tensor, proposals: torch.Tensor, score_threshold: float,
                         device: Optional[torch.device] = None):
    """The loop of detection/eval_detection/format_predictions.py:103-125 (format_predictions_epic.py:120-143) on the device:
    proposals [R, 2] (float64 seconds) rounded to 3 decimals, empty ones dropped, one detection per (proposal, class) with
    preds[r, c] > score_threshold, in (proposal, class) order. Returns device tensors (row [n] int64, cls [n] int64, score [n] fp32,
    segs [n, 2] fp32 - what torch.FloatTensor(segs) gives the reference's NMS)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else (preds.device if preds.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if device.type != "cuda":
        raise RuntimeError("tim_b200.postprocess runs on a CUDA device only (no CPU fallback)")
    preds = preds.to(device=device, dtype=torch.float32).contiguous()
    props = proposals.to(device=device, dtype=torch.float64).reshape(-1, 2).contiguous()
    R, C_ = int(preds.shape[0]), int(preds.shape[1])
    if props.shape[0] != R:
        raise ValueError("preds [R,C] and proposals [R,2] must agree on R")
    thr = float(score_threshold)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        counts = torch.zeros((R,), dtype=torch.int32, device=device)
        if R:
            _lib.check(lib.tim_det_count(_ptr(preds), _ptr(props), R, C_, thr, _ptr(counts), stream), None)
        ends = torch.cumsum(counts, 0, dtype=torch.int64)
        n = int(ends[-1]) if R else 0                       # the one host read: the output size
        row = torch.empty((n,), dtype=torch.int64, device=device)
        cls = torch.empty((n,), dtype=torch.int64, device=device)
        score = torch.empty((n,), dtype=torch.float32, device=device)
        segs = torch.empty((n, 2), dtype=torch.float32, device=device)
        if n:
            offs = (ends - counts).contiguous()
            _lib.check(lib.tim_det_emit(_ptr(preds), _ptr(props), R, C_, thr, _ptr(offs), _ptr(row), _ptr(cls), _ptr(score), _ptr(segs), stream), None)
    return row, cls, score, segs


def format_predictions(preds, proposals, video_idxs, score_threshold=0.03, sigma=0.25, iou_threshold=0.1, min_score=0.001, method=2,
                       nms="soft", device=None):
    """Thresholding + per-video, per-class NMS of a whole evaluation in three launches: what main() of
    detection/eval_detection/format_predictions.py:98-141 computes with a Python loop over proposals and a joblib pool over
    videos. preds [R, C] (sigmoid scores), proposals [R, 2] (seconds), video_idxs [R] (integer id of each proposal's video).
    Returns device tensors (segs [k,2], scores [k], cls [k], video [k]): videos ascending, inside a video by descending score
    (the order of `results[video]` in the reference's submission file)."""
    row, cls, score, segs = threshold_detections(preds, proposals, score_threshold, device=device)
    vid = torch.as_tensor(video_idxs).to(device=row.device, dtype=torch.int64)[row]
    return batched_nms_videos(segs, score, cls, vid, iou_threshold=iou_threshold, min_score=min_score, sigma=sigma, method=method,
                              nms=nms, device=row.device)
