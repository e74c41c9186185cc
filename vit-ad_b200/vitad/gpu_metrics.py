"""Detection / localisation metrics of ValidationHelper.calc_all_metrics (src/util/ValidationHelper.py:131-211)
computed on the device: one sort of the scores, cumulative true/false-positive counts at the distinct thresholds, and
the same trapezoid / threshold rules sklearn applies (roc_curve, roc_auc_score, precision_recall_curve + auc), with ties
handled as sklearn does (one curve point per distinct score).  Over N x 50 176 pixels per category the reference's
sklearn calls are the end-to-end bottleneck once scoring is fast.

The sort / scan / compaction are library primitives reached through torch (CUB radix sort, scan); counts are int64 and
areas float64, so the results agree with sklearn to rounding.  Functions take 1-D tensors on any device (the CPU suite
checks them against sklearn); `calc_all_metrics_device` is the validator-facing entry and requires CUDA.
"""
from __future__ import annotations

import numpy as np
import torch


def _binary_clf_curve(scores: torch.Tensor, labels: torch.Tensor):
    """sklearn.metrics._ranking._binary_clf_curve: (fps, tps, thresholds) at the distinct scores, descending."""
    s, order = torch.sort(scores.reshape(-1).to(torch.float32), descending=True)
    y = (labels.reshape(-1)[order] != 0).to(torch.int64)
    tps_all = torch.cumsum(y, 0)
    n = s.numel()
    boundary = torch.ones(n, dtype=torch.bool, device=s.device)
    boundary[:-1] = s[:-1] != s[1:]
    idx = boundary.nonzero().squeeze(1)
    tps = tps_all[idx]
    fps = idx + 1 - tps
    return fps, tps, s[idx]


def _trapz(x: torch.Tensor, y: torch.Tensor) -> float:
    x, y = x.to(torch.float64), y.to(torch.float64)
    return float(((x[1:] - x[:-1]) * (y[1:] + y[:-1])).sum() * 0.5)


def roc_auc_score(scores: torch.Tensor, labels: torch.Tensor) -> float:
    """metrics.roc_auc_score(y_true=labels, y_score=scores) for binary labels."""
    fps, tps, _ = _binary_clf_curve(scores, labels)
    if int(tps[-1]) == 0 or int(fps[-1]) == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    zero = torch.zeros(1, dtype=fps.dtype, device=fps.device)
    fpr = torch.cat((zero, fps)).to(torch.float64) / float(fps[-1])
    tpr = torch.cat((zero, tps)).to(torch.float64) / float(tps[-1])
    return _trapz(fpr, tpr)


def pr_auc_score(scores: torch.Tensor, labels: torch.Tensor) -> float:
    """metrics.auc(y=precision, x=recall) of metrics.precision_recall_curve(labels, scores)
    (ValidationHelper.py:176-179): the curve ends in (recall 0, precision 1)."""
    fps, tps, _ = _binary_clf_curve(scores, labels)
    ps = (tps + fps).to(torch.float64)
    precision = torch.where(ps > 0, tps.to(torch.float64) / ps.clamp_min(1.0), torch.zeros_like(ps))
    recall = tps.to(torch.float64) / float(tps[-1]) if int(tps[-1]) > 0 else torch.ones_like(ps)
    one = torch.ones(1, dtype=torch.float64, device=ps.device)
    # sklearn reverses both arrays (recall decreasing) and appends (1, 0); the area is direction-corrected by auc()
    precision = torch.cat((precision.flip(0), one))
    recall = torch.cat((recall.flip(0), torch.zeros_like(one)))
    return abs(_trapz(recall, precision))


def calc_threshold(scores: torch.Tensor, labels: torch.Tensor, fpr_threshold: float = 0.3) -> float:
    """ValidationHelper.calc_threshold (:70-88): of the ROC points with fpr <= fpr_threshold take the maximal tpr and
    return the threshold of the FIRST point reaching it (np.argmax), i.e. the highest such threshold.  roc_curve's
    leading (fpr 0, tpr 0) point has threshold inf; its drop_intermediate only removes collinear points, never the
    first point of a tpr level."""
    fps, tps, thr = _binary_clf_curve(scores, labels)
    fpr = fps.to(torch.float64) / max(float(fps[-1]), 1.0)
    ok = (fpr <= fpr_threshold).nonzero().squeeze(1)
    if ok.numel() == 0 or int(tps[ok[-1]]) == 0:
        return float("inf")
    best = tps[ok[-1]]  # tps is non-decreasing along the curve
    first = int((tps == best).nonzero()[0])
    return float(thr[first])


def calc_all_metrics_device(result: dict, fp_thres: float, dataset_name: str = "", device=None) -> dict:
    """Same keys as vitad.metrics.calc_all_metrics / the reference's W&B log (:196-208), computed on the GPU.
    `result` holds numpy arrays (validator output) or device tensors."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("calc_all_metrics_device needs a CUDA device (use vitad.metrics for the sklearn path)")
        device = torch.device("cuda", torch.cuda.current_device())

    def dev(a):
        return a.to(device) if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).to(device)

    out = {"dataset": dataset_name, "fp_thres": fp_thres}
    il, isc = dev(result["image_labels"]).reshape(-1), dev(result["image_scores"]).reshape(-1)
    if int((il != 0).any()) and int((il == 0).any()):
        out["image_auroc_score"] = roc_auc_score(isc, il)
        out["image_prauc_score"] = pr_auc_score(isc, il)
    pl, ps = dev(result["pixel_labels"]).reshape(-1), dev(result["pixel_scores"]).reshape(-1).to(torch.float32)
    if pl.numel() > 1 and int((pl != 0).any()) and int((pl == 0).any()):
        out["pixel_auroc_score"] = roc_auc_score(ps, pl)
        thr = calc_threshold(ps, pl, fp_thres)
        anomalies = torch.where(ps > thr, ps, torch.zeros_like(ps))  # predict_anomaly(..., "fluently") (:91-104)
        out[f"pro_score_{fp_thres}fp"] = roc_auc_score(anomalies, pl)
    return out
