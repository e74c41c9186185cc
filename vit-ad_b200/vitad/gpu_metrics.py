"""Detection / localisation metrics of ValidationHelper.calc_all_metrics (src/util/ValidationHelper.py:131-211)
computed on the device: ONE sort of the scores per (scores, labels) pair, cumulative true/false-positive counts at the
distinct thresholds, and the same trapezoid / threshold rules sklearn applies (roc_curve incl. drop_intermediate,
roc_auc_score, precision_recall_curve + auc), with ties handled as sklearn does (one curve point per distinct score).
Over N x 50 176 pixels per category the reference's sklearn calls are the end-to-end bottleneck once scoring is fast
(SURVEY.md §8 f2).

The reference sorts the pixel scores three times per category (pixel AUROC, calc_threshold's roc_curve, and the AUROC of
the thresholded map, :70-88,91-104,141-175).  The thresholded map `where(s > thr, s, 0)` keeps the order of the scores
above the threshold and ties everything else at 0, so its ROC curve is the head of the first curve plus one closing
point: all three metrics come from the same sorted curve here.

The sort / scan / compaction are library primitives reached through torch (CUB radix sort, scan); counts are int64 and
areas float64, so the results agree with sklearn to rounding.  Functions take 1-D tensors on any device (the CPU suite
checks them against sklearn); `calc_all_metrics_device` is the validator-facing entry and requires CUDA.
"""
from __future__ import annotations

import numpy as np
import torch


class Curve:
    """sklearn's _binary_clf_curve: (fps, tps, thresholds) at the distinct scores, descending thresholds."""

    def __init__(self, scores: torch.Tensor, labels: torch.Tensor):
        s, order = torch.sort(scores.reshape(-1).to(torch.float32), descending=True)
        y = (labels.reshape(-1)[order] != 0).to(torch.int64)
        tps_all = torch.cumsum(y, 0)
        n = s.numel()
        boundary = torch.ones(n, dtype=torch.bool, device=s.device)
        boundary[:-1] = s[:-1] != s[1:]
        idx = boundary.nonzero().squeeze(1)
        self.tps = tps_all[idx]
        self.fps = idx + 1 - self.tps
        self.thr = s[idx]
        self.n_pos, self.n_neg = int(self.tps[-1]), int(self.fps[-1])


def _binary_clf_curve(scores: torch.Tensor, labels: torch.Tensor):
    c = Curve(scores, labels)
    return c.fps, c.tps, c.thr


def _trapz(x: torch.Tensor, y: torch.Tensor) -> float:
    x, y = x.to(torch.float64), y.to(torch.float64)
    return float(((x[1:] - x[:-1]) * (y[1:] + y[:-1])).sum() * 0.5)


def _roc_auc(fps: torch.Tensor, tps: torch.Tensor) -> float:
    n_pos, n_neg = int(tps[-1]), int(fps[-1])
    if n_pos == 0 or n_neg == 0:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    # the trapezoid area in exact integer arithmetic: 2 * area * P * N = sum (fps_i - fps_{i-1}) * (tps_i + tps_{i-1}) with
    # (fps_0, tps_0) = (0, 0); the sum is < 2 * P * N <= 2^63 for any tensor that fits a GPU, and one division yields what
    # sklearn's float64 trapezoid sums to within an ulp — in two passes over the curve instead of six float64 ones
    df = torch.diff(fps, prepend=fps.new_zeros(1))
    st = tps + torch.cat((tps.new_zeros(1), tps[:-1]))
    return float(int((df * st).sum())) / (2.0 * n_pos * n_neg)


def roc_auc_score(scores: torch.Tensor, labels: torch.Tensor, curve: Curve | None = None) -> float:
    """metrics.roc_auc_score(y_true=labels, y_score=scores) for binary labels."""
    c = Curve(scores, labels) if curve is None else curve
    return _roc_auc(c.fps, c.tps)


def pr_auc_score(scores: torch.Tensor, labels: torch.Tensor, curve: Curve | None = None) -> float:
    """metrics.auc(y=precision, x=recall) of metrics.precision_recall_curve(labels, scores)
    (ValidationHelper.py:176-179): the curve ends in (recall 0, precision 1)."""
    c = Curve(scores, labels) if curve is None else curve
    fps, tps = c.fps, c.tps
    ps = (tps + fps).to(torch.float64)
    precision = torch.where(ps > 0, tps.to(torch.float64) / ps.clamp_min(1.0), torch.zeros_like(ps))
    recall = tps.to(torch.float64) / float(tps[-1]) if int(tps[-1]) > 0 else torch.ones_like(ps)
    one = torch.ones(1, dtype=torch.float64, device=ps.device)
    # sklearn reverses both arrays (recall decreasing) and appends (1, 0); the area is direction-corrected by auc()
    precision = torch.cat((precision.flip(0), one))
    recall = torch.cat((recall.flip(0), torch.zeros_like(one)))
    return abs(_trapz(recall, precision))


def calc_threshold(scores: torch.Tensor, labels: torch.Tensor, fpr_threshold: float = 0.3, curve: Curve | None = None) -> float:
    """ValidationHelper.calc_threshold (:70-88) over metrics.roc_curve(drop_intermediate=True): of the KEPT curve points
    with fpr <= fpr_threshold take the maximal tpr and return the threshold of the first point reaching it (np.argmax).
    roc_curve keeps the first and last point and every point where the second difference of fps or of tps is non-zero
    (a corner); a point in the middle of a straight segment — e.g. a run of ties with a constant positive share — is
    dropped, and with it its threshold, so the rule has to be applied to the kept points only.  The leading
    (fpr 0, tpr 0) point roc_curve prepends has threshold inf."""
    c = Curve(scores, labels) if curve is None else curve
    fps, tps, thr = c.fps, c.tps, c.thr
    n = fps.numel()
    keep = torch.ones(n, dtype=torch.bool, device=fps.device)
    if n > 2:
        d2f = fps[2:] - 2 * fps[1:-1] + fps[:-2]
        d2t = tps[2:] - 2 * tps[1:-1] + tps[:-2]
        keep[1:-1] = (d2f != 0) | (d2t != 0)
    fpr = fps.to(torch.float64) / float(c.n_neg) if c.n_neg > 0 else torch.full((n,), float("nan"), dtype=torch.float64,
                                                                                  device=fps.device)
    ok = (keep & (fpr <= fpr_threshold)).nonzero().squeeze(1)
    if ok.numel() == 0 or int(tps[ok[-1]]) == 0:
        return float("inf")  # the prepended (0, 0, inf) point wins the argmax
    best = tps[ok[-1]]  # tps is non-decreasing along the curve
    first = ok[(tps[ok] == best).nonzero()[0]]
    return float(thr[int(first)])


def thresholded_roc_auc(curve: Curve, thr: float) -> float:
    """roc_auc_score of predict_anomaly(scores, thr, "fluently") = where(scores > thr, scores, 0) (:91-104,166-175) from
    the curve of the raw scores: its distinct-score points with threshold > max(thr, 0) are unchanged (same order, same
    ties), every remaining sample ties at 0 and closes the curve in one step.  Scores are >= 0 on this path (anomaly
    maps); negative scores above a negative threshold would sort below the zeros and are handled by the generic path."""
    if not thr >= 0.0:
        raise ValueError("thresholded_roc_auc expects a non-negative threshold (anomaly maps are >= 0)")
    head = int((curve.thr > thr).sum())  # thr descending: the first `head` points survive
    if head == curve.thr.numel():
        return _roc_auc(curve.fps, curve.tps)
    last = curve.fps.new_tensor([curve.n_neg]), curve.tps.new_tensor([curve.n_pos])
    return _roc_auc(torch.cat((curve.fps[:head], last[0])), torch.cat((curve.tps[:head], last[1])))


def calc_all_metrics_device(result: dict, fp_thres: float, dataset_name: str = "", device=None) -> dict:
    """Same keys as vitad.metrics.calc_all_metrics / the reference's W&B log (:196-208), computed on the GPU.
    `result` holds numpy arrays (validator output) or device tensors (valid_loop_*(on_device=True), gather_results)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("calc_all_metrics_device needs a CUDA device (use vitad.metrics for the sklearn path)")
        device = torch.device("cuda", torch.cuda.current_device())

    def dev(a):
        return a.to(device) if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a)).to(device)

    out = {"dataset": dataset_name, "fp_thres": fp_thres}
    il, isc = dev(result["image_labels"]).reshape(-1), dev(result["image_scores"]).reshape(-1)
    if not bool(torch.isfinite(isc).all()):
        raise ValueError("image scores contain NaN/Inf (sklearn's roc_auc_score raises on such input as well)")
    if il.numel() > 1:
        c = Curve(isc, il)
        if c.n_pos > 0 and c.n_neg > 0:
            out["image_auroc_score"] = roc_auc_score(isc, il, c)
            out["image_prauc_score"] = pr_auc_score(isc, il, c)
    pl, ps = dev(result["pixel_labels"]).reshape(-1), dev(result["pixel_scores"]).reshape(-1).to(torch.float32)
    if pl.numel() > 1:
        c = Curve(ps, pl)
        if c.n_pos > 0 and c.n_neg > 0:
            out["pixel_auroc_score"] = roc_auc_score(ps, pl, c)
            thr = calc_threshold(ps, pl, fp_thres, c)
            if thr >= 0.0 and float(c.thr[-1]) >= 0.0:
                out[f"pro_score_{fp_thres}fp"] = thresholded_roc_auc(c, thr)
            else:  # negative scores: the zeros do not sort last, fall back to the literal definition
                anomalies = torch.where(ps > thr, ps, torch.zeros_like(ps))  # predict_anomaly(..., "fluently") (:91-104)
                out[f"pro_score_{fp_thres}fp"] = roc_auc_score(anomalies, pl)
    return out
