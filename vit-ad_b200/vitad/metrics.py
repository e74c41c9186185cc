"""Detection / localisation metrics of the validators (src/util/ValidationHelper.py:42-211) without the
W&B and matplotlib side effects: the same sklearn calls, returned as a dict.  Host-side boundary consumer,
not part of the CUDA hot path (GPU metrics are listed as the next widening step in DESIGN.md)."""
from __future__ import annotations

import numpy as np


def calc_threshold(anomaly_map: np.ndarray, test_labels: np.ndarray, fpr_threshold: float = 0.3) -> float:
    """ValidationHelper.py:70-88: threshold with maximal TPR subject to FPR <= fpr_threshold."""
    from sklearn import metrics

    fpr, tpr, thresholds = metrics.roc_curve(y_true=test_labels, y_score=anomaly_map)
    idx = np.where(fpr <= fpr_threshold)
    return float(thresholds[np.argmax(tpr[idx])])


def create_heatmap_from_scores(anomaly_map: np.ndarray, pixel_labels: np.ndarray, fpr_threshold: float):
    """ValidationHelper.py:107-128: scores below the threshold are zeroed, the others kept."""
    thr = calc_threshold(anomaly_map.flatten(), pixel_labels.flatten(), fpr_threshold)
    return np.where(anomaly_map > thr, anomaly_map, 0)


def calc_all_metrics(result: dict, fp_thres: float, dataset_name: str = "") -> dict:
    """ValidationHelper.py:131-211 → {image_auroc_score, image_prauc_score, pixel_auroc_score, pro_score}."""
    from sklearn import metrics

    out = {"dataset": dataset_name, "fp_thres": fp_thres}
    il, isc = np.asarray(result["image_labels"]).ravel(), np.asarray(result["image_scores"]).ravel()
    if len(np.unique(il)) > 1:
        out["image_auroc_score"] = float(metrics.roc_auc_score(y_true=il, y_score=isc))
        precision, recall, _ = metrics.precision_recall_curve(il, isc)
        out["image_prauc_score"] = float(metrics.auc(y=precision, x=recall))
    pl, ps = np.asarray(result["pixel_labels"]).ravel(), np.asarray(result["pixel_scores"]).ravel()
    if len(np.unique(pl)) > 1:
        out["pixel_auroc_score"] = float(metrics.roc_auc_score(y_true=pl, y_score=ps))
        anomalies = create_heatmap_from_scores(np.asarray(result["pixel_scores"]), np.asarray(result["pixel_labels"]),
                                               fp_thres)
        out[f"pro_score_{fp_thres}fp"] = float(metrics.roc_auc_score(y_true=pl, y_score=anomalies.ravel()))
    return out
