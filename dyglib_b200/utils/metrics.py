"""Link-prediction metrics next to the hot path: drop-in for ``utils/metrics.py``.

The reference copies every batch's predictions to the host and calls scikit-learn (``utils/metrics.py:5-20``, one device
synchronisation per batch, ``train_link_prediction.py:251-253``).  Here average precision and ROC-AUC are computed where the
tensors live (a sort and two prefix sums in float64, same definitions as ``sklearn.metrics.average_precision_score`` /
``roc_auc_score``: thresholds at the distinct scores, ties grouped), so a loop can keep per-batch results on the device
(``link_prediction_metrics_tensors``) and read them once per epoch.  torch ops only: plumbing, not a kernel of the path.
"""
from __future__ import annotations

import torch


def _curve_counts(predicts: torch.Tensor, labels: torch.Tensor):
    """Cumulative true / false positives at every distinct threshold (descending score), like sklearn's ``_binary_clf_curve``."""
    s = predicts.detach().reshape(-1).double()
    y = labels.detach().reshape(-1).double()
    order = torch.argsort(s, descending=True, stable=True)
    s, y = s[order], y[order]
    tps = torch.cumsum(y, 0)
    fps = torch.cumsum(1.0 - y, 0)
    last = torch.ones_like(s, dtype=torch.bool)      # last element of every run of equal scores
    last[:-1] = s[1:] != s[:-1]
    return tps[last], fps[last]


def link_prediction_metrics_tensors(predicts: torch.Tensor, labels: torch.Tensor):
    """(average_precision, roc_auc) as 0-d float64 tensors on the inputs' device (no host synchronisation)."""
    tps, fps = _curve_counts(predicts, labels)
    n_pos, n_neg = tps[-1], fps[-1]
    precision = tps / (tps + fps)
    recall = tps / n_pos
    prev_recall = torch.cat([recall.new_zeros(1), recall[:-1]])
    ap = torch.sum((recall - prev_recall) * precision)
    tpr = torch.cat([tps.new_zeros(1), tps / n_pos])
    fpr = torch.cat([fps.new_zeros(1), fps / n_neg])
    auc = torch.trapezoid(tpr, fpr)
    return ap, auc


def get_link_prediction_metrics(predicts: torch.Tensor, labels: torch.Tensor):
    """``get_link_prediction_metrics`` (``utils/metrics.py:5-20``): {'average_precision': float, 'roc_auc': float}."""
    ap, auc = link_prediction_metrics_tensors(predicts, labels)
    return {'average_precision': float(ap.item()), 'roc_auc': float(auc.item())}


def get_node_classification_metrics(predicts: torch.Tensor, labels: torch.Tensor):
    """``get_node_classification_metrics`` (``utils/metrics.py:23-36``)."""
    return {'roc_auc': float(link_prediction_metrics_tensors(predicts, labels)[1].item())}
