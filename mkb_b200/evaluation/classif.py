"""Triplet classification; mirror of mkb/evaluation/classif.py (find_threshold, accuracy).

Scores come from ``utils.make_prediction`` (the positives kernel); the threshold search is the ROC
criterion the reference takes from scikit-learn — the score value maximising TPR - FPR, first maximum in
descending-threshold order — computed directly with numpy.
"""
import numpy as np
import torch

from ..utils import make_prediction

__all__ = ["find_threshold", "accuracy"]


def _scores(model, X, batch_size, num_workers, device):
    with torch.no_grad():
        return make_prediction(model=model, dataset=X, batch_size=batch_size, num_workers=num_workers,
                               device=device).cpu().numpy()


def best_threshold(y_true, y_score):
    """Threshold t maximising TPR(t) - FPR(t) over the distinct scores (prediction = score >= t); ties go to
    the largest threshold, as ``roc_curve`` + ``argmax`` yield (classif.py:100-114)."""
    y_true = np.asarray(y_true)
    y_score = np.asarray(y_score)
    order = np.argsort(-y_score, kind="stable")
    s, pos = y_score[order], (y_true[order] > 0)
    last = np.r_[np.nonzero(np.diff(s))[0], s.size - 1]  # last index of every run of equal scores
    tp, fp = np.cumsum(pos)[last], np.cumsum(~pos)[last]
    n_pos, n_neg = max(int(pos.sum()), 1), max(int((~pos).sum()), 1)
    j = tp / n_pos - fp / n_neg
    best = int(np.argmax(j))
    if j[best] <= 0:  # roc_curve's first point (nothing predicted positive) wins: its threshold is +inf
        return np.inf
    return s[last][best]


def find_threshold(model, X, y, batch_size, num_workers=1, device="cuda"):
    """Best score threshold for "this triple exists" on (X, y) with y > 0 for true triples."""
    return best_threshold(y, _scores(model, X, batch_size, num_workers, device))


def _accuracy(y_pred, y_true, threshold):
    """Share of triples on the right side of the threshold (classif.py:117-140)."""
    y_pred, y_true = np.asarray(y_pred), np.asarray(y_true)
    return float((((y_pred >= threshold) & (y_true > 0)) | ((y_pred < threshold) & (y_true <= 0))).sum() / len(y_pred))


def accuracy(model, X, y, threshold, batch_size, num_workers=1, device="cuda"):
    return _accuracy(y_pred=_scores(model, X, batch_size, num_workers, device), y_true=y, threshold=threshold)
