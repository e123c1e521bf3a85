"""Validation scoring on the GPU: ``ValidationMonitor._get_validation_loss``
(callbacks.py:499-527) = sigmoid -> crop -> binarize at 21 thresholds -> IoU / IoUT (metrics.py:8-64) with an early-stopping
sweep.  One kernel (`salt_validation_counts`) reads each logit map once and produces, per image and per threshold, the
intersection and prediction pixel counts; everything after that is O(images x thresholds) host arithmetic that follows the
reference line by line.

For the binary masks of this competition ``get_segmentations`` yields at most one object per mask, so metrics.py reduces to
  both empty -> IoU 1;  exactly one empty -> IoU 0;  else |gt & pred| / |gt | pred|
  precision_at(t) = 1 if IoU >= t else 0  (tp / (tp + fp + fn) with one object on each side).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

IOUT_THRESHOLDS = (0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95)     # metrics.py:50
SWEEP_THRESHOLDS = np.linspace(0.5, 0.3, 21)                                  # callbacks.py:503


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def validation_counts(logits, gt, thresholds=SWEEP_THRESHOLDS, logits_flip=None):
    """logits fp32 CUDA [B,C,S,S] (class 1 = salt), gt uint8 CUDA [B,T,T] -> int32 CUDA tensors
    inter [B,K], pred [B,K], gtsum [B]."""
    if not torch.cuda.is_available():
        raise _lib.SaltEngineError('needs a CUDA device (B200); there is no CPU fallback')
    lib = _lib.load()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 4
    gt = gt.contiguous()
    assert gt.is_cuda and gt.dtype == torch.uint8 and gt.dim() == 3 and gt.shape[0] == logits.shape[0]
    b, k, s, _ = logits.shape
    t = gt.shape[1]
    thr = np.ascontiguousarray(np.asarray(thresholds, dtype=np.float64))
    n = int(thr.size)
    inter = torch.empty((b, n), dtype=torch.int32, device=logits.device)
    pred = torch.empty((b, n), dtype=torch.int32, device=logits.device)
    gtsum = torch.empty((b,), dtype=torch.int32, device=logits.device)
    _lib.check(lib.salt_validation_counts(_ptr(logits), _ptr(logits_flip), b, k, s, t, _ptr(gt),
                                          thr.ctypes.data_as(C.POINTER(C.c_double)), n, _ptr(inter), _ptr(pred), _ptr(gtsum),
                                          C.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)))
    return inter, pred, gtsum


def ious_from_counts(inter, pred, gtsum):
    """Per-image IoU [B,K] following metrics.py:21-35 compute_ious for single-object masks."""
    inter = np.asarray(inter, dtype=np.float64)
    pred = np.asarray(pred, dtype=np.float64)
    gts = np.asarray(gtsum, dtype=np.float64)[:, None]
    union = pred + gts - inter
    iou = np.where(union > 0, inter / np.maximum(union, 1.0), 0.0)
    both_empty = (pred == 0) & (gts == 0)
    return np.where(both_empty, 1.0, iou)


def iout_from_ious(iou):
    """metrics.py:47-52 compute_eval_metric per image, then the mean of metrics.py:64-68."""
    thr = np.asarray(IOUT_THRESHOLDS, dtype=np.float64)
    prec = (iou[..., None] >= thr).mean(axis=-1)
    return prec.mean(axis=0)


class ValidationScorer:
    """Accumulates the counts of a validation epoch batch by batch (on the device) and reproduces the reference's
    best-threshold selection: thresholds walk from 0.5 down to 0.3 and stop at the first one that does not improve IoUT
    (callbacks.py:502-512)."""

    def __init__(self, thresholds=SWEEP_THRESHOLDS, dp=None):
        self.thresholds = np.asarray(thresholds, dtype=np.float64)
        self.dp = dp                 # salt_b200.dist.DataParallelContext: ranks score disjoint shards of the validation set
        self._parts = []

    def update(self, logits, gt, logits_flip=None):
        self._parts.append(validation_counts(logits, gt, self.thresholds, logits_flip))

    def counts(self):
        inter = torch.cat([p[0] for p in self._parts]).cpu().numpy()
        pred = torch.cat([p[1] for p in self._parts]).cpu().numpy()
        gts = torch.cat([p[2] for p in self._parts]).cpu().numpy()
        if self.dp is not None and self.dp.world > 1:
            inter, pred, gts = self.dp.gather_rows(inter, pred, gts)
        return inter, pred, gts

    def result(self):
        return select_threshold(*self.counts(), thresholds=self.thresholds)


def select_threshold(inter, pred, gtsum, thresholds=SWEEP_THRESHOLDS):
    """-> {'threshold', 'iout', 'iou', 'iout_per_threshold'} as callbacks.py:499-520 computes them."""
    iou = ious_from_counts(inter, pred, gtsum)
    iout = iout_from_ious(iou)
    iout_best, k_best = 0.0, None
    for k in range(len(thresholds)):
        if iout[k] > iout_best:
            iout_best, k_best = float(iout[k]), k
        else:
            break
    if k_best is None:                       # callbacks.py:501: threshold_best starts at 0.5
        k_best = int(np.argmin(np.abs(np.asarray(thresholds) - 0.5)))
    return {'threshold': float(thresholds[k_best]), 'iout': float(iout[k_best]), 'iou': float(iou[:, k_best].mean()),
            'iout_per_threshold': iout}
