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

    def __init__(self, thresholds=SWEEP_THRESHOLDS, dp=None, counts_fn=None):
        self.thresholds = np.asarray(thresholds, dtype=np.float64)
        self.dp = dp                 # salt_b200.dist.DataParallelContext: ranks score disjoint shards of the validation set
        self._counts_fn = counts_fn or validation_counts     # (logits, gt, thresholds, logits_flip) -> inter, pred, gtsum tensors
        self._parts = []

    def update(self, logits, gt, logits_flip=None):
        self._parts.append(self._counts_fn(logits, gt, self.thresholds, logits_flip))

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


Y_COLUMN = 'file_path_mask'          # callbacks.py:25
ORIGINAL_SIZE = (101, 101)           # callbacks.py:26


def read_masks(masks_filepaths):
    """utils.py:82-88: mask PNG -> uint8 {0,1} array."""
    from PIL import Image
    masks = []
    for path in masks_filepaths:
        mask = Image.open(path)
        masks.append(np.asarray(mask.convert('L').point(lambda x: 0 if x < 128 else 1)).astype(np.uint8))
    return masks


class ValidationMonitor:
    """Drop-in for the reference's ``callbacks.ValidationMonitor`` (callbacks.py:455-527), same constructor arguments and callback
    surface (``set_params / on_train_begin / on_epoch_begin / on_epoch_end / on_batch_begin / on_batch_end / on_train_end /
    training_break / get_validation_loss``), writing the same ``transformer.validation_loss[epoch] = {'sum', 'iou', 'iout'}``.

    What changes underneath: ``_transform`` keeps the logits on the GPU, and the 21-threshold sweep (crop -> binarize ->
    IoUT, callbacks.py:499-520) is ONE counts kernel per validation batch (``salt_validation_counts``) plus O(images x thresholds)
    host arithmetic, instead of a device-to-host copy, a numpy sigmoid and up to 21 steppy post-processing pipelines with
    pycocotools calls per image.  ``y_true`` may be passed directly (list / array of uint8 [101,101] masks) instead of being read from
    ``meta_valid['file_path_mask']``."""

    def __init__(self, data_dir=None, loader_mode='resize_and_pad', epoch_every=None, batch_every=None, use_depth=False, y_true=None,
                 scorer_factory=None):
        self.epoch_every = False if epoch_every == 0 else epoch_every
        self.batch_every = False if batch_every == 0 else batch_every
        if loader_mode != 'resize_and_pad':
            raise NotImplementedError('only the crop post-processing of loader_mode resize_and_pad is implemented (callbacks.py:833-834)')
        if use_depth:
            raise NotImplementedError('use_depth models are outside the engine (callbacks.py:568-588)')
        self.data_dir, self.loader_mode, self.use_depth = data_dir, loader_mode, use_depth
        self.meta_valid, self.y_true = None, y_true
        self.epoch_id = self.batch_id = None
        self._scorer_factory = scorer_factory or (lambda dp: ValidationScorer(dp=dp))
        self.last_result = None

    # ---- callbacks.Callback surface (callbacks.py:29-75)
    def set_params(self, transformer, validation_datagen, meta_valid=None, *args, **kwargs):
        self.transformer = transformer
        self.model, self.optimizer = transformer.model, transformer.optimizer
        self.loss_function, self.output_names = transformer.loss_function, transformer.output_names
        self.activation_func = transformer.activation_func
        self.validation_datagen, self.meta_valid = validation_datagen, meta_valid
        if self.y_true is None and meta_valid is not None:
            self.y_true = read_masks(meta_valid[Y_COLUMN].values)

    def on_train_begin(self, *args, **kwargs):
        self.epoch_id, self.batch_id = 0, 0

    def on_train_end(self, *args, **kwargs):
        pass

    def on_epoch_begin(self, *args, **kwargs):
        pass

    def on_batch_begin(self, *args, **kwargs):
        pass

    def on_batch_end(self, *args, **kwargs):
        self.batch_id += 1

    def training_break(self, *args, **kwargs):
        return False

    def on_epoch_end(self, *args, **kwargs):
        if self.epoch_every and ((self.epoch_id % self.epoch_every) == 0):
            self.model.eval()
            self.get_validation_loss()
            self.model.train()
        self.epoch_id += 1

    def get_validation_loss(self):
        return self._get_validation_loss()

    # ---- callbacks.py:499-566
    def _get_validation_loss(self):
        if self.activation_func != 'sigmoid':
            raise Exception('Only softmax and sigmoid activations are allowed')        # callbacks.py:563 (softmax: no engine path)
        scorer = self._scorer_factory(getattr(self.transformer, 'dp', None))
        self.model.eval()
        batch_gen, steps = self.validation_datagen
        (name, loss_function_one, weight) = self.loss_function[0]
        partial_batch_losses, seen = [], 0
        for batch_id, data in enumerate(batch_gen):
            X, target = data[0], data[1]
            outputs_batch = self.model(X)
            partial_batch_losses.append(loss_function_one(outputs_batch, target).clone() * weight)
            b = int(outputs_batch.shape[0])
            gt = np.stack([np.asarray(m) for m in self.y_true[seen:seen + b]]).astype(np.uint8)
            if len(gt) != b:
                raise ValueError('validation masks (%d) do not cover the validation loader (batch %d at image %d)'
                                 % (len(self.y_true), b, seen))
            scorer.update(outputs_batch, torch.from_numpy(gt).to(outputs_batch.device))
            seen += b
            if batch_id == steps:
                break
        self.model.train()
        epoch_loss = sum(partial_batch_losses) / steps                                  # callbacks.py:552
        r = scorer.result()
        self.last_result = r
        if not self.transformer.validation_loss:
            self.transformer.validation_loss = {}
        self.transformer.validation_loss.setdefault(self.epoch_id, {'sum': epoch_loss,
                                                                    'iou': torch.Tensor([r['iou']]),
                                                                    'iout': torch.Tensor([r['iout']])})
        return self.transformer.validation_loss[self.epoch_id]
