"""CPU restatement of the data formats either side of the network.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy / pure Python, every function cites the reference lines it
follows.  Pinned by ``oracle/make_golden.py`` against the UNMODIFIED reference functions executed in the build container:
  * run_length_encoding / run_length_decoding        <- common_blocks/utils.py:99-132 (imported as is)
  * the tile adapter                                 <- torchvision Grayscale/ToTensor/Normalize + common_blocks/utils.py:494-500
                                                        AddDepthChannels + utils.py:308-313 get_crop_pad_sequence (imported as
                                                        is); the 'edge' pad itself is imgaug's iaa.Pad (package absent), which is
                                                        np.pad(mode='edge') - restated, "parity unpinned" for that one call
  * crop_image / binarize                            <- common_blocks/postprocessing.py:24-43: skimage is absent so the module is
                                                        imported with a stub for its unused `resize` import
  * compute_ious .. intersection_over_union_thresholds <- common_blocks/metrics.py:8-68 (imported as is); pycocotools is absent, its
                                                        mask.iou/encode are replaced by a 10-line dense-mask stub in
                                                        oracle/ref_shims.py ("parity unpinned" for the IoU primitive itself,
                                                        which is |a & b| / |a | b| for iscrowd = 0)
Fixtures: tests/golden/io_cases.npz.
"""
import numpy as np

from .synth import MEAN, STD, adapt_tiles  # noqa: F401  (adapt_tiles: loaders.py:607-612, utils.py:494-500, augmentation.py:247-281)

IOUT_THRESHOLDS = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]   # metrics.py:50


def get_crop_pad_sequence(vertical, horizontal):
    """utils.py:308-313 -> (top, right, bottom, left)."""
    top = int(vertical / 2)
    bottom = vertical - top
    right = int(horizontal / 2)
    left = horizontal - right
    return top, right, bottom, left


def adapt_tiles_hflip(tiles_u8, out_size=128):
    """TTA h-flip of the raw tile (augmentation.py:143-147) followed by the adapter."""
    return adapt_tiles(np.ascontiguousarray(tiles_u8[:, :, ::-1]), out_size)


def run_length_encoding(x):
    """utils.py:99-111: column-major flat order, pixels numbered from 1, [start, length, ...]."""
    flat = np.asarray(x).T.flatten() != 0
    rle = []
    prev = False
    for i, b in enumerate(flat):
        if b and not prev:
            rle.extend((i + 1, 0))
        if b:
            rle[-1] += 1
        prev = b
    return rle


def run_length_decoding(rle, shape):
    """utils.py:114-132 on the list form: -> uint8 mask (H, W)."""
    img = np.zeros(shape[0] * shape[1], dtype=np.uint8)
    for start, length in zip(rle[0::2], rle[1::2]):
        img[start - 1:start - 1 + length] = 1
    return img.reshape((shape[1], shape[0])).T


def sigmoid(x):
    """utils.py:173-174."""
    return 1. / (1 + np.exp(-x))


def crop_image(image, target_size):
    """postprocessing.py:24-38."""
    top, right, bottom, left = get_crop_pad_sequence(image.shape[1] - target_size[0], image.shape[2] - target_size[1])
    return image[:, top:image.shape[1] - bottom, left:image.shape[2] - right]


def binarize(image, threshold):
    """postprocessing.py:41-43."""
    return (image[1, :, :] > threshold).astype(np.uint8)


def compute_iou_single(gt, pred):
    """metrics.py:21-35 compute_ious for masks holding at most one object (labels 0/1) -> scalar IoU."""
    g, p = gt != 0, pred != 0
    if not g.any() and not p.any():
        return 1.0
    if not g.any() or not p.any():
        return 0.0
    return float(np.logical_and(g, p).sum()) / float(np.logical_or(g, p).sum())


def compute_eval_metric_single(gt, pred):
    """metrics.py:38-52: one object per side -> precision(t) = [IoU >= t]."""
    iou = compute_iou_single(gt, pred)
    return sum(1.0 if iou >= th else 0.0 for th in IOUT_THRESHOLDS) / len(IOUT_THRESHOLDS)


def validation_sweep(logits, y_true, target=101, thresholds=None):
    """callbacks.py:499-527 _get_validation_loss on raw network outputs: sigmoid (callbacks.py:571), crop + binarize per
    threshold (callbacks.py:832-866), IoUT with early stopping, then IoU / IoUT at the best threshold.
    logits: fp32 [B,2,S,S]; y_true: list/array of uint8 [target,target].  -> dict(threshold, iout, iou, iouts_seen)."""
    thresholds = np.linspace(0.5, 0.3, 21) if thresholds is None else thresholds
    probs = [sigmoid(np.squeeze(m)) for m in logits]

    def predict(thr):
        return [binarize(crop_image(p, (target, target)), thr) for p in probs]

    iout_best, threshold_best, seen = 0.0, 0.5, []
    for thr in thresholds:
        y_pred = predict(thr)
        iout = float(np.mean([compute_eval_metric_single(t, p) for t, p in zip(y_true, y_pred)]))
        seen.append(iout)
        if iout > iout_best:
            iout_best, threshold_best = iout, thr
        else:
            break
    y_pred = predict(threshold_best)
    iout = float(np.mean([compute_eval_metric_single(t, p) for t, p in zip(y_true, y_pred)]))
    iou = float(np.mean([compute_iou_single(t, p) for t, p in zip(y_true, y_pred)]))
    return {'threshold': float(threshold_best), 'iout': iout, 'iou': iou, 'iouts_seen': seen}


def validation_counts(logits, y_true, target=101, thresholds=None, logits_flip=None):
    """What salt_validation_counts returns: inter [B,K], pred [B,K], gtsum [B] (int64)."""
    thresholds = np.linspace(0.5, 0.3, 21) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    b = len(logits)
    inter = np.zeros((b, len(thresholds)), np.int64)
    pred = np.zeros((b, len(thresholds)), np.int64)
    gts = np.zeros((b,), np.int64)
    for i in range(b):
        p = sigmoid(np.asarray(logits[i], dtype=np.float32))
        if logits_flip is not None:
            pf = sigmoid(np.asarray(logits_flip[i], dtype=np.float32))[:, :, ::-1]       # augmentation.py:155-176 un-flip
            p = (p + pf) / np.float32(2)                                                 # loaders.py:751-760 mean
        p = crop_image(p, (target, target))
        g = np.asarray(y_true[i]) != 0
        gts[i] = g.sum()
        for k, thr in enumerate(thresholds):
            m = p[1] > thr
            pred[i, k] = m.sum()
            inter[i, k] = np.logical_and(m, g).sum()
    return inter, pred, gts
