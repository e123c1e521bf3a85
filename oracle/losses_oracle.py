"""CPU restatement of the loss / activation / post-processing arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference followed:
  common_blocks/lovasz_losses.py:21-33    lovasz_grad
  common_blocks/lovasz_losses.py:81-115   lovasz_hinge / lovasz_hinge_flat (ELU variant)
  common_blocks/lovasz_losses.py:133-145  flatten_binary_scores
  common_blocks/lovasz_losses.py:237-255  mean
  common_blocks/models.py:315-340,361-388 DiceLoss, mixed_dice_bce_loss, multiclass_dice_loss
  common_blocks/utils.py:173-174,308-320  sigmoid, get_crop_pad_sequence
  common_blocks/postprocessing.py:24-43   crop_image, binarize
  common_blocks/loaders.py:751-760        aggregate_augmentations (mean of probabilities)
  common_blocks/augmentation.py:155-176   inverse TTA (per-channel fliplr)
"""
import numpy as np
import torch
import torch.nn.functional as F


def lovasz_hinge_per_image(logits, target):
    """logits, target: [B,C,H,W] (target in {0,1}).  Each image is flattened over
    *all* of C*H*W (the reference zips over the batch dim only, lovasz_losses.py:89-91),
    sorted by hinge error descending, weighted by the Jaccard-extension gradient and
    averaged over the batch.  Differentiable w.r.t. logits."""
    b = logits.shape[0]
    lg = logits.reshape(b, -1)
    gt = target.reshape(b, -1).to(torch.long)
    total = logits.new_zeros(())
    for i in range(b):
        sign = 2.0 * gt[i].float() - 1.0
        err = 1.0 - lg[i] * sign
        err_sorted, order = torch.sort(err, dim=0, descending=True)
        g = gt[i][order].float()
        gsum = g.sum()
        inter = gsum - g.cumsum(0)
        union = gsum + (1.0 - g).cumsum(0)
        jac = 1.0 - inter / union
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
        total = total + torch.dot(F.elu(err_sorted), jac)
    return total / b


def bce_dice(logits, target, dice_weight=0.2, bce_weight=0.9, eps=1e-7):
    """0.2 * mean_c(1 - 2*sum(p_c*t_c)/(sum p_c + sum t_c + eps)) + 0.9 * BCEWithLogits(mean).
    Sums run over the whole batch (models.py:322-323); target is truncated to {0,1} ints."""
    c = logits.shape[1]
    t = target[:, :c].to(torch.long).float()
    p = torch.sigmoid(logits)
    dice = logits.new_zeros(())
    for k in range(c):
        pk, tk = p[:, k], t[:, k]
        dice = dice + (1.0 - (2.0 * (pk * tk).sum()) / (pk.sum() + tk.sum() + eps))
    dice = dice / c
    bce = F.binary_cross_entropy_with_logits(logits, t)
    return dice_weight * dice + bce_weight * bce


def sigmoid_np(x):
    return 1.0 / (1.0 + np.exp(-x))


def crop_bounds(size, target):
    """(top, bottom, left, right) crop offsets: rows [top, size-bottom), cols [left, size-right)."""
    d = size - target
    top = d // 2
    bottom = d - top
    right = d // 2
    left = d - right
    return top, bottom, left, right


def predict_masks(logits_orig, logits_flip=None, target=101, threshold=0.5):
    """numpy: logits [B,2,S,S] (and optionally the logits of the h-flipped copies) ->
    (probs [B,2,S,S] fp32 mean over TTA copies after un-flipping, masks u8 [B,target,target])."""
    p = sigmoid_np(logits_orig.astype(np.float32))
    if logits_flip is not None:
        pf = sigmoid_np(logits_flip.astype(np.float32))[..., ::-1]
        p = np.mean(np.stack([p, pf], axis=-1), axis=-1)
    s = p.shape[-1]
    top, bottom, left, right = crop_bounds(s, target)
    crop = p[:, :, top:s - bottom, left:s - right]
    mask = (crop[:, 1] > threshold).astype(np.uint8)
    return p.astype(np.float32), mask


def iou_masks(a, b):
    """mean IoU of two stacks of binary masks (empty-vs-empty counts as 1)."""
    a = a.reshape(a.shape[0], -1).astype(bool)
    b = b.reshape(b.shape[0], -1).astype(bool)
    inter = (a & b).sum(1).astype(np.float64)
    union = (a | b).sum(1).astype(np.float64)
    return float(np.mean(np.where(union == 0, 1.0, inter / np.maximum(union, 1))))
