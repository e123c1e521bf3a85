"""Data formats either side of the network on the GPU: the reference loader's tile adapter
and the submission's run-length encoder.  Thin ctypes calls into libsaltunet.so; torch tensors are only the device memory.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .synthetic import MEAN, STD


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda():
    if not torch.cuda.is_available():
        raise _lib.SaltEngineError('needs a CUDA device (B200); there is no CPU fallback')


def adapt_tiles(tiles, size=128, hflip=False, mean0=MEAN[0], std0=STD[0], out=None):
    """Raw grey tiles uint8 [B,h,w] (CUDA) -> network input fp32 [B,3,size,size], as the reference's inference loader builds it
    (loaders.py:607-612 Grayscale(3)/ToTensor/Normalize/AddDepthChannels, augmentation.py:247-281 InferencePad 'edge').
    hflip = the TTA's np.fliplr of the raw tile (augmentation.py:143-147)."""
    _need_cuda()
    lib = _lib.load()
    tiles = torch.as_tensor(tiles)
    if tiles.dtype != torch.uint8 or tiles.dim() != 3:
        raise _lib.SaltEngineError('tiles must be uint8 [B,h,w], got %s %s' % (tiles.dtype, tuple(tiles.shape)))
    if not tiles.is_cuda:
        tiles = tiles.cuda(non_blocking=True)
    tiles = tiles.contiguous()
    b, th, tw = tiles.shape
    if out is None:
        out = torch.empty((b, 3, size, size), dtype=torch.float32, device=tiles.device)
    _lib.check(lib.salt_adapt_tiles(_ptr(tiles), b, th, tw, size, mean0, std0, int(hflip), _ptr(out), _stream(tiles.device)))
    return out


def rle_encode_device(masks, cap_runs=None):
    """masks uint8 [B,H,W] (CUDA) -> (runs int32 [B,cap,2] of (start, length), nruns int32 [B]) on the device
    (utils.py:99-111 run_length_encoding for every mask of the batch in one launch)."""
    _need_cuda()
    lib = _lib.load()
    assert masks.is_cuda and masks.dtype == torch.uint8 and masks.dim() == 3
    masks = masks.contiguous()
    b, h, w = masks.shape
    if cap_runs is None:
        cap_runs = (h * w + 1) // 2            # the most runs a mask of h*w pixels can have
    runs = torch.empty((b, cap_runs, 2), dtype=torch.int32, device=masks.device)
    nruns = torch.empty((b,), dtype=torch.int32, device=masks.device)
    _lib.check(lib.salt_rle_encode(_ptr(masks), b, h, w, cap_runs, _ptr(runs), _ptr(nruns), _stream(masks.device)))
    return runs, nruns


def encode_rle(masks):
    """utils.py:78-79 encode_rle: list of flat [start, length, start, length, ...] python lists, one per mask."""
    masks = torch.as_tensor(np.stack(masks) if isinstance(masks, (list, tuple)) else masks)
    if masks.dtype != torch.uint8:
        masks = (masks != 0).to(torch.uint8)
    if not masks.is_cuda:
        _need_cuda()
        masks = masks.cuda()
    if masks.shape[0] == 0:
        return []
    runs, nruns = rle_encode_device(masks)
    n = nruns.cpu().numpy()
    r = runs[:, :max(int(n.max()), 1)].cpu().numpy()
    return [r[i, :n[i]].reshape(-1).tolist() for i in range(len(n))]


def create_submission(meta, predictions):
    """utils.py:68-75 create_submission: ``meta`` is the reference's metadata frame (``meta['id']``) or a plain sequence of ids;
    returns the same ``DataFrame(columns=['id', 'rle_mask']).astype(str)``, one 'start length start length ...' string per mask."""
    import pandas as pd
    ids = meta if isinstance(meta, (list, tuple, np.ndarray)) else meta['id'].values
    output = [[image_id, ' '.join(str(v) for v in rle)] for image_id, rle in zip(ids, encode_rle(predictions))]
    return pd.DataFrame(output, columns=['id', 'rle_mask']).astype(str)
