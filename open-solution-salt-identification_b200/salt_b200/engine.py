"""Python handle on one engine instance: owns the device memory (torch tensors, used purely as
allocations / views), exposes forward, losses, backward, optimiser step and prediction.

PyTorch is plumbing here: it allocates HBM, provides the current CUDA stream and (in dist.py) the NCCL
process group.  Every arithmetic operation runs in libsaltunet.so.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import SaltEngineError


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class UNetEngine:
    """UNetResNet (reference architectures/unet.py:22-109; encoder_depth 18/34), UNetSeResNet (unet.py:112-172;
    encoder_depth 50/101/152, architecture='UNetSeResNet') or UNetSeResNetXt (unet.py:175-236; SE-ResNeXt 32x4d encoder of depth
    50/101, architecture='UNetSeResNetXt' - its grouped 3x3 convolutions run densely over block-diagonal packed weights) on the
    CUDA engine.

    precision: 'fp32' (parity mode: fp32 storage, fp32 FMA) or 'bf16' (bf16 activations / weights copies,
    fp32 accumulation, fp32 master weights and optimiser state).
    """

    def __init__(self, encoder_depth=34, num_classes=2, max_batch=8, size=128, precision='fp32',
                 use_tensor_cores=True, training=True, device=None, architecture=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise SaltEngineError('UNetEngine needs a CUDA device (B200); there is no CPU fallback')
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.encoder_depth, self.num_classes, self.max_batch, self.size = encoder_depth, num_classes, max_batch, size
        self.precision = precision
        if architecture is None:
            architecture = 'UNetSeResNet' if encoder_depth >= 50 else 'UNetResNet'
        self.architecture = architecture
        arch_id = {'UNetResNet': _lib.ARCH_UNET_RESNET, 'UNetSeResNet': _lib.ARCH_UNET_SERESNET,
                   'UNetSeResNetXt': _lib.ARCH_UNET_SERESNEXT}[architecture]
        cfg = _lib.SaltConfig(arch_id, encoder_depth, num_classes, max_batch, size, size,
                              {'fp32': _lib.PREC_FP32, 'bf16': _lib.PREC_BF16}[precision], int(bool(use_tensor_cores)))
        h = C.c_void_p()
        _lib.check(self.lib.salt_create(C.byref(cfg), C.byref(h)))
        self.h = h
        n_p = self.lib.salt_param_floats(h)
        n_b = self.lib.salt_buffer_floats(h)
        ws = self.lib.salt_workspace_bytes(h)
        with torch.cuda.device(self.device):
            self.params = torch.zeros(n_p, dtype=torch.float32, device=self.device)
            self.buffers = torch.zeros(n_b, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros(n_p, dtype=torch.float32, device=self.device) if training else None
            self.adam_m = torch.zeros(n_p, dtype=torch.float32, device=self.device) if training else None
            self.adam_v = torch.zeros(n_p, dtype=torch.float32, device=self.device) if training else None
            self.workspace = torch.empty(ws, dtype=torch.uint8, device=self.device)
            self._loss = torch.zeros(1, dtype=torch.float32, device=self.device)
            self._sums = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.table = self._read_table()
        _lib.check(self.lib.salt_bind(h, _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                                      _ptr(self.buffers), _ptr(self.workspace), ws))
        self.step_count = 0
        self.profiling = False
        self.num_batches_tracked = 0
        self._init_buffers()

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.salt_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def _read_table(self):
        table = OrderedDict()
        name = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        ndim, isbuf = C.c_int(), C.c_int()
        off, numel = C.c_size_t(), C.c_size_t()
        for i in range(self.lib.salt_num_tensors(self.h)):
            _lib.check(self.lib.salt_tensor_info(self.h, i, name, 256, shape, C.byref(ndim), C.byref(off),
                                                 C.byref(numel), C.byref(isbuf)))
            table[name.value.decode()] = (tuple(shape[:ndim.value]), off.value, numel.value, bool(isbuf.value))
        return table

    def _init_buffers(self):
        for k, (shape, off, numel, isbuf) in self.table.items():
            if k.endswith('running_var'):
                self.buffers[off:off + numel] = 1.0

    def view(self, key, grad=False):
        """Zero-copy torch view of one state_dict entry (or of its gradient)."""
        shape, off, numel, isbuf = self.table[key]
        flat = self.buffers if isbuf else (self.grads if grad else self.params)
        return flat[off:off + numel].view(shape)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def load_state(self, state):
        """state: mapping of canonical reference keys -> array/tensor.  Unknown keys are ignored by the
        caller (see models.EngineModule.load_state_dict)."""
        for k, v in state.items():
            if k in self.table:
                t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v)
                self.view(k).copy_(t.to(self.device, torch.float32).reshape(self.table[k][0]))
        self.params_changed()

    def params_changed(self):
        _lib.check(self.lib.salt_params_changed(self.h))

    # ------------------------------------------------------------------ compute
    def forward(self, x, train=False, out=None):
        """x: fp32 CUDA tensor [B,3,S,S] -> logits fp32 [B,num_classes,S,S]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous(), 'x must be a contiguous fp32 CUDA tensor'
        b = x.shape[0]
        if tuple(x.shape[1:]) != (3, self.size, self.size):
            raise SaltEngineError('input must be [B,3,%d,%d], got %s' % (self.size, self.size, tuple(x.shape)))
        if out is None:
            out = torch.empty((b, self.num_classes, self.size, self.size), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.salt_forward(self.h, _ptr(x), b, _ptr(out), int(train), self._stream()))
        if train:
            self.num_batches_tracked += 1
        return out

    def forward_tiles(self, tiles, train=False, out=None, hflip=False, mean0=0.485, std0=0.229):
        """tiles: uint8 CUDA tensor [B,h,w] (raw grey tiles) -> logits fp32 [B,num_classes,S,S].  The reference loader's
        pad / normalise / depth-channel adapter (loaders.py:607-612, utils.py:494-500, augmentation.py:247-281) runs fused
        into the stem's im2col kernel; mean0/std0 default to reference main.py:55-56."""
        assert tiles.is_cuda and tiles.dtype == torch.uint8 and tiles.is_contiguous() and tiles.dim() == 3
        b, th, tw = tiles.shape
        if out is None:
            out = torch.empty((b, self.num_classes, self.size, self.size), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.salt_forward_tiles(self.h, _ptr(tiles), b, th, tw, mean0, std0, int(hflip), _ptr(out), int(train),
                                               self._stream()))
        if train:
            self.num_batches_tracked += 1
        return out

    def loss_lovasz(self, logits, target, dlogits=None):
        if dlogits is None:
            dlogits = torch.empty_like(logits)
        _lib.check(self.lib.salt_loss_lovasz(self.h, _ptr(logits), _ptr(target), logits.shape[0], _ptr(self._loss),
                                             _ptr(dlogits), self._stream()))
        return self._loss, dlogits

    def loss_bce_dice(self, logits, target, dlogits=None, group=None):
        """0.2*dice + 0.9*bce; with a process group the Dice/BCE sums are all-reduced so that the loss is the
        reference's whole-batch value."""
        if dlogits is None:
            dlogits = torch.empty_like(logits)
        b = logits.shape[0]
        _lib.check(self.lib.salt_loss_bce_dice_reduce(self.h, _ptr(logits), _ptr(target), b, _ptr(self._sums), self._stream()))
        world = 1
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
            dist.all_reduce(self._sums, group=group)
        total = float(logits.numel()) * world
        _lib.check(self.lib.salt_loss_bce_dice_finish(self.h, _ptr(logits), _ptr(target), b, _ptr(self._sums), total,
                                                      float(world), _ptr(self._loss), _ptr(dlogits), self._stream()))
        return self._loss, dlogits

    def backward(self, dlogits):
        _lib.check(self.lib.salt_backward(self.h, _ptr(dlogits), self._stream()))

    N_SEGMENTS = 3

    def backward_segment(self, dlogits, segment):
        """One of the three consecutive segments of the backward pass (0: final + decoder + center, 1: layer4 + layer3,
        2: layer2 + layer1 + stem); afterwards ``grad_segment(segment)`` of the flat gradient buffer is final."""
        _lib.check(self.lib.salt_backward_segment(self.h, _ptr(dlogits), int(segment), self._stream()))

    def grad_segment(self, segment):
        """Flat view of the gradients that ``backward_segment(segment)`` completes (a contiguous range of ``self.grads``)."""
        off, n = C.c_size_t(), C.c_size_t()
        _lib.check(self.lib.salt_grad_segment(self.h, int(segment), C.byref(off), C.byref(n)))
        return self.grads[off.value:off.value + n.value]

    def adam_step(self, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        self.step_count += 1
        _lib.check(self.lib.salt_adam_step(self.h, lr, weight_decay, betas[0], betas[1], eps, self.step_count,
                                           grad_scale, self._stream()))

    def predict(self, logits, logits_flip=None, crop=101, threshold=0.5, want_probs=True, want_mask=True):
        b = logits.shape[0]
        probs = torch.empty_like(logits) if want_probs else None
        mask = torch.empty((b, crop, crop), dtype=torch.uint8, device=self.device) if want_mask else None
        _lib.check(self.lib.salt_predict(self.h, _ptr(logits), _ptr(logits_flip), b, crop, threshold, _ptr(probs),
                                         _ptr(mask), self._stream()))
        return probs, mask

    def profile(self, on):
        self.profiling = bool(on)
        _lib.check(self.lib.salt_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{class: (ms, flops, launches)} for conv fwd / dgrad / wgrad since profile(True)."""
        out = {}
        for cls, name in enumerate(('conv_fwd', 'conv_dgrad', 'conv_wgrad')):
            ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
            _lib.check(self.lib.salt_profile_read(self.h, cls, C.byref(ms), C.byref(fl), C.byref(n)))
            out[name] = (ms.value, fl.value, n.value)
        return out

    PASS_CLASSES = ('bn_apply', 'bn_bwd_reduce', 'bn_bwd_apply', 'bn_finalize', 'gather_fwd', 'gather_bwd', 'scse', 'other')

    def profile_read_passes(self):
        """{class: (ms, algorithmic bytes, launches)} for the memory-bound passes since profile(True)."""
        out = {}
        for i, name in enumerate(self.PASS_CLASSES):
            ms, by, n = C.c_double(), C.c_double(), C.c_longlong()
            _lib.check(self.lib.salt_profile_read(self.h, 3 + i, C.byref(ms), C.byref(by), C.byref(n)))
            out[name] = (ms.value, by.value, n.value)
        return out

    PROFILE_GROUPS = ('stem', 'layer1', 'layer2', 'layer3', 'layer4', 'center', 'dec5', 'dec4', 'dec3', 'dec2', 'dec1', 'final')

    def profile_read_groups(self):
        """{group: {class: (ms, flops, launches)}} - the convolution launches of each layer group since profile(True)."""
        out = {}
        for gi, gname in enumerate(self.PROFILE_GROUPS):
            d = {}
            for cls, name in enumerate(('conv_fwd', 'conv_dgrad', 'conv_wgrad')):
                ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
                _lib.check(self.lib.salt_profile_read_group(self.h, cls, gi, C.byref(ms), C.byref(fl), C.byref(n)))
                d[name] = (ms.value, fl.value, n.value)
            out[gname] = d
        return out

    def profile_records(self):
        """[(class name, group name, work, ms)] for every bracketed launch since profile(True), in launch order."""
        n = self.lib.salt_profile_records(self.h, None, None, None, None, 0)
        cls, grp = (C.c_int * n)(), (C.c_int * n)()
        work, ms = (C.c_double * n)(), (C.c_double * n)()
        self.lib.salt_profile_records(self.h, cls, grp, work, ms, n)
        names = ('conv_fwd', 'conv_dgrad', 'conv_wgrad') + self.PASS_CLASSES
        return [(names[cls[i]], self.PROFILE_GROUPS[grp[i]], work[i], ms[i]) for i in range(n)]

    def activation(self, name):
        shape = (C.c_int * 4)()
        _lib.check(self.lib.salt_get_activation(self.h, name.encode(), C.c_void_p(0), shape, self._stream()))
        out = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.salt_get_activation(self.h, name.encode(), _ptr(out), shape, self._stream()))
        return out
