"""Drop-in replacement for the reference's segmentation transformer.

Mirrors ``common_blocks/models.py:67-208`` (``SegmentationModel``) - same constructor
(``architecture_config, training_config, callbacks_config``), same ``fit`` / ``transform`` / ``fit_transform``
/ ``load`` / ``persist`` surface, same loader dicts in, same ``{'mask_prediction': [ndarray (C,H,W)]}`` out,
same attributes for the callbacks (``model``, ``optimizer``, ``loss_function``, ``output_names``,
``validation_loss``, ``callbacks``) - so ``common_blocks/pipelines.py`` and ``main.py:network()`` can use it
unchanged.  Underneath, every array operation runs in libsaltunet.so.

Extra engine knobs come from the environment so reference callers need no change:
  SALT_ENGINE_PRECISION  'bf16' (default) | 'fp32'      SALT_ENGINE_MAX_BATCH  (default 128)
  SALT_ENGINE_SIZE       network input size (default 128)   SALT_ENGINE_LOSS  'lovasz' (default) | 'bce_dice'
  SALT_ENGINE_GRAPH      '1' (default): replay the forward and backward passes of a training step as CUDA graphs
"""
import os
from collections import OrderedDict

import torch

from . import dist as sdist
from .engine import UNetEngine

# reference models.py:15-24 - the entries this engine implements
ARCHITECTURES = {'UNetResNet': {'model_config': {'encoder_depth': 34, 'use_hypercolumn': True, 'dropout_2d': 0.0,
                                                 'pretrained': True, 'pool0': False},
                                'init_weights': False},
                 'UNetSeResNet': {'model_config': {'encoder_depth': 50, 'use_hypercolumn': True, 'dropout_2d': 0.0,
                                                   'pretrained': 'imagenet', 'pool0': False},
                                  'init_weights': False},
                 # SURVEY 8(f) N4, first entry: reference models.py:25-30 / unet.py:175-236 / encoders.py:86-118
                 'UNetSeResNetXt': {'model_config': {'encoder_depth': 50, 'use_hypercolumn': True, 'dropout_2d': 0.0,
                                                     'pretrained': 'imagenet', 'pool0': False},
                                    'init_weights': False}}


def _alias_map(table):
    """alias key -> canonical key for the duplicated registrations of reference encoders.py:21-36 (ResNet) and
    :59-74 (SE-ResNet: the stem lives in ``encoder.layer0``)."""
    amap = {}
    stems = (('encoders.encoder.conv1.', 'encoders.conv1.0.'), ('encoders.encoder.bn1.', 'encoders.conv1.1.'),
             ('encoders.encoder.layer0.conv1.', 'encoders.conv1.0.'), ('encoders.encoder.layer0.bn1.', 'encoders.conv1.1.'))
    for k in table:
        stem = [(a, b) for a, b in stems if k.startswith(a)]
        if stem:
            amap[stem[0][1] + k[len(stem[0][0]):]] = k
        else:
            for li in (1, 2, 3, 4):
                pre = 'encoders.encoder.layer%d.' % li
                if k.startswith(pre):
                    amap['encoders.encoder%d.%s' % (li + 1, k[len(pre):])] = k
    return amap


class EngineModule:
    """What the reference code sees as ``self.model`` (an nn.DataParallel-wrapped nn.Module):
    callable, ``train()/eval()``, ``state_dict()/load_state_dict()`` with ``module.``-prefixed keys
    (models.py:81-82,199-204, callbacks.py:776-794)."""

    def __init__(self, engine):
        self.engine = engine
        self.training = True
        self._extra = OrderedDict()      # tensors the engine has no use for (encoders.encoder.fc.*), kept for round trips
        self._aliases = _alias_map(engine.table)

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def cuda(self, *a, **k):
        return self

    def cpu(self):
        return self

    def parameters(self):
        return [self.engine.params]

    def __call__(self, x):
        x = torch.as_tensor(x)
        if not x.is_cuda:
            x = x.to(self.engine.device, non_blocking=True)
        if x.dtype == torch.uint8 and x.dim() == 3:
            # raw grey tiles [B,h,w]: the loader's adapter runs fused into the stem
            return self.engine.forward_tiles(x.contiguous(), train=self.training)
        return self.engine.forward(x.float().contiguous(), train=self.training)

    def state_dict(self, prefix='module.'):
        eng, out = self.engine, OrderedDict()
        nbt = torch.tensor(eng.num_batches_tracked, dtype=torch.int64)
        canon = OrderedDict()
        for k in eng.table:
            canon[k] = eng.view(k).detach().cpu().clone()
            if k.endswith('running_var'):
                canon[k[:-len('running_var')] + 'num_batches_tracked'] = nbt.clone()
        for k, v in canon.items():
            out[prefix + k] = v
        for k, v in self._extra.items():
            out[prefix + k] = v
        for alias, k in self._aliases.items():
            out[prefix + alias] = canon[k]
            if k.endswith('running_var'):
                out[prefix + alias[:-len('running_var')] + 'num_batches_tracked'] = nbt.clone()
        return out

    def load_state_dict(self, state, strict=False):
        eng, canon = self.engine, {}
        for k, v in state.items():
            if k.startswith('module.'):
                k = k[len('module.'):]
            k = self._aliases.get(k, k)
            if k in eng.table:
                canon[k] = v
            elif k.endswith('num_batches_tracked'):
                eng.num_batches_tracked = int(v)
            else:
                self._extra[k] = torch.as_tensor(v).clone()
        missing = [k for k in eng.table if k not in canon]
        if strict and missing:
            raise KeyError('missing keys: %s' % missing[:5])
        eng.load_state(canon)
        return self


class EngineAdam(torch.optim.Optimizer):
    """``self.optimizer``: torch.optim.Adam(weight_regularization(...), lr) of models.py:74-75, as a real
    ``torch.optim.Optimizer`` so the reference's schedulers accept it: ``ReduceLROnPlateauScheduler`` / ``ExponentialLRScheduler``
    build ``torch.optim.lr_scheduler.*(optimizer=self.optimizer)`` and mutate ``param_groups[0]['lr']`` (callbacks.py:181,
    219-241), loggers read ``state_dict()['param_groups'][0]['lr']``.  One parameter group over the engine's flat fp32
    parameter buffer; ``step()`` is ONE fused kernel (``salt_adam_step``: L2 term, moments, bias correction, update).  The
    moments live in engine-owned flat buffers, so ``self.state`` stays empty (the reference never saves optimiser state)."""

    def __init__(self, engine, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.engine = engine
        super().__init__([engine.params], dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps))

    def zero_grad(self, set_to_none=True):
        pass      # salt_backward() zeroes the flat gradient buffer itself

    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.engine.adam_step(lr=float(g['lr']), weight_decay=float(g['weight_decay']), betas=tuple(g['betas']), eps=float(g['eps']),
                              grad_scale=grad_scale)
        return loss


class _NullCallbacks:
    """Stand-in for cbk.CallbackList when the reference's callbacks module is not importable."""

    def set_params(self, *a, **k): pass
    def on_train_begin(self, *a, **k): pass
    def on_train_end(self, *a, **k): pass
    def on_epoch_begin(self, *a, **k): pass
    def on_epoch_end(self, *a, **k): pass
    def on_batch_begin(self, *a, **k): pass
    def on_batch_end(self, *a, **k): pass
    def training_break(self, *a, **k): return False


# callbacks that write files or talk to an experiment tracker: under torchrun only rank 0 keeps them
_RANK0_ONLY_CALLBACKS = ('ModelCheckpoint', 'NeptuneMonitor', 'ExperimentTiming', 'TrainingMonitor')


def callbacks_network(callbacks_config, dp=None):
    """models.py:300-312 when the reference package is importable, otherwise no callbacks (with a warning).  Only a missing
    reference package takes the fallback: configuration errors (a mistyped callbacks_config key, ...) propagate, so a model can
    never train silently without its checkpoint / early-stopping callbacks.  With more than one rank, the callbacks that write
    to disk or log (``_RANK0_ONLY_CALLBACKS``) are kept on rank 0 only; the ones that steer training (validation, LR scheduler,
    early stopping) run on every rank on identical parameters and therefore take identical decisions."""
    if not callbacks_config:                  # None / {}: no callbacks requested (extension; the reference always passes its 7 sections)
        return _NullCallbacks()
    try:
        from common_blocks.models import callbacks_network as ref_callbacks_network
    except ImportError as exc:
        import warnings
        warnings.warn('reference package common_blocks is not importable (%s): SegmentationModel runs WITHOUT callbacks '
                      '(no checkpoints, no early stopping, no LR schedule)' % (exc,))
        return _NullCallbacks()
    cbs = ref_callbacks_network(callbacks_config)
    if dp is not None and dp.world > 1 and dp.rank != 0 and hasattr(cbs, 'callbacks'):
        cbs.callbacks = [c for c in cbs.callbacks if type(c).__name__ not in _RANK0_ONLY_CALLBACKS]
    return cbs


class Model:
    """toolkit.pytorch_transformers.models.Model as used by models.py:67-76."""

    def __init__(self, architecture_config, training_config, callbacks_config):
        self.architecture_config = architecture_config
        self.training_config = training_config
        self.callbacks_config = callbacks_config
        self.model = None
        self.optimizer = None
        self.loss_function = None
        self.callbacks = None
        self.validation_loss = {}

    @property
    def output_names(self):
        return [name for (name, func, weight) in self.loss_function]

    def fit_transform(self, *args, **kwargs):
        self.fit(*args, **kwargs)
        return self.transform(*args, **kwargs)

    def persist(self, filepath):
        dp = getattr(self, 'dp', None)
        if dp is None or dp.rank == 0:           # every rank holds the same parameters: one writer
            d = os.path.dirname(filepath)
            if d:
                os.makedirs(d, exist_ok=True)
            torch.save(self.model.state_dict(), filepath)
        if dp is not None:
            dp.barrier()


class _EngineLoss:
    """(name, fn, weight) entry of ``loss_function``: callable on (logits, target) like the reference's
    lovasz_loss / mixed_dice_bce_loss, returning a 1-element tensor; the gradient w.r.t. the logits computed by
    the same kernel is kept for the backward pass."""

    def __init__(self, engine, kind, group=None):
        self.engine, self.kind, self.group = engine, kind, group
        self.dlogits = None

    def __call__(self, output, target):
        target = torch.as_tensor(target).to(self.engine.device, torch.float32).contiguous()
        if tuple(target.shape) != tuple(output.shape):
            raise ValueError('loss: target shape %s != output shape %s' % (tuple(target.shape), tuple(output.shape)))
        if output.shape[0] > self.engine.max_batch:
            raise ValueError('loss: batch %d exceeds the engine max_batch %d' % (output.shape[0], self.engine.max_batch))
        # the kernels write logits.shape[0] images of gradient: never hand them a buffer of another batch size
        if self.dlogits is None or tuple(self.dlogits.shape) != tuple(output.shape) or self.dlogits.device != output.device:
            self.dlogits = torch.empty_like(output)
        if self.kind == 'lovasz':
            loss, self.dlogits = self.engine.loss_lovasz(output, target, self.dlogits)
        else:
            loss, self.dlogits = self.engine.loss_bce_dice(output, target, self.dlogits, group=self.group)
        return loss.clone()          # a fresh 1-element tensor per call, as the reference returns (callbacks may keep it)


class _PinnedSink:
    """Device -> host path of ``transform``: the copy of batch i runs on a side stream into one of two pinned buffers while
    batch i+1 is computed (models.py:167 does a synchronous ``.data.cpu().numpy()`` per batch); ``finish()`` returns the
    per-image list of utils.py:316-320 ``get_list_of_image_predictions``."""

    def __init__(self, device, max_batch):
        self.device, self.max_batch = device, max_batch
        self.stream = torch.cuda.Stream(device)
        self.slots, self.n, self.pending, self.out = [None, None], 0, None, []

    def _drain(self):
        if self.pending is not None:
            ev, host, b = self.pending
            ev.synchronize()
            self.out.extend(list(host[:b].numpy().copy()))
            self.pending = None

    def push(self, probs):
        b, slot = int(probs.shape[0]), self.n & 1
        self.n += 1
        if self.slots[slot] is None or self.slots[slot].shape[0] < b:
            self.slots[slot] = torch.empty((max(b, self.max_batch),) + tuple(probs.shape[1:]), dtype=torch.float32, pin_memory=True)
        done = torch.cuda.Event()
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.slots[slot][:b].copy_(probs, non_blocking=True)
            probs.record_stream(self.stream)
            done.record(self.stream)
        self._drain()                   # batch i-1: its copy overlapped this batch's forward pass
        self.pending = (done, self.slots[slot], b)

    def finish(self):
        self._drain()
        return self.out


class SegmentationModel(Model):
    def __init__(self, architecture_config, training_config, callbacks_config):
        super().__init__(architecture_config, training_config, callbacks_config)
        self.activation_func = self.architecture_config['model_params']['activation']
        self.dp = sdist.DataParallelContext.from_env()
        self.set_model()
        self.set_loss()
        opt = dict(architecture_config.get('optimizer_params', {'lr': 1e-4}))
        reg = architecture_config.get('regularizer_params', {'regularize': True, 'weight_decay_conv2d': 1e-4})
        wd = reg.get('weight_decay_conv2d', 0.0) if reg.get('regularize', False) else 0.0
        self.optimizer = EngineAdam(self.engine, lr=opt.get('lr', 1e-4), weight_decay=wd)
        self.callbacks = callbacks_network(self.callbacks_config, self.dp)

    # models.py:179-184
    def set_model(self):
        mp = self.architecture_config['model_params']
        architecture = mp['architecture']
        if architecture not in ARCHITECTURES:
            raise NotImplementedError('architecture %r is not implemented by the B200 engine (have: %s)'
                                      % (architecture, sorted(ARCHITECTURES)))
        cfg = ARCHITECTURES[architecture]['model_config']
        self.engine = UNetEngine(architecture=architecture, encoder_depth=mp.get('encoder_depth', cfg['encoder_depth']),
                                 num_classes=mp['out_channels'],
                                 max_batch=int(os.environ.get('SALT_ENGINE_MAX_BATCH', mp.get('max_batch', 128))),
                                 size=int(os.environ.get('SALT_ENGINE_SIZE', mp.get('size', 128))),
                                 precision=os.environ.get('SALT_ENGINE_PRECISION', mp.get('precision', 'bf16')),
                                 device=self.dp.device)
        self.model = EngineModule(self.engine)
        self._initialize_model_weights = lambda: None
        if self.dp.world > 1:
            self.dp.broadcast(self.engine.params, self.engine.buffers)
            self.engine.params_changed()

    # models.py:186-194
    def set_loss(self):
        if self.activation_func == 'softmax':
            raise NotImplementedError('No softmax loss defined')
        elif self.activation_func == 'sigmoid':
            kind = os.environ.get('SALT_ENGINE_LOSS', self.architecture_config['model_params'].get('loss', 'lovasz'))
            loss_function = _EngineLoss(self.engine, kind, group=self.dp.group)
        else:
            raise Exception('Only softmax and sigmoid activations are allowed')
        self.loss_function = [('mask', loss_function, 1.0)]

    # models.py:78-103
    def fit(self, datagen, validation_datagen=None, meta_valid=None):
        self._initialize_model_weights()
        self.callbacks.set_params(self, validation_datagen=validation_datagen, meta_valid=meta_valid)
        self.callbacks.on_train_begin()
        batch_gen, steps = datagen
        for epoch_id in range(self.training_config['epochs']):
            self.callbacks.on_epoch_begin()
            for batch_id, data in enumerate(self._prefetch(batch_gen)):
                self.callbacks.on_batch_begin()
                metrics = self._fit_loop(data)
                self.callbacks.on_batch_end(metrics=metrics)
                if batch_id == steps:
                    break
            self.callbacks.on_epoch_end()
            if self.callbacks.training_break():
                break
        self.callbacks.on_train_end()
        return self

    # ------------------------------------------------------------------ input prefetch (fit only)
    # models.py:106-117 copies every batch to the GPU at the top of the step.  Here the copy of batch i+1 is issued on a side
    # stream, into one of two staging buffers, BEFORE step i is launched, so it runs under step i's kernels; step i+1 then
    # starts with a device-to-device copy of X into the graph's static input (25 MB, ~10 us) and reads its target in place.
    class _Staged:
        __slots__ = ('b', 'slot', 'ready')

    def _stage(self, data):
        if data is None or not self._graph_enabled() or not isinstance(data, (list, tuple)) or len(data) != 2:
            return data
        x, t = data
        if not (torch.is_tensor(x) and torch.is_tensor(t)) or x.is_cuda or x.dtype != torch.float32 or t.dtype != torch.float32:
            return data
        st, b = self._graph_state(), int(x.shape[0])
        if b not in st['graphs']:
            return data
        dev = self.engine.device
        if 'stage' not in st:
            st['stage'] = [{'x': torch.empty_like(st['x']), 't': torch.empty_like(st['t']), 'free': None} for _ in range(2)]
            st['stage_stream'], st['stage_next'] = torch.cuda.Stream(dev), 0
        slot = st['stage_next']
        st['stage_next'] ^= 1
        buf = st['stage'][slot]
        if buf['free'] is not None:
            st['stage_stream'].wait_event(buf['free'])        # the step that last read this slot has finished
        s = self._Staged()
        s.b, s.slot, s.ready = b, slot, torch.cuda.Event()
        with torch.cuda.stream(st['stage_stream']):
            buf['x'][:b].copy_(x, non_blocking=True)
            buf['t'][:b].copy_(t, non_blocking=True)
            s.ready.record()
        return s

    def _prefetch(self, batch_gen):
        it = iter(batch_gen)
        cur = self._stage(next(it, None))
        while cur is not None:
            nxt = self._stage(next(it, None))
            yield cur
            cur = nxt

    def _fit_staged(self, s):
        st, dev = self._graph_state(), self.engine.device
        buf, cur = st['stage'][s.slot], torch.cuda.current_stream(dev)
        cur.wait_event(s.ready)
        st['x'][:s.b].copy_(buf['x'][:s.b], non_blocking=True)
        out = self._train_step_graph(s.b, target=buf['t'][:s.b])
        buf['free'] = torch.cuda.Event()
        buf['free'].record(cur)
        return out

    # models.py:105-136
    def _fit_loop(self, data):
        torch.cuda.nvtx.range_push('salt.fit.step')          # host-side marker: one range per optimisation step
        try:
            return self._fit_loop_impl(data)
        finally:
            torch.cuda.nvtx.range_pop()

    def _fit_loop_impl(self, data):
        if isinstance(data, self._Staged):
            return self._fit_staged(data)
        dev = self.engine.device
        b = int(data[0].shape[0])
        if self._graph_enabled() and len(data) == 2 and b in self._graph_state()['graphs']:
            st = self._graph_state()               # host -> static device buffers, no intermediate tensor
            cur = torch.cuda.current_stream(dev)
            if 'copy_stream' not in st:
                st['copy_stream'], st['t_ready'] = torch.cuda.Stream(dev), torch.cuda.Event()
            # the target is not needed before the loss: its H2D copy runs on a side stream under the forward pass
            st['copy_stream'].wait_stream(cur)     # the previous step's loss kernels have finished reading st['t']
            with torch.cuda.stream(st['copy_stream']):
                st['t'][:b].copy_(torch.as_tensor(data[1]), non_blocking=True)
                st['t_ready'].record()
            st['x'][:b].copy_(torch.as_tensor(data[0]), non_blocking=True)
            return self._train_step_graph(b, wait=st['t_ready'])
        X = torch.as_tensor(data[0]).to(dev, torch.float32, non_blocking=True).contiguous()
        targets = [torch.as_tensor(t).to(dev, torch.float32, non_blocking=True).contiguous() for t in data[1:]]
        return self.train_step_device(X, targets)

    # ------------------------------------------------------------------ CUDA-graph replay of the two passes
    # A training step is ~470 kernel launches of 2-800 us; replaying forward and backward as two CUDA graphs removes the
    # per-launch host work and shortens the gaps between dependent kernels.  Loss, gradient all-reduce and Adam stay eager
    # (their arguments - Dice sums, learning rate, step count - change per step).  Graphs are keyed by batch size and are
    # captured after two eager steps of that size, over static input / logits / dlogits buffers.
    def _graph_state(self):
        st = getattr(self, '_gs', None)
        if st is None:
            eng = self.engine
            shape = (eng.max_batch, eng.num_classes, eng.size, eng.size)
            st = self._gs = {'x': torch.empty((eng.max_batch, 3, eng.size, eng.size), dtype=torch.float32, device=eng.device),
                             't': torch.empty(shape, dtype=torch.float32, device=eng.device),
                             'logits': torch.empty(shape, dtype=torch.float32, device=eng.device),
                             'dlogits': torch.empty(shape, dtype=torch.float32, device=eng.device),
                             'seen': {}, 'graphs': {}}
        return st

    def _bucketed(self):
        """Data parallel: SALT_DP_BUCKETS=1 all-reduces the gradient in three buckets overlapped with the backward segments.  Default
        off: measured on B200 (profiles/r2_notes.md section 8) the overlapped form is SLOWER at 2 and at 8 GPUs (14.95 vs 14.84 ms,
        15.45 vs 15.34 ms per step) - the NCCL kernels take whole SMs away from the persistent one-CTA-per-SM convolutions, whose
        static tile schedule then runs a second wave - so one all-reduce after the backward pass (0.12 ms at 2 GPUs) is the default."""
        return self.dp.world > 1 and os.environ.get('SALT_DP_BUCKETS', '0') == '1'

    def _graph_enabled(self):
        return (os.environ.get('SALT_ENGINE_GRAPH', '1') != '0' and not self.engine.profiling
                and not getattr(self, '_gs', {}).get('disabled', False))

    def _capture(self, b):
        """Capture forward and backward of batch size b.  Returns False (and switches graph replay off for this model) if the
        capture fails - the eager path is always valid."""
        from . import _lib
        eng, st = self.engine, self._graph_state()
        torch.cuda.synchronize(eng.device)
        eng.params_changed()                      # the weight re-pack must be part of the captured forward
        nbt = eng.num_batches_tracked
        # with a process group alive, NCCL's watchdog thread polls CUDA events: keep its calls from invalidating the capture
        mode = 'thread_local' if self.dp.world > 1 else 'global'
        try:
            n0 = _lib.launch_count()
            gf = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf, capture_error_mode=mode):
                eng.forward(st['x'][:b], train=True, out=st['logits'][:b])
            n1 = _lib.launch_count()
            if self._bucketed():
                # data parallel: one graph per backward segment, so that each segment's gradient bucket can be all-reduced
                # (NCCL, side stream) while the next segment computes - SURVEY.md 8(e)
                gb = []
                for seg in range(eng.N_SEGMENTS):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode=mode):
                        eng.backward_segment(st['dlogits'][:b], seg)
                    gb.append(g)
            else:
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, capture_error_mode=mode):
                    eng.backward(st['dlogits'][:b])
            n2 = _lib.launch_count()
        except Exception as exc:                  # pragma: no cover - depends on driver / NCCL state
            import warnings
            warnings.warn('CUDA-graph capture failed (%s); training continues with eager launches' % (exc,))
            st['disabled'] = True
            torch.cuda.synchronize(eng.device)
            eng.num_batches_tracked = nbt
            eng.params_changed()
            return False
        eng.num_batches_tracked = nbt             # capture does not execute
        st['graphs'][b] = (gf, gb, n1 - n0, n2 - n1)
        return True

    def _train_step_graph(self, b, wait=None, target=None):
        from . import _lib
        eng, st = self.engine, self._graph_state()
        gf, gb, nf, nb = st['graphs'][b]
        (name, loss_function, weight) = self.loss_function[0]
        self.model.train()
        gf.replay()
        eng.num_batches_tracked += 1
        if wait is not None:
            torch.cuda.current_stream(eng.device).wait_event(wait)
        loss_function.dlogits = st['dlogits'][:b]
        batch_loss = loss_function(st['logits'][:b], st['t'][:b] if target is None else target)
        if weight != 1.0:
            batch_loss = batch_loss * weight
            st['dlogits'][:b].mul_(weight)
        if isinstance(gb, list):
            works = []
            for seg, g in enumerate(gb):
                g.replay()
                works.append(self.dp.allreduce_async(eng.grad_segment(seg)))       # overlaps the next segment's kernels
            for w in works:
                w.wait()
            scale = 1.0 / self.dp.world
        else:
            gb.replay()
            scale = self.dp.allreduce_grads(eng.grads)
        _lib.count_replayed(nf + nb)
        self.optimizer.step(grad_scale=scale)
        return {'sum': batch_loss}

    def train_step_device(self, X, targets):
        """One optimisation step on device-resident fp32 tensors: forward (train BN), loss + dL/dlogits, backward,
        gradient all-reduce (data parallel), fused Adam."""
        self.model.train()
        b = int(X.shape[0])
        if self._graph_enabled() and X.dtype == torch.float32 and len(targets) == 1 and b <= self.engine.max_batch:
            st = self._graph_state()
            if b not in st['graphs'] and st['seen'].get(b, 0) >= 2 and len(st['graphs']) < 4:
                self._capture(b)
            if b in st['graphs']:
                if X.data_ptr() != st['x'].data_ptr():
                    st['x'][:b].copy_(X, non_blocking=True)
                if targets[0].data_ptr() != st['t'].data_ptr():
                    st['t'][:b].copy_(targets[0], non_blocking=True)
                return self._train_step_graph(b)
            st['seen'][b] = st['seen'].get(b, 0) + 1
        self.optimizer.zero_grad()
        outputs_batch = self.model(X)
        partial_batch_losses = {}
        (name, loss_function, weight), target = self.loss_function[0], targets[0]
        batch_loss = loss_function(outputs_batch, target) * weight
        partial_batch_losses['sum'] = batch_loss
        dlogits = loss_function.dlogits if weight == 1.0 else loss_function.dlogits * weight
        if self._bucketed():
            works = []
            for seg in range(self.engine.N_SEGMENTS):
                self.engine.backward_segment(dlogits, seg)
                works.append(self.dp.allreduce_async(self.engine.grad_segment(seg)))
            for w in works:
                w.wait()
            scale = 1.0 / self.dp.world
        else:
            self.engine.backward(dlogits)
            scale = self.dp.allreduce_grads(self.engine.grads)
        self.optimizer.step(grad_scale=scale)
        return partial_batch_losses

    # models.py:138-147
    def transform(self, datagen, validation_datagen=None, *args, **kwargs):
        outputs = self._transform(datagen, validation_datagen)
        if self.activation_func not in ('sigmoid',):
            raise Exception('Only softmax and sigmoid activations are allowed')
        return outputs

    # models.py:149-177 with the numpy sigmoid of utils.py:173 fused on the GPU; device->host through _PinnedSink
    def _transform(self, datagen, validation_datagen=None, **kwargs):
        self.model.eval()
        batch_gen, steps = datagen
        name = self.output_names[0]
        eng = self.engine
        sink = _PinnedSink(eng.device, eng.max_batch)
        for batch_id, data in enumerate(batch_gen):
            X = data[0] if isinstance(data, (list, tuple)) else data
            logits = self.model(X)
            probs, _ = eng.predict(logits, None, crop=min(101, eng.size), want_mask=False)
            sink.push(probs)
            if batch_id == steps:
                break
        preds = sink.finish()
        self.model.train()
        return {'{}_prediction'.format(name): preds}

    # models.py:196-208
    def load(self, filepath):
        self.model.eval()
        self.model.load_state_dict(torch.load(filepath, map_location='cpu'))
        return self
