"""CPU: the drop-in boundary driven by the reference's OWN, UNMODIFIED callers.

``common_blocks.models.callbacks_network`` (models.py:300-312: ExperimentTiming, TrainingMonitor, ValidationMonitor,
ModelCheckpoint, ReduceLROnPlateauScheduler, NeptuneMonitor, EarlyStopping) and ``common_blocks.utils.FineTuneStep``
(utils.py:415-486) are imported from /root/reference through ``oracle/ref_shims.py`` and run against
``salt_b200.models.SegmentationModel``.  Only the CUDA engine underneath is replaced by a tiny CPU stand-in (a 1x1
convolution with the engine's Python surface): what is under test is the host-side contract - ``optimizer`` is a real
``torch.optim.Optimizer`` whose learning rate the reference's ``ReduceLROnPlateau`` mutates, losses are fresh 1-element tensors,
``model`` supports eval()/train()/state_dict() with ``module.`` keys, ``persist``/``load`` round-trip through the step.

Skipped when the reference tree is absent (GPU box).
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shims, losses_oracle

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason='reference tree not present')

S, B, T = 128, 2, 101


class StubEngine:
    """CPU stand-in with UNetEngine's Python surface: logits = 1x1 conv (3 -> 2) of the input."""
    instances = []

    def __init__(self, architecture=None, encoder_depth=34, num_classes=2, max_batch=8, size=128, precision='bf16', device=None,
                 **kw):
        self.device = torch.device('cpu')
        self.num_classes, self.max_batch, self.size, self.precision = num_classes, max_batch, size, precision
        self.table = {'final.1.weight': ((num_classes, 3, 1, 1), 0, 3 * num_classes, False),
                      'final.1.bias': ((num_classes,), 8, num_classes, False)}
        g = torch.Generator().manual_seed(5)
        self.params = torch.randn(12, generator=g) * 0.5
        self.grads, self.buffers = torch.zeros(12), torch.zeros(0)
        self.m, self.v = torch.zeros(12), torch.zeros(12)
        self.step_count = self.num_batches_tracked = 0
        self.profiling = False
        self.lrs = []
        self._x = None
        StubEngine.instances.append(self)

    def view(self, key, grad=False):
        shape, off, numel, _ = self.table[key]
        return (self.grads if grad else self.params)[off:off + numel].view(shape)

    def params_changed(self):
        pass

    def load_state(self, state):
        for k, v in state.items():
            if k in self.table:
                self.view(k).copy_(torch.as_tensor(v).reshape(self.table[k][0]))

    def forward(self, x, train=False, out=None):
        self._x = x
        w, b = self.view('final.1.weight').view(self.num_classes, 3), self.view('final.1.bias')
        y = torch.einsum('kc,bchw->bkhw', w, x) + b.view(1, -1, 1, 1)
        if train:
            self.num_batches_tracked += 1
        if out is not None:
            out.copy_(y)
            return out
        return y

    def _loss(self, fn, logits, target, dlogits):
        lg = logits.detach().clone().requires_grad_(True)
        loss = fn(lg, target)
        loss.backward()
        dlogits.copy_(lg.grad)
        return loss.detach().reshape(1), dlogits

    def loss_lovasz(self, logits, target, dlogits=None):
        return self._loss(losses_oracle.lovasz_hinge_per_image, logits, target, dlogits if dlogits is not None else torch.empty_like(logits))

    def loss_bce_dice(self, logits, target, dlogits=None, group=None):
        return self._loss(losses_oracle.bce_dice, logits, target, dlogits if dlogits is not None else torch.empty_like(logits))

    def backward(self, dlogits):
        self.grads.zero_()
        self.view('final.1.weight', grad=True).copy_(torch.einsum('bkhw,bchw->kc', dlogits, self._x).view(self.num_classes, 3, 1, 1))
        self.view('final.1.bias', grad=True).copy_(dlogits.sum(dim=(0, 2, 3)))

    def adam_step(self, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        self.step_count += 1
        self.lrs.append(lr)
        g = self.grads * grad_scale + weight_decay * self.params
        self.m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        self.v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        mh, vh = self.m / (1 - betas[0] ** self.step_count), self.v / (1 - betas[1] ** self.step_count)
        self.params.sub_(lr * mh / (vh.sqrt() + eps))

    def predict(self, logits, logits_flip=None, crop=101, threshold=0.5, want_probs=True, want_mask=True):
        return torch.sigmoid(logits), None


class ListSink:
    def __init__(self, device, max_batch):
        self.out = []

    def push(self, probs):
        self.out.extend(list(probs.numpy().copy()))

    def finish(self):
        return self.out


ARCH = {'model_params': {'architecture': 'UNetResNet', 'encoder_depth': 34, 'in_channels': 3, 'out_channels': 2, 'activation': 'sigmoid'},
        'optimizer_params': {'lr': 1e-2}, 'regularizer_params': {'regularize': True, 'weight_decay_conv2d': 1e-4},
        'weights_init': {'function': 'xavier'}}


def _callbacks_config(tmp, patience=0):
    """reference main.py:251-281, on the temporary experiment directory"""
    return {'model_checkpoint': {'filepath': os.path.join(tmp, 'checkpoints', 'network', 'best.torch'), 'epoch_every': 1,
                                 'metric_name': 'sum', 'minimize': True},
            'exponential_lr_scheduler': {'gamma': 0.95, 'epoch_every': 1},
            'reduce_lr_on_plateau_scheduler': {'metric_name': 'sum', 'minimize': True, 'reduce_factor': 0.5,
                                               'reduce_patience': patience, 'min_lr': 1e-6},
            'training_monitor': {'batch_every': 0, 'epoch_every': 1},
            'experiment_timing': {'batch_every': 0, 'epoch_every': 1},
            'validation_monitor': {'epoch_every': 1, 'data_dir': tmp, 'loader_mode': 'resize_and_pad', 'use_depth': False},
            'neptune_monitor': {'model_name': 'network', 'image_nr': 16, 'image_resize': 1.0, 'image_every': None, 'use_depth': False},
            'early_stopping': {'patience': 1000, 'metric_name': 'sum', 'minimize': True}}


def _data(tmp, n_batches=2, seed=3):
    from PIL import Image
    import pandas as pd
    g = torch.Generator().manual_seed(seed)
    batches, paths = [], []
    for i in range(n_batches):
        x = torch.randn(B, 3, S, S, generator=g)
        salt = (x[:, 0:1] + 0.3 * torch.randn(B, 1, S, S, generator=g) > 0.2).float()
        batches.append([x, torch.cat([1 - salt, salt], dim=1)])
        for j in range(B):
            m = (salt[j, 0, 13:13 + T, 14:14 + T].numpy() * 255).astype(np.uint8)      # the crop of postprocessing.py:24-38
            p = os.path.join(tmp, 'mask_%d_%d.png' % (i, j))
            Image.fromarray(m).save(p)
            paths.append(p)
    return batches, pd.DataFrame({'file_path_mask': paths})


@pytest.fixture()
def patched(monkeypatch):
    monkeypatch.setenv('SALT_ENGINE_GRAPH', '0')
    monkeypatch.setenv('SALT_ENGINE_MAX_BATCH', str(B))
    monkeypatch.setenv('SALT_ENGINE_SIZE', str(S))
    monkeypatch.setenv('SALT_ENGINE_LOSS', 'lovasz')
    ref_shims.install()
    from salt_b200 import models
    monkeypatch.setattr(models, 'UNetEngine', StubEngine)
    monkeypatch.setattr(models, '_PinnedSink', ListSink)
    StubEngine.instances.clear()
    return models


def test_reference_callbacks_and_finetune_step_drive_the_dropin(patched, tmp_path):
    import common_blocks.callbacks as cbk
    from common_blocks.utils import FineTuneStep
    from steppy.adapter import Adapter, E
    tmp = str(tmp_path)
    train, _ = _data(tmp, 3, seed=1)
    valid, meta_valid = _data(tmp, 2, seed=2)
    model = patched.SegmentationModel(ARCH, {'epochs': 3}, _callbacks_config(tmp, patience=0))

    # the reference's own callback objects, in the reference's order (models.py:309-312)
    names = [type(c).__name__ for c in model.callbacks.callbacks]
    assert names == ['ExperimentTiming', 'TrainingMonitor', 'ValidationMonitor', 'ModelCheckpoint', 'ReduceLROnPlateauScheduler',
                     'NeptuneMonitor', 'EarlyStopping']
    assert all(type(c).__module__ == 'common_blocks.callbacks' for c in model.callbacks.callbacks)
    assert isinstance(model.optimizer, torch.optim.Optimizer)

    step = FineTuneStep(name='network', transformer=model, experiment_directory=tmp, input_data=['callback_input', 'loader'],
                        adapter=Adapter({'datagen': E('loader', 'datagen'), 'validation_datagen': E('loader', 'validation_datagen'),
                                         'meta_valid': E('callback_input', 'meta_valid')}),
                        is_trainable=True, fine_tuning=False)
    data = {'loader': {'datagen': (train, len(train)), 'validation_datagen': (valid, len(valid))},
            'callback_input': {'meta_valid': meta_valid}}
    out = step.fit_transform(data)

    eng = StubEngine.instances[-1]
    sched = [c for c in model.callbacks.callbacks if isinstance(c, cbk.ReduceLROnPlateauScheduler)][0]
    assert isinstance(sched.lr_scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau) and sched.optimizer is model.optimizer
    # 3 epochs x 3 batches trained; validation ran every epoch through the reference's ValidationMonitor
    assert eng.step_count == 9 and sorted(model.validation_loss) == [0, 1, 2]
    for ep, v in model.validation_loss.items():
        assert set(v) == {'sum', 'iou', 'iout'} and all(tuple(t.shape) == (1,) for t in v.values())
        assert 0.0 <= float(v['iou']) <= 1.0
    # validation losses are distinct tensors holding distinct values (the loss of a later call must not leak into an earlier one)
    vals = [float(model.validation_loss[e]['sum']) for e in (0, 1, 2)]
    assert len(set(vals)) == 3, vals
    # the scheduler's decisions reach the engine: lr used by step() follows optimizer.param_groups[0]['lr']
    assert eng.lrs[0] == pytest.approx(1e-2) and all(a >= b for a, b in zip(eng.lrs, eng.lrs[1:]))
    assert model.optimizer.state_dict()['param_groups'][0]['lr'] == model.optimizer.param_groups[0]['lr']
    # ModelCheckpoint -> persist_torch_model(self.model, filepath): DataParallel-style keys
    ck = torch.load(os.path.join(tmp, 'checkpoints', 'network', 'best.torch'))
    assert set(ck) == {'module.final.1.weight', 'module.final.1.bias'}
    # FineTuneStep persisted the transformer and returned the transform() dict: one (2,S,S) float32 probability map per image
    assert os.path.exists(step.exp_dir_transformers_step)
    preds = out['mask_prediction']
    assert len(preds) == 3 * B and preds[0].shape == (2, S, S) and preds[0].dtype == np.float32
    ref = torch.sigmoid(eng.forward(torch.cat([b[0] for b in train]))).numpy()
    assert np.allclose(np.stack(preds), ref, atol=1e-6)

    # second run: the transformer is cached -> FineTuneStep loads it into a fresh model and only transforms (utils.py:463-467)
    fresh = patched.SegmentationModel(ARCH, {'epochs': 3}, _callbacks_config(tmp))
    step2 = FineTuneStep(name='network', transformer=fresh, experiment_directory=tmp, input_data=['callback_input', 'loader'],
                         adapter=step.adapter, is_trainable=True, fine_tuning=False)
    out2 = step2.fit_transform(data)
    assert StubEngine.instances[-1].step_count == 0
    assert np.allclose(np.stack(out2['mask_prediction']), np.stack(preds), atol=1e-6)


def test_reduce_lr_on_plateau_mutates_engine_lr(patched, tmp_path):
    """A validation loss that cannot improve (lr = 0 -> frozen parameters) makes ReduceLROnPlateau halve the rate every epoch
    (patience 0); the halved rate is what the engine's Adam kernel receives."""
    tmp = str(tmp_path)
    train, _ = _data(tmp, 1, seed=1)
    valid, meta_valid = _data(tmp, 1, seed=2)
    arch = dict(ARCH, optimizer_params={'lr': 0.0})
    model = patched.SegmentationModel(arch, {'epochs': 3}, _callbacks_config(tmp, patience=0))
    model.optimizer.param_groups[0]['lr'] = 1e-3
    eng = StubEngine.instances[-1]
    eng.adam_step = lambda lr=0, **k: eng.lrs.append(lr)          # frozen parameters: the metric never improves
    model.fit(datagen=(train, 1), validation_datagen=(valid, 1), meta_valid=meta_valid)
    assert eng.lrs == pytest.approx([1e-3, 1e-3, 5e-4]), eng.lrs        # epoch 0 sets the best, epochs 1 and 2 are "plateau" epochs
    assert model.optimizer.param_groups[0]['lr'] == pytest.approx(2.5e-4)


def test_configuration_errors_propagate_and_rank_gating(patched, tmp_path):
    tmp = str(tmp_path)
    bad = _callbacks_config(tmp)
    del bad['early_stopping']
    with pytest.raises(KeyError):
        patched.SegmentationModel(ARCH, {'epochs': 1}, bad)          # a mistyped config must not silently train without callbacks
    assert isinstance(patched.callbacks_network({}), patched._NullCallbacks)

    class DP:
        world, rank = 2, 1
    kept = [type(c).__name__ for c in patched.callbacks_network(_callbacks_config(tmp), DP()).callbacks]
    assert kept == ['ValidationMonitor', 'ReduceLROnPlateauScheduler', 'EarlyStopping']
    DP.rank = 0
    assert len(patched.callbacks_network(_callbacks_config(tmp), DP()).callbacks) == 7


def test_loss_wrapper_never_reuses_a_buffer_of_another_batch_size(patched):
    model = patched.SegmentationModel(ARCH, {'epochs': 1}, {})
    name, fn, w = model.loss_function[0]
    g = torch.Generator().manual_seed(0)
    l1 = fn(torch.randn(1, 2, S, S, generator=g), torch.zeros(1, 2, S, S))
    d1 = fn.dlogits
    l2 = fn(torch.randn(2, 2, S, S, generator=g), torch.zeros(2, 2, S, S))
    assert tuple(fn.dlogits.shape) == (2, 2, S, S) and tuple(d1.shape) == (1, 2, S, S)
    assert l1.data_ptr() != l2.data_ptr() and tuple(l1.shape) == (1,)
    with pytest.raises(ValueError):
        fn(torch.randn(3, 2, S, S, generator=g), torch.zeros(3, 2, S, S))       # > max_batch


def test_architecture_names_of_the_reference(patched):
    """`model_params['architecture']` selects among the reference's ARCHITECTURES entries this engine implements (models.py:15-30):
    the engine receives the name and the entry's default encoder_depth; every other name - implemented by the reference or not -
    raises NotImplementedError like the reference does for an unknown one (models.py:188)."""
    import copy
    seen = []
    orig_init = StubEngine.__init__

    def spy(self, architecture=None, encoder_depth=34, **kw):
        seen.append((architecture, encoder_depth))
        orig_init(self, architecture=architecture, encoder_depth=encoder_depth, **kw)

    StubEngine.__init__ = spy
    try:
        for name, depth in (('UNetResNet', 34), ('UNetSeResNet', 50), ('UNetSeResNetXt', 50)):
            arch = copy.deepcopy(ARCH)
            arch['model_params']['architecture'] = name
            arch['model_params'].pop('encoder_depth', None)
            patched.SegmentationModel(arch, {'epochs': 1}, {})
            assert seen[-1] == (name, depth)
        for name in ('PSPNet', 'LargeKernelMatters', 'UNetDenseNet', 'NoSuchNet'):
            arch = copy.deepcopy(ARCH)
            arch['model_params']['architecture'] = name
            with pytest.raises(NotImplementedError):
                patched.SegmentationModel(arch, {'epochs': 1}, {})
    finally:
        StubEngine.__init__ = orig_init
