"""Import the UNMODIFIED reference modules from /root/reference in the build container.

TEST INFRASTRUCTURE, build-container only: /root/reference does not exist on the GPU
box, so nothing that runs there imports this file.  It is used by
``oracle/make_golden.py`` (to write tests/golden/*.npz) and by the CPU-only tests
that are skipped when the reference tree is absent.

The reference needs packages that are not installed offline (steppy, steppy-toolkit,
pretrainedmodels, attrdict, imgaug, neptune, skimage, pycocotools, matplotlib).  None of
them is on the arithmetic path; they are replaced by inert stubs.
"""
import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SALT_REFERENCE_ROOT', '/root/reference')


class _Anything:
    """Inert stand-in: callable, subscriptable, attribute-chainable, usable as a base class."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, item):
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return type(name, (_Anything,), {})


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = _StubModule(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent:
        setattr(_stub(parent), child, m)
    return m


class BaseTransformer:
    """steppy.base.BaseTransformer as inferred from its call sites (utils.py:444-486)."""

    def fit(self, *a, **k):
        return self

    def transform(self, *a, **k):
        raise NotImplementedError

    def fit_transform(self, *a, **k):
        self.fit(*a, **k)
        return self.transform(*a, **k)

    def load(self, path):
        return self

    def persist(self, path):
        pass


class Model(BaseTransformer):
    """toolkit.pytorch_transformers.models.Model as inferred from models.py:67-76:
    stores the three config dicts and nulls the attributes the subclass fills."""

    def __init__(self, architecture_config, training_config, callbacks_config):
        super().__init__()
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

    def persist(self, filepath):
        import torch
        torch.save(self.model.state_dict(), filepath)


def _install_cocomask_stub():
    """pycocotools.mask as used by utils.py:292-305 / metrics.py:21-35 (package absent): `encode` of a dense uint8 mask and
    `iou` of two lists of such encodings with iscrowd = 0, i.e. |a & b| / |a | b|.  Dense bytes stand in for the RLE string."""
    import numpy as np
    m = sys.modules['pycocotools.mask']
    if not isinstance(m, _StubModule):
        return

    def encode(mask):
        mask = np.asarray(mask)
        return {'size': list(mask.shape), 'counts': bytes(np.ascontiguousarray(mask).astype(np.uint8).tobytes())}

    def _dense(seg):
        c = seg['counts']
        c = c.encode('UTF-8') if isinstance(c, str) else c
        return np.frombuffer(c, dtype=np.uint8).reshape(seg['size']) != 0

    def iou(dt, gt, iscrowd):
        # cocoapi maskApi.c rleIou: rows follow the first list, columns the second
        if len(dt) == 0 or len(gt) == 0:
            return []
        out = np.zeros((len(dt), len(gt)))
        for i, a in enumerate(dt):
            for j, b in enumerate(gt):
                a_, b_ = _dense(a), _dense(b)
                u = np.logical_or(a_, b_).sum()
                out[i, j] = np.logical_and(a_, b_).sum() / u if u else 0.0
        return out

    m.__dict__['encode'] = encode
    m.__dict__['iou'] = iou


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'common_blocks'))


_installed = False


def install():
    """Put the stubs in sys.modules and the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    if not hasattr(collections, 'Iterable'):
        collections.Iterable = collections.abc.Iterable      # common_blocks/utils.py:7
    for name in ('pretrainedmodels', 'attrdict', 'imgaug', 'imgaug.augmenters', 'neptune', 'deepsense',
                 'deepsense.neptune', 'skimage', 'skimage.transform', 'skimage.morphology',
                 'pycocotools', 'pycocotools.mask', 'matplotlib', 'matplotlib.pyplot', 'cv2',
                 'steppy', 'steppy.base', 'steppy.adapter', 'steppy.utils',
                 'toolkit', 'toolkit.pytorch_transformers', 'toolkit.pytorch_transformers.models',
                 'toolkit.pytorch_transformers.utils', 'toolkit.pytorch_transformers.validation',
                 'toolkit.sklearn_transformers', 'toolkit.sklearn_transformers.models',
                 'toolkit.preprocessing', 'toolkit.preprocessing.misc'):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if isinstance(sys.modules['pretrainedmodels'], _StubModule):
        # encoders.py:52-53 looks the constructor up in the package __dict__; the package is absent, use the restatement
        from . import senet_restated
        for fn in ('se_resnet50', 'se_resnet101', 'se_resnet152', 'se_resnext50_32x4d', 'se_resnext101_32x4d'):
            sys.modules['pretrainedmodels'].__dict__[fn] = getattr(senet_restated, fn)
    _install_cocomask_stub()
    sys.modules['steppy.base'].BaseTransformer = BaseTransformer
    sys.modules['toolkit.pytorch_transformers.models'].Model = Model
    _install_callers()
    import joblib
    import sklearn
    ext = _stub('sklearn.externals')
    ext.joblib = joblib                                      # common_blocks/utils.py:19
    sklearn.externals = ext
    sys.modules['sklearn.externals.joblib'] = joblib
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_unet(depth, num_classes=2, arch=None):
    """common_blocks.architectures.unet.UNetResNet(pretrained=False, hypercolumn, pool0=False); depth 50 builds
    unet.UNetSeResNet - or, arch='UNetSeResNetXt', unet.UNetSeResNetXt - on top of oracle/senet_restated.py (the reference's own
    wrapper classes, unmodified)."""
    install()
    import warnings
    from common_blocks.architectures import unet
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if arch == 'UNetSeResNetXt':
            return unet.UNetSeResNetXt(encoder_depth=depth, num_classes=num_classes, dropout_2d=0.0, pretrained=None,
                                       use_hypercolumn=True, pool0=False)
        if depth >= 50:
            return unet.UNetSeResNet(encoder_depth=depth, num_classes=num_classes, dropout_2d=0.0, pretrained=None,
                                     use_hypercolumn=True, pool0=False)
        return unet.UNetResNet(encoder_depth=depth, num_classes=num_classes, dropout_2d=0.0,
                               pretrained=False, use_hypercolumn=True, pool0=False)


def load_numpy_state(module, sd_np, depth):
    """Copy a canonical numpy state into a reference nn.Module (aliases share storage)."""
    import torch
    own = module.state_dict()
    with torch.no_grad():
        for k, v in sd_np.items():
            own[k].copy_(torch.from_numpy(v))
    return module


# ---------------------------------------------------------------------------------------------------------------------
# steppy==0.1.6 / steppy-toolkit==0.1.5 pieces that the reference's CALLERS of the hot path need (both packages are absent from
# the image and from /root/reference).  Restated from their call sites, only as far as those call sites go:
#   utils.py:415-486 FineTuneStep(Step)            -> Step: constructor arguments, exp_dir_* paths, transformer_is_cached
#   callbacks.py:832-866 postprocessing_pipeline   -> Step.transform(data) recursion over input_steps, Adapter / E, IdentityOperation
#   callbacks.py:16-17,72-76,150-161,776-794       -> Averager, persist_torch_model, score_model
# They let tests/test_boundary_reference_cpu.py drive the UNMODIFIED callbacks_network / FineTuneStep against the drop-in.
# ---------------------------------------------------------------------------------------------------------------------
class E:
    """steppy.adapter.E(input_name, key): 'take `key` from the output dict of step / data entry `input_name`'."""

    def __init__(self, input_name, key):
        self.input_name, self.key = input_name, key


class Adapter:
    """steppy.adapter.Adapter({kwarg: E(...)}) - maps upstream outputs to the transformer's keyword arguments."""

    def __init__(self, adapting_recipes):
        self.adapting_recipes = adapting_recipes

    def adapt(self, all_ouputs):
        out = {}
        for name, recipe in self.adapting_recipes.items():
            if isinstance(recipe, E):
                out[name] = all_ouputs[recipe.input_name][recipe.key]
            elif isinstance(recipe, (list, tuple)):
                out[name] = type(recipe)(all_ouputs[r.input_name][r.key] if isinstance(r, E) else r for r in recipe)
            else:
                out[name] = recipe
        return out


class IdentityOperation(BaseTransformer):
    def transform(self, **kwargs):
        return kwargs


class Step:
    """steppy.base.Step as used by utils.py:415-486 and callbacks.py:832-866 (no output caching, no structure persistence)."""

    def __init__(self, name, transformer, experiment_directory, input_data=None, input_steps=None, adapter=None,
                 is_trainable=False, cache_output=False, persist_output=False, load_persisted_output=False, force_fitting=False,
                 persist_upstream_pipeline_structure=False):
        self.name, self.transformer = name, transformer
        self.input_steps, self.input_data, self.adapter = input_steps or [], input_data or [], adapter
        self.is_trainable, self.force_fitting = is_trainable, force_fitting
        self.cache_output, self.persist_output, self.load_persisted_output = cache_output, persist_output, load_persisted_output
        self.exp_dir = os.path.join(experiment_directory)
        self.exp_dir_transformers = os.path.join(self.exp_dir, 'transformers')
        self.exp_dir_outputs = os.path.join(self.exp_dir, 'outputs')
        self.exp_dir_cache = os.path.join(self.exp_dir, 'cache')
        for d in (self.exp_dir_transformers, self.exp_dir_outputs, self.exp_dir_cache):
            os.makedirs(d, exist_ok=True)
        self.exp_dir_transformers_step = os.path.join(self.exp_dir_transformers, name)
        self.exp_dir_outputs_step = os.path.join(self.exp_dir_outputs, '{}'.format(name))
        self.exp_dir_cache_step = os.path.join(self.exp_dir_cache, '{}'.format(name))

    @property
    def transformer_is_cached(self):
        return os.path.exists(self.exp_dir_transformers_step)

    def _step_inputs(self, data, how):
        outputs = {s.name: getattr(s, how)(data) for s in self.input_steps}
        for name in self.input_data:
            outputs[name] = data[name]
        if self.adapter is not None:
            return self.adapter.adapt(outputs)
        merged = {}
        for o in outputs.values():
            merged.update(o)
        return merged

    def fit_transform(self, data):
        return self._cached_fit_transform(self._step_inputs(data, 'fit_transform'))

    def transform(self, data):
        return self._cached_transform(self._step_inputs(data, 'transform'))

    def _cached_fit_transform(self, step_inputs):
        if self.is_trainable:
            out = self.transformer.fit_transform(**step_inputs)
            self.transformer.persist(self.exp_dir_transformers_step)
            return out
        return self.transformer.transform(**step_inputs)

    def _cached_transform(self, step_inputs):
        if self.is_trainable:
            self.transformer.load(self.exp_dir_transformers_step)
        return self.transformer.transform(**step_inputs)

    def _persist_output(self, output_data, filepath):
        import joblib
        joblib.dump(output_data, filepath)


class Averager:
    """toolkit.pytorch_transformers.utils.Averager: running mean fed by send() (callbacks.py:150-161,342-354)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.currently_sum, self.count = 0.0, 0

    def send(self, value):
        self.currently_sum += value
        self.count += 1

    @property
    def value(self):
        return self.currently_sum / self.count if self.count else 0.0


def persist_torch_model(model, path):
    """toolkit.pytorch_transformers.utils.persist_torch_model as called at callbacks.py:791."""
    import torch
    model.eval()
    torch.save(model.state_dict(), path)
    model.train()


def score_model(model, loss_function, datagen):
    """toolkit.pytorch_transformers.validation.score_model as called at callbacks.py:72-76: mean of the weighted loss over
    the validation batches -> {'sum': 1-element tensor}."""
    batch_gen, steps = datagen
    losses = []
    for batch_id, data in enumerate(batch_gen):
        X, targets = data[0], data[1:]
        outputs = model(X)
        (name, fn, weight), target = loss_function[0], targets[0]
        losses.append(fn(outputs, target) * weight)
        if batch_id == steps:
            break
    return {'sum': sum(losses) / steps}


def _install_callers():
    """the restated steppy / toolkit classes, so that common_blocks.callbacks / common_blocks.utils.FineTuneStep run (part of install())."""
    sb, sa = sys.modules['steppy.base'], sys.modules['steppy.adapter']
    sb.Step, sb.IdentityOperation, sb.BaseTransformer = Step, IdentityOperation, BaseTransformer
    sa.Adapter, sa.E = Adapter, E
    tu, tv = sys.modules['toolkit.pytorch_transformers.utils'], sys.modules['toolkit.pytorch_transformers.validation']
    tu.Averager, tu.persist_torch_model = Averager, persist_torch_model
    tv.score_model = score_model
