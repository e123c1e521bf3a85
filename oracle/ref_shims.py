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
        for fn in ('se_resnet50', 'se_resnet101', 'se_resnet152'):
            sys.modules['pretrainedmodels'].__dict__[fn] = getattr(senet_restated, fn)
    _install_cocomask_stub()
    sys.modules['steppy.base'].BaseTransformer = BaseTransformer
    sys.modules['toolkit.pytorch_transformers.models'].Model = Model
    import joblib
    import sklearn
    ext = _stub('sklearn.externals')
    ext.joblib = joblib                                      # common_blocks/utils.py:19
    sklearn.externals = ext
    sys.modules['sklearn.externals.joblib'] = joblib
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def reference_unet(depth, num_classes=2):
    """common_blocks.architectures.unet.UNetResNet(pretrained=False, hypercolumn, pool0=False); depth 50 builds
    unet.UNetSeResNet on top of oracle/senet_restated.py (the reference's own wrapper classes, unmodified)."""
    install()
    import warnings
    from common_blocks.architectures import unet
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
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
