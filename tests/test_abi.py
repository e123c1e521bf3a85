"""CPU: the C-ABI library builds, loads, and exports every symbol include/saltunet.h declares; plan
construction (no device work) reproduces the reference state_dict table; compute entry points fail loudly
without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    entry.build()
    from salt_b200 import _lib
    return _lib.load()


def test_exports_match_header(lib):
    from salt_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'saltunet.h')).read()
    declared = set(re.findall(r'\b(salt_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    for name in declared:
        assert hasattr(lib, name), 'library does not export %s' % name
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)


def test_plan_table_matches_reference_state_dict(lib):
    from salt_b200 import _lib
    from oracle import synth
    for depth in (18, 34):
        cfg = _lib.SaltConfig(0, depth, 2, 4, 128, 128, _lib.PREC_BF16, 1)
        h = C.c_void_p()
        _lib.check(lib.salt_create(C.byref(cfg), C.byref(h)))
        name = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        ndim, isbuf, off, numel = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        got = {}
        for i in range(lib.salt_num_tensors(h)):
            _lib.check(lib.salt_tensor_info(h, i, name, 256, shape, C.byref(ndim), C.byref(off), C.byref(numel), C.byref(isbuf)))
            got[name.value.decode()] = tuple(shape[:ndim.value])
        want = {n: s for n, s, _ in synth.param_specs(depth, 2)}
        assert got == want
        assert lib.salt_workspace_bytes(h) > 0
        lib.salt_destroy(h)


def test_plan_table_seresnet_depth_variants(lib):
    """UNetSeResNet with the three encoder depths the reference class accepts (encoders.py:52-59): state_dict keys and shapes."""
    from salt_b200 import _lib
    from oracle import synth
    for depth in (50, 101, 152):
        cfg = _lib.SaltConfig(_lib.ARCH_UNET_SERESNET, depth, 2, 2, 64, 64, _lib.PREC_BF16, 1)
        h = C.c_void_p()
        _lib.check(lib.salt_create(C.byref(cfg), C.byref(h)))
        name = C.create_string_buffer(256)
        shape = (C.c_int * 4)()
        ndim, isbuf, off, numel = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        got = {}
        for i in range(lib.salt_num_tensors(h)):
            _lib.check(lib.salt_tensor_info(h, i, name, 256, shape, C.byref(ndim), C.byref(off), C.byref(numel), C.byref(isbuf)))
            got[name.value.decode()] = tuple(shape[:ndim.value])
        assert got == {n: s for n, s, _ in synth.param_specs(depth, 2)}
        lib.salt_destroy(h)
    cfg = _lib.SaltConfig(_lib.ARCH_UNET_SERESNET, 34, 2, 2, 64, 64, _lib.PREC_BF16, 1)
    h = C.c_void_p()
    assert lib.salt_create(C.byref(cfg), C.byref(h)) != 0 and b'50, 101 or 152' in lib.salt_last_error()


def test_bad_config_is_rejected(lib):
    from salt_b200 import _lib
    h = C.c_void_p()
    cfg = _lib.SaltConfig(0, 50, 2, 4, 128, 128, 0, 0)
    assert lib.salt_create(C.byref(cfg), C.byref(h)) != 0
    assert b'18 and 34' in lib.salt_last_error()
    cfg = _lib.SaltConfig(0, 34, 2, 4, 100, 100, 0, 0)
    assert lib.salt_create(C.byref(cfg), C.byref(h)) != 0
    assert b'multiples of 32' in lib.salt_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib):
    from salt_b200 import _lib
    from salt_b200.engine import UNetEngine
    with pytest.raises(_lib.SaltEngineError):
        UNetEngine(18, 2, 2, 64)
    d = _lib.SaltConvDesc(1, 8, 8, 4, 8, 8, 4, 3, 1, 1, 0, 0)
    assert lib.salt_op_conv_forward(C.byref(d), None, None, None, None, None, None) != 0
    assert b'no CUDA device' in lib.salt_last_error()
