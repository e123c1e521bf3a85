"""ctypes binding of libsaltunet.so (C ABI declared in include/saltunet.h).

There is deliberately no fallback: if the CUDA library is missing the import raises, and if no GPU is
present every compute entry point returns an error that is raised as ``SaltEngineError``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SALT_LIB_PATH') or os.path.join(os.path.dirname(_HERE), 'libsaltunet.so')      # SALT_LIB_PATH: profiling builds

PREC_FP32, PREC_BF16 = 0, 1
ARCH_UNET_RESNET = 0
ARCH_UNET_SERESNET = 1
ARCH_UNET_SERESNEXT = 2


class SaltEngineError(RuntimeError):
    pass


class SaltConfig(C.Structure):
    _fields_ = [('arch', C.c_int), ('encoder_depth', C.c_int), ('num_classes', C.c_int), ('max_batch', C.c_int),
                ('height', C.c_int), ('width', C.c_int), ('precision', C.c_int), ('use_tensor_cores', C.c_int)]


class SaltConvDesc(C.Structure):
    _fields_ = [('batch', C.c_int), ('in_h', C.c_int), ('in_w', C.c_int), ('in_c', C.c_int),
                ('out_h', C.c_int), ('out_w', C.c_int), ('out_c', C.c_int),
                ('kernel', C.c_int), ('stride', C.c_int), ('pad', C.c_int),
                ('precision', C.c_int), ('use_tensor_cores', C.c_int)]


_vp, _fp, _dp, _sz, _i, _f = C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float

# name -> (restype, argtypes); mirrors include/saltunet.h one to one
PROTOTYPES = {
    'salt_last_error': (C.c_char_p, []),
    'salt_version': (C.c_char_p, []),
    'salt_create': (_i, [C.POINTER(SaltConfig), C.POINTER(_vp)]),
    'salt_destroy': (None, [_vp]),
    'salt_param_floats': (_sz, [_vp]),
    'salt_buffer_floats': (_sz, [_vp]),
    'salt_workspace_bytes': (_sz, [_vp]),
    'salt_num_tensors': (_i, [_vp]),
    'salt_tensor_info': (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz),
                              C.POINTER(_i)]),
    'salt_bind': (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _vp, _sz]),
    'salt_params_changed': (_i, [_vp]),
    'salt_forward': (_i, [_vp, _fp, _i, _fp, _i, _vp]),
    'salt_loss_lovasz': (_i, [_vp, _fp, _fp, _i, _fp, _fp, _vp]),
    'salt_loss_bce_dice_reduce': (_i, [_vp, _fp, _fp, _i, _dp, _vp]),
    'salt_loss_bce_dice_finish': (_i, [_vp, _fp, _fp, _i, _dp, C.c_double, _f, _fp, _fp, _vp]),
    'salt_backward': (_i, [_vp, _fp, _vp]),
    'salt_backward_segment': (_i, [_vp, _fp, _i, _vp]),
    'salt_grad_segment': (_i, [_vp, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    'salt_adam_step': (_i, [_vp, _f, _f, _f, _f, _f, _i, _f, _vp]),
    'salt_predict': (_i, [_vp, _fp, _fp, _i, _i, _f, _fp, _vp, _vp]),
    'salt_adapt_tiles': (_i, [_vp, _i, _i, _i, _i, _f, _f, _i, _fp, _vp]),
    'salt_forward_tiles': (_i, [_vp, _vp, _i, _i, _i, _f, _f, _i, _fp, _i, _vp]),
    'salt_rle_encode': (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    'salt_validation_counts': (_i, [_fp, _fp, _i, _i, _i, _i, _vp, C.POINTER(C.c_double), _i, _vp, _vp, _vp, _vp]),
    'salt_get_activation': (_i, [_vp, C.c_char_p, _fp, C.POINTER(_i), _vp]),
    'salt_launch_count': (C.c_ulonglong, []),
    'salt_cluster_launch_count': (C.c_ulonglong, []),
    'salt_profile_enable': (_i, [_vp, _i]),
    'salt_profile_read': (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    'salt_profile_read_group': (_i, [_vp, _i, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    'salt_profile_records': (C.c_longlong, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_longlong]),
    'salt_op_conv_forward': (_i, [C.POINTER(SaltConvDesc), _vp, _fp, _fp, _vp, _dp, _vp]),
    'salt_op_conv_dgrad': (_i, [C.POINTER(SaltConvDesc), _vp, _fp, _vp, _i, _vp]),
    'salt_op_conv_wgrad': (_i, [C.POINTER(SaltConvDesc), _vp, _vp, _fp, _vp]),
    'salt_op_adam': (_i, [_fp, _fp, _fp, _fp, _sz, _f, _f, _f, _f, _f, _i, _f, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises SaltEngineError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SaltEngineError('%s not found: build it with `python __graft_entry__.py` (nvcc, sm_100a); '
                              'there is no CPU fallback' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise SaltEngineError(load().salt_last_error().decode())


_replayed = 0


def count_replayed(n):
    """Kernel launches executed by CUDA-graph replays (the library counter only sees launches issued through its host code)."""
    global _replayed
    _replayed += int(n)


def launch_count():
    return int(load().salt_launch_count()) + _replayed
